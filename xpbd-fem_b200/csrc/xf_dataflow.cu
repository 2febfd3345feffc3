// XF_SCHEDULE_DATAFLOW — the persistent sweep without any barrier.
//
// The colour schedule (xf_prepare.cpp) is kept — same elements, same equivalent serial order, same bits — but the
// grid-wide barrier between colours is replaced by *versioned vertex records*.  A vertex record is one 32-byte L2
// sector {x, y, z, w, flags} that always moves with ONE 256-bit strong access (LDG/STG.E.ENL2.256.STRONG.GPU), so a
// reader sees either the old or the new record, never a mix.  The upper 24 bits of `flags` carry the *stage* that
// wrote the record:
//     stage(substep s, vertex phase) = base + s*(nC+1)
//     stage(substep s, colour c)     = base + s*(nC+1) + 1 + c
// Every writer of a vertex knows which stage wrote it last: for an element that is the colour of the previous element
// around the vertex (or the vertex phase), precomputed on the host and carried in the top byte of each vertex index;
// for the vertex phase it is the last colour around the vertex (`lastCode`).  A thread gathers its four records and
// simply re-gathers until all four carry the expected stage, then solves and scatters the records stamped with its own
// stage.  Elements touching one vertex are totally ordered by colour, so between "predecessor wrote" and "I write"
// nobody else touches the record: no fences, no flags, no atomics, no barrier of any scope.
//
// Deadlock freedom: the kernel is launched cooperatively (all CTAs co-resident); every thread walks its work in stage
// order and every item depends only on items of strictly earlier stages, so the unfinished item of minimal stage is
// always runnable by a thread that is not waiting for anything else.
//
// Why: at 1M tets a colour holds only ~41.6k elements (9 warps per SM); the per-colour cost of the barrier schedule was
// 1.2 us of grid barrier + arrival skew + ~1 us element latency, 25 times per substep.  Here the only serialisation left
// is the true data dependence: predecessor's store -> L2 -> my load.
//
// Covers the main sweep (all energies / solve modes, undamped in-constraint) + the fused vertex phase.  Calls with volume passes or
// post-solve damping sweeps run on k_substeps_dataflow_general (xf_dataflow_general.cu, same protocol); only in-constraint Rayleigh
// damping (which reads O of other threads' vertices) runs on XF_SCHEDULE_PERSISTENT, per call.
#include "xf_dataflow.cuh"

namespace xf {

template <int ENERGY, bool SIMUL, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_substeps_dataflow(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                              uint32_t verBase, uint32_t tuning) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t gsize = gridDim.x * blockDim.x;
	// work is dealt in warp-sized chunks round-robin over the CTAs (chunk k -> CTA k % grid, warp k / grid)
	const uint32_t warpSlot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u;
	const uint32_t nC = p.nColors;
	const uint32_t stride = nC + 1u;
	// packed tuning word: bits 0-15 vertex-phase back-off (ns), bits 16-31 element back-off (ns)
	const uint32_t sleepNs = tuning & 0x7fffu, elemSleepNs = tuning >> 16;
	const bool prefetch = (tuning & 0x8000u) == 0;
	// a warp whose first chunk lies beyond the largest colour never runs an element: it only shares the vertex phase, and must
	// not spend issue slots walking the colours next to the warps that are on the dependence chain
	const bool spare = warpSlot >= sc.maxColorSize;
	ElemRec rec;
	bool dead = false; // gave up on a stalled record (SpinGiveUp): touch nothing any more, just run out of the loops
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps; // closing post phase (locks, manipulator, velocities) of the last substep
		const uint32_t stageBase = verBase + s * stride;
		if (!closing && p.colorStart[0] + warpSlot + lane < p.colorStart[1]) { DataflowLoad<ENERGY, EXACT>(sc, p.colorStart[0] + warpSlot + lane, rec); }
		// vertex phase: post of the previous substep + predict, once the vertex's last element of that substep has written
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) { // warp-uniform trip count
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				dead = !DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs, s > 0 ? VaryRow(p, s - 1u) : nullptr);
			}
		}
		if (closing) { break; }
		if (spare) { continue; }
		for (uint32_t c = 0; c < nC; c++) {
			const uint32_t end = p.colorStart[c + 1];
			uint32_t e0 = p.colorStart[c] + warpSlot;
			if (e0 < end) {
				const bool has = e0 + lane < end && !dead;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) { dead = !DataflowElement<ENERGY, SIMUL, EXACT>(sc, p, rec, mask, stageBase, c, elemSleepNs); }
			}
			for (e0 += gsize; e0 < end; e0 += gsize) {
				const bool has = e0 + lane < end && !dead;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) {
					ElemRec more;
					DataflowLoad<ENERGY, EXACT>(sc, e0 + lane, more);
					dead = !DataflowElement<ENERGY, SIMUL, EXACT>(sc, p, more, mask, stageBase, c, elemSleepNs);
				}
			}
			if (c + 1 < nC && p.colorStart[c + 1] + warpSlot + lane < p.colorStart[c + 2]) {
				DataflowLoad<ENERGY, EXACT>(sc, p.colorStart[c + 1] + warpSlot + lane, rec);
			}
			if (prefetch) {
				const uint32_t c2 = c + 2 < nC ? c + 2 : c + 2 - nC; // wraps into the next substep
				if (p.colorStart[c2] + warpSlot + lane < p.colorStart[c2 + 1]) { DataflowPrefetch<ENERGY, EXACT>(sc, p.colorStart[c2] + warpSlot + lane); }
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Clustered variant (HostMesh::groupSize > 1, see ColorClustered in xf_prepare.cpp): one thread solves the elements of
// one cluster back to back.  The colours come in groups of G = groupSize; element k of colour C*G + t is the t-th
// element of cluster k of cluster-colour C.  A vertex is gathered from L2 (waiting for its stage tag) at its FIRST use in
// the cluster, lives in the thread's private shared-memory slots between uses, and is scattered - stamped with the stage
// of the element that used it last, which is what every later reader expects - at its LAST use.  Dependences inside a
// cluster never leave the thread; the chain through L2 has one link per cluster colour (8 on a MeshGen lattice)
// instead of one per colour (24), and a vertex costs one gather and one scatter per cluster instead of per element.
// ------------------------------------------------------------------------------------------------
constexpr int kClusterSlots = 8;

template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool ClusterElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, uint32_t info, VertexRec* cache,
                                               unsigned mask, uint32_t stageBase, uint32_t c) {
	const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	uint32_t vid[4], expectTag[4];
	VertexRec* slot[4];
	bool first[4], last[4];
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		vid[n] = raw[n] & 0x00ffffffu;
		expectTag[n] = (stageBase + (raw[n] >> 24)) << 8;
		const uint32_t bits = info >> (5 * n);
		slot[n] = cache + (bits & 7u) * blockDim.x;
		first[n] = (bits & 8u) != 0;
		last[n] = (bits & 16u) != 0;
		if (first[n]) {
			v[n] = LoadVertex(sc.Xw, vid[n]);
		} else {
			const VertexRec r = *slot[n];
			v[n].x[0] = r.x; v[n].x[1] = r.y; v[n].x[2] = r.z;
			v[n].w = r.w;
			v[n].flags = r.flags;
		}
	}
	const ElemCompliance ec = DataflowCompliance<EXACT>(sc, p, rec);
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = !first[n] || (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!ok[n]) { v[n] = LoadVertex(sc.Xw, vid[n]); }
		}
	}
	const uint32_t newTag = (stageBase + 1u + c) << 8;
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n].flags = (v[n].flags & 0xffu) | newTag; }
	ElemRec r = rec;
	r.idx = make_uint4(vid[0], vid[1], vid[2], vid[3]);
	SolveElementGathered<ENERGY, SIMUL, EXACT, false>(NoStore{}, p, r, v, ec);
#pragma unroll
	for (int n = 0; n < 4; n++) {
		if (last[n]) {
			StoreVertex(sc.Xw, vid[n], v[n]);
		} else {
			*slot[n] = VertexRec{ v[n].x[0], v[n].x[1], v[n].x[2], v[n].w, v[n].flags };
		}
	}
	return true;
}

template <int ENERGY, bool SIMUL, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_substeps_cluster(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                             uint32_t verBase, uint32_t tuning) {
	extern __shared__ __align__(32) unsigned char clusterSmem[];
	VertexRec* cache = reinterpret_cast<VertexRec*>(clusterSmem) + threadIdx.x; // slot s of this thread: cache[s * blockDim.x]
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t gsize = gridDim.x * blockDim.x;
	const uint32_t warpSlot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u;
	const uint32_t nC = p.nColors, G = sc.groupSize;
	const uint32_t stride = nC + 1u;
	const uint32_t sleepNs = tuning & 0x7fffu;
	bool dead = false;
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		const uint32_t stageBase = verBase + s * stride;
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				dead = !DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs, s > 0 ? VaryRow(p, s - 1u) : nullptr);
			}
		}
		if (closing) { break; }
		for (uint32_t c0 = 0; c0 < nC; c0 += G) {
			const uint32_t n0 = p.colorStart[c0 + 1] - p.colorStart[c0]; // clusters of this cluster colour
			for (uint32_t k0 = warpSlot; k0 < n0; k0 += gsize) {
				// the records of this cluster (and of this thread's cluster of the next cluster colour) into L1
				for (uint32_t t = 0; t < G; t++) {
					const uint32_t e = p.colorStart[c0 + t] + k0 + lane;
					if (e < p.colorStart[c0 + t + 1]) {
						DataflowPrefetch<ENERGY, EXACT>(sc, e);
						asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eK + e));
					}
				}
				for (uint32_t t = 0; t < G; t++) {
					const uint32_t c = c0 + t;
					const uint32_t nt = p.colorStart[c + 1] - p.colorStart[c]; // clusters with more than t elements (sizes descend)
					if (k0 >= nt) { break; }
					const bool has = k0 + lane < nt && !dead;
					const unsigned mask = __ballot_sync(0xffffffffu, has);
					if (has) {
						const uint32_t e = p.colorStart[c] + k0 + lane;
						ElemRec rec;
						DataflowLoad<ENERGY, EXACT>(sc, e, rec);
						dead = !ClusterElement<ENERGY, SIMUL, EXACT>(sc, p, rec, __ldg(sc.eK + e), cache, mask, stageBase, c);
					}
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Chained variant (XF_GROUPING_CHAINS, ChainInfo in xf_prepare.cpp).  Same colouring, same stages, same one element per
// thread per colour as k_substeps_dataflow, and the same dependence chain: what changes is the traffic.  At 1M tets a stage
// is a burst of 41.6k elements x 8 records x 32 B = 10.6 MB through L2, ~1700 cycles at the ~6300 B/cycle the L2 slices
// deliver, out of a 4200-cycle stage.  The thread that ran position j of colour c runs position j of colour c+1; with the
// ring-ordered lattice hint that is the next tet around the same cell diagonal, which shares a face with the previous one.
// The three shared records (nobody else writes them in between: the previous-writer code says so) stay in the thread's four
// private shared-memory slots; only the fourth is gathered (and waited for) and only the record that leaves is scattered.
// Per cell: 9 gathers + 9 scatters instead of 24 + 24.  Slots are [slot][half][thread] 16-byte units: conflict-free
// LDS/STS.128, never touched by another thread, so no synchronisation.
// ------------------------------------------------------------------------------------------------
constexpr int kChainSlots = 4;

__device__ __forceinline__ VertexRegs ChainLoad(const uint4* cache, uint32_t slot) {
	const uint4 a = cache[(2u * slot) * 256u], b = cache[(2u * slot + 1u) * 256u];
	VertexRegs v;
	v.x[0] = __hiloint2double((int)a.y, (int)a.x);
	v.x[1] = __hiloint2double((int)a.w, (int)a.z);
	v.x[2] = __hiloint2double((int)b.y, (int)b.x);
	v.w = __uint_as_float(b.z);
	v.flags = b.w;
	return v;
}
__device__ __forceinline__ void ChainStore(uint4* cache, uint32_t slot, const VertexRegs& v) {
	cache[(2u * slot) * 256u] = make_uint4((uint32_t)__double2loint(v.x[0]), (uint32_t)__double2hiint(v.x[0]), (uint32_t)__double2loint(v.x[1]),
	                                       (uint32_t)__double2hiint(v.x[1]));
	cache[(2u * slot + 1u) * 256u] = make_uint4((uint32_t)__double2loint(v.x[2]), (uint32_t)__double2hiint(v.x[2]), __float_as_uint(v.w), v.flags);
}

template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool ChainElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, uint32_t info, uint4* cache,
                                             unsigned mask, uint32_t stageBase, uint32_t c, uint32_t sleepNs) {
	const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	uint32_t vid[4], expectTag[4], slot[4];
	bool first[4], last[4];
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		vid[n] = raw[n] & 0x00ffffffu;
		expectTag[n] = (stageBase + (raw[n] >> 24)) << 8;
		const uint32_t bits = info >> (5 * n);
		slot[n] = bits & 3u;
		first[n] = (bits & 8u) != 0;
		last[n] = (bits & 16u) != 0;
	}
	// the gathers go out first; the slot reads and the compliance divisions sit in their shadow
#pragma unroll
	for (int n = 0; n < 4; n++) {
		if (first[n]) { v[n] = LoadVertex(sc.Xw, vid[n]); }
	}
#pragma unroll
	for (int n = 0; n < 4; n++) {
		if (!first[n]) { v[n] = ChainLoad(cache, slot[n]); }
	}
	const ElemCompliance ec = DataflowCompliance<EXACT>(sc, p, rec);
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = !first[n] || (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
		if (sleepNs) { __nanosleep(sleepNs); }
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!ok[n]) { v[n] = LoadVertex(sc.Xw, vid[n]); }
		}
	}
	const uint32_t newTag = (stageBase + 1u + c) << 8;
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n].flags = (v[n].flags & 0xffu) | newTag; }
	ElemRec r = rec;
	r.idx = make_uint4(vid[0], vid[1], vid[2], vid[3]);
	SolveElementGathered<ENERGY, SIMUL, EXACT, false>(NoStore{}, p, r, v, ec);
	// records that leave the thread first (somebody is waiting for them), then the ones it keeps
#pragma unroll
	for (int n = 0; n < 4; n++) {
		if (last[n]) { StoreVertex(sc.Xw, vid[n], v[n]); }
	}
#pragma unroll
	for (int n = 0; n < 4; n++) {
		if (!last[n]) { ChainStore(cache, slot[n], v[n]); }
	}
	return true;
}

template <int ENERGY, bool SIMUL, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_substeps_chain(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                           uint32_t verBase, uint32_t tuning) {
	extern __shared__ __align__(16) unsigned char chainSmem[];
	uint4* cache = reinterpret_cast<uint4*>(chainSmem) + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t gsize = gridDim.x * blockDim.x;
	const uint32_t warpSlot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u;
	const uint32_t mine = warpSlot + lane; // this thread's position in every colour (the host made sure no colour is larger than the grid)
	const uint32_t nC = p.nColors;
	const uint32_t stride = nC + 1u;
	const uint32_t sleepNs = tuning & 0x7fffu, elemSleepNs = tuning >> 16;
	const bool prefetch = (tuning & 0x8000u) == 0;
	const bool spare = warpSlot >= sc.maxColorSize; // never runs an element: vertex phase only
	ElemRec rec;
	uint32_t info = 0;
	bool dead = false;
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		const uint32_t stageBase = verBase + s * stride;
		if (!closing && p.colorStart[0] + mine < p.colorStart[1]) {
			DataflowLoad<ENERGY, EXACT>(sc, p.colorStart[0] + mine, rec);
			info = __ldg(sc.eK + p.colorStart[0] + mine);
		}
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				dead = !DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs, s > 0 ? VaryRow(p, s - 1u) : nullptr);
			}
		}
		if (closing) { break; }
		if (spare) { continue; }
		for (uint32_t c = 0; c < nC; c++) {
			const uint32_t end = p.colorStart[c + 1];
			const uint32_t e0 = p.colorStart[c] + warpSlot;
			if (e0 < end) {
				const bool has = e0 + lane < end && !dead;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) { dead = !ChainElement<ENERGY, SIMUL, EXACT>(sc, p, rec, info, cache, mask, stageBase, c, elemSleepNs); }
			}
			if (c + 1 < nC && p.colorStart[c + 1] + mine < p.colorStart[c + 2]) {
				DataflowLoad<ENERGY, EXACT>(sc, p.colorStart[c + 1] + mine, rec);
				info = __ldg(sc.eK + p.colorStart[c + 1] + mine);
			}
			if (prefetch) {
				const uint32_t c2 = c + 2 < nC ? c + 2 : c + 2 - nC; // wraps into the next substep
				if (p.colorStart[c2] + mine < p.colorStart[c2 + 1]) {
					DataflowPrefetch<ENERGY, EXACT>(sc, p.colorStart[c2] + mine);
					asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eK + p.colorStart[c2] + mine));
				}
			}
		}
	}
}

namespace {
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct ClusterRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, uint32_t verBase, uint32_t tuning,
	                       cudaStream_t st, uint64_t* launches) {
		auto fn = k_substeps_cluster<ENERGY, SIMUL, EXACT>;
		const size_t smem = sizeof(VertexRec) * kClusterSlots * 256;
		static OccupancyCache cache; // per instantiation AND per device: the shared-memory opt-in is a per-device attribute
		int perSm = 0;
		cudaError_t e0 = cache.Get((const void*)fn, 256, smem, true, &perSm);
		if (e0 != cudaSuccess) { return e0; }
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&tuning };
		cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)(perSm * smCount)), dim3(256), args, smem, st);
		++*launches;
		return e;
	}
};
}  // namespace

cudaError_t LaunchSubstepsCluster(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                  uint32_t tuning, cudaStream_t stream, uint64_t* launchCount) {
	return DispatchConfig<ClusterRunner>(p.energy, p.simultaneous != 0, exact, false, sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount);
}

namespace {
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct DataflowRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, uint32_t verBase, uint32_t sleepNs,
	                       cudaStream_t st, uint64_t* launches) {
		auto fn = k_substeps_dataflow<ENERGY, SIMUL, EXACT>;
		static OccupancyCache cache; // per instantiation and per device
		int perSm = 0;
		cudaError_t e0 = cache.Get((const void*)fn, 256, 0, false, &perSm);
		if (e0 != cudaSuccess) { return e0; }
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&sleepNs };
		// cooperative launch: not for grid.sync (there is none) but because it guarantees that all CTAs are co-resident
		const int grid = perSm * smCount;
		cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)grid), dim3((unsigned)DataflowBlockThreads(sc, grid)), args, 0, st);
		++*launches;
		return e;
	}
};
}  // namespace

namespace {
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct ChainRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, uint32_t verBase, uint32_t tuning,
	                       cudaStream_t st, uint64_t* launches) {
		auto fn = k_substeps_chain<ENERGY, SIMUL, EXACT>;
		const size_t smem = 32u * kChainSlots * 256u;
		static OccupancyCache cache;
		int perSm = 0;
		cudaError_t e0 = cache.Get((const void*)fn, 256, smem, false, &perSm);
		if (e0 != cudaSuccess) { return e0; }
		// one position per thread and colour: a colour larger than the co-resident grid runs on the plain kernel
		if ((uint64_t)sc.maxColorSize > (uint64_t)perSm * (uint64_t)smCount * 256u) {
			return DataflowRunner<ENERGY, SIMUL, EXACT, DAMPED>::Run(sc, p, nSubsteps, smCount, verBase, tuning, st, launches);
		}
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&tuning };
		const int grid = perSm * smCount;
		cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)grid), dim3((unsigned)DataflowBlockThreads(sc, grid)), args, smem, st);
		++*launches;
		return e;
	}
};
}  // namespace

cudaError_t LaunchSubstepsChain(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                uint32_t tuning, cudaStream_t stream, uint64_t* launchCount) {
	return DispatchConfig<ChainRunner>(p.energy, p.simultaneous != 0, exact, false, sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount);
}

cudaError_t LaunchSubstepsDataflow(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                   uint32_t sleepNs, cudaStream_t stream, uint64_t* launchCount) {
	return DispatchConfig<DataflowRunner>(p.energy, p.simultaneous != 0, exact, false, sc, p, nSubsteps, smCount, verBase, sleepNs, stream, launchCount);
}

}  // namespace xf
