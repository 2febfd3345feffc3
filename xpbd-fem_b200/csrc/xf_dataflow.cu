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
// Covers the main sweep (all energies / solve modes, undamped in-constraint) + the fused vertex phase.  Volume passes,
// damping sweeps and in-constraint Rayleigh damping (which reads O of other threads' vertices) run on
// XF_SCHEDULE_PERSISTENT; xf_substep falls back per call.
#include "xf_dispatch.cuh"
#include "xf_element.cuh"
#include "xf_phase.cuh"

namespace xf {

namespace {

constexpr uint32_t kVerMask = 0xffffff00u;
// A record that never reaches the expected stage means a broken schedule (or a caller that rewrote the state while a
// launch was in flight): fail loudly (launch error) instead of hanging the device.  2^24 polls is seconds, a healthy
// wait is a few polls.
constexpr uint32_t kSpinLimit = 1u << 24;

// Waiting is warp-uniform on purpose.  A lane that left a spin loop early would be parked at the loop's reconvergence
// point, and independent thread scheduling releases parked lanes when the spinning ones yield (that is how it guarantees
// progress): the warp would then run the ~750-instruction element body once per group of lanes.  Measured: 2x slower at
// every mesh size.  With a vote over the lanes that have work, the warp leaves the loop as one.
template <bool EXACT>
__device__ __forceinline__ void DataflowVertex(const DeviceScene& sc, const SubstepParams& p, uint32_t i, unsigned mask, bool doPost, bool doPredict,
                                               bool wait, uint32_t expectTag, uint32_t newTag, uint32_t sleepNs) {
	VertexRegs v = LoadVertex(sc.Xw, i);
	if (wait) {
		for (uint32_t spins = 0;; spins++) {
			const bool ok = (v.flags & kVerMask) == expectTag;
			if (__all_sync(mask, ok)) { break; }
			if (spins > kSpinLimit) { __trap(); }
			if (sleepNs) { __nanosleep(sleepNs); }
			if (!ok) { v = LoadVertex(sc.Xw, i); }
		}
	}
	VertexPhaseBody<EXACT>(sc, p, i, v, doPost, doPredict);
	v.flags = (v.flags & 0xffu) | newTag;
	StoreVertex(sc.Xw, i, v);
}

// One element: spin-gather the four versioned records, solve, scatter with this stage's tag.  `mask` = the lanes of
// this warp that run an element in this step (all of them call this function together).
template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ void DataflowElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, unsigned mask, uint32_t stageBase,
                                                uint32_t c, uint32_t sleepNs) {
	const GlobalStore vs = StoreOf(sc);
	const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	uint32_t vid[4], expectTag[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		vid[n] = raw[n] & 0x00ffffffu;
		expectTag[n] = (stageBase + (raw[n] >> 24)) << 8;
	}
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n] = vs.LoadX(vid[n]); }
	const ElemCompliance ec = ComplianceOf<EXACT>(p, rec.volume); // two chained divisions, in the shadow of the gather
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (spins > kSpinLimit) { __trap(); }
		if (sleepNs) { __nanosleep(sleepNs); }
		// only the stale records are read again (a poll costs L1TEX wavefronts, the resource the sweep runs on)
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!ok[n]) { v[n] = vs.LoadX(vid[n]); }
		}
	}
	const uint32_t newTag = (stageBase + 1u + c) << 8;
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n].flags = (v[n].flags & 0xffu) | newTag; }
	ElemRec r = rec;
	r.idx = make_uint4(vid[0], vid[1], vid[2], vid[3]);
	SolveElementGathered<ENERGY, SIMUL, EXACT, false>(vs, p, r, v, ec);
}

// The record of the stage after next: pulled into L1 now (the planes are read with ld.global.nc, which allocates in L1),
// so that the load issued right before it is needed costs an L1 hit instead of an L2 / HBM round trip in the
// stage-to-stage dependence chain.
template <int ENERGY, bool EXACT>
__device__ __forceinline__ void DataflowPrefetch(const DeviceScene& sc, uint32_t e) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eAd + e));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eB + e));
	if (kPrefactored && EXACT) { asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eC + e)); }
}

template <int ENERGY, bool EXACT>
__device__ __forceinline__ void DataflowLoad(const DeviceScene& sc, uint32_t e, ElemRec& rec) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	LoadElementFrom<kPrefactored, EXACT>(sc.eAd, sc, e, rec);
}

}  // namespace

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
	ElemRec rec;
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps; // closing post phase (locks, manipulator, velocities) of the last substep
		const uint32_t stageBase = verBase + s * stride;
		if (!closing && p.colorStart[0] + warpSlot + lane < p.colorStart[1]) { DataflowLoad<ENERGY, EXACT>(sc, p.colorStart[0] + warpSlot + lane, rec); }
		// vertex phase: post of the previous substep + predict, once the vertex's last element of that substep has written
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) { // warp-uniform trip count
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs);
			}
		}
		if (closing) { break; }
		for (uint32_t c = 0; c < nC; c++) {
			const uint32_t end = p.colorStart[c + 1];
			uint32_t e0 = p.colorStart[c] + warpSlot;
			if (e0 < end) {
				const bool has = e0 + lane < end;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) { DataflowElement<ENERGY, SIMUL, EXACT>(sc, p, rec, mask, stageBase, c, elemSleepNs); }
			}
			for (e0 += gsize; e0 < end; e0 += gsize) {
				const bool has = e0 + lane < end;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) {
					ElemRec more;
					DataflowLoad<ENERGY, EXACT>(sc, e0 + lane, more);
					DataflowElement<ENERGY, SIMUL, EXACT>(sc, p, more, mask, stageBase, c, elemSleepNs);
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
struct NoStore { // the clustered kernel routes the updated records itself
	__device__ __forceinline__ void StoreX(uint32_t, const VertexRegs&) const {}
	__device__ __forceinline__ void LoadO(uint32_t, double*) const {}
	__device__ __forceinline__ void LoadV(uint32_t, double*) const {}
	__device__ __forceinline__ void StoreV(uint32_t, const double*) const {}
};

constexpr int kClusterSlots = 8;

template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ void ClusterElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, uint32_t info, VertexRec* cache,
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
	const ElemCompliance ec = ComplianceOf<EXACT>(p, rec.volume);
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = !first[n] || (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (spins > kSpinLimit) { __trap(); }
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
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		const uint32_t stageBase = verBase + s * stride;
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs);
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
					const bool has = k0 + lane < nt;
					const unsigned mask = __ballot_sync(0xffffffffu, has);
					if (has) {
						const uint32_t e = p.colorStart[c] + k0 + lane;
						ElemRec rec;
						DataflowLoad<ENERGY, EXACT>(sc, e, rec);
						ClusterElement<ENERGY, SIMUL, EXACT>(sc, p, rec, __ldg(sc.eK + e), cache, mask, stageBase, c);
					}
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
		static int perSm = 0;
		if (perSm == 0) {
			cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess) { return e; }
			e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 256, smem);
			if (e != cudaSuccess) { return e; }
			if (perSm < 1) { return cudaErrorLaunchOutOfResources; }
		}
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
		static int perSm = 0; // per instantiation; the occupancy of a kernel does not change
		if (perSm == 0) {
			cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 256, 0);
			if (e != cudaSuccess) { return e; }
			if (perSm < 1) { return cudaErrorLaunchOutOfResources; }
		}
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&sleepNs };
		// cooperative launch: not for grid.sync (there is none) but because it guarantees that all CTAs are co-resident
		cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)(perSm * smCount)), dim3(256), args, 0, st);
		++*launches;
		return e;
	}
};
}  // namespace

cudaError_t LaunchSubstepsDataflow(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                   uint32_t sleepNs, cudaStream_t stream, uint64_t* launchCount) {
	return DispatchConfig<DataflowRunner>(p.energy, p.simultaneous != 0, exact, false, sc, p, nSubsteps, smCount, verBase, sleepNs, stream, launchCount);
}

}  // namespace xf
