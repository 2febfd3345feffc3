// XF_SCHEDULE_BRICKS — the persistent cooperative kernel with brick-resident vertices.
//
// The sweep of one mesh is bound by L2 *random-sector requests*: every element gathers and scatters four 32-byte
// vertex records and, because a vertex is rewritten by another SM between colours, nothing can live in L1.  Here the
// elements are grouped into spatially compact bricks (Morton chunks, host side: BuildBricks), one brick per CTA
// of the persistent grid.  A vertex touched by a single brick is *private* to that CTA and stays in its shared
// memory for the whole launch (all substeps), so only the brick-surface vertices travel through L2
// (≈1/3 of the accesses at 1M tets, ≈1/5 at 8M).  The colour order and the per-colour grid barriers are unchanged,
// hence so is the equivalent serial order: results are bit-identical to the other schedules.
// MEASURED (round 1, B200): 99.6 us/substep vs 92.5 us for XF_SCHEDULE_PERSISTENT at 998 250 tets, 1.47e10 vs 1.62e10
// element-substeps/s at 8M tets - the sweep turned out to be bound by the barrier and by dependent-issue latency at
// ~9 warps/SM, not by L2 sectors, and the smem/global select adds instructions.  Kept selectable, not the default.
#include <cooperative_groups.h>

#include "xf_dispatch.cuh"
#include "xf_element.cuh"
#include "xf_phase.cuh"

namespace xf {

namespace cg = cooperative_groups;

// Vertex store that resolves an index either to a shared-memory slot (bit 31 set) or to the global record.
struct BrickStore {
	VertexRec* priv; // shared memory
	GlobalStore base;
	__device__ __forceinline__ VertexRegs LoadX(uint32_t i) const {
		if (i & 0x80000000u) {
			const VertexRec& r = priv[i & 0x7fffffffu];
			VertexRegs v;
			v.x[0] = r.x; v.x[1] = r.y; v.x[2] = r.z;
			v.w = r.w;
			v.flags = r.flags;
			return v;
		}
		return base.LoadX(i);
	}
	__device__ __forceinline__ void StoreX(uint32_t i, const VertexRegs& v) const {
		if (i & 0x80000000u) {
			VertexRec& r = priv[i & 0x7fffffffu];
			r.x = v.x[0]; r.y = v.x[1]; r.z = v.x[2]; // w / flags never change inside a sweep
		} else {
			base.StoreX(i, v);
		}
	}
	// O and V stay in global memory, addressed by global vertex id; the damped variants need the id of a private
	// vertex, which the caller resolves through `privVerts` (see GlobalIdOf)
	const uint32_t* privVerts; // this brick's slot -> global id
	__device__ __forceinline__ uint32_t GlobalIdOf(uint32_t i) const { return (i & 0x80000000u) ? __ldg(privVerts + (i & 0x7fffffffu)) : i; }
	__device__ __forceinline__ void LoadO(uint32_t i, double* o) const { base.LoadO(GlobalIdOf(i), o); }
	__device__ __forceinline__ void LoadV(uint32_t i, double* o) const { base.LoadV(GlobalIdOf(i), o); }
	__device__ __forceinline__ void StoreV(uint32_t i, const double* v) const { base.StoreV(GlobalIdOf(i), v); }
};

template <int KIND, int ENERGY, bool EXACT>
__device__ __forceinline__ void BrickLoad(const DeviceScene& sc, uint32_t e, ElemRec& rec) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	if (KIND == 3) { rec.idx = __ldg(reinterpret_cast<const uint4*>(sc.eAb + e)); return; }
	LoadElementFrom<(KIND != 1) && kPrefactored, EXACT>(sc.eAb, sc, e, rec); // eAb = eA with brick-encoded vertex indices
}
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__device__ __forceinline__ void BrickRun(const DeviceScene& sc, const BrickStore& vs, const SubstepParams& p, uint32_t e, const ElemRec& rec) {
	if (KIND == 0) { SolveElement<ENERGY, SIMUL, EXACT, DAMPED>(vs, p, rec); }
	if (KIND == 1) { SolveVolumeOnly<EXACT>(vs, p, rec); }
	if (KIND == 2) { DampElement<ENERGY, SIMUL, EXACT>(vs, p, rec); }
	if (KIND == 3) { PbdDampElement<EXACT>(vs, p, __ldg(sc.eArea + e), rec.idx); }
}

// vertex phase of this CTA: its private vertices (position in shared memory) + a grid-stride share of the shared ones
template <bool EXACT>
__device__ __forceinline__ void BrickVertexPhase(const DeviceScene& sc, const SubstepParams& p, VertexRec* priv, uint32_t privBegin, uint32_t privCount,
                                                 bool doPost, bool doPredict) {
	for (uint32_t s = threadIdx.x; s < privCount; s += blockDim.x) {
		const uint32_t i = __ldg(sc.privVerts + privBegin + s);
		VertexRec& r = priv[s];
		VertexRegs v;
		v.x[0] = r.x; v.x[1] = r.y; v.x[2] = r.z;
		v.w = r.w;
		v.flags = r.flags;
		VertexPhaseBody<EXACT>(sc, p, i, v, doPost, doPredict);
		r.x = v.x[0]; r.y = v.x[1]; r.z = v.x[2];
		r.w = v.w;
	}
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	for (uint32_t k = gtid; k < sc.nSharedVerts; k += gsize) { VertexPhase<EXACT>(sc, p, __ldg(sc.sharedVerts + k), doPost, doPredict); }
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__global__ void __launch_bounds__(256, 2) k_substeps_bricks(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps) {
	extern __shared__ __align__(32) unsigned char smemRaw[];
	VertexRec* priv = reinterpret_cast<VertexRec*>(smemRaw);
	cg::grid_group grid = cg::this_grid();
	const uint32_t b = blockIdx.x, nB = sc.nBricks, nC = p.nColors;
	const uint32_t privBegin = __ldg(sc.privStart + b), privCount = __ldg(sc.privStart + b + 1) - privBegin;
	const bool anyDamp = p.doDamp || p.doPbdDamp;
	BrickStore vs{ priv, StoreOf(sc), sc.privVerts + privBegin };
	for (uint32_t s = threadIdx.x; s < privCount; s += blockDim.x) { // private vertices move in once per launch
		const VertexRegs v = LoadVertex(sc.Xw, __ldg(sc.privVerts + privBegin + s));
		priv[s] = VertexRec{ v.x[0], v.x[1], v.x[2], v.w, v.flags };
	}
	__syncthreads();
	ElemRec rec;
	for (uint32_t s = 0; s < nSubsteps; s++) {
		uint32_t rb = __ldg(sc.brickStart + b), re = __ldg(sc.brickStart + b + 1); // (colour 0, brick b)
		if (rb + threadIdx.x < re) { BrickLoad<0, ENERGY, EXACT>(sc, rb + threadIdx.x, rec); }
		BrickVertexPhase<EXACT>(sc, p, priv, privBegin, privCount, (s > 0) && !anyDamp, true);
		grid.sync();
		for (uint32_t c = 0; c < nC; c++) {
			uint32_t e = rb + threadIdx.x;
			if (e < re) { BrickRun<0, ENERGY, SIMUL, EXACT, DAMPED>(sc, vs, p, e, rec); }
			for (e += blockDim.x; e < re; e += blockDim.x) {
				ElemRec more;
				BrickLoad<0, ENERGY, EXACT>(sc, e, more);
				BrickRun<0, ENERGY, SIMUL, EXACT, DAMPED>(sc, vs, p, e, more);
			}
			if (c + 1 < nC) { // next colour's first record: its latency hides behind the barrier
				rb = __ldg(sc.brickStart + (size_t)(c + 1) * nB + b);
				re = __ldg(sc.brickStart + (size_t)(c + 1) * nB + b + 1);
				if (rb + threadIdx.x < re) { BrickLoad<0, ENERGY, EXACT>(sc, rb + threadIdx.x, rec); }
			}
			grid.sync();
		}
		for (uint32_t pass = 0; pass < p.volumePasses; pass++) {
			for (uint32_t c = 0; c < nC; c++) {
				const uint32_t vb = __ldg(sc.brickStart + (size_t)c * nB + b), ve = __ldg(sc.brickStart + (size_t)c * nB + b + 1);
				for (uint32_t e = vb + threadIdx.x; e < ve; e += blockDim.x) {
					ElemRec more;
					BrickLoad<1, ENERGY, EXACT>(sc, e, more);
					BrickRun<1, ENERGY, SIMUL, EXACT, false>(sc, vs, p, e, more);
				}
				grid.sync();
			}
		}
		if (anyDamp) {
			BrickVertexPhase<EXACT>(sc, p, priv, privBegin, privCount, true, false);
			grid.sync();
			uint32_t lo, hi;
			DampSlice(p, sc.nT, p.tickId + s, lo, hi);
			for (int kind = 2; kind <= 3; kind++) {
				if (kind == 2 ? !p.doDamp : !p.doPbdDamp) { continue; }
				for (uint32_t c = 0; c < nC; c++) {
					if (!(p.colorStart[c] < hi && p.colorStart[c + 1] > lo)) { continue; }
					const uint32_t db = __ldg(sc.brickStart + (size_t)c * nB + b), de = __ldg(sc.brickStart + (size_t)c * nB + b + 1);
					for (uint32_t e = db + threadIdx.x; e < de; e += blockDim.x) {
						if (!InSlice(sc, e, lo, hi)) { continue; }
						ElemRec more;
						if (kind == 2) {
							BrickLoad<2, ENERGY, EXACT>(sc, e, more);
							BrickRun<2, ENERGY, SIMUL, EXACT, false>(sc, vs, p, e, more);
						} else {
							BrickLoad<3, ENERGY, EXACT>(sc, e, more);
							BrickRun<3, ENERGY, SIMUL, EXACT, false>(sc, vs, p, e, more);
						}
					}
					grid.sync();
				}
			}
		}
	}
	if (!anyDamp && nSubsteps > 0) { BrickVertexPhase<EXACT>(sc, p, priv, privBegin, privCount, true, false); }
	__syncthreads();
	for (uint32_t s = threadIdx.x; s < privCount; s += blockDim.x) { // private vertices move out once per launch
		const VertexRec& r = priv[s];
		VertexRegs v;
		v.x[0] = r.x; v.x[1] = r.y; v.x[2] = r.z;
		v.w = r.w;
		v.flags = r.flags;
		StoreVertex(sc.Xw, __ldg(sc.privVerts + privBegin + s), v);
	}
}

namespace {
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct BrickRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, cudaStream_t st, uint64_t* launches) {
		auto fn = k_substeps_bricks<ENERGY, SIMUL, EXACT, DAMPED>;
		const size_t smem = sizeof(VertexRec) * (size_t)std::max(sc.maxPrivPerBrick, 1u);
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) { return e; }
		int perSm = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 256, smem);
		if (e != cudaSuccess) { return e; }
		if ((uint32_t)(perSm * smCount) < sc.nBricks) { return cudaErrorCooperativeLaunchTooLarge; } // bricks must be co-resident
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps };
		e = cudaLaunchCooperativeKernel((const void*)fn, dim3(sc.nBricks), dim3(256), args, smem, st);
		++*launches;
		return e;
	}
};
}  // namespace

cudaError_t LaunchSubstepsBricks(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, cudaStream_t stream,
                                 uint64_t* launchCount) {
	const bool damped = p.damping > 0.0f && p.rayleigh < XF_RAYLEIGH_POST;
	return DispatchConfig<BrickRunner>(p.energy, p.simultaneous != 0, exact, damped, sc, p, nSubsteps, smCount, stream, launchCount);
}

}  // namespace xf
