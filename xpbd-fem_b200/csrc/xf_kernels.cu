// sm_100a kernels of the substep: vertex phases (predict / post), the colour sweeps, the post-solve
// damping sweeps, and the persistent cooperative kernel that runs whole substeps with grid barriers
// between colours.  The path is gather/scatter- and latency-bound (≈5 flop/B), so no tensor cores:
// what matters is one 32-byte sector per vertex gather, coalesced SoA element streams, a grid sized to
// the 148 SMs, and as few grid-wide barriers as the colouring allows.
#include <cooperative_groups.h>

#include "xf_element.cuh"
#include "xf_dispatch.cuh"
#include "xf_phase.cuh"

namespace xf {

namespace cg = cooperative_groups;

template <bool EXACT>
__global__ void __launch_bounds__(256) k_vertex_phase(const DeviceScene sc, const __grid_constant__ SubstepParams p, int doPost, int doPredict) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < sc.nV) { VertexPhase<EXACT>(sc, p, i, doPost != 0, doPredict != 0); }
}

// ------------------------------------------------------------------------------------------------
// Sweeps over one colour (elements [begin, end) of the colour-sorted planes).
//   KIND 0: main constraint solve   KIND 1: volume-only pass   KIND 2: Rayleigh damp (V)   KIND 3: PBD damp (V)
// ------------------------------------------------------------------------------------------------
template <int KIND, int ENERGY, bool EXACT>
__device__ __forceinline__ void SweepLoad(const DeviceScene& sc, uint32_t e, ElemRec& rec) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	if (KIND == 3) { rec.idx = LoadElementIdx(sc, e); return; }
	LoadElement<(KIND != 1) && kPrefactored, EXACT>(sc, e, rec);
}
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__device__ __forceinline__ void SweepRun(const DeviceScene& sc, const SubstepParams& p, uint32_t e, const ElemRec& rec) {
	const GlobalStore vs = StoreOf(sc);
	if (KIND == 0) { SolveElement<ENERGY, SIMUL, EXACT, DAMPED>(vs, p, rec); }
	if (KIND == 1) { SolveVolumeOnly<EXACT>(vs, p, rec); }
	if (KIND == 2) { DampElement<ENERGY, SIMUL, EXACT>(vs, p, rec); }
	if (KIND == 3) { PbdDampElement<EXACT>(vs, p, __ldg(sc.eArea + e), rec.idx); }
}
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__device__ __forceinline__ void SweepOne(const DeviceScene& sc, const SubstepParams& p, uint32_t e) {
	ElemRec rec;
	SweepLoad<KIND, ENERGY, EXACT>(sc, e, rec);
	SweepRun<KIND, ENERGY, SIMUL, EXACT, DAMPED>(sc, p, e, rec);
}

template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__global__ void __launch_bounds__(256) k_sweep_color(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t begin, uint32_t end,
                                                     uint32_t lo, uint32_t hi) {
	uint32_t e = begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (e < end && (KIND < 2 || InSlice(sc, e, lo, hi))) { SweepOne<KIND, ENERGY, SIMUL, EXACT, DAMPED>(sc, p, e); }
}

// ------------------------------------------------------------------------------------------------
// Grid barrier for the persistent kernel: monotonically increasing arrival counter in L2.
// bar.sync orders the CTA's writes before thread 0's gpu-scope fence + relaxed atomic (release
// pattern); the spin uses an acquire load; vertex data is read with ld.global.cg afterwards.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int LoadAcquire(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void GridBarrier(unsigned int* counter, unsigned int& target) {
	__syncthreads();
	target += gridDim.x;
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(counter, 1u);
		while (LoadAcquire(counter) < target) { }
		__threadfence();
	}
	__syncthreads();
}

// One colour range [b, end) swept by the whole grid.  Work is dealt out in warp-sized chunks round-robin over
// the CTAs (chunk k -> CTA k % gridDim, warp slot k / gridDim) so that a colour smaller than the grid still
// loads every SM equally.  `rec` holds this thread's first element of the range, loaded before the barrier
// that precedes the sweep; the record of the first element of [nb, nend) is loaded before returning, so its
// HBM/L2 latency overlaps the next barrier.
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED, int NEXT_KIND>
__device__ __forceinline__ void SweepRange(const DeviceScene& sc, const SubstepParams& p, uint32_t b, uint32_t end, uint32_t slot, uint32_t gsize,
                                           ElemRec& rec, uint32_t nb, uint32_t nend) {
	uint32_t e = b + slot;
	if (e < end) { SweepRun<KIND, ENERGY, SIMUL, EXACT, DAMPED>(sc, p, e, rec); }
	for (e += gsize; e < end; e += gsize) { SweepOne<KIND, ENERGY, SIMUL, EXACT, DAMPED>(sc, p, e); }
	if (NEXT_KIND >= 0 && nb + slot < nend) { SweepLoad<(NEXT_KIND < 0 ? 0 : NEXT_KIND), ENERGY, EXACT>(sc, nb + slot, rec); }
}

#ifndef XF_PERSIST_THREADS
#define XF_PERSIST_THREADS 256
#endif
#ifndef XF_PERSIST_MIN_BLOCKS
#define XF_PERSIST_MIN_BLOCKS 2
#endif
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__global__ void __launch_bounds__(XF_PERSIST_THREADS, XF_PERSIST_MIN_BLOCKS) k_substeps_persistent(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps) {
	cg::grid_group grid = cg::this_grid();
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t gsize = gridDim.x * blockDim.x;
	// element slot: warp-sized chunks dealt round-robin over CTAs
	const uint32_t slot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u);
	const bool anyDamp = p.doDamp || p.doPbdDamp;
	const uint32_t nC = p.nColors;
	ElemRec rec;
	for (uint32_t s = 0; s < nSubsteps; s++) {
		// prefetch the first record of colour 0, then the vertex phase: post of the previous substep (unless a
		// damping sweep already closed it) + predict
		if (p.colorStart[0] + slot < p.colorStart[1]) { SweepLoad<0, ENERGY, EXACT>(sc, p.colorStart[0] + slot, rec); }
		const bool fusePost = (s > 0) && !anyDamp;
		for (uint32_t i = gtid; i < sc.nV; i += gsize) { VertexPhase<EXACT>(sc, p, i, fusePost, true, fusePost ? VaryRow(p, s - 1u) : nullptr); }
		grid.sync();
		for (uint32_t c = 0; c < nC; c++) {
			const bool last = c + 1 == nC;
			if (!last) {
				SweepRange<0, ENERGY, SIMUL, EXACT, DAMPED, 0>(sc, p, p.colorStart[c], p.colorStart[c + 1], slot, gsize, rec, p.colorStart[c + 1], p.colorStart[c + 2]);
			} else if (p.volumePasses > 0) {
				SweepRange<0, ENERGY, SIMUL, EXACT, DAMPED, 1>(sc, p, p.colorStart[c], p.colorStart[c + 1], slot, gsize, rec, p.colorStart[0], p.colorStart[1]);
			} else {
				SweepRange<0, ENERGY, SIMUL, EXACT, DAMPED, -1>(sc, p, p.colorStart[c], p.colorStart[c + 1], slot, gsize, rec, 0, 0);
			}
			grid.sync();
		}
		for (uint32_t pass = 0; pass < p.volumePasses; pass++) {
			for (uint32_t c = 0; c < nC; c++) {
				const bool last = (c + 1 == nC) && (pass + 1 == p.volumePasses);
				const uint32_t nc = (c + 1) % nC;
				if (!last) {
					SweepRange<1, ENERGY, SIMUL, EXACT, false, 1>(sc, p, p.colorStart[c], p.colorStart[c + 1], slot, gsize, rec, p.colorStart[nc], p.colorStart[nc + 1]);
				} else {
					SweepRange<1, ENERGY, SIMUL, EXACT, false, -1>(sc, p, p.colorStart[c], p.colorStart[c + 1], slot, gsize, rec, 0, 0);
				}
				grid.sync();
			}
		}
		if (anyDamp) {
			for (uint32_t i = gtid; i < sc.nV; i += gsize) { VertexPhase<EXACT>(sc, p, i, true, false, VaryRow(p, s)); }
			grid.sync();
			uint32_t lo, hi;
			DampSlice(p, sc.nT, p.tickId + s, lo, hi);
			if (p.doDamp) {
				for (uint32_t c = 0; c < nC; c++) {
					if (p.colorStart[c] < hi && p.colorStart[c + 1] > lo) {
						for (uint32_t e = p.colorStart[c] + slot; e < p.colorStart[c + 1]; e += gsize) {
							if (InSlice(sc, e, lo, hi)) { SweepOne<2, ENERGY, SIMUL, EXACT, false>(sc, p, e); }
						}
						grid.sync();
					}
				}
			}
			if (p.doPbdDamp) {
				for (uint32_t c = 0; c < nC; c++) {
					if (p.colorStart[c] < hi && p.colorStart[c + 1] > lo) {
						for (uint32_t e = p.colorStart[c] + slot; e < p.colorStart[c + 1]; e += gsize) {
							if (InSlice(sc, e, lo, hi)) { SweepOne<3, ENERGY, SIMUL, EXACT, false>(sc, p, e); }
						}
						grid.sync();
					}
				}
			}
		}
	}
	if (!anyDamp && nSubsteps > 0) {
		for (uint32_t i = gtid; i < sc.nV; i += gsize) { VertexPhase<EXACT>(sc, p, i, true, false, VaryRow(p, nSubsteps - 1u)); }
	}
}

// ------------------------------------------------------------------------------------------------
// Auxiliary kernels
// ------------------------------------------------------------------------------------------------
// Per-element (1/6) det[X0-X3, X1-X3, X2-X3] in *stream* order (Fem.cpp:1067); summed on the host in the
// reference's order so xf_volume is bit-identical to GeoLinear3d::CalculateVolume.
__global__ void __launch_bounds__(256) k_element_volumes(const DeviceScene sc) {
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= sc.nT) { return; }
	uint32_t e = sc.streamToSorted[s];
	uint4 idx = LoadElementIdx(sc, e);
	VertexRegs v[4] = { LoadVertex(sc.Xw, idx.x), LoadVertex(sc.Xw, idx.y), LoadVertex(sc.Xw, idx.z), LoadVertex(sc.Xw, idx.w) };
	float P[3][3];
	Edges<true>(v, P);
	typedef Op<true> O;
	// determinant(mat3) with columns P0,P1,P2, vectormath.cpp:34-39
	float a = P[0][0], b = P[1][0], c = P[2][0];
	float d = P[0][1], e1 = P[1][1], f = P[2][1];
	float g = P[0][2], h = P[1][2], i = P[2][2];
	float det = O::add(O::sub(O::mul(a, O::sub(O::mul(e1, i), O::mul(f, h))), O::mul(b, O::sub(O::mul(d, i), O::mul(f, g)))),
	                   O::mul(c, O::sub(O::mul(d, h), O::mul(e1, g))));
	sc.eScratch[s] = O::mul(1.0f / 6.0f, det);
}

// alpha plane of the barrier-free kernels (DeviceScene::eAlpha): ComplianceOf's operations, once per element and settings change
template <bool EXACT>
__global__ void __launch_bounds__(256) k_element_alpha(const DeviceScene sc, const __grid_constant__ SubstepParams p) {
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= sc.nT) { return; }
	const ElemCompliance c = ComplianceOf<EXACT>(p, sc.eB[e].volume);
	sc.eAlpha[e] = make_float2(c.alpha0, c.alpha1);
}

// Geo3d::Transform, Geo.cpp:358-364: positions narrowed to fp32, through the mat4, widened; O = X.
__global__ void __launch_bounds__(256) k_transform(const DeviceScene sc, float c00, float c01, float c02, float c10, float c11, float c12, float c30,
                                                   float c31) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= sc.nV) { return; }
	typedef Op<true> O;
	VertexRegs v = LoadVertex(sc.Xw, i);
	float x = __double2float_rn(v.x[0]), y = __double2float_rn(v.x[1]), z = __double2float_rn(v.x[2]);
	// rows of t4 = columns (c0,0) (c1,0) (0,0,1,0) (c3.xy,0,1) dotted with (x,y,z,1), left-assoc
	float rx = O::add(O::add(O::add(O::mul(c00, x), O::mul(c10, y)), O::mul(0.0f, z)), O::mul(c30, 1.0f));
	float ry = O::add(O::add(O::add(O::mul(c01, x), O::mul(c11, y)), O::mul(0.0f, z)), O::mul(c31, 1.0f));
	float rz = O::add(O::add(O::add(O::mul(c02, x), O::mul(c12, y)), O::mul(1.0f, z)), O::mul(0.0f, 1.0f));
	v.x[0] = (double)rx; v.x[1] = (double)ry; v.x[2] = (double)rz;
	StoreVertex(sc.Xw, i, v);
	StoreD3(sc.O, i, v.x);
}

__global__ void __launch_bounds__(256) k_pack_state(const DeviceScene sc, double* __restrict__ X, double* __restrict__ V, float* __restrict__ W) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= sc.nV) { return; }
	VertexRegs v = LoadVertex(sc.Xw, i);
	const size_t x = sc.extOfInt ? __ldg(sc.extOfInt + i) : i; // the caller's numbering
	if (X) { X[3 * x] = v.x[0]; X[3 * x + 1] = v.x[1]; X[3 * x + 2] = v.x[2]; }
	if (V) { double vel[3]; LoadD3(sc.V, i, vel); V[3 * x] = vel[0]; V[3 * x + 1] = vel[1]; V[3 * x + 2] = vel[2]; }
	if (W) { W[x] = v.w; }
}
__global__ void __launch_bounds__(256) k_unpack_state(const DeviceScene sc, const double* __restrict__ X, const double* __restrict__ V,
                                                      const float* __restrict__ W) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= sc.nV) { return; }
	VertexRegs v = LoadVertex(sc.Xw, i);
	const size_t x = sc.extOfInt ? __ldg(sc.extOfInt + i) : i; // the caller's numbering
	if (X) { v.x[0] = X[3 * x]; v.x[1] = X[3 * x + 1]; v.x[2] = X[3 * x + 2]; }
	if (W) { v.w = W[x]; }
	StoreVertex(sc.Xw, i, v);
	if (V) { double vel[3] = { V[3 * x], V[3 * x + 1], V[3 * x + 2] }; StoreD3(sc.V, i, vel); }
}

// fp64 statistics: volume, kinetic, gravitational, deviatoric, volumetric, non-finite count.
__device__ __forceinline__ double WarpSum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
	return v;
}
__global__ void __launch_bounds__(256) k_stats(const DeviceScene sc, const __grid_constant__ SubstepParams p, double gx, double gy) {
	double acc[6] = { 0, 0, 0, 0, 0, 0 };
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	for (uint32_t i = gtid; i < sc.nV; i += gsize) {
		VertexRegs v = LoadVertex(sc.Xw, i);
		double vel[3];
		LoadD3(sc.V, i, vel);
		if (!(isfinite(v.x[0]) && isfinite(v.x[1]) && isfinite(v.x[2]))) { acc[5] += 1.0; }
		if (v.w > 0.0f) {
			double m = 1.0 / (double)v.w;
			acc[1] += 0.5 * m * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
			acc[2] -= m * (gx * v.x[0] + gy * v.x[1]);
		}
	}
	for (uint32_t e = gtid; e < sc.nT; e += gsize) {
		ElemRec r;
		LoadElement<false, true>(sc, e, r);
		VertexRegs v[4] = { LoadVertex(sc.Xw, r.idx.x), LoadVertex(sc.Xw, r.idx.y), LoadVertex(sc.Xw, r.idx.z), LoadVertex(sc.Xw, r.idx.w) };
		float P[3][3], F[3][3], adj[3][3];
		Edges<true>(v, P);
		DeformationGradient<true>(r, P, F);
		float J = AdjugateAndDet<true>(F, adj);
		double I1 = (double)Op<true>::add(Op<true>::add(Op<true>::dot(F[0], F[0]), Op<true>::dot(F[1], F[1])), Op<true>::dot(F[2], F[2]));
		double IM = I1 - 3.0;
		double U0 = IM;
		if (p.energy == XF_ENERGY_YEOH_SKIN || p.energy == XF_ENERGY_YEOH_SKIN_FAST) { U0 = 0.1095 * IM + 14.95 * IM * IM + 4.595 * IM * IM * IM; }
		float comp0 = __fdiv_rn(p.invMu, r.volume), comp1 = __fdiv_rn(p.invLambda, r.volume);
		acc[0] += (double)r.volume * (double)J;
		acc[3] += U0 / (double)comp0;
		if (comp1 > 0.0f) { double d = (double)J - (double)p.a; acc[4] += d * d / (double)comp1; }
	}
	__shared__ double sm[6][8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 6; k++) {
		double s = WarpSum(acc[k]);
		if (lane == 0) { sm[k][warp] = s; }
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		double s = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); w++) { s += sm[threadIdx.x][w]; }
		atomicAdd(sc.statScratch + threadIdx.x, s);
	}
}

// ------------------------------------------------------------------------------------------------
// Host-side dispatch
// ------------------------------------------------------------------------------------------------
namespace {

inline dim3 GridFor(uint32_t n, int threads) { return dim3((n + threads - 1) / threads); }

template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
cudaError_t LaunchSweepT(const DeviceScene& sc, const SubstepParams& p, uint32_t b, uint32_t e, cudaStream_t st, uint32_t lo = 0, uint32_t hi = 0xffffffffu) {
	if (b >= e) { return cudaSuccess; }
	k_sweep_color<KIND, ENERGY, SIMUL, EXACT, DAMPED><<<GridFor(e - b, 256), 256, 0, st>>>(sc, p, b, e, lo, hi);
	return cudaGetLastError();
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct PerColorRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p0, uint32_t nSubsteps, cudaStream_t st, uint64_t* launches) {
		SubstepParams p = p0;
		const bool anyDamp = p.doDamp || p.doPbdDamp;
		const dim3 vgrid = GridFor(sc.nV, 256);
		cudaError_t err = cudaSuccess;
		auto ok = [&](cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) { err = e; } return e == cudaSuccess; };
		for (uint32_t s = 0; s < nSubsteps && err == cudaSuccess; s++) {
			const bool fusePost = (s > 0) && !anyDamp;
			k_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(sc, p, fusePost ? 1 : 0, 1);
			ok(cudaGetLastError()); ++*launches;
			for (uint32_t c = 0; c < p.nColors; c++) {
				ok(LaunchSweepT<0, ENERGY, SIMUL, EXACT, DAMPED>(sc, p, p.colorStart[c], p.colorStart[c + 1], st)); ++*launches;
			}
			for (uint32_t pass = 0; pass < p.volumePasses; pass++) {
				for (uint32_t c = 0; c < p.nColors; c++) {
					ok(LaunchSweepT<1, ENERGY, SIMUL, EXACT, false>(sc, p, p.colorStart[c], p.colorStart[c + 1], st)); ++*launches;
				}
			}
			if (anyDamp) {
				k_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(sc, p, 1, 0);
				ok(cudaGetLastError()); ++*launches;
				uint32_t lo = 0, hi = sc.nT;
				if (p.rayleigh == XF_RAYLEIGH_POST_AMORTIZED) {
					uint32_t k = (p0.tickId + s) % XF_AMORTIZATION_PERIOD;
					lo = (uint32_t)(((uint64_t)sc.nT * k) / XF_AMORTIZATION_PERIOD);
					hi = (uint32_t)(((uint64_t)sc.nT * (k + 1)) / XF_AMORTIZATION_PERIOD);
				}
				// the serial order is colour-major too: colour c intersects the slice iff the ranges overlap
				if (p.doDamp) {
					for (uint32_t c = 0; c < p.nColors; c++) {
						if (p.colorStart[c] < hi && p.colorStart[c + 1] > lo) {
							ok(LaunchSweepT<2, ENERGY, SIMUL, EXACT, false>(sc, p, p.colorStart[c], p.colorStart[c + 1], st, lo, hi)); ++*launches;
						}
					}
				}
				if (p.doPbdDamp) {
					for (uint32_t c = 0; c < p.nColors; c++) {
						if (p.colorStart[c] < hi && p.colorStart[c + 1] > lo) {
							ok(LaunchSweepT<3, ENERGY, SIMUL, EXACT, false>(sc, p, p.colorStart[c], p.colorStart[c + 1], st, lo, hi)); ++*launches;
						}
					}
				}
			}
		}
		if (!anyDamp && nSubsteps > 0 && err == cudaSuccess) {
			k_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(sc, p, 1, 0);
			ok(cudaGetLastError()); ++*launches;
		}
		return err;
	}
};

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct PersistentRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, const LaunchShape& shape, cudaStream_t st,
	                       uint64_t* launches) {
		cudaError_t e = cudaMemsetAsync(sc.barrier, 0, sizeof(unsigned int), st);
		if (e != cudaSuccess) { return e; }
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps };
		e = cudaLaunchCooperativeKernel((const void*)k_substeps_persistent<ENERGY, SIMUL, EXACT, DAMPED>, dim3(shape.gridBlocks),
		                                dim3(shape.blockThreads), args, 0, st);
		++*launches;
		return e;
	}
};

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct OccupancyQuery {
	static cudaError_t Run(int* blocksPerSm, int threads) {
		return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocksPerSm, k_substeps_persistent<ENERGY, SIMUL, EXACT, DAMPED>, threads, 0);
	}
};

}  // namespace

cudaError_t QueryLaunchShape(int device, uint32_t energy, bool exact, LaunchShape* shape) {
	cudaDeviceProp prop;
	cudaError_t e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) { return e; }
	shape->smCount = prop.multiProcessorCount;
	shape->blockThreads = XF_PERSIST_THREADS;
	// the most register-hungry variants bound the co-resident grid for all of them
	int worst = 1 << 30;
	for (int simul = 0; simul < 2; simul++) {
		for (int damped = 0; damped < 2; damped++) {
			int b = 0;
			e = DispatchConfig<OccupancyQuery>(energy, simul != 0, exact, damped != 0, &b, shape->blockThreads);
			if (e != cudaSuccess) { return e; }
			if (b < worst) { worst = b; }
		}
	}
	if (worst < 1) { return cudaErrorLaunchOutOfResources; }
	shape->gridBlocks = worst * shape->smCount;
	return cudaSuccess;
}

cudaError_t LaunchSubstepsPerColor(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, cudaStream_t stream,
                                   uint64_t* launchCount) {
	const bool damped = p.damping > 0.0f && p.rayleigh < XF_RAYLEIGH_POST;
	return DispatchConfig<PerColorRunner>(p.energy, p.simultaneous != 0, exact, damped, sc, p, nSubsteps, stream, launchCount);
}

cudaError_t LaunchSubstepsPersistent(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, const LaunchShape& shape,
                                     cudaStream_t stream, uint64_t* launchCount) {
	const bool damped = p.damping > 0.0f && p.rayleigh < XF_RAYLEIGH_POST;
	return DispatchConfig<PersistentRunner>(p.energy, p.simultaneous != 0, exact, damped, sc, p, nSubsteps, shape, stream, launchCount);
}

cudaError_t LaunchTransform(const DeviceScene& sc, const float* m9, cudaStream_t stream, uint64_t* launchCount) {
	k_transform<<<GridFor(sc.nV, 256), 256, 0, stream>>>(sc, m9[0], m9[1], m9[2], m9[3], m9[4], m9[5], m9[6], m9[7]);
	++*launchCount;
	return cudaGetLastError();
}

cudaError_t LaunchStats(const DeviceScene& sc, const SubstepParams& p, double gx, double gy, int smCount, cudaStream_t stream, uint64_t* launchCount) {
	cudaError_t e = cudaMemsetAsync(sc.statScratch, 0, 6 * sizeof(double), stream);
	if (e != cudaSuccess) { return e; }
	k_stats<<<dim3(smCount * 4), 256, 0, stream>>>(sc, p, gx, gy);
	++*launchCount;
	return cudaGetLastError();
}

cudaError_t LaunchPackState(const DeviceScene& sc, double* dX, double* dV, float* dW, cudaStream_t stream, uint64_t* launchCount) {
	k_pack_state<<<GridFor(sc.nV, 256), 256, 0, stream>>>(sc, dX, dV, dW);
	++*launchCount;
	return cudaGetLastError();
}
cudaError_t LaunchUnpackState(const DeviceScene& sc, const double* dX, const double* dV, const float* dW, cudaStream_t stream,
                              uint64_t* launchCount) {
	k_unpack_state<<<GridFor(sc.nV, 256), 256, 0, stream>>>(sc, dX, dV, dW);
	++*launchCount;
	return cudaGetLastError();
}

cudaError_t LaunchElementAlpha(const DeviceScene& sc, const SubstepParams& p, bool exact, cudaStream_t stream, uint64_t* launchCount) {
	if (exact) {
		k_element_alpha<true><<<GridFor(sc.nT, 256), 256, 0, stream>>>(sc, p);
	} else {
		k_element_alpha<false><<<GridFor(sc.nT, 256), 256, 0, stream>>>(sc, p);
	}
	++*launchCount;
	return cudaGetLastError();
}

cudaError_t LaunchElementVolumes(const DeviceScene& sc, cudaStream_t stream, uint64_t* launchCount) {
	k_element_volumes<<<GridFor(sc.nT, 256), 256, 0, stream>>>(sc);
	++*launchCount;
	return cudaGetLastError();
}

}  // namespace xf

// ------------------------------------------------------------------------------------------------
// Debug/measurement hook (not part of the public header): cost of the bare grid barrier, several variants.
// ------------------------------------------------------------------------------------------------
namespace xf {
template <int VARIANT>
__device__ __forceinline__ void GridBarrierV(unsigned int* counter, unsigned int& target) {
	__syncthreads();
	target += gridDim.x;
	if (threadIdx.x == 0) {
		if (VARIANT == 0) {
			__threadfence();
			atomicAdd(counter, 1u);
			while (LoadAcquire(counter) < target) { }
			__threadfence();
		} else if (VARIANT == 1) {
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			while (LoadAcquire(counter) < target) { }
		} else if (VARIANT == 2) {
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			unsigned int v;
			do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
		} else if (VARIANT == 3) {
			asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
			unsigned int v;
			do {
				asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
				if (v < target) { __nanosleep(40); }
			} while (v < target);
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
		}
	}
	__syncthreads();
}
// Two-level barrier: CTAs arrive on one of `groups` counters (separate 128-byte lines => separate L2 slices);
// the last arriver of a group arrives on the root; the last root arriver bumps the release word everybody polls.
__device__ __forceinline__ void GridBarrierTree(unsigned int* base, unsigned int& epoch, uint32_t groupSize) {
	__syncthreads();
	epoch += 1;
	if (threadIdx.x == 0) {
		const uint32_t nGroups = (gridDim.x + groupSize - 1) / groupSize;
		const uint32_t g = blockIdx.x / groupSize;
		const uint32_t members = min(groupSize, gridDim.x - g * groupSize);
		unsigned int* groupCtr = base + 32 * (2 + g);
		unsigned int* rootCtr = base + 32;
		unsigned int* release = base;
		__threadfence();
		if (atomicAdd(groupCtr, 1u) == members * epoch - 1) {
			if (atomicAdd(rootCtr, 1u) == nGroups * epoch - 1) {
				asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(release), "r"(epoch) : "memory");
			}
		}
		unsigned int v;
		do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(release) : "memory"); } while (v < epoch);
		asm volatile("fence.acq_rel.gpu;" ::: "memory");
	}
	__syncthreads();
}
template <int VARIANT>
__global__ void __launch_bounds__(256, 2) k_barrier_only(unsigned int* counter, uint32_t iterations) {
	unsigned int target = 0;
	if (VARIANT >= 5) {
		const uint32_t groupSize = VARIANT == 5 ? 8 : (VARIANT == 6 ? 16 : 32);
		for (uint32_t i = 0; i < iterations; i++) { GridBarrierTree(counter, target, groupSize); }
	} else if (VARIANT == 4) {
		cg::grid_group g = cg::this_grid();
		for (uint32_t i = 0; i < iterations; i++) { g.sync(); }
	} else {
		for (uint32_t i = 0; i < iterations; i++) { GridBarrierV<VARIANT>(counter, target); }
	}
}
}  // namespace xf

extern "C" int xf_debug_barrier_us(int device, int variant, int blocksPerSm, int threads, uint32_t iterations, float* outUsPerBarrier) {
	cudaSetDevice(device);
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	unsigned int* counter = nullptr;
	cudaMalloc(&counter, 128 * 64);
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	float best = 1e30f;
	const void* fn = nullptr;
	switch (variant) {
	case 0: fn = (const void*)xf::k_barrier_only<0>; break;
	case 1: fn = (const void*)xf::k_barrier_only<1>; break;
	case 2: fn = (const void*)xf::k_barrier_only<2>; break;
	case 3: fn = (const void*)xf::k_barrier_only<3>; break;
	case 4: fn = (const void*)xf::k_barrier_only<4>; break;
	case 5: fn = (const void*)xf::k_barrier_only<5>; break;
	case 6: fn = (const void*)xf::k_barrier_only<6>; break;
	default: fn = (const void*)xf::k_barrier_only<7>; break;
	}
	for (int rep = 0; rep < 5; rep++) {
		cudaMemset(counter, 0, 128 * 64);
		void* args[] = { (void*)&counter, (void*)&iterations };
		cudaEventRecord(a);
		cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(prop.multiProcessorCount * blocksPerSm), dim3(threads), args, 0, 0);
		cudaEventRecord(b);
		if (e != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) { cudaFree(counter); return -2; }
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, a, b);
		best = ms < best ? ms : best;
	}
	cudaFree(counter);
	*outUsPerBarrier = best * 1000.0f / (float)iterations;
	return 0;
}
