// Barrier-free schedule, TWO WARPS PER 32 ELEMENTS (k_substeps_pair).
//
// On the barrier-free schedule (xf_dataflow.cu) a colour stage costs one pass of the element's dependence chain, and
// that chain is ~1.2 us of arithmetic for ONE thread: ~630 instructions whose dependence graph is only ~2.3 wide, issued
// in order.  Nothing else can run in that time - every vertex is busy in every stage - so the only way to shorten a stage
// is to put more lanes on one element.  Lanes of one warp would diverge (the two halves of the work are different code),
// so the two halves go to two WARPS of the same CTA that meet at named barriers and trade operands through shared memory:
//     warp D (deviatoric): corners 0,1 | prefactored I1 / Yeoh term  -> U0, g0[4] | updates and scatters corners 0,1
//     warp V (volumetric): corners 2,3 | F, adj F, J                 -> U1, g1[4] | updates and scatters corners 2,3
//   exchange 1: the four corner positions and inverse masses (each warp gathered - and waited for - two of them)
//   exchange 2: the two energies and gradient sets; both warps then form the same three w-sums and solve the same 2x2
// Every value is still produced by exactly the reference's operation sequence (each by one warp), so results stay
// bit-identical; the instruction total per element is about the same (~2 x 330), the chain per element about half.
// Covers the prefactored energies (MixedSel, YeohSkinFast) in simultaneous mode, undamped - the headline configuration;
// everything else runs on k_substeps_dataflow / k_substeps_persistent.
#include "xf_dispatch.cuh"
#include "xf_element.cuh"
#include "xf_phase.cuh"

namespace xf {

namespace {

constexpr uint32_t kPairVerMask = 0xffffff00u;
constexpr uint32_t kPairSpinLimit = 1u << 24;
constexpr int kPairThreads = 192; // 6 warps = 3 pairs per CTA, 3 CTAs per SM: 9 pairs = 288 element slots per SM

// Shared-memory exchange area of one warp pair, [value][lane] so that every access is conflict-free.
struct PairExchange {
	double x[4][3][32]; // corner positions
	float w[4][32];     // inverse masses
	float g[2][12][32]; // g[0] = deviatoric gradients of corners 0..3 (x,y,z), g[1] = volumetric
	float U[2][32];
};

__device__ __forceinline__ void PairSync(uint32_t pairInCta) {
	asm volatile("bar.sync %0, 64;" ::"r"(pairInCta + 1u) : "memory");
}

template <bool EXACT>
__device__ __forceinline__ void PairVertex(const DeviceScene& sc, const SubstepParams& p, uint32_t i, unsigned mask, bool doPost, bool doPredict,
                                           bool wait, uint32_t expectTag, uint32_t newTag, uint32_t sleepNs) {
	VertexRegs v = LoadVertex(sc.Xw, i);
	if (wait) {
		for (uint32_t spins = 0;; spins++) {
			const bool ok = (v.flags & kPairVerMask) == expectTag;
			if (__all_sync(mask, ok)) { break; }
			if (spins > kPairSpinLimit) { __trap(); }
			if (sleepNs) { __nanosleep(sleepNs); }
			if (!ok) { v = LoadVertex(sc.Xw, i); }
		}
	}
	VertexPhaseBody<EXACT>(sc, p, i, v, doPost, doPredict);
	v.flags = (v.flags & 0xffu) | newTag;
	StoreVertex(sc.Xw, i, v);
}

// One element, one of its two warps.  ROLE 0 = deviatoric (corners 0,1), ROLE 1 = volumetric (corners 2,3).
// `has` = this lane has an element (tail chunk); lanes without one still take part in the two barriers.
template <int ENERGY, bool EXACT, int ROLE>
__device__ __forceinline__ void PairElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, bool has, unsigned mask, PairExchange& ex,
                                            uint32_t pairInCta, uint32_t lane, uint32_t stageBase, uint32_t c) {
	typedef Op<EXACT> O;
	constexpr int kMine0 = ROLE == 0 ? 0 : 2, kMine1 = kMine0 + 1, kOther0 = ROLE == 0 ? 2 : 0, kOther1 = kOther0 + 1;
	const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	VertexRegs mine[2];
	uint32_t vid[2];
	ElemCompliance ec;
	if (has) {
		uint32_t expectTag[2];
		vid[0] = raw[kMine0] & 0x00ffffffu;
		vid[1] = raw[kMine1] & 0x00ffffffu;
		expectTag[0] = (stageBase + (raw[kMine0] >> 24)) << 8;
		expectTag[1] = (stageBase + (raw[kMine1] >> 24)) << 8;
		mine[0] = LoadVertex(sc.Xw, vid[0]);
		mine[1] = LoadVertex(sc.Xw, vid[1]);
		ec = ComplianceOf<EXACT>(p, rec.volume);
		for (uint32_t spins = 0;; spins++) {
			const bool ok0 = (mine[0].flags & kPairVerMask) == expectTag[0], ok1 = (mine[1].flags & kPairVerMask) == expectTag[1];
			if (__all_sync(mask, ok0 && ok1)) { break; }
			if (spins > kPairSpinLimit) { __trap(); }
			if (!ok0) { mine[0] = LoadVertex(sc.Xw, vid[0]); }
			if (!ok1) { mine[1] = LoadVertex(sc.Xw, vid[1]); }
		}
#pragma unroll
		for (int k = 0; k < 3; k++) {
			ex.x[kMine0][k][lane] = mine[0].x[k];
			ex.x[kMine1][k][lane] = mine[1].x[k];
		}
		ex.w[kMine0][lane] = mine[0].w;
		ex.w[kMine1][lane] = mine[1].w;
	}
	PairSync(pairInCta); // exchange 1: positions and inverse masses
	float w[4], gMine[4][3], UMine = 0.0f;
	if (has) {
		VertexRegs v[4];
		v[kMine0] = mine[0];
		v[kMine1] = mine[1];
#pragma unroll
		for (int k = 0; k < 3; k++) {
			v[kOther0].x[k] = ex.x[kOther0][k][lane];
			v[kOther1].x[k] = ex.x[kOther1][k][lane];
		}
		w[kMine0] = mine[0].w;
		w[kMine1] = mine[1].w;
		w[kOther0] = ex.w[kOther0][lane];
		w[kOther1] = ex.w[kOther1][lane];
		float P[3][3];
		Edges<EXACT>(v, P);
		if (ROLE == 0) {
			UMine = PrefactoredI1<EXACT>(rec, P, gMine);
			if (ENERGY == XF_ENERGY_YEOH_SKIN_FAST) {
				const float IM = O::sub(UMine, 3.0f);
				UMine = fmaxf(0.0001f, YeohEnergy<EXACT>(IM));
				const float gScale = YeohSlope<EXACT>(IM);
#pragma unroll
				for (int n = 0; n < 4; n++) {
#pragma unroll
					for (int k = 0; k < 3; k++) { gMine[n][k] = O::mul(gMine[n][k], gScale); }
				}
			}
		} else {
			float F[3][3];
			DeformationGradient<EXACT>(rec, P, F);
			UMine = VolumetricFromF<EXACT>(rec, F, p.a, gMine);
		}
#pragma unroll
		for (int n = 0; n < 4; n++) {
#pragma unroll
			for (int k = 0; k < 3; k++) { ex.g[ROLE][3 * n + k][lane] = gMine[n][k]; }
		}
		ex.U[ROLE][lane] = UMine;
	}
	PairSync(pairInCta); // exchange 2: energies and gradients
	if (has) {
		float gOther[4][3];
#pragma unroll
		for (int n = 0; n < 4; n++) {
#pragma unroll
			for (int k = 0; k < 3; k++) { gOther[n][k] = ex.g[1 - ROLE][3 * n + k][lane]; }
		}
		const float UOther = ex.U[1 - ROLE][lane];
		const float(&g0)[4][3] = ROLE == 0 ? gMine : gOther;
		const float(&g1)[4][3] = ROLE == 0 ? gOther : gMine;
		const float U0 = ROLE == 0 ? UMine : UOther, U1 = ROLE == 0 ? UOther : UMine;
		// EnergyXpbdConstrainSimultaneous<.., 2>, undamped (Xpbd.h:154-179): both warps evaluate it, each updates its two corners
		float w00 = 1.0e-22f, w10 = 1.0e-22f, w11 = 1.0e-22f;
#pragma unroll
		for (int n = 0; n < 4; n++) { w00 = O::add(w00, O::mul(w[n], O::dot(g0[n], g0[n]))); }
#pragma unroll
		for (int n = 0; n < 4; n++) { w10 = O::add(w10, O::mul(w[n], O::dot(g1[n], g0[n]))); }
#pragma unroll
		for (int n = 0; n < 4; n++) { w11 = O::add(w11, O::mul(w[n], O::dot(g1[n], g1[n]))); }
		const float A0 = O::add(w00, O::mul(O::mul(2.0f, U0), ec.alpha0));
		const float b0 = O::mul(-2.0f, U0);
		const float A2 = O::add(w11, O::mul(O::mul(2.0f, U1), ec.alpha1));
		const float b1 = O::mul(-2.0f, U1);
		float l0, l1;
		Cramer2<EXACT>(A0, w10, A2, b0, b1, l0, l1);
		const uint32_t newTag = (stageBase + 1u + c) << 8;
#pragma unroll
		for (int m = 0; m < 2; m++) {
			const int n = kMine0 + m;
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float acc = O::add(O::mul(l0, g0[n][k]), O::mul(l1, g1[n][k]));
				mine[m].x[k] = O::dadd(mine[m].x[k], (double)O::mul(w[n], acc));
			}
			mine[m].flags = (mine[m].flags & 0xffu) | newTag;
			StoreVertex(sc.Xw, vid[m], mine[m]);
		}
	}
}

template <int ENERGY, bool EXACT>
__device__ __forceinline__ void PairLoad(const DeviceScene& sc, uint32_t e, ElemRec& rec) {
	LoadElementFrom<true, EXACT>(sc.eAd, sc, e, rec);
}
template <bool EXACT>
__device__ __forceinline__ void PairPrefetch(const DeviceScene& sc, uint32_t e) {
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eAd + e));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eB + e));
	if (EXACT) { asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eC + e)); }
}

}  // namespace

template <int ENERGY, bool EXACT>
__global__ void __maxnreg__(112) k_substeps_pair(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                                   uint32_t verBase, uint32_t tuning) {
	extern __shared__ __align__(16) unsigned char pairSmem[];
	const uint32_t lane = threadIdx.x & 31u, warpInCta = threadIdx.x >> 5;
	const uint32_t pairInCta = warpInCta >> 1, role = warpInCta & 1u, pairsPerCta = blockDim.x >> 6;
	PairExchange& ex = reinterpret_cast<PairExchange*>(pairSmem)[pairInCta];
	const uint32_t gsize = gridDim.x * blockDim.x;
	// vertex work: warp-sized chunks round-robin over CTAs; element work: the same, per warp PAIR
	const uint32_t warpSlot = (warpInCta * gridDim.x + blockIdx.x) * 32u;
	const uint32_t pairSlot = (pairInCta * gridDim.x + blockIdx.x) * 32u, pairStride = gridDim.x * pairsPerCta * 32u;
	const uint32_t nC = p.nColors;
	const uint32_t stride = nC + 1u;
	const uint32_t sleepNs = tuning & 0x7fffu;
	ElemRec rec;
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		const uint32_t stageBase = verBase + s * stride;
		if (!closing && p.colorStart[0] + pairSlot + lane < p.colorStart[1]) { PairLoad<ENERGY, EXACT>(sc, p.colorStart[0] + pairSlot + lane, rec); }
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
				PairVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs);
			}
		}
		if (closing) { break; }
		for (uint32_t c = 0; c < nC; c++) {
			const uint32_t end = p.colorStart[c + 1];
			bool first = true;
			for (uint32_t e0 = p.colorStart[c] + pairSlot; e0 < end; e0 += pairStride) {
				const bool has = e0 + lane < end;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (!first && has) { PairLoad<ENERGY, EXACT>(sc, e0 + lane, rec); }
				first = false;
				if (role == 0) {
					PairElement<ENERGY, EXACT, 0>(sc, p, rec, has, mask, ex, pairInCta, lane, stageBase, c);
				} else {
					PairElement<ENERGY, EXACT, 1>(sc, p, rec, has, mask, ex, pairInCta, lane, stageBase, c);
				}
			}
			if (c + 1 < nC && p.colorStart[c + 1] + pairSlot + lane < p.colorStart[c + 2]) {
				PairLoad<ENERGY, EXACT>(sc, p.colorStart[c + 1] + pairSlot + lane, rec);
			}
			const uint32_t c2 = c + 2 < nC ? c + 2 : c + 2 - nC; // wraps into the next substep
			if (p.colorStart[c2] + pairSlot + lane < p.colorStart[c2 + 1]) { PairPrefetch<EXACT>(sc, p.colorStart[c2] + pairSlot + lane); }
		}
	}
}

namespace {
template <int ENERGY, bool EXACT>
cudaError_t RunPair(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, uint32_t verBase, uint32_t tuning, cudaStream_t st,
                    uint64_t* launches) {
	auto fn = k_substeps_pair<ENERGY, EXACT>;
	const size_t smem = sizeof(PairExchange) * (kPairThreads / 64);
	static int perSm = 0; // per instantiation
	if (perSm == 0) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) { return e; }
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, kPairThreads, smem);
		if (e != cudaSuccess) { return e; }
		if (perSm < 1) { return cudaErrorLaunchOutOfResources; }
	}
	void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&tuning };
	cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)(perSm * smCount)), dim3(kPairThreads), args, smem, st);
	++*launches;
	return e;
}
}  // namespace

bool PairKernelCovers(const SubstepParams& p) {
	return p.simultaneous != 0 && (p.energy == XF_ENERGY_MIXED_SEL || p.energy == XF_ENERGY_YEOH_SKIN_FAST);
}

cudaError_t LaunchSubstepsPair(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                               uint32_t tuning, cudaStream_t stream, uint64_t* launchCount) {
	if (p.energy == XF_ENERGY_MIXED_SEL) {
		return exact ? RunPair<XF_ENERGY_MIXED_SEL, true>(sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount)
		             : RunPair<XF_ENERGY_MIXED_SEL, false>(sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount);
	}
	return exact ? RunPair<XF_ENERGY_YEOH_SKIN_FAST, true>(sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount)
	             : RunPair<XF_ENERGY_YEOH_SKIN_FAST, false>(sc, p, nSubsteps, smCount, verBase, tuning, stream, launchCount);
}

}  // namespace xf
