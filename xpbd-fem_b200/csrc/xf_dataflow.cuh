// Device functions shared by the barrier-free kernels (xf_dataflow.cu, xf_dataflow_general.cu): tag waits on versioned
// vertex records, the element step, record prefetch.  See xf_dataflow.cu for the protocol.
#pragma once

#include "xf_dispatch.cuh"
#include "xf_element.cuh"
#include "xf_phase.cuh"

namespace xf {

namespace {

constexpr uint32_t kVerMask = 0xffffff00u;

// alpha of the undamped solves: precomputed per element for the call's settings (DeviceScene::eAlpha, k_element_alpha).
// Measured at 1M tets (profiles/r2_bench_alpha_plane_ab.log): 2.407 vs 2.510 ms per 50 substeps - the two chained IEEE divisions
// with their slow-path checks were ~25 of ~700 instructions of every element, and the stage time follows the instruction count.
template <bool EXACT>
__device__ __forceinline__ ElemCompliance DataflowCompliance(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec) {
	return ElemCompliance{ 0.0f, 0.0f, rec.alpha0, rec.alpha1 }; // comp0 / comp1 only feed in-constraint damping, which never runs here
}

// A record that never reaches the expected stage means a broken schedule (or a caller that rewrote the state while a
// launch was in flight): report it (DeviceScene::errDev / errHost -> XF_ERR_CUDA from xf_sync and the getters) and drain the
// kernel instead of hanging the device or trapping (a trap would poison the CUDA context of every scene of the process).
// DeviceScene::spinLimit = 2^24 polls is seconds, a healthy wait is a few polls.  A thread that gave up is `dead`: it walks
// the rest of its loops without touching anything.

// Waiting is warp-uniform on purpose.  A lane that left a spin loop early would be parked at the loop's reconvergence
// point, and independent thread scheduling releases parked lanes when the spinning ones yield (that is how it guarantees
// progress): the warp would then run the ~750-instruction element body once per group of lanes.  Measured: 2x slower at
// every mesh size.  With a vote over the lanes that have work, the warp leaves the loop as one.
template <bool EXACT>
__device__ __forceinline__ bool DataflowVertex(const DeviceScene& sc, const SubstepParams& p, uint32_t i, unsigned mask, bool doPost, bool doPredict,
                                               bool wait, uint32_t expectTag, uint32_t newTag, uint32_t sleepNs, const float* vary = nullptr) {
	VertexRegs v = LoadVertex(sc.Xw, i);
	if (wait) {
		for (uint32_t spins = 0;; spins++) {
			const bool ok = (v.flags & kVerMask) == expectTag;
			if (__all_sync(mask, ok)) { break; }
			if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
			if (sleepNs) { __nanosleep(sleepNs); }
			if (!ok) { v = LoadVertex(sc.Xw, i); }
		}
	}
	VertexPhaseBody<EXACT>(sc, p, i, v, doPost, doPredict, 0.0, vary);
	v.flags = (v.flags & 0xffu) | newTag;
	StoreVertex(sc.Xw, i, v);
	return true;
}

// One element: spin-gather the four versioned records, solve, scatter with this stage's tag.  `mask` = the lanes of
// this warp that run an element in this step (all of them call this function together).
template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool DataflowElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, unsigned mask, uint32_t stageBase,
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
	const ElemCompliance ec = DataflowCompliance<EXACT>(sc, p, rec); // precomputed, or two chained divisions in the shadow of the gather
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
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
	return true;
}

// The record of the stage after next: pulled into L1 now (the planes are read with ld.global.nc, which allocates in L1),
// so that the load issued right before it is needed costs an L1 hit instead of an L2 / HBM round trip in the
// stage-to-stage dependence chain.
template <int ENERGY, bool EXACT>
__device__ __forceinline__ void DataflowPrefetch(const DeviceScene& sc, uint32_t e) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eAd + e));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eAlpha + e));
	asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eB + e));
	if (kPrefactored && EXACT) { asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eC + e)); }
}

template <int ENERGY, bool EXACT>
__device__ __forceinline__ void DataflowLoad(const DeviceScene& sc, uint32_t e, ElemRec& rec) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	LoadElementFrom<kPrefactored, EXACT>(sc.eAd, sc, e, rec);
	const float2 al = __ldg(sc.eAlpha + e);
	rec.alpha0 = al.x;
	rec.alpha1 = al.y;
}

}  // namespace

}  // namespace xf
