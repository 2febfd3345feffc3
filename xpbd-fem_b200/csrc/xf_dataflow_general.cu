// The barrier-free schedule (xf_dataflow.cu) for calls with volume passes (Geo.cpp:779-782) and post-solve damping sweeps
// (Geo.cpp:346-355, 790-811: Rayleigh_Post / Rayleigh_PostAmortized, PbdDamp) - the reference's default scene
// (wasm/ui.js:76-88) - which used to fall back to the grid-barrier kernel.
//
// Stages of one substep, S = 1 + nC * (1 + volumePasses) + (damping ? 1 : 0) tags per substep on the position records:
//     predict | main sweep, colours 0..nC-1 | volume pass 0, colours 0..nC-1 | ... | post | Damp sweep | PbdDamp sweep
// Without damping sweeps post and predict stay fused into one vertex phase between substeps (as in k_substeps_dataflow).
//
// Position records: the same versioned 32-byte records.  An element of pass q expects, per corner, the colour of the
// previous element around that vertex in the same pass (the code in the index's top byte) or, for the first element
// around the vertex, the LAST element around it in the previous pass (lastCode[vertex]; pass 0: the predict stage).
//
// Velocity records {vx, vy, vz, tag} are versioned by a COUNT: tag = (substep number of the scene's life) << 16 | writes so far.
// The post phase stores count 0.  The damping sweeps act on a slice [lo, hi) of the serial order (amortisation,
// Geo.cpp:794-797), so who wrote last depends on the slice; the count does not need to know: an element expects
//     sweepIndex * inSlice(v) + rank(e, v) - below(v),
// rank = its rank among the elements around v (eRank), below / inSlice = how many of them lie below lo / inside the slice
// (vSlice: eight boundary counts per vertex).  The next predict waits for (number of sweeps) * inSlice(v): every in-slice
// element around v has then read v's position record, so the predict may overwrite it.  Damping elements only READ
// position records (they wait for the post stage's tag) and never touch O.
// In-constraint Rayleigh damping (Paper / Limit) reads O of other threads' vertices and stays on the grid-barrier kernel.
#include "xf_dataflow.cuh"

namespace xf {

namespace {

struct VelRegs {
	double v[3];
	unsigned long long tag;
};
__device__ __forceinline__ VelRegs LoadVel(const double4* V, uint32_t i) {
	VelRegs r;
	double t;
	Load32B(V + i, r.v[0], r.v[1], r.v[2], t);
	r.tag = (unsigned long long)__double_as_longlong(t);
	return r;
}
__device__ __forceinline__ void StoreVel(double4* V, uint32_t i, const double* v, unsigned long long tag) {
	Store32B(V + i, v[0], v[1], v[2], __longlong_as_double((long long)tag));
}

// elements around vertex i below the slice / inside it; slice k of 8, or k == 8: the whole mesh
__device__ __forceinline__ void SliceCounts(const DeviceScene& sc, uint32_t i, uint32_t k, uint32_t& below, uint32_t& inside) {
	const uint2 w = __ldg(sc.vSlice + i);
	const unsigned long long bits = ((unsigned long long)w.y << 32) | w.x;
	if (k >= 8u) {
		below = 0;
		inside = (uint32_t)(bits >> 56) & 0xffu;
	} else {
		below = k == 0 ? 0u : (uint32_t)(bits >> (8u * (k - 1u))) & 0xffu;
		inside = ((uint32_t)(bits >> (8u * k)) & 0xffu) - below;
	}
}

// One element of a volume pass or of the main sweep with general expectations (`expect` computed by the caller).
template <int KIND, int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool GeneralElement(const DeviceScene& sc, const SubstepParams& p, const ElemRec& rec, unsigned mask, const uint32_t (&vid)[4],
                                               const uint32_t (&expectTag)[4], uint32_t newTag, uint32_t sleepNs) {
	typedef Op<EXACT> O;
	const GlobalStore vs = StoreOf(sc);
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n] = vs.LoadX(vid[n]); }
	const ElemCompliance ec = DataflowCompliance<EXACT>(sc, p, rec);
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) { ok[n] = (v[n].flags & kVerMask) == expectTag[n]; }
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
		if (sleepNs) { __nanosleep(sleepNs); }
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!ok[n]) { v[n] = vs.LoadX(vid[n]); }
		}
	}
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n].flags = (v[n].flags & 0xffu) | newTag; }
	ElemRec r = rec;
	r.idx = make_uint4(vid[0], vid[1], vid[2], vid[3]);
	if (KIND == 0) {
		SolveElementGathered<ENERGY, SIMUL, EXACT, false>(vs, p, r, v, ec);
	} else { // SolveVolumeOnly, Fem.cpp:840-867, on the gathered records
		float comp = O::mul(p.compliance, r.volume);
		float P[3][3], F[3][3], g[4][3];
		Edges<EXACT>(v, P);
		DeformationGradient<EXACT>(r, P, F);
		float U = VolumetricFromF<EXACT>(r, F, 1.0f, g);
		ConstrainOne<EXACT, false>(vs, p, r.idx, v, U, g, comp, XF_DIV_MAYBE_ZERO(O, comp, p.dt2), 0.0f);
#pragma unroll
		for (int n = 0; n < 4; n++) { vs.StoreX(vid[n], v[n]); }
	}
	return true;
}

// One element of a damping sweep: KIND 2 = DampElement (Rayleigh, post-solve), KIND 3 = PbdDamp.
template <int KIND, int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool DampingElement(const DeviceScene& sc, const SubstepParams& p, uint32_t e, unsigned mask, uint32_t postTag,
                                               unsigned long long vBase, uint32_t sweepIndex, uint32_t slice, uint32_t sleepNs) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	ElemRec rec;
	if (KIND == 3) {
		rec.idx = LoadElementIdx(sc, e);
	} else {
		LoadElement<kPrefactored, EXACT>(sc, e, rec);
	}
	const uint32_t vid[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	const uint32_t rank = __ldg(sc.eRank + e);
	VertexRegs x[4];
	VelRegs vr[4];
	unsigned long long expectV[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		x[n] = LoadVertex(sc.Xw, vid[n]);
		vr[n] = LoadVel(sc.V, vid[n]);
	}
#pragma unroll
	for (int n = 0; n < 4; n++) {
		uint32_t below, inside;
		SliceCounts(sc, vid[n], slice, below, inside);
		expectV[n] = vBase | (unsigned long long)(sweepIndex * inside + ((rank >> (8 * n)) & 0xffu) - below);
	}
	for (uint32_t spins = 0;; spins++) {
		bool okX[4], okV[4];
		bool all = true;
#pragma unroll
		for (int n = 0; n < 4; n++) {
			okX[n] = (x[n].flags & kVerMask) == postTag;
			okV[n] = vr[n].tag == expectV[n];
			all = all && okX[n] && okV[n];
		}
		if (__all_sync(mask, all)) { break; }
		if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { return false; }
		if (sleepNs) { __nanosleep(sleepNs); }
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!okX[n]) { x[n] = LoadVertex(sc.Xw, vid[n]); }
			if (!okV[n]) { vr[n] = LoadVel(sc.V, vid[n]); }
		}
	}
	double vel[4][3];
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { vel[n][k] = vr[n].v[k]; }
	}
	if (KIND == 2) {
		DampElementGathered<ENERGY, SIMUL, EXACT>(p, rec, x, vel);
	} else {
		PbdDampGathered<EXACT>(p, __ldg(sc.eArea + e), x, vel);
	}
#pragma unroll
	for (int n = 0; n < 4; n++) { StoreVel(sc.V, vid[n], vel[n], expectV[n] + 1ull); }
	return true;
}

}  // namespace

template <int ENERGY, bool SIMUL, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_substeps_dataflow_general(const DeviceScene sc, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                                      uint32_t verBase, unsigned long long vEpoch, uint32_t tuning) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t gsize = gridDim.x * blockDim.x;
	const uint32_t warpSlot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u;
	const uint32_t nC = p.nColors, VP = p.volumePasses;
	const bool anyDamp = p.doDamp || p.doPbdDamp;
	const uint32_t nSweeps = (p.doDamp ? 1u : 0u) + (p.doPbdDamp ? 1u : 0u);
	const uint32_t stride = 1u + nC * (1u + VP) + (anyDamp ? 1u : 0u);
	const uint32_t sleepNs = tuning & 0x7fffu, elemSleepNs = tuning >> 16;
	const bool spare = warpSlot >= sc.maxColorSize;
	bool dead = false;
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		if (closing && anyDamp) { break; } // the last substep's post phase already ran, before its damping sweeps
		const uint32_t stageBase = verBase + s * stride;             // tag of this substep's predict stage
		const uint32_t prevLast = stageBase - stride + nC * VP;      // + lastCode: last writer of the previous substep's last pass
		// ---- vertex phase at the head of the substep: [post of the previous substep +] predict
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (!has) { continue; }
			if (!anyDamp) {
				const uint32_t lc = __ldg(sc.lastCode + i);
				const uint32_t expectTag = (lc ? prevLast + lc : stageBase - stride) << 8;
				dead = !DataflowVertex<EXACT>(sc, p, i, mask, s > 0, !closing, s > 0, expectTag, stageBase << 8, sleepNs, s > 0 ? VaryRow(p, s - 1u) : nullptr);
			} else {
				// predict only; the damping sweeps of the previous substep must be through with this vertex (they read its position)
				if (s > 0) {
					const uint32_t prevSlice = p.rayleigh == XF_RAYLEIGH_POST_AMORTIZED ? (p.tickId + s - 1u) % XF_AMORTIZATION_PERIOD : 8u;
					uint32_t below, inside;
					SliceCounts(sc, i, prevSlice, below, inside);
					const unsigned long long want = ((vEpoch + s - 1ull) << 16) | (unsigned long long)(nSweeps * inside);
					VelRegs vr = LoadVel(sc.V, i);
					for (uint32_t spins = 0;; spins++) {
						const bool ok = vr.tag == want;
						if (__all_sync(mask, ok)) { break; }
						if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { dead = true; break; }
						if (sleepNs) { __nanosleep(sleepNs); }
						if (!ok) { vr = LoadVel(sc.V, i); }
					}
					if (dead) { continue; }
				}
				VertexRegs v = LoadVertex(sc.Xw, i);
				VertexPhaseBody<EXACT>(sc, p, i, v, false, true);
				v.flags = (v.flags & 0xffu) | (stageBase << 8);
				StoreVertex(sc.Xw, i, v);
			}
		}
		if (closing) { break; }
		// ---- main sweep and volume passes.  The element record is loaded where it is used (carrying the next stage's record in
		// registers across the solve, as k_substeps_dataflow does, cost more in spills here than it hid: 4.17 vs 3.96 ms per 50
		// damped substeps at 1M tets); the records of the next two stages are pulled into L1 instead.
		if (!spare) {
			const uint32_t nStages = nC * (1u + VP);
			for (uint32_t t = 0; t < nStages; t++) {
				const uint32_t q = t / nC, c = t - q * nC;
				const uint32_t passBase = stageBase + nC * q; // + 1 + colour = the tag an element of this pass writes
				const uint32_t end = p.colorStart[c + 1];
				for (uint32_t e0 = p.colorStart[c] + warpSlot; e0 < end; e0 += gsize) {
					const bool has = e0 + lane < end && !dead;
					const unsigned mask = __ballot_sync(0xffffffffu, has);
					if (!has) { continue; }
					ElemRec rec;
					DataflowLoad<ENERGY, EXACT>(sc, e0 + lane, rec);
					const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
					uint32_t vid[4], expectTag[4];
#pragma unroll
					for (int n = 0; n < 4; n++) {
						vid[n] = raw[n] & 0x00ffffffu;
						const uint32_t pc = raw[n] >> 24;
						// first element around the vertex in this pass: the last one of the previous pass, or the predict stage
						const uint32_t first = q == 0 ? stageBase : passBase - nC + (uint32_t)__ldg(sc.lastCode + vid[n]);
						expectTag[n] = (pc ? passBase + pc : first) << 8;
					}
					const uint32_t newTag = (passBase + 1u + c) << 8;
					if (q == 0) {
						dead = !GeneralElement<0, ENERGY, SIMUL, EXACT>(sc, p, rec, mask, vid, expectTag, newTag, elemSleepNs);
					} else {
						dead = !GeneralElement<1, ENERGY, SIMUL, EXACT>(sc, p, rec, mask, vid, expectTag, newTag, elemSleepNs);
					}
				}
				for (uint32_t ahead = 1; ahead <= 2u; ahead++) { // wraps into the next pass / substep
					const uint32_t ca = (c + ahead) % nC;
					if (p.colorStart[ca] + warpSlot + lane < p.colorStart[ca + 1]) { DataflowPrefetch<ENERGY, EXACT>(sc, p.colorStart[ca] + warpSlot + lane); }
				}
			}
		}
		if (!anyDamp) { continue; }
		// ---- post phase: ground, locks, manipulator, handles, velocities (count 0 of this substep's V tags)
		const uint32_t postStage = stageBase + 1u + nC * (1u + VP);
		const unsigned long long vBase = (vEpoch + s) << 16;
		uint32_t lo, hi;
		DampSlice(p, sc.nT, p.tickId + s, lo, hi);
		if (!spare) { // this thread's damping elements (first chunk of every colour that meets the slice): records into L1 ahead of the post phase
			for (uint32_t c = 0; c < nC; c++) {
				const uint32_t e = max(p.colorStart[c], lo) + warpSlot + lane;
				if (e < min(p.colorStart[c + 1], hi)) {
					asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eA + e));
					asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eRank + e));
					if (p.doDamp) {
						asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eB + e));
						asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eC + e));
					}
					if (p.doPbdDamp) { asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.eArea + e)); }
				}
			}
		}
		for (uint32_t i0 = warpSlot; i0 < sc.nV; i0 += gsize) {
			const uint32_t i = i0 + lane;
			const bool has = i < sc.nV && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (!has) { continue; }
			const uint32_t lc = __ldg(sc.lastCode + i);
			const uint32_t expectTag = (lc ? stageBase + nC * VP + lc : stageBase) << 8;
			VertexRegs v = LoadVertex(sc.Xw, i);
			for (uint32_t spins = 0;; spins++) {
				const bool ok = (v.flags & kVerMask) == expectTag;
				if (__all_sync(mask, ok)) { break; }
				if (SpinGiveUp(sc.errDev, sc.errHost, mask, spins, sc.spinLimit)) { dead = true; break; }
				if (sleepNs) { __nanosleep(sleepNs); }
				if (!ok) { v = LoadVertex(sc.Xw, i); }
			}
			if (dead) { continue; }
			VertexPhaseBody<EXACT>(sc, p, i, v, true, false, __longlong_as_double((long long)vBase), VaryRow(p, s));
			v.flags = (v.flags & 0xffu) | (postStage << 8);
			StoreVertex(sc.Xw, i, v);
		}
		if (spare) { continue; }
		// ---- damping sweeps over the slice [lo, hi) of the serial order (= device order on this schedule)
		const uint32_t slice = p.rayleigh == XF_RAYLEIGH_POST_AMORTIZED ? (p.tickId + s) % XF_AMORTIZATION_PERIOD : 8u;
		for (uint32_t sweep = 0; sweep < 2u; sweep++) {
			if (sweep == 0 ? !p.doDamp : !p.doPbdDamp) { continue; }
			const uint32_t sweepIndex = sweep == 0 ? 0u : (p.doDamp ? 1u : 0u);
			for (uint32_t c = 0; c < nC; c++) {
				const uint32_t b = max(p.colorStart[c], lo), end = min(p.colorStart[c + 1], hi);
				for (uint32_t e0 = b + warpSlot; e0 < end; e0 += gsize) { // empty when the colour misses the slice
					const bool has = e0 + lane < end && !dead;
					const unsigned mask = __ballot_sync(0xffffffffu, has);
					if (!has) { continue; }
					if (sweep == 0) {
						dead = !DampingElement<2, ENERGY, SIMUL, EXACT>(sc, p, e0 + lane, mask, postStage << 8, vBase, sweepIndex, slice, elemSleepNs);
					} else {
						dead = !DampingElement<3, ENERGY, SIMUL, EXACT>(sc, p, e0 + lane, mask, postStage << 8, vBase, sweepIndex, slice, elemSleepNs);
					}
				}
			}
		}
	}
}

namespace {
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct GeneralRunner {
	static cudaError_t Run(const DeviceScene& sc, const SubstepParams& p, uint32_t nSubsteps, int smCount, uint32_t verBase, unsigned long long vEpoch,
	                       uint32_t tuning, cudaStream_t st, uint64_t* launches) {
		auto fn = k_substeps_dataflow_general<ENERGY, SIMUL, EXACT>;
		static OccupancyCache cache;
		int perSm = 0;
		cudaError_t e0 = cache.Get((const void*)fn, 256, 0, false, &perSm);
		if (e0 != cudaSuccess) { return e0; }
		void* args[] = { (void*)&sc, (void*)&p, (void*)&nSubsteps, (void*)&verBase, (void*)&vEpoch, (void*)&tuning };
		const int grid = perSm * smCount;
		cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)grid), dim3((unsigned)DataflowBlockThreads(sc, grid)), args, 0, st);
		++*launches;
		return e;
	}
};
}  // namespace

uint32_t DataflowGeneralStride(const SubstepParams& p) {
	return 1u + p.nColors * (1u + p.volumePasses) + ((p.doDamp || p.doPbdDamp) ? 1u : 0u);
}

cudaError_t LaunchSubstepsDataflowGeneral(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                          uint64_t vEpoch, uint32_t tuning, cudaStream_t stream, uint64_t* launchCount) {
	return DispatchConfig<GeneralRunner>(p.energy, p.simultaneous != 0, exact, false, sc, p, nSubsteps, smCount, verBase, (unsigned long long)vEpoch, tuning,
	                                     stream, launchCount);
}

}  // namespace xf
