// Device functions shared by the sweep kernels (xf_kernels.cu, xf_dataflow*.cu): the vertex
// phase and the damping-slice helpers.
#pragma once

#include "xf_element.cuh"

namespace xf {

// ------------------------------------------------------------------------------------------------
// Vertex phases.  post = ground (x1) -> locks -> manipulator -> handles (x2) -> velocity update
// (Geo.cpp:318-344); predict = Geo.cpp:307-312.  Fused into one pass between substeps when no damping
// sweep separates them.
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void DragTowards(VertexRegs& v, const float* target, float c18) {
	typedef Op<EXACT> O;
	float k = O::div(v.w, O::add(fmaxf(0.000001f, v.w), c18));
#pragma unroll
	for (int c = 0; c < 3; c++) {
		float d = O::mul(O::sub(target[c], __double2float_rn(v.x[c])), k);
		v.x[c] = O::dadd(v.x[c], (double)d);
	}
}

// Body of the vertex phase for vertex `i` (global id) whose position record is already in registers; O, V (and X0
// for the right lock) are read/written in global memory.  The caller stores `v` back to wherever it lives.
// `vTag` goes into the spare fourth word of the V record: the barrier-free kernels with damping sweeps version V records
// the way they version position records (xf_dataflow_general.cu); 0.0 everywhere else.
// `vary`: this substep's row of SubstepParams::vary (lock transform, manipulator target), or nullptr.
__device__ __forceinline__ const float* VaryRow(const SubstepParams& p, uint32_t substep) {
	return p.vary ? p.vary + (size_t)kVaryFloats * substep : nullptr;
}
template <bool EXACT>
__device__ __forceinline__ void VertexPhaseBody(const DeviceScene& sc, const SubstepParams& p, uint32_t i, VertexRegs& v, bool doPost, bool doPredict,
                                                double vTag = 0.0, const float* vary = nullptr) {
	typedef Op<EXACT> O;
	double o[3], vel[3];
	LoadD3(sc.O, i, o);
	if (doPost) {
		if (p.groundOn) {
			double y0 = (double)p.groundY;
			if (v.x[1] < y0) {
				double keepT = (double)p.groundKeep;
				v.x[1] = y0;
				v.x[0] = O::dadd(o[0], O::dmul(O::dsub(v.x[0], o[0]), keepT));
				v.x[2] = O::dadd(o[2], O::dmul(O::dsub(v.x[2], o[2]), keepT));
			}
		}
		if (p.lockLeft && (v.flags & XF_VERT_LEFT)) {
			v.x[0] = o[0]; v.x[1] = o[1]; v.x[2] = o[2];
			v.w = 0.0f;
		}
		if (p.lockRight && (v.flags & XF_VERT_RIGHT)) {
			double x0d[3];
			LoadD3(sc.X0, i, x0d);
			float x0[3] = { __double2float_rn(x0d[0]), __double2float_rn(x0d[1]), __double2float_rn(x0d[2]) };
#pragma unroll
			for (int r = 0; r < 3; r++) {
				const float t0 = vary ? __ldg(vary + 0 + r) : p.lockT[0 + r], t1 = vary ? __ldg(vary + 4 + r) : p.lockT[4 + r],
				            t2 = vary ? __ldg(vary + 8 + r) : p.lockT[8 + r];
				float t = O::dot(t0, t1, t2, x0[0], x0[1], x0[2]);
				double q = (double)O::add(p.origin[r], t);
				v.x[r] = q;
				o[r] = q;
			}
			v.w = 0.0f;
		}
		if (p.manipOn && i == p.manipIdx) {
			const float target[3] = { vary ? __ldg(vary + 12) : p.manipTarget[0], vary ? __ldg(vary + 13) : p.manipTarget[1],
				                      vary ? __ldg(vary + 14) : p.manipTarget[2] };
			DragTowards<EXACT>(v, target, p.c18);
		}
		for (uint32_t h = 0; h < p.handleCount; h++) {
			if (p.handleIdx[h] == i) { DragTowards<EXACT>(v, p.handleTarget[h], p.c18); }
		}
		double invDt = (double)p.invDt;
#pragma unroll
		for (int k = 0; k < 3; k++) { vel[k] = O::dmul(O::dsub(v.x[k], o[k]), invDt); }
	} else {
		LoadD3(sc.V, i, vel);
	}
	if (doPredict) {
		double g[3] = { (double)p.gdtX, (double)p.gdtY, 0.0 };
		double keep = (double)p.keep;
		double ddt = (double)p.dt;
#pragma unroll
		for (int k = 0; k < 3; k++) {
			vel[k] = O::dadd(vel[k], g[k]);
			vel[k] = O::dmul(vel[k], keep);
			o[k] = v.x[k];
			v.x[k] = O::dadd(v.x[k], O::dmul(vel[k], ddt));
		}
	}
	StoreD3(sc.O, i, o);
	Store32B(sc.V + i, vel[0], vel[1], vel[2], vTag);
}

template <bool EXACT>
__device__ __forceinline__ void VertexPhase(const DeviceScene& sc, const SubstepParams& p, uint32_t i, bool doPost, bool doPredict,
                                            const float* vary = nullptr) {
	VertexRegs v = LoadVertex(sc.Xw, i);
	VertexPhaseBody<EXACT>(sc, p, i, v, doPost, doPredict, 0.0, vary);
	StoreVertex(sc.Xw, i, v);
}

// Damping sweeps act on the slice [lo, hi) of the SERIAL order (Geo.cpp:794-797); on a partitioned mesh the local planes hold a
// subset of it, so membership is decided per element from its (global) serial position.
__device__ __forceinline__ bool InSlice(const DeviceScene& sc, uint32_t e, uint32_t lo, uint32_t hi) {
	const uint32_t pos = __ldg(sc.canonPos + e);
	return pos >= lo && pos < hi;
}

// Amortised damping slice [count*k/8, count*(k+1)/8) of the serial order, Geo.cpp:794-797.
__device__ __forceinline__ void DampSlice(const SubstepParams& p, uint32_t nT, uint32_t tick, uint32_t& lo, uint32_t& hi) {
	if (p.rayleigh == XF_RAYLEIGH_POST_AMORTIZED) {
		uint32_t k = tick % XF_AMORTIZATION_PERIOD;
		lo = (uint32_t)(((uint64_t)nT * k) / XF_AMORTIZATION_PERIOD);
		hi = (uint32_t)(((uint64_t)nT * (k + 1)) / XF_AMORTIZATION_PERIOD);
	} else {
		lo = 0;
		hi = nT;
	}
}

}  // namespace xf
