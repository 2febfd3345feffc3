// Per-element device math of the small-step XPBD linear-tet solve (one thread = one element).
//
// Restates, for Nodes == 4 / Points == 1, what the reference does in
//   SolveElementMixed            Fem.cpp:437-564
//   CalculateIncompressibleNeoHookeanEnergyAndGradients  Fem.cpp:163-192
//   CalculateDeformationGradient / adjugate / ApplyEnergyGradient   Fem.cpp:284-354
//   EnergyXpbdConstrain / EnergyXpbdConstrainSimultaneous / RayleighDamp / PbdDamp   Xpbd.h:86-348
//   SolveVolumeOnly              Fem.cpp:840-867
// but as straight-line register code: no vec/mat classes, no arrays in local memory.
//
// EXACT == true : every fp32/fp64 operation goes through a round-to-nearest intrinsic
//                 (__fmul_rn/__fadd_rn/__fdiv_rn/__dadd_rn...), which nvcc never contracts into FMAs,
//                 in the reference's association order => bit-identical to the reference compiled with
//                 -ffp-contract=off.
// EXACT == false: same formulas with ordinary operators (FMA contraction allowed), QQ/QR recomputed from Qi.
#pragma once

#include "xf_scene.h"

namespace xf {

// x / y where x is a per-call constant that is exactly +0 at nu == 0.5 (1/lambda): 0 / y == +0 for y > 0, so the
// IEEE division (and its slow-path call for zero numerators) can be skipped without changing a bit.
#define XF_DIV_MAYBE_ZERO(O, x, y) (((x) == 0.0f && (y) > 0.0f) ? 0.0f : O::div((x), (y)))

template <bool EXACT>
struct Op {
	static __device__ __forceinline__ float mul(float a, float b) { if (EXACT) { return __fmul_rn(a, b); } else { return a * b; } }
	static __device__ __forceinline__ float add(float a, float b) { if (EXACT) { return __fadd_rn(a, b); } else { return a + b; } }
	static __device__ __forceinline__ float sub(float a, float b) { if (EXACT) { return __fsub_rn(a, b); } else { return a - b; } }
	static __device__ __forceinline__ float div(float a, float b) { if (EXACT) { return __fdiv_rn(a, b); } else { return a / b; } }
	// 1.0f / x: the correctly rounded reciprocal is the correctly rounded quotient, with a shorter instruction sequence
	static __device__ __forceinline__ float rcp(float x) { if (EXACT) { return __frcp_rn(x); } else { return 1.0f / x; } }
	static __device__ __forceinline__ double dmul(double a, double b) { if (EXACT) { return __dmul_rn(a, b); } else { return a * b; } }
	static __device__ __forceinline__ double dadd(double a, double b) { if (EXACT) { return __dadd_rn(a, b); } else { return a + b; } }
	static __device__ __forceinline__ double dsub(double a, double b) { if (EXACT) { return __dsub_rn(a, b); } else { return a - b; } }
	// (a0*b0 + a1*b1) + a2*b2, vectormath.h:367
	static __device__ __forceinline__ float dot(const float* a, const float* b) {
		return add(add(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2]));
	}
	static __device__ __forceinline__ float dot(float a0, float a1, float a2, float b0, float b1, float b2) {
		return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2));
	}
};

// Registers holding one element's constants.
struct ElemRec {
	uint4 idx;
	float Qi[3][3]; // [col][row]
	float volume;
	float QQ[3], QR[3];
	float alpha0, alpha1; // comp / dt^2 of the call's settings (DeviceScene::eAlpha), filled by the barrier-free kernels' loads
};

// 256-bit read-only load (element planes never change after upload).
__device__ __forceinline__ void LoadConst32B(const void* p, uint32_t (&r)[8]) {
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
	             : "l"(p));
}

template <bool NEED_PREFACTORED, bool EXACT>
__device__ __forceinline__ void LoadElementFrom(const ElemRecA* planeA, const DeviceScene& sc, uint32_t e, ElemRec& r) {
	uint32_t a[8], b[8];
	LoadConst32B(planeA + e, a);
	LoadConst32B(sc.eB + e, b);
	r.idx = make_uint4(a[0], a[1], a[2], a[3]);
	r.Qi[0][0] = __uint_as_float(a[4]); r.Qi[0][1] = __uint_as_float(a[5]); r.Qi[0][2] = __uint_as_float(a[6]);
	r.Qi[1][0] = __uint_as_float(a[7]); r.Qi[1][1] = __uint_as_float(b[0]); r.Qi[1][2] = __uint_as_float(b[1]);
	r.Qi[2][0] = __uint_as_float(b[2]); r.Qi[2][1] = __uint_as_float(b[3]); r.Qi[2][2] = __uint_as_float(b[4]);
	r.volume = __uint_as_float(b[5]);
	if (NEED_PREFACTORED && EXACT) {
		float4 q = __ldg(sc.eC + e);
		r.QQ[0] = __uint_as_float(b[6]); r.QQ[1] = __uint_as_float(b[7]); r.QQ[2] = q.x;
		r.QR[0] = q.y; r.QR[1] = q.z; r.QR[2] = q.w;
	}
	// FAST recomputes QQ/QR from Qi at the point of use (PrefactoredI1): doing it here would make the prefetching
	// thread wait for the record before it reaches the barrier and serialise the HBM latency into every phase.
}
template <bool NEED_PREFACTORED, bool EXACT>
__device__ __forceinline__ void LoadElement(const DeviceScene& sc, uint32_t e, ElemRec& r) {
	LoadElementFrom<NEED_PREFACTORED, EXACT>(sc.eA, sc, e, r);
}
__device__ __forceinline__ uint4 LoadElementIdx(const DeviceScene& sc, uint32_t e) { return __ldg(reinterpret_cast<const uint4*>(sc.eA + e)); }

// Vertex gather: L2-only loads (ld.global.cg) because other SMs rewrite positions between colours.
struct VertexRegs {
	double x[3];
	float w;
	uint32_t flags;
};
// 32-byte records move as ONE 256-bit L2 request (LDG.E.256 / STG.E.256, sm_100+) instead of two 128-bit ones:
// the sweep is bound by L2 request/sector throughput in the bursts that follow each barrier.
// The accesses are STRONG (relaxed, gpu scope): the barrier-free schedules race on these records by design (the tag inside
// the record is the synchronisation), and a weak racing access would be undefined in the PTX model.  What the model does NOT
// promise is single-copy atomicity beyond 8 bytes: that one 32-byte-aligned 256-bit access is one L2 sector transaction is a
// property of sm_100 hardware.  tools/check_sass.py (run by the Makefile) fails the build if these do not come out as
// LDG/STG.E.ENL2.256.STRONG.GPU, and tests/test_gpu_hardening.py stresses it (xf_debug_torn_records).
__device__ __forceinline__ void Load32B(const void* p, double& a, double& b, double& c, double& d) {
	asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void Store32B(void* p, double a, double b, double c, double d) {
	asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ VertexRegs LoadVertex(const VertexRec* Xw, uint32_t i) {
	VertexRegs v;
	double packed;
	Load32B(Xw + i, v.x[0], v.x[1], v.x[2], packed);
	long long bits = __double_as_longlong(packed);
	v.w = __int_as_float((int)(bits & 0xffffffffll));
	v.flags = (uint32_t)((unsigned long long)bits >> 32);
	return v;
}
__device__ __forceinline__ void StoreVertex(VertexRec* Xw, uint32_t i, const VertexRegs& v) {
	long long bits = (long long)(((unsigned long long)v.flags << 32) | (unsigned long long)(uint32_t)__float_as_int(v.w));
	Store32B(Xw + i, v.x[0], v.x[1], v.x[2], __longlong_as_double(bits));
}
__device__ __forceinline__ void LoadD3(const double4* A, uint32_t i, double* out) {
	double pad;
	Load32B(A + i, out[0], out[1], out[2], pad);
}
__device__ __forceinline__ void StoreD3(double4* A, uint32_t i, const double* v) { Store32B(A + i, v[0], v[1], v[2], 0.0); }

// Give-up test of a tag wait, called by every lane of the waiting group with the same `spins` (the groups leave their spin
// loops as one, see xf_dataflow.cu).  Every 1024 polls: past the limit, report the stall (device word for the other waiters,
// pinned host word for the host); in any case stop waiting once somebody has reported one.
__device__ __forceinline__ bool SpinGiveUp(unsigned int* errDev, unsigned int* errHost, unsigned mask, uint32_t spins, uint32_t limit) {
	if ((spins & 1023u) != 1023u) { return false; }
	if (spins > limit) {
		atomicExch(errDev, 1u);
		if (errHost) {
			*reinterpret_cast<volatile unsigned int*>(errHost) = 1u;
			__threadfence_system();
		}
	}
	unsigned int e;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(errDev) : "memory");
	return __any_sync(mask, e != 0u) != 0;
}

// Store policy of the kernels that route the updated records themselves.
struct NoStore {
	__device__ __forceinline__ void StoreX(uint32_t, const VertexRegs&) const {}
	__device__ __forceinline__ void LoadO(uint32_t, double*) const {}
	__device__ __forceinline__ void LoadV(uint32_t, double*) const {}
	__device__ __forceinline__ void StoreV(uint32_t, const double*) const {}
};

// Vertex-store policies: where an element's vertices live.  GlobalStore = the HBM/L2 arrays of a DeviceScene
// (one mesh per device); SmemStore = one small scene resident in shared memory (batched scenes, xf_batch.cu).
struct GlobalStore {
	VertexRec* Xw;
	double4* O;
	double4* V;
	__device__ __forceinline__ VertexRegs LoadX(uint32_t i) const { return LoadVertex(Xw, i); }
	__device__ __forceinline__ void StoreX(uint32_t i, const VertexRegs& v) const { StoreVertex(Xw, i, v); }
	__device__ __forceinline__ void LoadO(uint32_t i, double* o) const { LoadD3(O, i, o); }
	__device__ __forceinline__ void LoadV(uint32_t i, double* o) const { LoadD3(V, i, o); }
	__device__ __forceinline__ void StoreV(uint32_t i, const double* v) const { StoreD3(V, i, v); }
};
__device__ __forceinline__ GlobalStore StoreOf(const DeviceScene& sc) { return GlobalStore{ sc.Xw, sc.O, sc.V }; }

struct SmemStore {
	double* X; // 3 per vertex
	double* O;
	double* V;
	float* w;
	__device__ __forceinline__ VertexRegs LoadX(uint32_t i) const {
		VertexRegs v;
		v.x[0] = X[3 * i]; v.x[1] = X[3 * i + 1]; v.x[2] = X[3 * i + 2];
		v.w = w[i];
		v.flags = 0;
		return v;
	}
	__device__ __forceinline__ void StoreX(uint32_t i, const VertexRegs& v) const { X[3 * i] = v.x[0]; X[3 * i + 1] = v.x[1]; X[3 * i + 2] = v.x[2]; }
	__device__ __forceinline__ void LoadO(uint32_t i, double* o) const { o[0] = O[3 * i]; o[1] = O[3 * i + 1]; o[2] = O[3 * i + 2]; }
	__device__ __forceinline__ void LoadV(uint32_t i, double* o) const { o[0] = V[3 * i]; o[1] = V[3 * i + 1]; o[2] = V[3 * i + 2]; }
	__device__ __forceinline__ void StoreV(uint32_t i, const double* v) const { V[3 * i] = v[0]; V[3 * i + 1] = v[1]; V[3 * i + 2] = v[2]; }
};

// P[n] = Vec(X[n] - X[3]): fp64 difference, then narrowed.  Fem.cpp:453
template <bool EXACT>
__device__ __forceinline__ void Edges(const VertexRegs (&v)[4], float (&P)[3][3]) {
#pragma unroll
	for (int n = 0; n < 3; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { P[n][k] = __double2float_rn(Op<EXACT>::dsub(v[n].x[k], v[3].x[k])); }
	}
}

// F[c][r] = Qi[0][r]*P[0][c] + Qi[1][r]*P[1][c] + Qi[2][r]*P[2][c].  Fem.cpp:311-323, vectormath.h:508-514
template <bool EXACT>
__device__ __forceinline__ void DeformationGradient(const ElemRec& e, const float (&P)[3][3], float (&F)[3][3]) {
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int r = 0; r < 3; r++) { F[c][r] = Op<EXACT>::dot(e.Qi[0][r], e.Qi[1][r], e.Qi[2][r], P[0][c], P[1][c], P[2][c]); }
	}
}

// g[n] = column n of (G * Qi), n < 3.  Fem.cpp:338-354
template <bool EXACT>
__device__ __forceinline__ void ApplyGradient(const ElemRec& e, const float (&G)[3][3], float (&g)[4][3]) {
#pragma unroll
	for (int n = 0; n < 3; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { g[n][k] = Op<EXACT>::dot(G[0][k], G[1][k], G[2][k], e.Qi[n][0], e.Qi[n][1], e.Qi[n][2]); }
	}
}

// g[3] = ((0 - g[0]) - g[1]) - g[2]
template <bool EXACT>
__device__ __forceinline__ void CloseGradient(float (&g)[4][3]) {
#pragma unroll
	for (int k = 0; k < 3; k++) { g[3][k] = Op<EXACT>::sub(Op<EXACT>::sub(Op<EXACT>::sub(0.0f, g[0][k]), g[1][k]), g[2][k]); }
}

// adj(F) and J = F[0] . (adj[0][0], adj[1][0], adj[2][0]).  Fem.cpp:284-306
template <bool EXACT>
__device__ __forceinline__ float AdjugateAndDet(const float (&m)[3][3], float (&adj)[3][3]) {
	typedef Op<EXACT> O;
	adj[0][0] = O::sub(O::mul(m[1][1], m[2][2]), O::mul(m[1][2], m[2][1]));
	adj[0][1] = -O::sub(O::mul(m[0][1], m[2][2]), O::mul(m[0][2], m[2][1]));
	adj[0][2] = O::sub(O::mul(m[0][1], m[1][2]), O::mul(m[0][2], m[1][1]));
	adj[1][0] = -O::sub(O::mul(m[1][0], m[2][2]), O::mul(m[1][2], m[2][0]));
	adj[1][1] = O::sub(O::mul(m[0][0], m[2][2]), O::mul(m[0][2], m[2][0]));
	adj[1][2] = -O::sub(O::mul(m[0][0], m[1][2]), O::mul(m[0][2], m[1][0]));
	adj[2][0] = O::sub(O::mul(m[1][0], m[2][1]), O::mul(m[1][1], m[2][0]));
	adj[2][1] = -O::sub(O::mul(m[0][0], m[2][1]), O::mul(m[0][1], m[2][0]));
	adj[2][2] = O::sub(O::mul(m[0][0], m[1][1]), O::mul(m[0][1], m[1][0]));
	return O::dot(m[0][0], m[0][1], m[0][2], adj[0][0], adj[1][0], adj[2][0]);
}

// Volumetric constraint U1 = (J - target)^2 and its gradient (weight == 1 for T4).  Fem.cpp:479-485, 537-547
template <bool EXACT>
__device__ __forceinline__ float VolumetricFromF(const ElemRec& e, const float (&F)[3][3], float target, float (&g1)[4][3]) {
	typedef Op<EXACT> O;
	float adjF[3][3];
	float J = AdjugateAndDet<EXACT>(F, adjF);
	float d = O::sub(J, target);
	float U1 = O::mul(d, d);
	float s = O::mul(2.0f, d);
	float GJ[3][3];
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int r = 0; r < 3; r++) { GJ[c][r] = O::mul(adjF[c][r], s); }
	}
	ApplyGradient<EXACT>(e, GJ, g1);
	CloseGradient<EXACT>(g1);
	return U1;
}

// Prefactored I1 and gradient.  Fem.cpp:163-192
template <bool EXACT>
__device__ __forceinline__ float PrefactoredI1(const ElemRec& eIn, const float (&P)[3][3], float (&g)[4][3]) {
	typedef Op<EXACT> O;
	ElemRec e = eIn;
	if (!EXACT) {
		// QQ_i = |col_i(Qi)|^2, QR = 2 col_i . col_j  (what Fem.cpp:131-161 integrates, up to rounding)
		e.QQ[0] = Op<false>::dot(e.Qi[0], e.Qi[0]);
		e.QQ[1] = Op<false>::dot(e.Qi[1], e.Qi[1]);
		e.QQ[2] = Op<false>::dot(e.Qi[2], e.Qi[2]);
		e.QR[0] = 2.0f * Op<false>::dot(e.Qi[0], e.Qi[1]);
		e.QR[1] = 2.0f * Op<false>::dot(e.Qi[0], e.Qi[2]);
		e.QR[2] = 2.0f * Op<false>::dot(e.Qi[1], e.Qi[2]);
	}
	float U = 0.0f;
#pragma unroll
	for (int i = 0; i < 3; i++) {
		U = O::add(U, O::mul(e.QQ[i], O::dot(P[i], P[i])));
		float s = O::mul(2.0f, e.QQ[i]);
#pragma unroll
		for (int k = 0; k < 3; k++) { g[i][k] = O::mul(s, P[i][k]); }
	}
	// pairs (0,1) (0,2) (1,2)
	U = O::add(U, O::mul(e.QR[0], O::dot(P[0], P[1])));
#pragma unroll
	for (int k = 0; k < 3; k++) { g[0][k] = O::add(g[0][k], O::mul(e.QR[0], P[1][k])); }
#pragma unroll
	for (int k = 0; k < 3; k++) { g[1][k] = O::add(g[1][k], O::mul(e.QR[0], P[0][k])); }
	U = O::add(U, O::mul(e.QR[1], O::dot(P[0], P[2])));
#pragma unroll
	for (int k = 0; k < 3; k++) { g[0][k] = O::add(g[0][k], O::mul(e.QR[1], P[2][k])); }
#pragma unroll
	for (int k = 0; k < 3; k++) { g[2][k] = O::add(g[2][k], O::mul(e.QR[1], P[0][k])); }
	U = O::add(U, O::mul(e.QR[2], O::dot(P[1], P[2])));
#pragma unroll
	for (int k = 0; k < 3; k++) { g[1][k] = O::add(g[1][k], O::mul(e.QR[2], P[2][k])); }
#pragma unroll
	for (int k = 0; k < 3; k++) { g[2][k] = O::add(g[2][k], O::mul(e.QR[2], P[1][k])); }
	CloseGradient<EXACT>(g);
	return U;
}

// Yeoh polynomial pieces, C = {0.1095, 14.95, 4.595}.  Fem.cpp:470-474, 525-531
template <bool EXACT>
__device__ __forceinline__ float YeohEnergy(float IM) {
	typedef Op<EXACT> O;
	const float C0 = 0.1095f, C1 = 14.95f, C2 = 4.595f;
	return O::add(O::add(O::mul(C0, IM), O::mul(O::mul(C1, IM), IM)), O::mul(O::mul(O::mul(C2, IM), IM), IM));
}
template <bool EXACT>
__device__ __forceinline__ float YeohSlope(float IM) {
	typedef Op<EXACT> O;
	const float C0 = 0.1095f, C1 = 14.95f, C2 = 4.595f;
	return O::add(O::add(C0, O::mul(O::mul(2.0f, C1), IM)), O::mul(O::mul(O::mul(3.0f, C2), IM), IM));
}

// (X[n] - O[n]) narrowed, for the in-constraint damping terms.  Xpbd.h:96, 139
template <bool EXACT, typename VS>
__device__ __forceinline__ void Displacements(const VS& vs, const uint4& idx, const VertexRegs (&v)[4], float (&d)[4][3]) {
	const uint32_t is[4] = { idx.x, idx.y, idx.z, idx.w };
#pragma unroll
	for (int n = 0; n < 4; n++) {
		double o[3];
		vs.LoadO(is[n], o);
#pragma unroll
		for (int k = 0; k < 3; k++) { d[n][k] = __double2float_rn(Op<EXACT>::dsub(v[n].x[k], o[k])); }
	}
}

// EnergyXpbdConstrain, Xpbd.h:86-120.  Updates the register copies of the positions.
template <bool EXACT, bool DAMPED, typename VS, typename PARAMS>
__device__ __forceinline__ void ConstrainOne(const VS& vs, const PARAMS& p, const uint4& idx, VertexRegs (&v)[4], float U,
                                             const float (&g)[4][3], float compliance, float alpha, float dampingGamma) {
	typedef Op<EXACT> O;
	float wgg = 1.0e-22f;
#pragma unroll
	for (int n = 0; n < 4; n++) { wgg = O::add(wgg, O::mul(v[n].w, O::dot(g[n], g[n]))); }
	float lambdaPrime;
	float twoU = O::mul(2.0f, U);
	if (DAMPED) {
		float d[4][3];
		Displacements<EXACT>(vs, idx, v, d);
		float gV = 0.0f;
#pragma unroll
		for (int n = 0; n < 4; n++) { gV = O::add(gV, O::dot(d[n], g[n])); }
		float gamma = O::div(dampingGamma, p.dt);
		if (p.rayleigh == XF_RAYLEIGH_PAPER) {
			float A = O::add(O::mul(wgg, O::add(1.0f, gamma)), O::mul(twoU, alpha));
			lambdaPrime = O::div(O::sub(O::mul(-2.0f, U), O::mul(gamma, gV)), A);
		} else {
			lambdaPrime = O::div(O::mul(-2.0f, U), O::add(wgg, O::mul(twoU, alpha)));
			float invBeta = O::div(compliance, O::mul(p.dt, dampingGamma));
			float A = O::add(wgg, O::mul(twoU, invBeta));
			float b = O::sub(-gV, O::mul(wgg, lambdaPrime));
			b = O::mul(b, O::div(A, fmaxf(A, O::mul(4.0f, wgg))));
			lambdaPrime = O::add(lambdaPrime, O::div(b, A));
		}
	} else {
		lambdaPrime = O::div(O::mul(-2.0f, U), O::add(wgg, O::mul(twoU, alpha)));
	}
#pragma unroll
	for (int n = 0; n < 4; n++) {
		float s = O::mul(v[n].w, lambdaPrime);
#pragma unroll
		for (int k = 0; k < 3; k++) { v[n].x[k] = O::dadd(v[n].x[k], (double)O::mul(s, g[n][k])); }
	}
}

// 2x2 Cramer, Xpbd.h:173-179
template <bool EXACT>
__device__ __forceinline__ void Cramer2(float A0, float A1, float A2, float b0, float b1, float& l0, float& l1) {
	typedef Op<EXACT> O;
	float invA00 = O::rcp(A0);
	float invA11 = O::rcp(A2);
	float p0 = O::mul(A1, invA00);
	float p1 = O::mul(A1, invA11);
	float invDet = O::rcp(fmaxf(0.00000001f, O::sub(1.0f, O::mul(p0, p1))));
	float q0 = O::mul(b0, invA00);
	float q1 = O::mul(b1, invA11);
	l0 = O::mul(invDet, O::sub(q0, O::mul(p0, q1)));
	l1 = O::mul(invDet, O::sub(q1, O::mul(q0, p1)));
}

// EnergyXpbdConstrainSimultaneous<.., 2>, Xpbd.h:122-214
template <bool EXACT, bool DAMPED, typename VS, typename PARAMS>
__device__ __forceinline__ void ConstrainBoth(const VS& vs, const PARAMS& p, const uint4& idx, VertexRegs (&v)[4], float U0,
                                              float U1, const float (&g0)[4][3], const float (&g1)[4][3], float comp0, float comp1,
                                              float alpha0, float alpha1, float dampingGamma) {
	typedef Op<EXACT> O;
	float w00 = 1.0e-22f, w10 = 1.0e-22f, w11 = 1.0e-22f;
#pragma unroll
	for (int n = 0; n < 4; n++) { w00 = O::add(w00, O::mul(v[n].w, O::dot(g0[n], g0[n]))); }
#pragma unroll
	for (int n = 0; n < 4; n++) { w10 = O::add(w10, O::mul(v[n].w, O::dot(g1[n], g0[n]))); }
#pragma unroll
	for (int n = 0; n < 4; n++) { w11 = O::add(w11, O::mul(v[n].w, O::dot(g1[n], g1[n]))); }
	float A0, A2, b0, b1;
	float gV0 = 0.0f, gV1 = 0.0f;
	float gamma = 0.0f;
	if (DAMPED) {
		float d[4][3];
		Displacements<EXACT>(vs, idx, v, d);
#pragma unroll
		for (int n = 0; n < 4; n++) {
			gV0 = O::add(gV0, O::dot(g0[n], d[n]));
			gV1 = O::add(gV1, O::dot(g1[n], d[n]));
		}
		gamma = O::div(dampingGamma, p.dt);
	}
	if (DAMPED && p.rayleigh == XF_RAYLEIGH_PAPER) {
		float onePlus = O::add(1.0f, gamma);
		A0 = O::add(w00, O::div(O::mul(O::mul(2.0f, U0), alpha0), onePlus));
		b0 = O::div(O::sub(O::mul(-2.0f, U0), O::mul(gamma, gV0)), onePlus);
		A2 = O::add(w11, O::div(O::mul(O::mul(2.0f, U1), alpha1), onePlus));
		b1 = O::div(O::sub(O::mul(-2.0f, U1), O::mul(gamma, gV1)), onePlus);
	} else {
		A0 = O::add(w00, O::mul(O::mul(2.0f, U0), alpha0));
		b0 = O::mul(-2.0f, U0);
		A2 = O::add(w11, O::mul(O::mul(2.0f, U1), alpha1));
		b1 = O::mul(-2.0f, U1);
	}
	float l0, l1;
	Cramer2<EXACT>(A0, w10, A2, b0, b1, l0, l1);
	if (DAMPED && p.rayleigh == XF_RAYLEIGH_LIMIT) {
		float invBeta0 = O::div(comp0, O::mul(p.dt, dampingGamma));
		float invBeta1 = O::div(comp1, O::mul(p.dt, dampingGamma));
		float B0 = O::add(w00, O::mul(O::mul(2.0f, U0), invBeta0));
		float B2 = O::add(w11, O::mul(O::mul(2.0f, U1), invBeta1));
		float c0 = O::mul(b0, O::sub(-1.0f, O::mul(alpha0, l0)));
		float c1 = O::mul(b1, O::sub(-1.0f, O::mul(alpha1, l1)));
		c0 = O::add(c0, -gV0);
		c1 = O::add(c1, -gV1);
		c0 = O::mul(c0, O::div(B0, fmaxf(B0, O::mul(8.0f, w00))));
		c1 = O::mul(c1, O::div(B2, fmaxf(B2, O::mul(8.0f, w11))));
		float e0, e1;
		Cramer2<EXACT>(B0, w10, B2, c0, c1, e0, e1);
		l0 = O::add(l0, e0);
		l1 = O::add(l1, e1);
	}
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) {
			// Vec(0) + l0*g0 + l1*g1: the leading "+ 0" only normalises -0, which X += (double)... cannot observe
			float acc = O::add(O::mul(l0, g0[n][k]), O::mul(l1, g1[n][k]));
			v[n].x[k] = O::dadd(v[n].x[k], (double)O::mul(v[n].w, acc));
		}
	}
}

// Energies + gradients of both constraints from the current edge vectors (the part SolveElementMixed
// shares between the constraint solve and DampingMode::On).
template <int ENERGY, bool EXACT>
__device__ __forceinline__ void DeviatoricTerm(const ElemRec& e, const float (&P)[3][3], float& U0, float (&g0)[4][3], float (&F)[3][3],
                                               bool& haveF) {
	typedef Op<EXACT> O;
	if (ENERGY == XF_ENERGY_MIXED || ENERGY == XF_ENERGY_YEOH_SKIN) {
		DeformationGradient<EXACT>(e, P, F);
		haveF = true;
		float I1 = O::add(O::add(O::dot(F[0], F[0]), O::dot(F[1], F[1])), O::dot(F[2], F[2]));
		float scale;
		if (ENERGY == XF_ENERGY_MIXED) {
			U0 = I1;
			scale = 2.0f;
		} else {
			float IM = O::sub(I1, 3.0f);
			U0 = YeohEnergy<EXACT>(IM);
			scale = O::mul(YeohSlope<EXACT>(IM), 2.0f);
		}
		float G[3][3];
#pragma unroll
		for (int c = 0; c < 3; c++) {
#pragma unroll
			for (int r = 0; r < 3; r++) { G[c][r] = O::mul(F[r][c], scale); }
		}
		ApplyGradient<EXACT>(e, G, g0);
		U0 = fmaxf(0.0001f, U0);
		CloseGradient<EXACT>(g0);
	} else {
		haveF = false;
		U0 = PrefactoredI1<EXACT>(e, P, g0);
		if (ENERGY == XF_ENERGY_YEOH_SKIN_FAST) {
			float IM = O::sub(U0, 3.0f);
			U0 = fmaxf(0.0001f, YeohEnergy<EXACT>(IM));
			float gScale = YeohSlope<EXACT>(IM);
#pragma unroll
			for (int n = 0; n < 4; n++) {
#pragma unroll
				for (int k = 0; k < 3; k++) { g0[n][k] = O::mul(g0[n][k], gScale); }
			}
		}
	}
}

// comp = {1/mu/vol, 1/lambda/vol} (Fem.cpp:449) and, for the undamped solves, alpha = comp / dt^2 (Xpbd.h:88, 154): two
// chained IEEE divisions that depend on the element record only.  The barrier-free kernels evaluate them BEFORE they wait
// for the vertex records, so they are off the vertex-to-vertex dependence chain (same operations, same bits).
struct ElemCompliance {
	float comp0, comp1, alpha0, alpha1;
};
template <bool EXACT, typename PARAMS>
__device__ __forceinline__ ElemCompliance ComplianceOf(const PARAMS& p, float volume) {
	typedef Op<EXACT> O;
	ElemCompliance c;
	c.comp0 = O::div(p.invMu, volume);
	c.comp1 = XF_DIV_MAYBE_ZERO(O, p.invLambda, volume);
	c.alpha0 = O::div(c.comp0, p.dt2);
	c.alpha1 = XF_DIV_MAYBE_ZERO(O, c.comp1, p.dt2);
	return c;
}

}  // namespace xf
#include "xf_element_packed.cuh" // the same arithmetic two-wide (FMUL2 / FADD2) for the headline configurations
namespace xf {

// One element of the main sweep with its four vertex records already in registers (`v` is updated and stored).
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED, typename VS, typename PARAMS>
__device__ __forceinline__ void SolveElementGathered(const VS& vs, const PARAMS& p, const ElemRec& e, VertexRegs (&v)[4], const ElemCompliance& ec) {
	typedef Op<EXACT> O;
	const uint32_t is[4] = { e.idx.x, e.idx.y, e.idx.z, e.idx.w };
	if (UsePacked<ENERGY, SIMUL, EXACT, DAMPED>::value) {
		SolvePrefactoredSimulPacked<ENERGY>(p, e, v, ec);
#pragma unroll
		for (int n = 0; n < 4; n++) { vs.StoreX(is[n], v[n]); }
		return;
	}
	const float comp0 = ec.comp0, comp1 = ec.comp1;
	float P[3][3], F[3][3], g0[4][3], g1[4][3];
	float U0, U1;
	bool haveF;
	Edges<EXACT>(v, P);
	DeviatoricTerm<ENERGY, EXACT>(e, P, U0, g0, F, haveF);
	if (SIMUL) {
		if (!(ENERGY == XF_ENERGY_MIXED || ENERGY == XF_ENERGY_YEOH_SKIN)) { DeformationGradient<EXACT>(e, P, F); }
		U1 = VolumetricFromF<EXACT>(e, F, p.a, g1);
		ConstrainBoth<EXACT, DAMPED>(vs, p, e.idx, v, U0, U1, g0, g1, comp0, comp1, ec.alpha0, ec.alpha1, p.damping);
	} else {
		ConstrainOne<EXACT, DAMPED>(vs, p, e.idx, v, U0, g0, comp0, ec.alpha0, p.damping);
		Edges<EXACT>(v, P);
		DeformationGradient<EXACT>(e, P, F);
		U1 = VolumetricFromF<EXACT>(e, F, p.a, g1);
		ConstrainOne<EXACT, DAMPED>(vs, p, e.idx, v, U1, g1, comp1, ec.alpha1, p.damping);
	}
#pragma unroll
	for (int n = 0; n < 4; n++) { vs.StoreX(is[n], v[n]); }
}
template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED, typename VS, typename PARAMS>
__device__ __forceinline__ void SolveElementGathered(const VS& vs, const PARAMS& p, const ElemRec& e, VertexRegs (&v)[4]) {
	SolveElementGathered<ENERGY, SIMUL, EXACT, DAMPED>(vs, p, e, v, ComplianceOf<EXACT>(p, e.volume));
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED, typename VS, typename PARAMS>
__device__ __forceinline__ void SolveElement(const VS& vs, const PARAMS& p, const ElemRec& e) {
	const uint32_t is[4] = { e.idx.x, e.idx.y, e.idx.z, e.idx.w };
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n] = vs.LoadX(is[n]); }
	SolveElementGathered<ENERGY, SIMUL, EXACT, DAMPED>(vs, p, e, v);
}

// SolveVolumeOnly, Fem.cpp:840-867 (J -> 1 with compliance * volume, never damped).
template <bool EXACT, typename VS, typename PARAMS>
__device__ __forceinline__ void SolveVolumeOnly(const VS& vs, const PARAMS& p, const ElemRec& e) {
	typedef Op<EXACT> O;
	const uint32_t is[4] = { e.idx.x, e.idx.y, e.idx.z, e.idx.w };
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n] = vs.LoadX(is[n]); }
	float comp = O::mul(p.compliance, e.volume);
	float P[3][3], F[3][3], g[4][3];
	Edges<EXACT>(v, P);
	DeformationGradient<EXACT>(e, P, F);
	// U = weight*(J-1)*(J-1) with weight == 1: (1*(J-1))*(J-1) == (J-1)^2
	float U = VolumetricFromF<EXACT>(e, F, 1.0f, g);
	ConstrainOne<EXACT, false>(vs, p, e.idx, v, U, g, comp, XF_DIV_MAYBE_ZERO(O, comp, p.dt2), 0.0f);
#pragma unroll
	for (int n = 0; n < 4; n++) { vs.StoreX(is[n], v[n]); }
}

// DampElement: SolveElementMixed in DampingMode::On, Fem.cpp:910-929, 555-563; RayleighDamp Xpbd.h:216-263.
// The arithmetic with the four position records and velocities already in registers (`vel` is updated in place); the
// barrier-free kernels gather and scatter the versioned records themselves.
template <int ENERGY, bool SIMUL, bool EXACT, typename PARAMS>
__device__ __forceinline__ void DampElementGathered(const PARAMS& p, const ElemRec& e, const VertexRegs (&v)[4], double (&vel)[4][3]) {
	typedef Op<EXACT> O;
	float fv[4][3];
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { fv[n][k] = __double2float_rn(vel[n][k]); }
	}
	float comp0 = O::div(p.invMu, e.volume);
	float comp1 = XF_DIV_MAYBE_ZERO(O, p.invLambda, e.volume);
	float dg = p.dampDamping;
	float P[3][3], F[3][3], g0[4][3], g1[4][3];
	float U0, U1;
	bool haveF;
	Edges<EXACT>(v, P);
	DeviatoricTerm<ENERGY, EXACT>(e, P, U0, g0, F, haveF);
	if (!(ENERGY == XF_ENERGY_MIXED || ENERGY == XF_ENERGY_YEOH_SKIN)) { DeformationGradient<EXACT>(e, P, F); }
	U1 = VolumetricFromF<EXACT>(e, F, p.a, g1);
	float invBeta0 = O::div(comp0, O::mul(p.dt, dg));
	float invBeta1 = O::div(comp1, O::mul(p.dt, dg));
	if (SIMUL) {
		float A0 = 1.0e-22f, A1 = 1.0e-22f, A2 = 1.0e-22f, b0 = 0.0f, b1 = 0.0f;
#pragma unroll
		for (int n = 0; n < 4; n++) {
			A0 = O::add(A0, O::mul(v[n].w, O::dot(g0[n], g0[n])));
			A1 = O::add(A1, O::mul(v[n].w, O::dot(g0[n], g1[n])));
			A2 = O::add(A2, O::mul(v[n].w, O::dot(g1[n], g1[n])));
			b0 = O::sub(b0, O::dot(g0[n], fv[n]));
			b1 = O::sub(b1, O::dot(g1[n], fv[n]));
		}
		A0 = O::add(A0, O::mul(O::mul(2.0f, U0), invBeta0));
		A2 = O::add(A2, O::mul(O::mul(2.0f, U1), invBeta1));
		float l0, l1;
		Cramer2<EXACT>(A0, A1, A2, b0, b1, l0, l1);
#pragma unroll
		for (int n = 0; n < 4; n++) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float acc = O::add(O::mul(l0, g0[n][k]), O::mul(l1, g1[n][k]));
				vel[n][k] = O::dadd(vel[n][k], (double)O::mul(v[n].w, acc));
			}
		}
	} else {
#pragma unroll
		for (int c = 0; c < 2; c++) {
			const float(&g)[4][3] = c == 0 ? g0 : g1;
			float U = c == 0 ? U0 : U1;
			float invBeta = c == 0 ? invBeta0 : invBeta1;
			float wgg = 1.0e-22f, gV = 0.0f;
#pragma unroll
			for (int n = 0; n < 4; n++) {
				wgg = O::add(wgg, O::mul(v[n].w, O::dot(g[n], g[n])));
				gV = O::add(gV, O::dot(g[n], fv[n]));
			}
			float lambda = O::div(-gV, O::add(O::mul(O::mul(2.0f, U), invBeta), wgg));
#pragma unroll
			for (int n = 0; n < 4; n++) {
				float s = O::mul(v[n].w, lambda);
#pragma unroll
				for (int k = 0; k < 3; k++) {
					vel[n][k] = O::dadd(vel[n][k], (double)O::mul(s, g[n][k]));
					fv[n][k] = __double2float_rn(vel[n][k]);
				}
			}
		}
	}
}
template <int ENERGY, bool SIMUL, bool EXACT, typename VS, typename PARAMS>
__device__ __forceinline__ void DampElement(const VS& vs, const PARAMS& p, const ElemRec& e) {
	const uint32_t is[4] = { e.idx.x, e.idx.y, e.idx.z, e.idx.w };
	VertexRegs v[4];
	double vel[4][3];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		v[n] = vs.LoadX(is[n]);
		vs.LoadV(is[n], vel[n]);
	}
	DampElementGathered<ENERGY, SIMUL, EXACT>(p, e, v, vel);
#pragma unroll
	for (int n = 0; n < 4; n++) { vs.StoreV(is[n], vel[n]); }
}

// inverse(mat3) with fp64 cofactors, vectormath.cpp:41-60 (PbdDamp's inertia tensor).
template <bool EXACT>
__device__ __forceinline__ void InverseViaDouble(const float (&m)[3][3], float (&out)[3][3]) {
	typedef Op<EXACT> O;
	double d[3][3];
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int r = 0; r < 3; r++) { d[c][r] = (double)m[c][r]; }
	}
	float adj[3][3];
	adj[0][0] = __double2float_rn(O::dsub(O::dmul(d[1][1], d[2][2]), O::dmul(d[1][2], d[2][1])));
	adj[0][1] = __double2float_rn(-O::dsub(O::dmul(d[0][1], d[2][2]), O::dmul(d[0][2], d[2][1])));
	adj[0][2] = __double2float_rn(O::dsub(O::dmul(d[0][1], d[1][2]), O::dmul(d[0][2], d[1][1])));
	adj[1][0] = __double2float_rn(-O::dsub(O::dmul(d[1][0], d[2][2]), O::dmul(d[1][2], d[2][0])));
	adj[1][1] = __double2float_rn(O::dsub(O::dmul(d[0][0], d[2][2]), O::dmul(d[0][2], d[2][0])));
	adj[1][2] = __double2float_rn(-O::dsub(O::dmul(d[0][0], d[1][2]), O::dmul(d[0][2], d[1][0])));
	adj[2][0] = __double2float_rn(O::dsub(O::dmul(d[1][0], d[2][1]), O::dmul(d[1][1], d[2][0])));
	adj[2][1] = __double2float_rn(-O::dsub(O::dmul(d[0][0], d[2][1]), O::dmul(d[0][1], d[2][0])));
	adj[2][2] = __double2float_rn(O::dsub(O::dmul(d[0][0], d[1][1]), O::dmul(d[0][1], d[1][0])));
	float det = O::dot(m[0][0], m[0][1], m[0][2], adj[0][0], adj[1][0], adj[2][0]);
	float s = O::rcp(det);
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int r = 0; r < 3; r++) { out[c][r] = O::mul(adj[c][r], s); }
	}
}

// PbdDamp<4>, Xpbd.h:309-348
template <bool EXACT, typename PARAMS>
__device__ __forceinline__ void PbdDampGathered(const PARAMS& p, float surfaceArea, const VertexRegs (&v)[4], double (&vel)[4][3]) {
	typedef Op<EXACT> O;
	float damping = fminf(1.0f, O::div(p.pbdDamping, surfaceArea));
	float X[4][3], Vf[4][3], M[4];
	float Wsum = 1.0e-8f;
#pragma unroll
	for (int n = 0; n < 4; n++) { Wsum = O::add(Wsum, v[n].w); }
	float Wavg = O::mul(Wsum, 0.25f);
	float Wmin = O::mul(0.0001f, Wavg);
	float Xcm[3] = { 0.0f, 0.0f, 0.0f }, Vcm[3] = { 0.0f, 0.0f, 0.0f };
	float Msum = 0.0f;
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) {
			X[n][k] = __double2float_rn(O::dsub(v[n].x[k], v[3].x[k]));
			Vf[n][k] = __double2float_rn(vel[n][k]);
		}
		M[n] = O::div(Wavg, fmaxf(Wmin, v[n].w));
#pragma unroll
		for (int k = 0; k < 3; k++) {
			Xcm[k] = O::add(Xcm[k], O::mul(X[n][k], M[n]));
			Vcm[k] = O::add(Vcm[k], O::mul(Vf[n][k], M[n]));
		}
		Msum = O::add(Msum, M[n]);
	}
#pragma unroll
	for (int k = 0; k < 3; k++) { Xcm[k] = O::div(Xcm[k], Msum); Vcm[k] = O::div(Vcm[k], Msum); }
	float r[4][3], L[3] = { 0.0f, 0.0f, 0.0f };
	float I[3][3] = { { 0.0f, 0.0f, 0.0f }, { 0.0f, 0.0f, 0.0f }, { 0.0f, 0.0f, 0.0f } };
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { r[n][k] = O::sub(X[n][k], Xcm[k]); }
		float cr[3] = { O::sub(O::mul(r[n][1], Vf[n][2]), O::mul(Vf[n][1], r[n][2])), O::sub(O::mul(r[n][2], Vf[n][0]), O::mul(Vf[n][2], r[n][0])),
			            O::sub(O::mul(r[n][0], Vf[n][1]), O::mul(Vf[n][0], r[n][1])) };
#pragma unroll
		for (int k = 0; k < 3; k++) { L[k] = O::add(L[k], O::mul(M[n], cr[k])); }
		float rr[3] = { O::mul(r[n][0], r[n][0]), O::mul(r[n][1], r[n][1]), O::mul(r[n][2], r[n][2]) };
		float rp[3] = { O::mul(r[n][0], r[n][1]), O::mul(r[n][1], r[n][2]), O::mul(r[n][2], r[n][0]) };
		float T[3][3] = { { O::add(rr[2], rr[1]), -rp[0], -rp[2] }, { -rp[0], O::add(rr[2], rr[0]), -rp[1] }, { -rp[2], -rp[1], O::add(rr[1], rr[0]) } };
#pragma unroll
		for (int c = 0; c < 3; c++) {
#pragma unroll
			for (int q = 0; q < 3; q++) { I[c][q] = O::add(I[c][q], O::mul(T[c][q], M[n])); }
		}
	}
	float Ii[3][3];
	InverseViaDouble<EXACT>(I, Ii);
	float wv[3];
#pragma unroll
	for (int q = 0; q < 3; q++) { wv[q] = O::dot(Ii[0][q], Ii[1][q], Ii[2][q], L[0], L[1], L[2]); }
#pragma unroll
	for (int n = 0; n < 4; n++) {
		float cr[3] = { O::sub(O::mul(wv[1], r[n][2]), O::mul(r[n][1], wv[2])), O::sub(O::mul(wv[2], r[n][0]), O::mul(r[n][2], wv[0])),
			            O::sub(O::mul(wv[0], r[n][1]), O::mul(r[n][0], wv[1])) };
#pragma unroll
		for (int k = 0; k < 3; k++) {
			float dV = O::sub(O::add(Vcm[k], cr[k]), Vf[n][k]);
			vel[n][k] = O::dadd(vel[n][k], (double)O::mul(damping, dV));
		}
	}
}
template <bool EXACT, typename VS, typename PARAMS>
__device__ __forceinline__ void PbdDampElement(const VS& vs, const PARAMS& p, float surfaceArea, const uint4& idx) {
	const uint32_t is[4] = { idx.x, idx.y, idx.z, idx.w };
	VertexRegs v[4];
	double vel[4][3];
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n] = vs.LoadX(is[n]); vs.LoadV(is[n], vel[n]); }
	PbdDampGathered<EXACT>(p, surfaceArea, v, vel);
#pragma unroll
	for (int n = 0; n < 4; n++) { vs.StoreV(is[n], vel[n]); }
}

}  // namespace xf
