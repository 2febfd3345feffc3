// Two-wide fp32 arithmetic for the bit-exact element solve (sm_100 FMUL2 / FADD2: add.rn.f32x2, mul.rn.f32x2).
//
// XF_PRECISION_EXACT may not contract a multiply and an add into an FMA, so its element body is ~460 separate FMUL / FADD
// out of ~600 instructions, and the sweep is bound by instruction issue on the schedulers that carry a colour's third warp
// (DESIGN section 6).  A packed instruction rounds each half exactly like the scalar one (IEEE round-to-nearest per lane, no
// contraction), so every scalar below is computed by the SAME operations in the SAME association order as in xf_element.cuh
// - same bits - in fewer issue slots.  Measured on B200 (tools/probes/fp32_rate.cu, profiles/r2_fp32_rate.log): a mix of
// FMUL2 and FADD2 issues at 0.85 per clock per scheduler = 1.8x the scalar flop rate; packed and scalar latency are equal.
//
// Layout: 3-vectors are {xy pair, z}.  Edge vectors P[i] and the deviatoric gradient g0 are paired over x,y; the deformation
// gradient is paired over its COLUMN index, F_r = {F[0][r], F[1][r]}, F[2][r], which is what P's pairs produce and what the
// adjugate's cofactor pairs consume; the cofactor pairs come out as {adj[c][1], adj[c][0]} (signs folded into the scale), so
// the volumetric gradient g1 is paired {y, x}.  Commutativity (a*b == b*a, a+b == b+a bit for bit) is used freely,
// associativity never.
//
// Covers the prefactored energies (MixedSel, YeohSkinFast) in simultaneous mode without in-constraint damping: the web
// demo's default and the benchmark's headline.  Everything else runs the scalar code.
#pragma once

#include "xf_element.cuh"

namespace xf {

#ifndef XF_NO_PACKED

struct V3p {
	float2 xy;
	float z;
};
// Every packed operation is issued as ONE fma.rn.f32x2 whose third (or second) operand is a constant the assembler cannot see
// (constant memory): a*b + (-0) == a*b, a*1 + b == a + b, b*(-1) + a == a - b, each rounded once - bit for bit the scalar
// result.  Why not mul.rn.f32x2 / add.rn.f32x2 (or the __fmul2_rn / __fadd2_rn intrinsics): ptxas 12.9 contracts a packed
// multiply feeding a packed add into FFMA2 even with explicit .rn and --fmad=false (it never does that to the scalar forms),
// and it does the same after folding visible constants 1 and -0; the parity tests caught the changed bits.
static __constant__ float kPackNegZero = -0.0f;
static __constant__ float kPackOne = 1.0f;
static __constant__ float kPackNegOne = -1.0f;
__device__ __forceinline__ float2 Fma2(float2 a, float2 b, float2 c) {
	float2 d;
	asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
	    "mov.b64 ra, {%2, %3};\n\t"
	    "mov.b64 rb, {%4, %5};\n\t"
	    "mov.b64 rc, {%6, %7};\n\t"
	    "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
	    "mov.b64 {%0, %1}, rd; }"
	    : "=f"(d.x), "=f"(d.y)
	    : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
	return d;
}
__device__ __forceinline__ float2 Mul2(float2 a, float2 b) { return Fma2(a, b, make_float2(kPackNegZero, kPackNegZero)); }
__device__ __forceinline__ float2 Add2(float2 a, float2 b) { return Fma2(a, make_float2(kPackOne, kPackOne), b); }
__device__ __forceinline__ float2 Sub2(float2 a, float2 b) { return Fma2(b, make_float2(kPackNegOne, kPackNegOne), a); }
__device__ __forceinline__ float2 Bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ V3p Scale(float s, const V3p& a) { return V3p{ Mul2(Bc(s), a.xy), __fmul_rn(s, a.z) }; }
__device__ __forceinline__ V3p AddV(const V3p& a, const V3p& b) { return V3p{ Add2(a.xy, b.xy), __fadd_rn(a.z, b.z) }; }
__device__ __forceinline__ V3p SubV(const V3p& a, const V3p& b) { return V3p{ Sub2(a.xy, b.xy), __fsub_rn(a.z, b.z) }; }
// (a0*b0 + a1*b1) + a2*b2 for vectors paired the same way (either order of the pair: the first sum commutes)
__device__ __forceinline__ float Dot(const V3p& a, const V3p& b) {
	const float2 m = Mul2(a.xy, b.xy);
	return __fadd_rn(__fadd_rn(m.x, m.y), __fmul_rn(a.z, b.z));
}

// SolveElementMixed for Energy_MixedSel / Energy_YeohSkinFast, simultaneous solve, undamped (Fem.cpp:523-552, Xpbd.h:122-214).
template <int ENERGY, typename PARAMS>
__device__ __forceinline__ void SolvePrefactoredSimulPacked(const PARAMS& p, const ElemRec& e, VertexRegs (&v)[4], const ElemCompliance& ec) {
	typedef Op<true> O;
	// P[i] = Vec(X[i] - X[3]), Fem.cpp:453
	V3p P[3];
#pragma unroll
	for (int i = 0; i < 3; i++) {
		P[i].xy = make_float2(__double2float_rn(__dsub_rn(v[i].x[0], v[3].x[0])), __double2float_rn(__dsub_rn(v[i].x[1], v[3].x[1])));
		P[i].z = __double2float_rn(__dsub_rn(v[i].x[2], v[3].x[2]));
	}
	// prefactored I1 and its gradient, Fem.cpp:163-192
	float U = 0.0f;
	V3p g0[4];
	const float s2[3] = { O::mul(2.0f, e.QQ[0]), O::mul(2.0f, e.QQ[1]), O::mul(2.0f, e.QQ[2]) };
#pragma unroll
	for (int i = 0; i < 3; i++) {
		U = O::add(U, O::mul(e.QQ[i], Dot(P[i], P[i])));
		g0[i].xy = Mul2(Bc(s2[i]), P[i].xy);
	}
	U = O::add(U, O::mul(e.QR[0], Dot(P[0], P[1])));
	g0[0].xy = Add2(g0[0].xy, Mul2(Bc(e.QR[0]), P[1].xy));
	g0[1].xy = Add2(g0[1].xy, Mul2(Bc(e.QR[0]), P[0].xy));
	U = O::add(U, O::mul(e.QR[1], Dot(P[0], P[2])));
	g0[0].xy = Add2(g0[0].xy, Mul2(Bc(e.QR[1]), P[2].xy));
	g0[2].xy = Add2(g0[2].xy, Mul2(Bc(e.QR[1]), P[0].xy));
	U = O::add(U, O::mul(e.QR[2], Dot(P[1], P[2])));
	g0[1].xy = Add2(g0[1].xy, Mul2(Bc(e.QR[2]), P[2].xy));
	g0[2].xy = Add2(g0[2].xy, Mul2(Bc(e.QR[2]), P[1].xy));
	{	// z components: g0[0].z = (2QQ0*P0z + QR0*P1z) + QR1*P2z, g0[1].z = (2QQ1*P1z + QR0*P0z) + QR2*P2z with the products of rows 0
		// and 1 paired, g0[2].z = (2QQ2*P2z + QR1*P0z) + QR2*P1z with its two QR products paired
		const float2 Z01 = make_float2(P[0].z, P[1].z);
		const float2 T1 = Mul2(make_float2(s2[0], s2[1]), Z01);          // {2QQ0*P0z, 2QQ1*P1z}
		const float2 T2 = Mul2(Bc(e.QR[0]), Z01);                        // {QR0*P0z, QR0*P1z}
		const float2 T3 = Mul2(make_float2(e.QR[1], e.QR[2]), Bc(P[2].z)); // {QR1*P2z, QR2*P2z}
		const float2 T4 = Mul2(make_float2(e.QR[1], e.QR[2]), Z01);      // {QR1*P0z, QR2*P1z}
		g0[0].z = O::add(O::add(T1.x, T2.y), T3.x);
		g0[1].z = O::add(O::add(T1.y, T2.x), T3.y);
		g0[2].z = O::add(O::add(O::mul(s2[2], P[2].z), T4.x), T4.y);
	}
	{
		const V3p zero = { make_float2(0.0f, 0.0f), 0.0f };
		g0[3] = SubV(SubV(SubV(zero, g0[0]), g0[1]), g0[2]);
	}
	float U0 = U;
	if (ENERGY == XF_ENERGY_YEOH_SKIN_FAST) { // Fem.cpp:525-531
		// YeohEnergy = (C0*IM + (C1*IM)*IM) + ((C2*IM)*IM)*IM, YeohSlope = (C0 + (2C1)*IM) + ((3C2)*IM)*IM: the products that differ only
		// in their constant run as pairs
		const float IM = O::sub(U, 3.0f);
		const float C0 = 0.1095f, C1 = 14.95f, C2 = 4.595f;
		const float2 A = Mul2(make_float2(C1, O::mul(2.0f, C1)), Bc(IM));                     // {C1*IM, (2C1)*IM}
		const float2 B = Mul2(Mul2(make_float2(C2, O::mul(3.0f, C2)), Bc(IM)), Bc(IM));       // {(C2*IM)*IM, ((3C2)*IM)*IM}
		U0 = fmaxf(0.0001f, O::add(O::add(O::mul(C0, IM), O::mul(A.x, IM)), O::mul(B.x, IM)));
		const float gScale = O::add(O::add(C0, A.y), B.y);
#pragma unroll
		for (int n = 0; n < 4; n++) { g0[n] = Scale(gScale, g0[n]); }
	}
	// F[c][r] = (Qi[0][r]*P[0][c] + Qi[1][r]*P[1][c]) + Qi[2][r]*P[2][c], Fem.cpp:311-323; Fr[r] = {F[0][r], F[1][r]}, F[2][r]
	V3p Fr[3];
#pragma unroll
	for (int r = 0; r < 3; r++) { Fr[r] = AddV(AddV(Scale(e.Qi[0][r], P[0]), Scale(e.Qi[1][r], P[1])), Scale(e.Qi[2][r], P[2])); }
	// adjugate, Fem.cpp:284-300.  Q0 = {-adj[0][1], adj[0][0]}, Q1 = {adj[1][1], -adj[1][0]}, Q2 = {-adj[2][1], adj[2][0]}
	const float2 Q0 = Sub2(Mul2(Fr[1].xy, Bc(Fr[2].z)), Mul2(Fr[2].xy, Bc(Fr[1].z)));
	const float2 Q1 = Sub2(Mul2(Fr[0].xy, Bc(Fr[2].z)), Mul2(Fr[2].xy, Bc(Fr[0].z)));
	const float2 Q2 = Sub2(Mul2(Fr[0].xy, Bc(Fr[1].z)), Mul2(Fr[1].xy, Bc(Fr[0].z)));
	const float a02 = O::sub(O::mul(Fr[1].xy.x, Fr[2].xy.y), O::mul(Fr[2].xy.x, Fr[1].xy.y));  //  adj[0][2]
	const float n12 = O::sub(O::mul(Fr[0].xy.x, Fr[2].xy.y), O::mul(Fr[2].xy.x, Fr[0].xy.y));  // -adj[1][2]
	const float a22 = O::sub(O::mul(Fr[0].xy.x, Fr[1].xy.y), O::mul(Fr[1].xy.x, Fr[0].xy.y));  //  adj[2][2]
	// J = (F[0][0]*adj[0][0] + F[0][1]*adj[1][0]) + F[0][2]*adj[2][0], Fem.cpp:304-306; adj[1][0] = -Q1.y
	const float J = O::add(O::sub(O::mul(Fr[0].xy.x, Q0.y), O::mul(Fr[1].xy.x, Q1.y)), O::mul(Fr[2].xy.x, Q2.y));
	const float d = O::sub(J, p.a);
	const float U1 = O::mul(d, d);
	const float s = O::mul(2.0f, d);
	// GJ[c][r] = adj[c][r] * s, paired {GJ[c][1], GJ[c][0]}: the sign of a cofactor moves into the scale, (-x)*s == x*(-s)
	const float2 sNP = make_float2(-s, s), sPN = make_float2(s, -s);
	const float2 G0 = Mul2(Q0, sNP), G1 = Mul2(Q1, sPN), G2 = Mul2(Q2, sNP);
	const float G0z = O::mul(a02, s), G1z = O::mul(n12, -s), G2z = O::mul(a22, s);
	// g1[n][k] = (GJ[0][k]*Qi[n][0] + GJ[1][k]*Qi[n][1]) + GJ[2][k]*Qi[n][2], Fem.cpp:338-354; g1[n].xy = {g1[n][1], g1[n][0]}
	V3p g1[4];
#pragma unroll
	for (int n = 0; n < 3; n++) {
		g1[n].xy = Add2(Add2(Mul2(G0, Bc(e.Qi[n][0])), Mul2(G1, Bc(e.Qi[n][1]))), Mul2(G2, Bc(e.Qi[n][2])));
		g1[n].z = O::add(O::add(O::mul(G0z, e.Qi[n][0]), O::mul(G1z, e.Qi[n][1])), O::mul(G2z, e.Qi[n][2]));
	}
	{
		const V3p zero = { make_float2(0.0f, 0.0f), 0.0f };
		g1[3] = SubV(SubV(SubV(zero, g1[0]), g1[1]), g1[2]);
	}
	// EnergyXpbdConstrainSimultaneous<2>, Xpbd.h:154-170: w00 and w11 run as one pair
	float2 W = make_float2(1.0e-22f, 1.0e-22f);
	float w10 = 1.0e-22f;
#pragma unroll
	for (int n = 0; n < 4; n++) {
		const float2 m0 = Mul2(g0[n].xy, g0[n].xy), m1 = Mul2(g1[n].xy, g1[n].xy);
		const float2 S = make_float2(O::add(m0.x, m0.y), O::add(m1.y, m1.x));
		const float2 Z = make_float2(g0[n].z, g1[n].z);
		W = Add2(W, Mul2(Bc(v[n].w), Add2(S, Mul2(Z, Z))));
	}
#pragma unroll
	for (int n = 0; n < 4; n++) { // dot(g1[n], g0[n]) = (g1x*g0x + g1y*g0y) + g1z*g0z
		const float dt = O::add(O::add(O::mul(g1[n].xy.y, g0[n].xy.x), O::mul(g1[n].xy.x, g0[n].xy.y)), O::mul(g1[n].z, g0[n].z));
		w10 = O::add(w10, O::mul(v[n].w, dt));
	}
	const float A0 = O::add(W.x, O::mul(O::mul(2.0f, U0), ec.alpha0));
	const float b0 = O::mul(-2.0f, U0);
	const float A2 = O::add(W.y, O::mul(O::mul(2.0f, U1), ec.alpha1));
	const float b1 = O::mul(-2.0f, U1);
	float l0, l1;
	Cramer2<true>(A0, w10, A2, b0, b1, l0, l1);
	// X[n] += double(w[n] * (l0*g0[n] + l1*g1[n])), Xpbd.h:209-213
#pragma unroll
	for (int n = 0; n < 4; n++) {
		const float2 a = Mul2(g0[n].xy, Bc(l0)), b = Mul2(g1[n].xy, Bc(l1));
		const float2 acc = make_float2(O::add(a.x, b.y), O::add(a.y, b.x));
		const float2 dxy = Mul2(Bc(v[n].w), acc);
		const float2 lz = Mul2(make_float2(g0[n].z, g1[n].z), make_float2(l0, l1)); // {l0*g0z, l1*g1z}
		const float dz = O::mul(v[n].w, O::add(lz.x, lz.y));
		v[n].x[0] = __dadd_rn(v[n].x[0], (double)dxy.x);
		v[n].x[1] = __dadd_rn(v[n].x[1], (double)dxy.y);
		v[n].x[2] = __dadd_rn(v[n].x[2], (double)dz);
	}
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct UsePacked {
	static constexpr bool value = EXACT && SIMUL && !DAMPED && (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
};

#else

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct UsePacked {
	static constexpr bool value = false;
};

#endif

}  // namespace xf
