// Warp-cooperative element solve: FOUR lanes per element (north_star's "one warp-cooperative thread group per element"), for the
// prefactored energies (MixedSel, YeohSkinFast) in simultaneous mode without in-constraint damping - the headline configuration.
//
// Same contract as SolvePrefactoredSimulPacked / SolveElementGathered (xf_element.cuh): every scalar of the reference's
// SolveElementMixed (Fem.cpp:437-564) -> EnergyXpbdConstrainSimultaneous<2> (Xpbd.h:122-214) is computed ONCE, by one lane, with the
// same round-to-nearest operations in the same association order, so the result is bit-identical; what changes is which lane
// computes which scalar and how the scalars travel.
//
//   vertex lanes     lane n owns vertex n: its record {x, w}; the edge P[n] = Vec(X[n] - X[3]) (Fem.cpp:453), the three weighted
//                    gradient dots of vertex n (Xpbd.h:154-170) and the update of X[n] (Xpbd.h:209-213)
//   component lanes  lane c < 3 owns coordinate c: P[.][c], column c of F (Fem.cpp:311-323), the cofactors adj[.][c]
//                    (Fem.cpp:284-300), and component c of the gradients g0[0..3][c], g1[0..3][c] including the closing
//                    g[3] = ((0 - g[0]) - g[1]) - g[2] (Fem.cpp:338-354, 163-192) - every formula is uniform over c, no lane
//                    needs its own coefficient pattern
//   every lane       the scalar chains: I1 / Yeoh polynomial (Fem.cpp:525-531) and the 2x2 Cramer solve (Xpbd.h:173-179),
//                    redundantly (a broadcast would cost a shuffle stage on the dependence chain and save no issue slot: all
//                    four lanes of a quad issue together anyway)
//
// Exchanges per element: x3 broadcast (3 fp64 shuffles; none if every lane gathered vertex 3 itself), all-gather of P (9), the two foreign columns of F for the cofactors (6,
// per-lane source), J (1), the 4 x 3 transposition of g0 and g1 from component lanes to vertex lanes through shared memory (8
// scalar stores, 2 x 128-bit loads per lane), all-gather of the 4 x 3 weighted dots (12).
//
// This file is written against a LANE POLICY `L` (shuffles, the quad's shared-memory row, round-to-nearest arithmetic) so that the
// very same source runs (a) on the device (DevLane4 below) and (b) on the host with four threads playing the four lanes
// (tests/coop_emu/coop_emu.cpp), where it is checked bit for bit against the C oracle without a GPU.  It is measured by
// xf_probe_coop.cu (xf_debug_coop_element) against the one-thread solve; the stepping kernels do not use it - DESIGN section 6
// says why, with the numbers.
#pragma once

#ifndef XF_COOP_FN
#define XF_COOP_FN __device__ __forceinline__
#endif
#ifdef __CUDACC__
#define XF_COOP_UNROLL _Pragma("unroll")
#else
#define XF_COOP_UNROLL
#endif

namespace xf {

// floats of shared memory per quad: 4 rows {g0x g0y g0z pad g1x g1y g1z pad}, padded to 36 so that the eight quads of a warp
// start 4 banks apart (scalar stores and 128-bit loads conflict-free)
constexpr int kCoopQuadFloats = 36;

template <class L>
XF_COOP_FN float CoopDot3(float a0, float a1, float a2, float b0, float b1, float b2) {
	typedef typename L::O O;
	return O::add(O::add(O::mul(a0, b0), O::mul(a1, b1)), O::mul(a2, b2)); // vectormath.h:367
}

// `x`, `w`: the calling lane's vertex record (lane n of the quad holds vertex n of the element); `x` is updated in place.
// `e`: the element's constants (same in the four lanes); `a` = 1 + mu/lambda (Fem.cpp:447); `alpha0/1` = compliance / dt^2.
// X3_GATHERED: every lane also gathered vertex 3's position itself (`x3in`; a second 256-bit load per lane next to its own
// record, in parallel with it) instead of receiving it from lane 3 - one shuffle stage less on the dependence chain.  After the
// solve, `x3in` is stale in lanes 0..2 (lane 3 holds the updated vertex).
template <int ENERGY, bool X3_GATHERED, class L, class REC>
XF_COOP_FN void SolvePrefactoredSimulCoop4(const L& ln, float a, const REC& e, float alpha0, float alpha1, double (&x)[3], float w,
                                           const double (&x3in)[3]) {
	typedef typename L::O O;
	const int q = ln.q();
	const int c = q < 2 ? q : 2; // component owned in the component-lane phase (lane 3 shadows lane 2, its results are not used)

	// ---- vertex lanes: P[q] = Vec(X[q] - X[3]), Fem.cpp:453
	float Pown[3];
	{
		double x3[3];
		XF_COOP_UNROLL
		for (int k = 0; k < 3; k++) { x3[k] = X3_GATHERED ? x3in[k] : ln.shfl(x[k], 3); }
		XF_COOP_UNROLL
		for (int k = 0; k < 3; k++) { Pown[k] = O::d2f(O::dsub(x[k], x3[k])); }
	}
	float P[3][3];
	XF_COOP_UNROLL
	for (int i = 0; i < 3; i++) {
		XF_COOP_UNROLL
		for (int k = 0; k < 3; k++) { P[i][k] = ln.shfl(Pown[k], i); }
	}

	// ---- every lane: prefactored I1 (Fem.cpp:163-192, the energy part) and the Yeoh polynomial (Fem.cpp:525-531)
	float U = 0.0f;
	U = O::add(U, O::mul(e.QQ[0], CoopDot3<L>(P[0][0], P[0][1], P[0][2], P[0][0], P[0][1], P[0][2])));
	U = O::add(U, O::mul(e.QQ[1], CoopDot3<L>(P[1][0], P[1][1], P[1][2], P[1][0], P[1][1], P[1][2])));
	U = O::add(U, O::mul(e.QQ[2], CoopDot3<L>(P[2][0], P[2][1], P[2][2], P[2][0], P[2][1], P[2][2])));
	U = O::add(U, O::mul(e.QR[0], CoopDot3<L>(P[0][0], P[0][1], P[0][2], P[1][0], P[1][1], P[1][2])));
	U = O::add(U, O::mul(e.QR[1], CoopDot3<L>(P[0][0], P[0][1], P[0][2], P[2][0], P[2][1], P[2][2])));
	U = O::add(U, O::mul(e.QR[2], CoopDot3<L>(P[1][0], P[1][1], P[1][2], P[2][0], P[2][1], P[2][2])));
	float U0 = U, gScale = 1.0f;
	if (ENERGY == XF_ENERGY_YEOH_SKIN_FAST) {
		const float C0 = 0.1095f, C1 = 14.95f, C2 = 4.595f;
		const float IM = O::sub(U, 3.0f);
		U0 = O::max(0.0001f, O::add(O::add(O::mul(C0, IM), O::mul(O::mul(C1, IM), IM)), O::mul(O::mul(O::mul(C2, IM), IM), IM)));
		gScale = O::add(O::add(C0, O::mul(O::mul(2.0f, C1), IM)), O::mul(O::mul(O::mul(3.0f, C2), IM), IM));
	}

	// ---- component lanes: coordinate c of the three edges
	// (selects, not branches: the compiler turned the plain ternaries into a divergent region on the dependence chain)
	const float pc0 = O::sel(c == 0, P[0][0], O::sel(c == 1, P[0][1], P[0][2]));
	const float pc1 = O::sel(c == 0, P[1][0], O::sel(c == 1, P[1][1], P[1][2]));
	const float pc2 = O::sel(c == 0, P[2][0], O::sel(c == 1, P[2][1], P[2][2]));
	// gradient of the prefactored I1, component c of g0[0..3] (Fem.cpp:171-191): g[i] = 2QQ_i P_i, then the pairs (0,1) (0,2) (1,2)
	float g0c[4];
	g0c[0] = O::add(O::add(O::mul(O::mul(2.0f, e.QQ[0]), pc0), O::mul(e.QR[0], pc1)), O::mul(e.QR[1], pc2));
	g0c[1] = O::add(O::add(O::mul(O::mul(2.0f, e.QQ[1]), pc1), O::mul(e.QR[0], pc0)), O::mul(e.QR[2], pc2));
	g0c[2] = O::add(O::add(O::mul(O::mul(2.0f, e.QQ[2]), pc2), O::mul(e.QR[1], pc0)), O::mul(e.QR[2], pc1));
	g0c[3] = O::sub(O::sub(O::sub(0.0f, g0c[0]), g0c[1]), g0c[2]);
	if (ENERGY == XF_ENERGY_YEOH_SKIN_FAST) {
		XF_COOP_UNROLL
		for (int n = 0; n < 4; n++) { g0c[n] = O::mul(g0c[n], gScale); }
	}
	// column c of F: m[r] = F[c][r] = (Qi[0][r]*P[0][c] + Qi[1][r]*P[1][c]) + Qi[2][r]*P[2][c], Fem.cpp:311-323
	float m[3];
	XF_COOP_UNROLL
	for (int r = 0; r < 3; r++) { m[r] = O::add(O::add(O::mul(e.Qi[0][r], pc0), O::mul(e.Qi[1][r], pc1)), O::mul(e.Qi[2][r], pc2)); }
	// cofactors with second index c (Fem.cpp:284-300) need the two OTHER columns of F: A = the lower, B = the higher of them
	float A[3], B[3];
	{
		const int srcA = c == 0 ? 1 : 0, srcB = c == 2 ? 1 : 2;
		XF_COOP_UNROLL
		for (int r = 0; r < 3; r++) {
			A[r] = ln.shfl(m[r], srcA);
			B[r] = ln.shfl(m[r], srcB);
		}
	}
	// raw_i = adj[i][c] up to the sign (-1)^(i+c), which moves into the scale below ((-x)*s == x*(-s) bit for bit)
	const float raw0 = O::sub(O::mul(A[1], B[2]), O::mul(A[2], B[1]));
	const float raw1 = O::sub(O::mul(A[0], B[2]), O::mul(A[2], B[0]));
	const float raw2 = O::sub(O::mul(A[0], B[1]), O::mul(A[1], B[0]));
	// J = (F[0][0]*adj[0][0] + F[0][1]*adj[1][0]) + F[0][2]*adj[2][0], Fem.cpp:304-306: lane 0 has column 0 and adj[.][0]
	// (adj[1][0] = -raw1: x + (-y) == x - y)
	const float Jl = O::add(O::sub(O::mul(m[0], raw0), O::mul(m[1], raw1)), O::mul(m[2], raw2));
	const float J = ln.shfl(Jl, 0);
	const float d = O::sub(J, a);
	const float U1 = O::mul(d, d);            // Fem.cpp:537-540 (weight == 1)
	const float s = O::mul(2.0f, d);
	const float se = O::sel((c & 1) != 0, -s, s); // scale of the cofactors with even first index
	const float G0 = O::mul(raw0, se), G1 = O::mul(raw1, -se), G2 = O::mul(raw2, se); // GJ[i][c]
	// component c of g1[n] = (GJ[0][c]*Qi[n][0] + GJ[1][c]*Qi[n][1]) + GJ[2][c]*Qi[n][2], Fem.cpp:338-354
	float g1c[4];
	XF_COOP_UNROLL
	for (int n = 0; n < 3; n++) { g1c[n] = O::add(O::add(O::mul(G0, e.Qi[n][0]), O::mul(G1, e.Qi[n][1])), O::mul(G2, e.Qi[n][2])); }
	g1c[3] = O::sub(O::sub(O::sub(0.0f, g1c[0]), g1c[1]), g1c[2]);

	// ---- component lanes -> vertex lanes: row n of the quad's shared memory = {g0[n][0..2], -, g1[n][0..2], -}
	// (lane 3 writes its shadow values into the two padding columns: no divergence around the stores)
	XF_COOP_UNROLL
	for (int n = 0; n < 4; n++) {
		ln.sts(8 * n + q, g0c[n]);
		ln.sts(8 * n + 4 + q, g1c[n]);
	}
	ln.sync();
	float g0v[3], g1v[3];
	ln.lds3(8 * q, g0v);
	ln.lds3(8 * q + 4, g1v);
	ln.sync(); // the row is rewritten by the next element of this quad

	// ---- vertex lanes: the weighted dots of vertex q, Xpbd.h:154-170, then summed in vertex order by every lane
	const float t00 = O::mul(w, CoopDot3<L>(g0v[0], g0v[1], g0v[2], g0v[0], g0v[1], g0v[2]));
	const float t10 = O::mul(w, CoopDot3<L>(g1v[0], g1v[1], g1v[2], g0v[0], g0v[1], g0v[2]));
	const float t11 = O::mul(w, CoopDot3<L>(g1v[0], g1v[1], g1v[2], g1v[0], g1v[1], g1v[2]));
	float w00 = 1.0e-22f, w10 = 1.0e-22f, w11 = 1.0e-22f;
	XF_COOP_UNROLL
	for (int n = 0; n < 4; n++) {
		w00 = O::add(w00, ln.shfl(t00, n));
		w10 = O::add(w10, ln.shfl(t10, n));
		w11 = O::add(w11, ln.shfl(t11, n));
	}
	const float A0 = O::add(w00, O::mul(O::mul(2.0f, U0), alpha0));
	const float b0 = O::mul(-2.0f, U0);
	const float A2 = O::add(w11, O::mul(O::mul(2.0f, U1), alpha1));
	const float b1 = O::mul(-2.0f, U1);
	// 2x2 Cramer, Xpbd.h:173-179
	const float invA00 = O::rcp(A0);
	const float invA11 = O::rcp(A2);
	const float p0 = O::mul(w10, invA00);
	const float p1 = O::mul(w10, invA11);
	const float invDet = O::rcp(O::max(0.00000001f, O::sub(1.0f, O::mul(p0, p1))));
	const float q0 = O::mul(b0, invA00);
	const float q1 = O::mul(b1, invA11);
	const float l0 = O::mul(invDet, O::sub(q0, O::mul(p0, q1)));
	const float l1 = O::mul(invDet, O::sub(q1, O::mul(q0, p1)));
	// X[q] += double(w[q] * (l0*g0[q] + l1*g1[q])), Xpbd.h:209-213
	XF_COOP_UNROLL
	for (int k = 0; k < 3; k++) {
		const float acc = O::add(O::mul(l0, g0v[k]), O::mul(l1, g1v[k]));
		x[k] = O::dadd(x[k], O::f2d(O::mul(w, acc)));
	}
}

#ifdef __CUDACC__
// Device lane policy: quads are aligned groups of four lanes of a warp; `sm` = the quad's kCoopQuadFloats floats of shared memory.
struct DevOps {
	static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
	static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
	static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
	static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
	static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
	static __device__ __forceinline__ float sel(bool cond, float a, float b) {
		float r;
		asm("{ .reg .pred p;\n\t"
		    "setp.ne.s32 p, %3, 0;\n\t"
		    "selp.f32 %0, %1, %2, p; }"
		    : "=f"(r)
		    : "f"(a), "f"(b), "r"((int)cond));
		return r;
	}
	static __device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
	static __device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
	static __device__ __forceinline__ float d2f(double a) { return __double2float_rn(a); }
	static __device__ __forceinline__ double f2d(float a) { return (double)a; }
};
struct DevLane4 {
	typedef DevOps O;
	unsigned mask;
	int lane4;
	float* sm;
	__device__ __forceinline__ int q() const { return lane4; }
	__device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(mask, v, src, 4); }
	__device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(mask, v, src, 4); }
	__device__ __forceinline__ void sts(int i, float v) const { sm[i] = v; }
	__device__ __forceinline__ void lds3(int i, float (&out)[3]) const {
		const float4 r = *reinterpret_cast<const float4*>(sm + i);
		out[0] = r.x; out[1] = r.y; out[2] = r.z;
	}
	__device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
#endif

}  // namespace xf
