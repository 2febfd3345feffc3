/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's small-step XPBD linear-tet path, written so that with
 * -ffp-contract=off it reproduces the reference's fp32/fp64 rounding operation by operation.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/XPBDFEM).  Parity status: PINNED BY EXECUTION against the unmodified reference
 * (see xpbd_oracle.h).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or load this file; the product (xpbd-fem_b200/) never does.
 *
 * Conventions: v3 = float[3]; a 3x3 matrix is stored column-major like the reference's mat3,
 * m[c][r] == reference m[c][r]; vertex state is double[3] per vertex (reference dvec3).
 */
#define _POSIX_C_SOURCE 199309L
#include "xpbd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- flag word (Settings.h:9-75) ---- */
#define XO_ENERGY_BIT 6
#define XO_ENERGY_MASK 31u
#define XO_SOLVE_BIT 11
#define XO_RAYLEIGH_BIT 20
#define XO_RAYLEIGH_MASK 3u
#define XO_LOCK_LEFT (1u << 26)
#define XO_LOCK_RIGHT (1u << 27)
enum { EN_PIXAR = 0, EN_MIXED = 3, EN_MIXED_SEL = 4, EN_YEOH = 5, EN_YEOH_SEL = 6, EN_YEOH_FAST = 7, EN_CONT_FIRST = 8, EN_CONT_LAST = 10 };
enum { RAY_PAPER = 0, RAY_LIMIT = 1, RAY_POST = 2, RAY_POST_AMORTIZED = 3 };
#define XO_AMORTIZATION_PERIOD 8u
#define XO_FLAG_LEFT 1
#define XO_FLAG_RIGHT 2
#define XO_FLAG_PICKABLE 16

typedef struct {
	uint32_t i[4];
	float Qi[3][3]; /* [col][row] */
	float QQ[3], QR[3];
	float volume, surfaceArea;
} xo_tet;

struct xo_scene {
	uint32_t nV, nT;
	double *X0, *X, *O, *V; /* 3 per vertex */
	float* w;
	uint8_t* flags;
	xo_tet* t;
	uint32_t* tOrder;
	float origin[3];
	int groundOn;
	float groundY, groundFriction;
	int stateF32; /* study mode (xo_set_state_precision): X, O, V are rounded to fp32 after every write, as a 16-byte record would hold them */
	uint32_t handleCount;
	uint32_t handleIdx[64];
	float handleTarget[64][3];
};

/* ---- tiny vector helpers; each mirrors the operation order of vectormath.h ---- */
static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; } /* vectormath.h:367 */

/* mat3 * vec3, vectormath.h:508-513: row-dot-vector, rows gathered from the columns */
static void mat_vec(float m[3][3], const float* v, float* out) {
	for (int r = 0; r < 3; r++) { out[r] = m[0][r] * v[0] + m[1][r] * v[1] + m[2][r] * v[2]; }
}
/* mat3 * mat3, vectormath.h:514: column by column */
static void mat_mat(float x[3][3], float y[3][3], float out[3][3]) {
	for (int c = 0; c < 3; c++) { mat_vec(x, y[c], out[c]); }
}
/* determinant(mat3), vectormath.cpp:34-39 */
static float det3(float m[3][3]) {
	float a = m[0][0], b = m[1][0], c = m[2][0];
	float d = m[0][1], e = m[1][1], f = m[2][1];
	float g = m[0][2], h = m[1][2], i = m[2][2];
	return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}
/* inverse(mat3), vectormath.cpp:41-60: cofactors in fp64, narrowed, then scaled by 1/det in fp32 */
static void inverse3(float m[3][3], float out[3][3]) {
	double m00 = m[0][0], m01 = m[0][1], m02 = m[0][2];
	double m10 = m[1][0], m11 = m[1][1], m12 = m[1][2];
	double m20 = m[2][0], m21 = m[2][1], m22 = m[2][2];
	float adj[3][3];
	adj[0][0] = (float)+(m11 * m22 - m12 * m21);
	adj[0][1] = (float)-(m01 * m22 - m02 * m21);
	adj[0][2] = (float)+(m01 * m12 - m02 * m11);
	adj[1][0] = (float)-(m10 * m22 - m12 * m20);
	adj[1][1] = (float)+(m00 * m22 - m02 * m20);
	adj[1][2] = (float)-(m00 * m12 - m02 * m10);
	adj[2][0] = (float)+(m10 * m21 - m11 * m20);
	adj[2][1] = (float)-(m00 * m21 - m01 * m20);
	adj[2][2] = (float)+(m00 * m11 - m01 * m10);
	float c0[3] = { adj[0][0], adj[1][0], adj[2][0] };
	float det = dot3(m[0], c0);
	float s = 1.0f / det; /* mat3 / float == m * (1.0f / s), vectormath.h:507 */
	for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { out[c][r] = adj[c][r] * s; } }
}
/* adjugate(mat3), Fem.cpp:284-300 (all fp32) */
static void adjugate3(float m[3][3], float adj[3][3]) {
	float m00 = m[0][0], m01 = m[0][1], m02 = m[0][2];
	float m10 = m[1][0], m11 = m[1][1], m12 = m[1][2];
	float m20 = m[2][0], m21 = m[2][1], m22 = m[2][2];
	adj[0][0] = +(m11 * m22 - m12 * m21);
	adj[0][1] = -(m01 * m22 - m02 * m21);
	adj[0][2] = +(m01 * m12 - m02 * m11);
	adj[1][0] = -(m10 * m22 - m12 * m20);
	adj[1][1] = +(m00 * m22 - m02 * m20);
	adj[1][2] = -(m00 * m12 - m02 * m10);
	adj[2][0] = +(m10 * m21 - m11 * m20);
	adj[2][1] = -(m00 * m21 - m01 * m20);
	adj[2][2] = +(m00 * m11 - m01 * m10);
}

/* P[n] = Vec(X[is[n]] - X[is[3]]): difference in fp64, then narrowed (Fem.cpp:453, 202) */
static void gather_edges(const double* X, const uint32_t* is, float P[3][3]) {
	const double* r = X + 3 * (size_t)is[3];
	for (int n = 0; n < 3; n++) {
		const double* p = X + 3 * (size_t)is[n];
		for (int k = 0; k < 3; k++) { P[n][k] = (float)(p[k] - r[k]); }
	}
}

/* ------------------------------------------------------------------------------------------
 * Mesh generation: GenerateHexBlock + GenerateTetBlock, MeshGen.cpp:156-244.
 * The three RandF() calls inside `vec3(a, b, c)` are unsequenced in C++; g++ (the compiler the
 * strict reference build uses here) evaluates them right to left, i.e. z, then y, then x.
 * tests/test_oracle_ref.py pins this against the reference build.
 * ------------------------------------------------------------------------------------------ */
void xo_generate_tet_block(uint32_t width, uint32_t height, uint32_t depth, float sx, float sy, float sz,
                           uint32_t pattern, float wonkiness, float* nodes, uint32_t* idxStream) {
	float scale[3] = { sx, sy, sz };
	float minX[3] = { sx * (-0.5f * (float)width), sy * (-0.5f * (float)height), sz * (-0.5f * (float)depth) };
	uint32_t randState = 1;
	for (uint32_t z = 0; z < depth + 1; z++) {
		int zEdge = (z == 0 || z == depth);
		for (uint32_t y = 0; y < height + 1; y++) {
			int yEdge = (y == 0 || y == height);
			for (uint32_t x = 0; x < width + 1; x++) {
				int xEdge = (x == 0 || x == width);
				uint32_t i = x + y * (width + 1) + z * (width + 1) * (height + 1);
				float r[3];
				int edge[3] = { xEdge, yEdge, zEdge };
				for (int k = 2; k >= 0; k--) { /* right-to-left argument evaluation */
					if (edge[k]) { r[k] = 0.0f; }
					else {
						randState = (uint32_t)(((uint64_t)randState * 48271u) % 0x7fffffffu);
						r[k] = (float)randState * 9.3132258e-10f - 1.0f;
					}
				}
				float g[3] = { (float)x, (float)y, (float)z };
				for (int k = 0; k < 3; k++) {
					float wiggle = wonkiness * r[k];
					nodes[3 * i + k] = minX[k] + scale[k] * (g[k] + 0.5f * wiggle);
				}
			}
		}
	}
	static const int tetCorners[6][4] = { { 0, 1, 5, 7 }, { 0, 7, 3, 1 }, { 0, 2, 3, 7 }, { 0, 7, 6, 2 }, { 0, 4, 6, 7 }, { 0, 7, 5, 4 } };
	uint32_t ie = 0;
	for (uint32_t z = 0; z < depth; z++) {
		for (uint32_t y = 0; y < height; y++) {
			for (uint32_t x = 0; x < width; x++) {
				uint32_t ih[8];
				for (uint32_t k = 0; k < 2; k++) {
					for (uint32_t j = 0; j < 2; j++) {
						for (uint32_t i = 0; i < 2; i++) {
							if (pattern == 0) {
								ih[i + j * 2 + k * 4] = (x + i) + (y + j) * (width + 1) + (z + k) * (width + 1) * (height + 1);
							} else { /* Pattern_Mirrored, MeshGen.cpp:196-205 */
								uint32_t xx = x % 2 == 0 ? i : (1 - i);
								uint32_t yy = y % 2 == 0 ? j : (1 - j);
								uint32_t zz = z % 2 == 0 ? k : (1 - k);
								uint32_t hi = i + j * 2 + k * 4;
								if ((x + y + z) % 2 == 1) { hi = 7 - hi; }
								ih[hi] = (x + xx) + (y + yy) * (width + 1) + (z + zz) * (width + 1) * (height + 1);
							}
						}
					}
				}
				for (int t = 0; t < 6; t++) {
					uint32_t* rec = idxStream + 30 * (size_t)ie + 5 * t;
					rec[0] = 4;
					for (int c = 0; c < 4; c++) { rec[1 + c] = ih[tetCorners[t][c]]; }
				}
				++ie;
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * Element init: InitTriTetFiniteElement<Tet>, Fem.cpp:196-224, with
 * GeneratePrefactoredCoefficients<Tet, vec3, 4, 4>, Fem.cpp:131-161 (4-point rule, weights 1/4).
 * The T4 shape-function gradients are {e0, e1, e2, (-1,-1,-1)} (Fem.cpp:56-57) so the map Jacobian
 * J = sum_n dN[n] (x) P[n] has J[c][r] = P[r][c] (the P[3] = 0 term adds -0).
 * ------------------------------------------------------------------------------------------ */
static void init_tet(const double* X, xo_tet* t, float density, float* m) {
	float P[4][3];
	const double* r = X + 3 * (size_t)t->i[3];
	for (int n = 0; n < 4; n++) {
		const double* p = X + 3 * (size_t)t->i[n];
		for (int k = 0; k < 3; k++) { P[n][k] = (float)(p[k] - r[k]); }
	}
	float J[3][3];
	for (int c = 0; c < 3; c++) { for (int rr = 0; rr < 3; rr++) { J[c][rr] = P[rr][c]; } }
	inverse3(J, t->Qi);

	float volume = 0.0f;
	float QQ[3] = { 0.0f, 0.0f, 0.0f }, QR[3] = { 0.0f, 0.0f, 0.0f };
	for (int p = 0; p < 4; p++) {
		float Ji[3][3];
		inverse3(J, Ji);
		float w = det3(J) * (1.0f / 6.0f) * 0.25f;
		/* Qi[i] = Ji * dN[i] = column i of Ji */
		volume += w;
		for (int i = 0; i < 3; i++) { QQ[i] += w * dot3(Ji[i], Ji[i]); }
		int head = 0;
		for (int i = 0; i < 3; i++) {
			for (int j = i + 1; j < 3; j++) { QR[head++] += w * 2.0f * dot3(Ji[i], Ji[j]); }
		}
	}
	for (int i = 0; i < 3; i++) { t->QQ[i] = QQ[i] / volume; t->QR[i] = QR[i] / volume; }

	t->volume = (1.0f / 6.0f) * det3(J);
	/* FaceArea(i0,i1,i2) = 0.5f * |cross(P[i0]-P[i2], P[i1]-P[i2])|, Fem.cpp:210-213 */
	static const int faces[4][3] = { { 0, 1, 2 }, { 1, 3, 2 }, { 0, 2, 3 }, { 3, 1, 0 } };
	float area = 0.0f;
	for (int f = 0; f < 4; f++) {
		float a[3], b[3], c[3];
		for (int k = 0; k < 3; k++) { a[k] = P[faces[f][0]][k] - P[faces[f][2]][k]; b[k] = P[faces[f][1]][k] - P[faces[f][2]][k]; }
		c[0] = a[1] * b[2] - b[1] * a[2];
		c[1] = a[2] * b[0] - b[2] * a[0];
		c[2] = a[0] * b[1] - b[0] * a[1];
		float fa = 0.5f * sqrtf(dot3(c, c));
		area = (f == 0) ? fa : area + fa;
	}
	t->surfaceArea = area;
	/* mass lumping: (Lumps[n] / Lumps[4]) * density * volume with Lumps = {1,1,1,1,4} */
	for (int n = 0; n < 4; n++) { m[t->i[n]] += (1.0f / 4.0f) * density * t->volume; }
}

/* GeoLinear3d::Init, Geo.cpp:697-772 (tet records only; rings/edges are rendering data) */
xo_scene* xo_create(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount,
                    float density, int autoResize) {
	xo_scene* s = (xo_scene*)calloc(1, sizeof(xo_scene));
	uint32_t nT = 0, maxVert = 0;
	for (uint32_t i = 0; i < idxCount; i += 5) { /* Connectivity3d::Init, Connectivity.cpp:245-266 */
		if (idxStream[i] != 4) { free(s); return NULL; }
		for (int j = 0; j < 4; j++) { if (idxStream[i + 1 + j] > maxVert) { maxVert = idxStream[i + 1 + j]; } }
		nT++;
	}
	uint32_t nV = 1 + maxVert;
	if (nV != nodeFloatCount / 3) { free(s); return NULL; }
	s->nV = nV; s->nT = nT;
	s->X0 = (double*)malloc(sizeof(double) * 3 * nV);
	s->X = (double*)malloc(sizeof(double) * 3 * nV);
	s->O = (double*)malloc(sizeof(double) * 3 * nV);
	s->V = (double*)calloc(3 * (size_t)nV, sizeof(double));
	s->w = (float*)calloc(nV, sizeof(float));
	s->flags = (uint8_t*)malloc(nV);
	s->t = (xo_tet*)malloc(sizeof(xo_tet) * nT);
	s->tOrder = (uint32_t*)malloc(sizeof(uint32_t) * nT);

	float mn[3] = { 1.0e24f, 1.0e24f, 1.0e24f }, mx[3] = { -1.0e24f, -1.0e24f, -1.0e24f };
	for (uint32_t i = 0; i < nV; i++) {
		for (int k = 0; k < 3; k++) {
			double v = (double)nodeXYZ[3 * i + k];
			s->X0[3 * i + k] = s->X[3 * i + k] = s->O[3 * i + k] = v;
			float f = (float)v;
			mn[k] = fminf(mn[k], f);
			mx[k] = fmaxf(mx[k], f);
		}
		s->flags[i] = XO_FLAG_PICKABLE;
	}
	float dims[3] = { mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2] };
	double maxDim = (double)fmaxf(dims[0], fmaxf(dims[1], dims[2]));
	for (uint32_t i = 0; i < nV; i++) {
		float fx = (float)s->X[3 * i];
		if ((fx - mn[0]) / (mx[0] - mn[0]) < 0.02f) { s->flags[i] |= XO_FLAG_LEFT; }
		if ((fx - mn[0]) / (mx[0] - mn[0]) > 0.98f) { s->flags[i] |= XO_FLAG_RIGHT; }
		if (autoResize) { /* 0.085 * ((X - 0.5 * dvec3(aabbMin + aabbMax)) / maxDim), Geo.cpp:727 */
			for (int k = 0; k < 3; k++) {
				double c = 0.5 * (double)(mn[k] + mx[k]);
				double v = 0.085 * ((s->X[3 * i + k] - c) / maxDim);
				s->X0[3 * i + k] = s->X[3 * i + k] = s->O[3 * i + k] = v;
			}
		}
	}
	for (uint32_t e = 0; e < nT; e++) {
		for (int j = 0; j < 4; j++) { s->t[e].i[j] = idxStream[5 * (size_t)e + 1 + j]; }
		init_tet(s->X, &s->t[e], density, s->w);
	}
	for (uint32_t i = 0; i < nV; i++) { s->w[i] = 1.0f / s->w[i]; }
	/* tOrder: Lehmer LCG "swap i with rand % n", Geo.cpp:759-769 */
	for (uint32_t i = 0; i < nT; i++) { s->tOrder[i] = i; }
	uint32_t randState = 1;
	for (uint32_t i = 0; i < nT; i++) {
		randState = (uint32_t)(((uint64_t)randState * 48271u) % 0x7fffffffu);
		uint32_t j = randState % nT;
		uint32_t c = s->tOrder[i]; s->tOrder[i] = s->tOrder[j]; s->tOrder[j] = c;
	}
	return s;
}

void xo_destroy(xo_scene* s) {
	if (!s) { return; }
	free(s->X0); free(s->X); free(s->O); free(s->V); free(s->w); free(s->flags); free(s->t); free(s->tOrder);
	free(s);
}
uint32_t xo_vert_count(const xo_scene* s) { return s->nV; }
uint32_t xo_tet_count(const xo_scene* s) { return s->nT; }
void xo_get_order(const xo_scene* s, uint32_t* order) { memcpy(order, s->tOrder, sizeof(uint32_t) * s->nT); }
void xo_set_order(xo_scene* s, const uint32_t* order) { memcpy(s->tOrder, order, sizeof(uint32_t) * s->nT); }
void xo_get_state(const xo_scene* s, double* X, double* V, float* w) {
	if (X) { memcpy(X, s->X, sizeof(double) * 3 * s->nV); }
	if (V) { memcpy(V, s->V, sizeof(double) * 3 * s->nV); }
	if (w) { memcpy(w, s->w, sizeof(float) * s->nV); }
}
void xo_set_state(xo_scene* s, const double* X, const double* V, const float* w) {
	if (X) { memcpy(s->X, X, sizeof(double) * 3 * s->nV); }
	if (V) { memcpy(s->V, V, sizeof(double) * 3 * s->nV); }
	if (w) { memcpy(s->w, w, sizeof(float) * s->nV); }
}
void xo_get_rest(const xo_scene* s, double* X0, double* O, uint8_t* flags) {
	if (X0) { memcpy(X0, s->X0, sizeof(double) * 3 * s->nV); }
	if (O) { memcpy(O, s->O, sizeof(double) * 3 * s->nV); }
	if (flags) { memcpy(flags, s->flags, s->nV); }
}
void xo_get_elements(const xo_scene* s, uint32_t* idx4, float* Qi9, float* QQ3, float* QR3, float* volume, float* area) {
	for (uint32_t e = 0; e < s->nT; e++) {
		const xo_tet* t = &s->t[e];
		for (int j = 0; j < 4; j++) { if (idx4) { idx4[4 * e + j] = t->i[j]; } }
		for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { if (Qi9) { Qi9[9 * e + 3 * c + r] = t->Qi[c][r]; } } }
		for (int j = 0; j < 3; j++) { if (QQ3) { QQ3[3 * e + j] = t->QQ[j]; } if (QR3) { QR3[3 * e + j] = t->QR[j]; } }
		if (volume) { volume[e] = t->volume; }
		if (area) { area[e] = t->surfaceArea; }
	}
}
/* Test plumbing for sub-meshes of a larger scene (tests/part_worker.py): element constants and rest positions taken from the
 * scene they were cut from instead of being re-derived from rounded coordinates. */
void xo_set_elements(xo_scene* s, const float* Qi9, const float* QQ3, const float* QR3, const float* volume, const float* area) {
	for (uint32_t e = 0; e < s->nT; e++) {
		xo_tet* t = &s->t[e];
		for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { t->Qi[c][r] = Qi9[9 * e + 3 * c + r]; } }
		for (int j = 0; j < 3; j++) { t->QQ[j] = QQ3[3 * e + j]; t->QR[j] = QR3[3 * e + j]; }
		t->volume = volume[e];
		t->surfaceArea = area[e];
	}
}
void xo_set_rest(xo_scene* s, const double* X0) {
	memcpy(s->X0, X0, sizeof(double) * 3 * s->nV);
	memcpy(s->X, X0, sizeof(double) * 3 * s->nV);
	memcpy(s->O, X0, sizeof(double) * 3 * s->nV);
}
void xo_set_flags(xo_scene* s, const uint8_t* flags) { memcpy(s->flags, flags, s->nV); }
void xo_get_origin(const xo_scene* s, float* o) { o[0] = s->origin[0]; o[1] = s->origin[1]; o[2] = s->origin[2]; }

/* ------------------------------------------------------------------------------------------
 * Per-element energies and gradients.  g[c][n] is the gradient of constraint c at node n.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float mu, lambda, a, comp[2], damp[2]; } xo_material;

/* Fem.cpp:445-450 */
static xo_material material_of(const xo_settings* st, const xo_tet* t) {
	xo_material m;
	m.mu = 1.0f / st->compliance;
	m.lambda = (2.0f * m.mu * st->poissonsRatio) / (1.0f - 2.0f * st->poissonsRatio);
	m.a = 1.0f + m.mu / m.lambda;
	m.comp[0] = 1.0f / m.mu / t->volume;
	m.comp[1] = 1.0f / m.lambda / t->volume;
	m.damp[0] = st->damping;
	m.damp[1] = st->damping;
	return m;
}

/* F = Qi * (sum_n e_n (x) P[n]), Fem.cpp:311-323 + 86-94: F[c][r] = sum_n Qi[n][r] * P[n][c] */
static void deformation_gradient(const xo_tet* t, float P[3][3], float F[3][3]) {
	float M[3][3];
	for (int c = 0; c < 3; c++) { for (int n = 0; n < 3; n++) { M[c][n] = P[n][c]; } }
	mat_mat((float(*)[3])t->Qi, M, F);
}

/* g[n] += (G * Qi) * e_n for n < 3, Fem.cpp:338-354 */
static void apply_gradient(const xo_tet* t, float G[3][3], float g[4][3]) {
	float GQ[3][3];
	mat_mat(G, (float(*)[3])t->Qi, GQ);
	for (int n = 0; n < 3; n++) { for (int k = 0; k < 3; k++) { g[n][k] += GQ[n][k]; } }
}

/* g[3] = ((0 - g[0]) - g[1]) - g[2], Fem.cpp:502 / 507 / 547 */
static void close_gradient(float g[4][3]) {
	for (int n = 0; n < 3; n++) { for (int k = 0; k < 3; k++) { g[3][k] -= g[n][k]; } }
}

/* volumetric term U1 = (J - a)^2 and its gradient, Fem.cpp:479-485 (weight == 1.0f for T4) */
static void volumetric_term(const xo_tet* t, float P[3][3], float a, float* U1, float g1[4][3]) {
	const float weight = 1.0f;
	float F[3][3], adjF[3][3], GJ[3][3];
	deformation_gradient(t, P, F);
	adjugate3(F, adjF);
	float c0[3] = { adjF[0][0], adjF[1][0], adjF[2][0] };
	float J = dot3(F[0], c0); /* determinantFromAdjugate, Fem.cpp:304-306 */
	*U1 += weight * ((J - a) * (J - a));
	float s = weight * 2.0f * (J - a);
	for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { GJ[c][r] = adjF[c][r] * s; } }
	apply_gradient(t, GJ, g1);
}

/* CalculateIncompressibleNeoHookeanEnergyAndGradients<vec3,4>, Fem.cpp:163-192 */
static void prefactored_i1(const xo_tet* t, float P[3][3], float* U, float g[4][3]) {
	float u = 0.0f;
	for (int i = 0; i < 3; i++) {
		u += t->QQ[i] * dot3(P[i], P[i]);
		float s = 2.0f * t->QQ[i];
		for (int k = 0; k < 3; k++) { g[i][k] = s * P[i][k]; }
	}
	int kk = 0;
	for (int i = 0; i < 3; i++) {
		for (int j = i + 1; j < 3; j++) {
			u += t->QR[kk] * dot3(P[i], P[j]);
			for (int k = 0; k < 3; k++) { g[i][k] = g[i][k] + t->QR[kk] * P[j][k]; }
			for (int k = 0; k < 3; k++) { g[j][k] = g[j][k] + t->QR[kk] * P[i][k]; }
			++kk;
		}
	}
	for (int k = 0; k < 3; k++) { g[3][k] = 0.0f; }
	close_gradient(g);
	*U = u;
}

/* EnergyXpbdConstrain, Xpbd.h:86-120 */
static void energy_constrain(double* X, const double* O, const float* w, const uint32_t* is, float U, float g[4][3],
                             float dt, float compliance, float dampingGamma, uint32_t dampingType) {
	float alpha = compliance / (dt * dt);
	float gamma = dampingGamma / dt;
	float lambdaPrime;
	if (dampingGamma > 0.0f && dampingType < RAY_POST) {
		float wgg = 1.0e-22f, gV = 0.0f;
		for (int i = 0; i < 4; i++) {
			wgg += w[is[i]] * dot3(g[i], g[i]);
			float d[3];
			for (int k = 0; k < 3; k++) { d[k] = (float)(X[3 * (size_t)is[i] + k] - O[3 * (size_t)is[i] + k]); }
			gV += dot3(d, g[i]);
		}
		if (dampingType == RAY_PAPER) {
			float A = wgg * (1.0f + gamma) + (2.0f * U * alpha);
			lambdaPrime = (-2.0f * U - gamma * gV) / A;
		} else {
			lambdaPrime = (-2.0f * U) / (wgg + (2.0f * U * alpha));
			float invBeta = compliance / (dt * dampingGamma);
			float A = wgg + 2.0f * U * invBeta;
			float b = -gV - wgg * lambdaPrime;
			b *= A / fmaxf(A, 4.0f * wgg);
			lambdaPrime += b / A;
		}
	} else {
		float wgg = 1.0e-22f;
		for (int i = 0; i < 4; i++) { wgg += w[is[i]] * dot3(g[i], g[i]); }
		lambdaPrime = (-2.0f * U) / (wgg + (2.0f * U * alpha));
	}
	for (int i = 0; i < 4; i++) {
		float s = w[is[i]] * lambdaPrime;
		for (int k = 0; k < 3; k++) { X[3 * (size_t)is[i] + k] += (double)(s * g[i][k]); }
	}
}

/* 2x2 Cramer, Xpbd.h:173-179 */
static void cramer2(const float* A, const float* b, float* out) {
	float invA00 = 1.0f / A[0];
	float invA11 = 1.0f / A[2];
	float invDet = 1.0f / fmaxf(0.00000001f, 1.0f - (A[1] * invA00) * (A[1] * invA11));
	out[0] = invDet * ((b[0] * invA00) - (A[1] * invA00) * (b[1] * invA11));
	out[1] = invDet * ((b[1] * invA11) - (b[0] * invA00) * (A[1] * invA11));
}

/* EnergyXpbdConstrainSimultaneous<.., 2>, Xpbd.h:122-214 */
static void energy_constrain_simultaneous(double* X, const double* O, const float* w, const uint32_t* is, const float* U,
                                          float g[2][4][3], float dt, const float* compliance, const float* dampingGamma,
                                          uint32_t dampingType) {
	float alpha[2], gamma[2];
	int anyDamping = 0;
	for (int i = 0; i < 2; i++) {
		alpha[i] = compliance[i] / (dt * dt);
		gamma[i] = dampingGamma[i] / dt;
		anyDamping = anyDamping || dampingGamma[i] > 0.0f;
	}
	float gV[2] = { 0.0f, 0.0f };
	if (anyDamping && dampingType < RAY_POST) {
		for (int n = 0; n < 4; n++) {
			float v[3];
			for (int k = 0; k < 3; k++) { v[k] = (float)(X[3 * (size_t)is[n] + k] - O[3 * (size_t)is[n] + k]); }
			for (int i = 0; i < 2; i++) { gV[i] += dot3(g[i][n], v); }
		}
	}
	float wgg[3], A[3], b[2];
	int k = 0;
	for (int i = 0; i < 2; i++) {
		for (int j = 0; j <= i; j++) {
			wgg[k] = 1.0e-22f;
			for (int n = 0; n < 4; n++) { wgg[k] += w[is[n]] * dot3(g[i][n], g[j][n]); }
			A[k] = wgg[k];
			++k;
		}
		if (dampingType == RAY_PAPER && gamma[i] > 0.0f) {
			A[k - 1] += (2.0f * U[i] * alpha[i]) / (1.0f + gamma[i]);
			b[i] = (-2.0f * U[i] - gamma[i] * gV[i]) / (1.0f + gamma[i]);
		} else {
			A[k - 1] += 2.0f * U[i] * alpha[i];
			b[i] = -2.0f * U[i];
		}
	}
	float lambdaPrime[2];
	cramer2(A, b, lambdaPrime);
	if (dampingType == RAY_LIMIT && gamma[0] > 0.0f) {
		int kk = 0;
		for (int i = 0; i < 2; i++) {
			kk += i + 1;
			float invBeta = compliance[i] / (dt * dampingGamma[i]);
			A[kk - 1] = wgg[kk - 1] + 2.0f * U[i] * invBeta;
			b[i] *= (-1.0f - alpha[i] * lambdaPrime[i]);
			b[i] += -gV[i];
			b[i] *= A[kk - 1] / fmaxf(A[kk - 1], 8.0f * wgg[kk - 1]);
		}
		float extra[2];
		cramer2(A, b, extra);
		lambdaPrime[0] += extra[0];
		lambdaPrime[1] += extra[1];
	}
	for (int n = 0; n < 4; n++) {
		float acc[3] = { 0.0f, 0.0f, 0.0f };
		for (int j = 0; j < 2; j++) { for (int c = 0; c < 3; c++) { acc[c] = acc[c] + lambdaPrime[j] * g[j][n][c]; } }
		float wn = w[is[n]];
		for (int c = 0; c < 3; c++) { X[3 * (size_t)is[n] + c] += (double)(wn * acc[c]); }
	}
}

/* RayleighDamp (one constraint), Xpbd.h:216-229 */
static void rayleigh_damp1(double* V, const float* w, const uint32_t* is, float U, float g[4][3], float dt, float compliance,
                           float dampingGamma) {
	float invBeta = compliance / (dt * dampingGamma);
	float wgg = 1.0e-22f, gV = 0.0f;
	for (int n = 0; n < 4; n++) {
		wgg += w[is[n]] * dot3(g[n], g[n]);
		float v[3] = { (float)V[3 * (size_t)is[n]], (float)V[3 * (size_t)is[n] + 1], (float)V[3 * (size_t)is[n] + 2] };
		gV += dot3(g[n], v);
	}
	float lambda = -gV / (2.0f * U * invBeta + wgg);
	for (int n = 0; n < 4; n++) {
		float s = w[is[n]] * lambda;
		for (int k = 0; k < 3; k++) { V[3 * (size_t)is[n] + k] += (double)(s * g[n][k]); }
	}
}

/* RayleighDamp (two coupled constraints), Xpbd.h:231-263 */
static void rayleigh_damp2(double* V, const float* w, const uint32_t* is, const float* U, float g[2][4][3], float dt,
                           const float* compliance, const float* dampingGamma) {
	float invBeta[2] = { compliance[0] / (dt * dampingGamma[0]), compliance[1] / (dt * dampingGamma[1]) };
	float A[3] = { 1.0e-22f, 1.0e-22f, 1.0e-22f };
	float b[2] = { 0.0f, 0.0f };
	for (int n = 0; n < 4; n++) {
		float wn = w[is[n]];
		A[0] += wn * dot3(g[0][n], g[0][n]);
		A[1] += wn * dot3(g[0][n], g[1][n]);
		A[2] += wn * dot3(g[1][n], g[1][n]);
		float v[3] = { (float)V[3 * (size_t)is[n]], (float)V[3 * (size_t)is[n] + 1], (float)V[3 * (size_t)is[n] + 2] };
		b[0] -= dot3(g[0][n], v);
		b[1] -= dot3(g[1][n], v);
	}
	A[0] += 2.0f * U[0] * invBeta[0];
	A[2] += 2.0f * U[1] * invBeta[1];
	float lp[2];
	cramer2(A, b, lp);
	for (int n = 0; n < 4; n++) {
		float wn = w[is[n]];
		for (int k = 0; k < 3; k++) {
			float acc = lp[0] * g[0][n][k] + lp[1] * g[1][n][k];
			V[3 * (size_t)is[n] + k] += (double)(wn * acc);
		}
	}
}

/* SolveElementMixed<Options> for Nodes == 4, Fem.cpp:437-564.  damping != 0 selects
 * DampingMode::On (velocity-space Rayleigh damping; V is written, X is not). */
static void solve_element_mixed(uint32_t energy, int dampingOn, float dt, double* X, const double* O, double* V, const float* w,
                                const xo_tet* t, const xo_settings* st) {
	uint32_t rayleighType = (st->flags >> XO_RAYLEIGH_BIT) & XO_RAYLEIGH_MASK;
	int simultaneous = (st->flags & (1u << XO_SOLVE_BIT)) != 0;
	if (dampingOn) { simultaneous = 1; }
	xo_material m = material_of(st, t);
	const uint32_t* is = t->i;

	float P[3][3];
	gather_edges(X, is, P);
	float U[2] = { 0.0f, 0.0f };
	float g[2][4][3];
	memset(g, 0, sizeof(g));

	if (energy == EN_MIXED || energy == EN_YEOH) {
		const float weight = 1.0f;
		float F[3][3];
		deformation_gradient(t, P, F);
		float I1 = dot3(F[0], F[0]) + dot3(F[1], F[1]) + dot3(F[2], F[2]); /* traceXTX, Fem.cpp:309 */
		float G[3][3];
		float scale;
		if (energy == EN_MIXED) {
			U[0] += weight * I1;
			scale = weight * 2.0f;
		} else {
			const float C[3] = { 0.1095f, 14.95f, 4.595f };
			float IM = I1 - 3.0f;
			U[0] += weight * (C[0] * IM + C[1] * IM * IM + C[2] * IM * IM * IM);
			scale = weight * (C[0] + 2.0f * C[1] * IM + 3.0f * C[2] * IM * IM) * 2.0f;
		}
		for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { G[c][r] = F[r][c] * scale; } } /* scale * transpose(F) */
		apply_gradient(t, G, g[0]);
		if (simultaneous) {
			/* same F; the reference recomputes adj/J from it inside the same quadrature loop */
			float adjF[3][3], GJ[3][3];
			adjugate3(F, adjF);
			float c0[3] = { adjF[0][0], adjF[1][0], adjF[2][0] };
			float J = dot3(F[0], c0);
			U[1] += weight * ((J - m.a) * (J - m.a));
			float s = weight * 2.0f * (J - m.a);
			for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { GJ[c][r] = adjF[c][r] * s; } }
			apply_gradient(t, GJ, g[1]);
		}
		U[0] = fmaxf(0.0001f, U[0]);
		if (simultaneous) {
			for (int n = 0; n < 3; n++) { for (int k = 0; k < 3; k++) { g[0][3][k] -= g[0][n][k]; g[1][3][k] -= g[1][n][k]; } }
			if (!dampingOn) { energy_constrain_simultaneous(X, O, w, is, U, g, dt, m.comp, m.damp, rayleighType); }
		} else {
			close_gradient(g[0]);
			energy_constrain(X, O, w, is, U[0], g[0], dt, m.comp[0], m.damp[0], rayleighType);
			gather_edges(X, is, P);
			volumetric_term(t, P, m.a, &U[1], g[1]);
			close_gradient(g[1]);
			energy_constrain(X, O, w, is, U[1], g[1], dt, m.comp[1], m.damp[1], rayleighType);
		}
	} else { /* MixedSelective / YeohSkinFast (and the reference's default case), Fem.cpp:523-552 */
		prefactored_i1(t, P, &U[0], g[0]);
		if (energy == EN_YEOH_FAST) {
			const float C[3] = { 0.1095f, 14.95f, 4.595f };
			float IM = U[0] - 3.0f;
			U[0] = fmaxf(0.0001f, C[0] * IM + C[1] * IM * IM + C[2] * IM * IM * IM);
			float gScale = C[0] + 2.0f * C[1] * IM + 3.0f * C[2] * IM * IM;
			for (int n = 0; n < 4; n++) { for (int k = 0; k < 3; k++) { g[0][n][k] *= gScale; } }
		}
		if (!simultaneous) {
			energy_constrain(X, O, w, is, U[0], g[0], dt, m.comp[0], m.damp[0], rayleighType);
			gather_edges(X, is, P);
		}
		volumetric_term(t, P, m.a, &U[1], g[1]);
		close_gradient(g[1]);
		if (!simultaneous) {
			energy_constrain(X, O, w, is, U[1], g[1], dt, m.comp[1], m.damp[1], rayleighType);
		} else if (!dampingOn) {
			energy_constrain_simultaneous(X, O, w, is, U, g, dt, m.comp, m.damp, rayleighType);
		}
	}

	if (dampingOn) { /* Fem.cpp:555-563 */
		int dampSimultaneous = (st->flags & (1u << XO_SOLVE_BIT)) != 0;
		if (dampSimultaneous) {
			rayleigh_damp2(V, w, is, U, g, dt, m.comp, m.damp);
		} else {
			rayleigh_damp1(V, w, is, U[0], g[0], dt, m.comp[0], m.damp[0]);
			rayleigh_damp1(V, w, is, U[1], g[1], dt, m.comp[1], m.damp[1]);
		}
	}
}

static int energy_supported(uint32_t energy) {
	return energy == EN_MIXED || energy == EN_MIXED_SEL || energy == EN_YEOH || energy == EN_YEOH_FAST;
}

/* SolveVolumeOnly, Fem.cpp:840-867 */
static void solve_volume_only(float dt, double* X, const double* O, const float* w, const xo_tet* t, const xo_settings* st) {
	float comp = st->compliance * t->volume;
	float P[3][3];
	gather_edges(X, t->i, P);
	float U = 0.0f;
	float g[4][3];
	memset(g, 0, sizeof(g));
	const float weight = 1.0f;
	float F[3][3], adjF[3][3], GJ[3][3];
	deformation_gradient(t, P, F);
	adjugate3(F, adjF);
	float c0[3] = { adjF[0][0], adjF[1][0], adjF[2][0] };
	float J = dot3(F[0], c0);
	U += weight * (J - 1.0f) * (J - 1.0f);
	float s = weight * 2.0f * (J - 1.0f);
	for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { GJ[c][r] = adjF[c][r] * s; } }
	apply_gradient(t, GJ, g);
	close_gradient(g);
	energy_constrain(X, O, w, t->i, U, g, dt, comp, 0.0f, RAY_POST);
}

/* PbdDamp<4> for dvec3, Xpbd.h:309-348 */
static void pbd_damp(const double* Xd, double* Vd, const float* W, const uint32_t* is, float damping) {
	float X[4][3], V[4][3], M[4];
	float Xcm[3] = { 0, 0, 0 }, Vcm[3] = { 0, 0, 0 };
	float Msum = 0.0f;
	float Wsum = 1.0e-8f;
	for (int i = 0; i < 4; i++) { Wsum += W[is[i]]; }
	float Waverage = Wsum * (1.0f / (float)4);
	float Wmin = 0.0001f * Waverage;
	for (int i = 0; i < 4; i++) {
		for (int k = 0; k < 3; k++) {
			X[i][k] = (float)(Xd[3 * (size_t)is[i] + k] - Xd[3 * (size_t)is[3] + k]);
			V[i][k] = (float)Vd[3 * (size_t)is[i] + k];
		}
		M[i] = Waverage / fmaxf(Wmin, W[is[i]]);
		for (int k = 0; k < 3; k++) { Xcm[k] = Xcm[k] + X[i][k] * M[i]; Vcm[k] = Vcm[k] + V[i][k] * M[i]; }
		Msum += M[i];
	}
	for (int k = 0; k < 3; k++) { Xcm[k] = Xcm[k] / Msum; Vcm[k] = Vcm[k] / Msum; }
	float r[4][3], L[3] = { 0, 0, 0 };
	float I[3][3];
	memset(I, 0, sizeof(I));
	for (int i = 0; i < 4; i++) {
		for (int k = 0; k < 3; k++) { r[i][k] = X[i][k] - Xcm[k]; }
		float cr[3] = { r[i][1] * V[i][2] - V[i][1] * r[i][2], r[i][2] * V[i][0] - V[i][2] * r[i][0], r[i][0] * V[i][1] - V[i][0] * r[i][1] };
		for (int k = 0; k < 3; k++) { L[k] = L[k] + M[i] * cr[k]; }
		float rr[3] = { r[i][0] * r[i][0], r[i][1] * r[i][1], r[i][2] * r[i][2] };
		float rp[3] = { r[i][0] * r[i][1], r[i][1] * r[i][2], r[i][2] * r[i][0] }; /* r * r.yzx */
		float T[3][3] = { { rr[2] + rr[1], -rp[0], -rp[2] }, { -rp[0], rr[2] + rr[0], -rp[1] }, { -rp[2], -rp[1], rr[1] + rr[0] } };
		/* M[i] * mat3 == mat3 * M[i]; I += ... */
		for (int c = 0; c < 3; c++) { for (int q = 0; q < 3; q++) { I[c][q] = I[c][q] + T[c][q] * M[i]; } }
	}
	float Ii[3][3], wv[3];
	inverse3(I, Ii);
	mat_vec(Ii, L, wv);
	for (int i = 0; i < 4; i++) {
		float cr[3] = { wv[1] * r[i][2] - r[i][1] * wv[2], wv[2] * r[i][0] - r[i][2] * wv[0], wv[0] * r[i][1] - r[i][0] * wv[1] };
		for (int k = 0; k < 3; k++) {
			float dV = Vcm[k] + cr[k] - V[i][k];
			Vd[3 * (size_t)is[i] + k] += (double)(damping * dV);
		}
	}
}

/* Study mode only (not part of the reference's algorithm): what an fp32 {x, y, z} vertex record would keep. */
void xo_set_state_precision(xo_scene* s, int f32) { s->stateF32 = f32; }
/* mode bit 0: positions (the record elements share); bit 1: velocities and previous positions too */
static int rounds(const xo_scene* s, const double* A) { return A == s->X ? (s->stateF32 & 1) : (s->stateF32 & 2); }
static void round_tet(const xo_scene* s, double* A, const xo_tet* t) {
	if (!rounds(s, A)) { return; }
	for (int n = 0; n < 4; n++) {
		for (int k = 0; k < 3; k++) { size_t q = 3 * (size_t)t->i[n] + k; A[q] = (double)(float)A[q]; }
	}
}
static void round_all(const xo_scene* s, double* A) {
	if (!rounds(s, A)) { return; }
	for (size_t q = 0; q < 3 * (size_t)s->nV; q++) { A[q] = (double)(float)A[q]; }
}

/* GeoLinear3d::Constrain, Geo.cpp:774-788 (tets only) */
static void constrain(xo_scene* s, const xo_settings* st, float dt) {
	uint32_t energy = (st->flags >> XO_ENERGY_BIT) & XO_ENERGY_MASK;
	if (!energy_supported(energy)) { energy = EN_MIXED_SEL; }
	for (uint32_t i = 0; i < s->nT; i++) {
		solve_element_mixed(energy, 0, dt, s->X, s->O, NULL, s->w, &s->t[s->tOrder[i]], st);
		round_tet(s, s->X, &s->t[s->tOrder[i]]);
	}
	for (uint32_t itr = 0; itr < st->volumePasses; itr++) {
		for (uint32_t i = 0; i < s->nT; i++) {
			solve_volume_only(dt, s->X, s->O, s->w, &s->t[s->tOrder[i]], st);
			round_tet(s, s->X, &s->t[s->tOrder[i]]);
		}
	}
}

/* GeoLinear3d::Damp, Geo.cpp:790-811 (tets only) */
static void damp(xo_scene* s, const xo_settings* st, float dt) {
	uint32_t energy = (st->flags >> XO_ENERGY_BIT) & XO_ENERGY_MASK;
	if (!energy_supported(energy)) { energy = EN_MIXED_SEL; }
	uint32_t rayleighType = (st->flags >> XO_RAYLEIGH_BIT) & XO_RAYLEIGH_MASK;
	int amortized = rayleighType == RAY_POST_AMORTIZED;
	uint32_t begin = amortized ? s->nT * (st->tickId % XO_AMORTIZATION_PERIOD) / XO_AMORTIZATION_PERIOD : 0;
	uint32_t end = amortized ? s->nT * ((st->tickId % XO_AMORTIZATION_PERIOD) + 1) / XO_AMORTIZATION_PERIOD : s->nT;
	if (rayleighType >= RAY_POST) {
		for (uint32_t i = begin; i < end; i++) {
			if (st->damping <= 0.0f) { break; } /* TDampElement early-out, Fem.cpp:911 */
			solve_element_mixed(energy, 1, dt, s->X, NULL, s->V, s->w, &s->t[s->tOrder[i]], st);
			round_tet(s, s->V, &s->t[s->tOrder[i]]);
		}
	}
	if (st->pbdDamping > 0.0f) {
		for (uint32_t i = begin; i < end; i++) {
			const xo_tet* t = &s->t[s->tOrder[i]];
			pbd_damp(s->X, s->V, s->w, t->i, fminf(1.0f, st->volumeAndTimeCorrectedPbdDamping / t->surfaceArea));
			round_tet(s, s->V, t);
		}
	}
}

void xo_set_ground(xo_scene* s, int enabled, float y0, float friction) { s->groundOn = enabled; s->groundY = y0; s->groundFriction = friction; }
void xo_set_handles(xo_scene* s, uint32_t count, const uint32_t* vertIdx, const float* targetXYZ) {
	s->handleCount = count > 64 ? 64 : count;
	for (uint32_t k = 0; k < s->handleCount; k++) {
		s->handleIdx[k] = vertIdx[k];
		for (int j = 0; j < 3; j++) { s->handleTarget[k][j] = targetXYZ[3 * k + j]; }
	}
}

/* X[i] += dvec3((target - vec3(X[i])) * (w / (max(1e-6f, w) + 1.8f / (dt*dt)))), Geo.cpp:338 */
static void drag_vertex(xo_scene* s, uint32_t i, const float* target, float dt) {
	float wi = s->w[i];
	float k = wi / (fmaxf(0.000001f, wi) + 1.8f / (dt * dt));
	for (int c = 0; c < 3; c++) {
		float d = (target[c] - (float)s->X[3 * (size_t)i + c]) * k;
		s->X[3 * (size_t)i + c] += (double)d;
	}
}

/* ---- Geo3d::Substep, Geo.cpp:305-356, in three phases so tests can emulate a partitioned run ---- */
/* predict, Geo.cpp:307-312 */
void xo_phase_predict(xo_scene* s, const xo_settings* st, float dt) {
	double gdt[3] = { (double)(st->gravity[0] * dt), (double)(st->gravity[1] * dt), (double)0.0f };
	double keep = (double)(1.0f - st->timeCorrectedDrag);
	double ddt = (double)dt;
	for (uint32_t i = 0; i < s->nV; i++) {
		for (int k = 0; k < 3; k++) {
			size_t q = 3 * (size_t)i + k;
			s->V[q] = s->V[q] + gdt[k];
			s->V[q] = s->V[q] * keep;
			s->O[q] = s->X[q];
			s->X[q] = s->X[q] + s->V[q] * ddt;
		}
	}
}

/* elements tOrder[begin..end) of the main sweep of GeoLinear3d::Constrain, Geo.cpp:777 */
void xo_phase_sweep(xo_scene* s, const xo_settings* st, float dt, uint32_t begin, uint32_t end) {
	uint32_t energy = (st->flags >> XO_ENERGY_BIT) & XO_ENERGY_MASK;
	if (!energy_supported(energy)) { energy = EN_MIXED_SEL; }
	for (uint32_t i = begin; i < end && i < s->nT; i++) { solve_element_mixed(energy, 0, dt, s->X, s->O, NULL, s->w, &s->t[s->tOrder[i]], st); }
}

/* An explicit list of elements (indices into the scene's tets, in the given order) of one sweep, for emulating a partitioned run:
 * kind 0 = main solve (Geo.cpp:777), 1 = volume-only pass (Geo.cpp:779-782), 2 = Rayleigh damp, 3 = PBD damp (Geo.cpp:790-811, with
 * the amortised settings of Geo.cpp:346-355 when the Rayleigh type is PostAmortized). */
void xo_phase_elems(xo_scene* s, const xo_settings* stIn, float dt, int kind, const uint32_t* elems, uint32_t count) {
	xo_settings st = *stIn;
	uint32_t energy = (st.flags >> XO_ENERGY_BIT) & XO_ENERGY_MASK;
	if (!energy_supported(energy)) { energy = EN_MIXED_SEL; }
	uint32_t rayleighType = (st.flags >> XO_RAYLEIGH_BIT) & XO_RAYLEIGH_MASK;
	if (kind >= 2 && rayleighType == RAY_POST_AMORTIZED) {
		st.damping *= (float)XO_AMORTIZATION_PERIOD;
		st.volumeAndTimeCorrectedPbdDamping = stIn->amortizedVolumeAndTimeCorrectedPbdDamping;
	}
	for (uint32_t k = 0; k < count; k++) {
		const xo_tet* t = &s->t[elems[k]];
		if (kind == 0) { solve_element_mixed(energy, 0, dt, s->X, s->O, NULL, s->w, t, &st); }
		if (kind == 1) { solve_volume_only(dt, s->X, s->O, s->w, t, &st); }
		if (kind == 2 && rayleighType >= RAY_POST && st.damping > 0.0f) { solve_element_mixed(energy, 1, dt, s->X, NULL, s->V, s->w, t, &st); }
		if (kind == 3 && st.pbdDamping > 0.0f) { pbd_damp(s->X, s->V, s->w, t->i, fminf(1.0f, st.volumeAndTimeCorrectedPbdDamping / t->surfaceArea)); }
	}
}

/* ground (x1) -> locks -> manipulator -> handles (x2) -> velocity update, Geo.cpp:318-344 */
void xo_phase_post(xo_scene* s, const xo_settings* stp, const xo_manipulator* manip, float dt) {
	const xo_settings st = *stp;
	if (s->groundOn) {
		double y0 = (double)s->groundY;
		double keepT = (double)(1.0f - s->groundFriction);
		for (uint32_t i = 0; i < s->nV; i++) {
			size_t q = 3 * (size_t)i;
			if (s->X[q + 1] < y0) {
				s->X[q + 1] = y0;
				s->X[q + 0] = s->O[q + 0] + (s->X[q + 0] - s->O[q + 0]) * keepT;
				s->X[q + 2] = s->O[q + 2] + (s->X[q + 2] - s->O[q + 2]) * keepT;
			}
		}
	}
	if (st.flags & XO_LOCK_LEFT) {
		for (uint32_t i = 0; i < s->nV; i++) {
			if (s->flags[i] & XO_FLAG_LEFT) { for (int k = 0; k < 3; k++) { s->X[3 * (size_t)i + k] = s->O[3 * (size_t)i + k]; } s->w[i] = 0.0f; }
		}
	}
	if (st.flags & XO_LOCK_RIGHT) {
		for (uint32_t i = 0; i < s->nV; i++) {
			if (s->flags[i] & XO_FLAG_RIGHT) {
				float x0[3] = { (float)s->X0[3 * (size_t)i], (float)s->X0[3 * (size_t)i + 1], (float)s->X0[3 * (size_t)i + 2] };
				for (int r = 0; r < 3; r++) {
					/* mat3 * vec3 with 16-byte padded columns */
					float v = st.lockedRightTransform3d[0 + r] * x0[0] + st.lockedRightTransform3d[4 + r] * x0[1] + st.lockedRightTransform3d[8 + r] * x0[2];
					double p = (double)(s->origin[r] + v);
					s->X[3 * (size_t)i + r] = p;
					s->O[3 * (size_t)i + r] = p;
				}
				s->w[i] = 0.0f;
			}
		}
	}
	if (manip && manip->picked) {
		const float* nrm = manip->manipPlaneNormal;
		float d[3] = { manip->pick0[0] - manip->pos[0], manip->pick0[1] - manip->pos[1], manip->pick0[2] - manip->pos[2] };
		float t = dot3(nrm, d) / dot3(nrm, manip->pickDirTarget);
		float target[3];
		for (int c = 0; c < 3; c++) { target[c] = manip->pos[c] + t * manip->pickDirTarget[c]; }
		drag_vertex(s, manip->pickedPointIdx, target, dt);
	}
	for (uint32_t hd = 0; hd < s->handleCount; hd++) { drag_vertex(s, s->handleIdx[hd], s->handleTarget[hd], dt); }
	double invDt = (double)(1.0f / dt);
	for (uint32_t i = 0; i < s->nV; i++) {
		for (int k = 0; k < 3; k++) { size_t q = 3 * (size_t)i + k; s->V[q] = (s->X[q] - s->O[q]) * invDt; }
	}
}

/* Geo3d::Substep, Geo.cpp:305-356 */
void xo_substep(xo_scene* s, const xo_settings* settingsIn, const xo_manipulator* manip, float dt, uint32_t n) {
	xo_settings st = *settingsIn;
	for (uint32_t step = 0; step < n; step++) {
		xo_phase_predict(s, &st, dt);
		round_all(s, s->X); round_all(s, s->V); round_all(s, s->O);
		constrain(s, &st, dt);
		xo_phase_post(s, &st, manip, dt);
		round_all(s, s->X); round_all(s, s->V);
		/* damping, Geo.cpp:346-355 */
		uint32_t rayleighType = (st.flags >> XO_RAYLEIGH_BIT) & XO_RAYLEIGH_MASK;
		if (rayleighType == RAY_POST_AMORTIZED) {
			xo_settings am = st;
			am.damping *= (float)XO_AMORTIZATION_PERIOD;
			am.volumeAndTimeCorrectedPbdDamping = st.amortizedVolumeAndTimeCorrectedPbdDamping;
			damp(s, &am, dt);
		} else {
			damp(s, &st, dt);
		}
		st.tickId++;
	}
}

/* Geo3d::Transform, Geo.cpp:358-364: positions are narrowed to fp32, pushed through a mat4, widened */
void xo_transform(xo_scene* s, const float* m9) {
	/* t4 columns: (t3[0],0) (t3[1],0) (0,0,1,0) (t3[2].xy,0,1) */
	float c0[3] = { m9[0], m9[1], m9[2] }, c1[3] = { m9[3], m9[4], m9[5] };
	float c2[3] = { 0.0f, 0.0f, 1.0f }, c3[3] = { m9[6], m9[7], 0.0f };
	for (uint32_t i = 0; i < s->nV; i++) {
		float v[3] = { (float)s->X[3 * (size_t)i], (float)s->X[3 * (size_t)i + 1], (float)s->X[3 * (size_t)i + 2] };
		for (int r = 0; r < 3; r++) {
			float o = c0[r] * v[0] + c1[r] * v[1] + c2[r] * v[2] + c3[r] * 1.0f;
			s->X[3 * (size_t)i + r] = s->O[3 * (size_t)i + r] = (double)o;
		}
	}
	float o3[3] = { s->origin[0], s->origin[1], s->origin[2] };
	for (int r = 0; r < 3; r++) { s->origin[r] = c0[r] * o3[0] + c1[r] * o3[1] + c2[r] * o3[2] + c3[r] * 1.0f; }
}

/* CalculateElementVolume (T4) summed in element order, Fem.cpp:1067 + Geo.cpp:827-832 */
float xo_volume(const xo_scene* s) {
	float volume = 0.0f;
	for (uint32_t e = 0; e < s->nT; e++) {
		float P[3][3];
		gather_edges(s->X, s->t[e].i, P);
		volume += (1.0f / 6.0f) * det3(P); /* mat3(P0,P1,P2): columns are the edges */
	}
	return volume;
}

/* Energies for the long-trajectory drift statistic (defined by this repo):
 * E_c = U_c / comp_c is the potential XPBD's energy constraint minimises (Xpbd.h:86-120 derivation). */
void xo_energy(const xo_scene* s, const xo_settings* st, double* kinetic, double* gravitational, double* elasticDev, double* elasticVol) {
	double ke = 0.0, pe = 0.0, ed = 0.0, ev = 0.0;
	for (uint32_t i = 0; i < s->nV; i++) {
		if (s->w[i] <= 0.0f) { continue; }
		double m = 1.0 / (double)s->w[i];
		const double* v = s->V + 3 * (size_t)i;
		const double* x = s->X + 3 * (size_t)i;
		ke += 0.5 * m * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
		pe -= m * ((double)st->gravity[0] * x[0] + (double)st->gravity[1] * x[1]);
	}
	uint32_t energy = (st->flags >> XO_ENERGY_BIT) & XO_ENERGY_MASK;
	for (uint32_t e = 0; e < s->nT; e++) {
		const xo_tet* t = &s->t[e];
		xo_material m = material_of(st, t);
		float P[3][3], F[3][3], adjF[3][3];
		gather_edges(s->X, t->i, P);
		deformation_gradient(t, P, F);
		double I1 = (double)(dot3(F[0], F[0]) + dot3(F[1], F[1]) + dot3(F[2], F[2]));
		adjugate3(F, adjF);
		float c0[3] = { adjF[0][0], adjF[1][0], adjF[2][0] };
		double J = (double)dot3(F[0], c0);
		double U0 = I1 - 3.0;
		if (energy == EN_YEOH || energy == EN_YEOH_FAST) { double IM = I1 - 3.0; U0 = 0.1095 * IM + 14.95 * IM * IM + 4.595 * IM * IM * IM; }
		ed += U0 / (double)m.comp[0];
		if (m.comp[1] > 0.0f) { ev += (J - (double)m.a) * (J - (double)m.a) / (double)m.comp[1]; }
	}
	if (kinetic) { *kinetic = ke; }
	if (gravitational) { *gravitational = pe; }
	if (elasticDev) { *elasticDev = ed; }
	if (elasticVol) { *elasticVol = ev; }
}

double xo_time_substeps(xo_scene* s, const xo_settings* settings, float dt, uint32_t n) {
	struct timespec a, b;
	clock_gettime(CLOCK_MONOTONIC, &a);
	xo_substep(s, settings, NULL, dt, n);
	clock_gettime(CLOCK_MONOTONIC, &b);
	return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
