// Host-side preparation (runs once per scene, or once per xf_substep call for the parameter block):
//   * synthetic tet blocks                          (MeshGen.cpp:156-244)
//   * per-element constants, lumped masses, flags   (GeoLinear3d::Init Geo.cpp:697-744, InitTriTetFiniteElement Fem.cpp:196-224)
//   * conflict-free element colouring + the equivalent serial order
//   * the per-call constant block the kernels read  (Fem.cpp:445-450, Geo.cpp:305-356)
// Compile this file with -ffp-contract=off: the fp32 expressions below must round exactly like the
// reference's so that XF_PRECISION_EXACT is bit-identical end to end.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "xf_scene.h"

namespace xf {

namespace {

struct F3 {
	float x, y, z;
	float& operator[](int i) { return (&x)[i]; }
	float operator[](int i) const { return (&x)[i]; }
};
inline F3 operator-(F3 a, F3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline float Dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 Cross(F3 a, F3 b) { return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }

// 3x3 with the reference's mat3 convention: c[col][row].
struct M3 {
	F3 c[3];
};

// determinant(mat3), vectormath.cpp:34-39
inline float Det(const M3& m) {
	float a = m.c[0].x, b = m.c[1].x, cc = m.c[2].x;
	float d = m.c[0].y, e = m.c[1].y, f = m.c[2].y;
	float g = m.c[0].z, h = m.c[1].z, i = m.c[2].z;
	return a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
}

// inverse(mat3), vectormath.cpp:41-60: fp64 cofactors narrowed to fp32, fp32 determinant, scale by 1.0f/det
M3 Inverse(const M3& m) {
	const double m00 = m.c[0].x, m01 = m.c[0].y, m02 = m.c[0].z;
	const double m10 = m.c[1].x, m11 = m.c[1].y, m12 = m.c[1].z;
	const double m20 = m.c[2].x, m21 = m.c[2].y, m22 = m.c[2].z;
	M3 adj;
	adj.c[0] = { (float)(m11 * m22 - m12 * m21), (float)-(m01 * m22 - m02 * m21), (float)(m01 * m12 - m02 * m11) };
	adj.c[1] = { (float)-(m10 * m22 - m12 * m20), (float)(m00 * m22 - m02 * m20), (float)-(m00 * m12 - m02 * m10) };
	adj.c[2] = { (float)(m10 * m21 - m11 * m20), (float)-(m00 * m21 - m01 * m20), (float)(m00 * m11 - m01 * m10) };
	const float det = Dot(m.c[0], F3{ adj.c[0].x, adj.c[1].x, adj.c[2].x });
	const float s = 1.0f / det;
	M3 out;
	for (int k = 0; k < 3; k++) { out.c[k] = { adj.c[k].x * s, adj.c[k].y * s, adj.c[k].z * s }; }
	return out;
}

uint32_t Lehmer(uint32_t& state) {
	state = (uint32_t)(((uint64_t)state * 48271u) % 0x7fffffffu);
	return state;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Element constants for one linear tet from the rest positions.
// ------------------------------------------------------------------------------------------------
static void InitTet(const double* X, const uint32_t* is, float density, float* mass, float* Qi9, float* QQ3, float* QR3, float* volume,
                    float* area) {
	// P[n] = float(X[is[n]] - X[is[3]]); P[3] == 0
	F3 P[4];
	for (int n = 0; n < 4; n++) {
		for (int k = 0; k < 3; k++) { P[n][k] = (float)(X[3 * (size_t)is[n] + k] - X[3 * (size_t)is[3] + k]); }
	}
	// map Jacobian: J[col c][row r] = P[r][c]  (dN = e0,e1,e2,-1; the -1 term multiplies P[3] == 0)
	M3 J;
	for (int c = 0; c < 3; c++) { J.c[c] = { P[0][c], P[1][c], P[2][c] }; }
	const M3 Ji = Inverse(J);
	for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) { Qi9[3 * c + r] = Ji.c[c][r]; } }

	// prefactored I1 coefficients: 4-point rule, each point weighs det(J)/6 * 1/4, columns of Ji are the
	// mapped shape gradients (Fem.cpp:131-161)
	float volSum = 0.0f, qq[3] = { 0, 0, 0 }, qr[3] = { 0, 0, 0 };
	for (int pt = 0; pt < 4; pt++) {
		const float wgt = Det(J) * (1.0f / 6.0f) * 0.25f;
		volSum += wgt;
		for (int i = 0; i < 3; i++) { qq[i] += wgt * Dot(Ji.c[i], Ji.c[i]); }
		qr[0] += wgt * 2.0f * Dot(Ji.c[0], Ji.c[1]);
		qr[1] += wgt * 2.0f * Dot(Ji.c[0], Ji.c[2]);
		qr[2] += wgt * 2.0f * Dot(Ji.c[1], Ji.c[2]);
	}
	for (int i = 0; i < 3; i++) { QQ3[i] = qq[i] / volSum; QR3[i] = qr[i] / volSum; }

	*volume = (1.0f / 6.0f) * Det(J);
	auto faceArea = [&](int i0, int i1, int i2) {
		const F3 n = Cross(P[i0] - P[i2], P[i1] - P[i2]);
		return 0.5f * sqrtf(Dot(n, n));
	};
	*area = faceArea(0, 1, 2) + faceArea(1, 3, 2) + faceArea(0, 2, 3) + faceArea(3, 1, 0);
	for (int n = 0; n < 4; n++) { mass[is[n]] += (1.0f / 4.0f) * density * *volume; }
}

// ------------------------------------------------------------------------------------------------
// Colouring.  Two elements conflict when they share a vertex.  The schedule processes colour 0, 1, ...
// with a barrier in between and the elements of a colour in parallel, which equals the serial
// Gauss-Seidel sweep in the order (colour ascending, stream index ascending).
// ------------------------------------------------------------------------------------------------
namespace {

struct ColorSet { // which colours already touch a vertex
	uint64_t bits[kMaxColors / 64];
	ColorSet() { memset(bits, 0, sizeof(bits)); }
	void Add(uint32_t c) { bits[c >> 6] |= (uint64_t)1 << (c & 63); }
	bool Has(uint32_t c) const { return (bits[c >> 6] >> (c & 63)) & 1; }
};

int FirstFree(const ColorSet& a, const ColorSet& b, const ColorSet& c, const ColorSet& d) {
	for (int wd = 0; wd < kMaxColors / 64; wd++) {
		uint64_t used = a.bits[wd] | b.bits[wd] | c.bits[wd] | d.bits[wd];
		if (~used) { return wd * 64 + __builtin_ctzll(~used); }
	}
	return -1;
}

// First-fit greedy over `visit`; returns the colour count or -1.
int GreedyColor(const std::vector<uint32_t>& idx, uint32_t nV, const std::vector<uint32_t>& visit, std::vector<uint32_t>& color) {
	std::vector<ColorSet> seen(nV);
	int count = 0;
	for (uint32_t e : visit) {
		const uint32_t* v = &idx[4 * (size_t)e];
		int c = FirstFree(seen[v[0]], seen[v[1]], seen[v[2]], seen[v[3]]);
		if (c < 0) { return -1; }
		color[e] = (uint32_t)c;
		for (int j = 0; j < 4; j++) { seen[v[j]].Add((uint32_t)c); }
		count = std::max(count, c + 1);
	}
	return count;
}

bool ColoringIsValid(const std::vector<uint32_t>& idx, uint32_t nV, const uint32_t* color, uint32_t nT, uint32_t* outCount) {
	std::vector<ColorSet> seen(nV);
	uint32_t count = 0;
	for (uint32_t e = 0; e < nT; e++) {
		uint32_t c = color[e];
		if (c >= (uint32_t)kMaxColors) { return false; }
		for (int j = 0; j < 4; j++) {
			ColorSet& s = seen[idx[4 * (size_t)e + j]];
			if (s.Has(c)) { return false; }
			s.Add(c);
		}
		count = std::max(count, c + 1);
	}
	*outCount = count;
	return true;
}

// Generic colouring for arbitrary tet meshes: first-fit in stream order, then a few rounds of
// iterated greedy (revisit the elements class by class — never increases the colour count and usually
// shaves the ragged tail classes), keeping the best.
int ColorGeneric(const std::vector<uint32_t>& idx, uint32_t nV, uint32_t nT, std::vector<uint32_t>& color) {
	std::vector<uint32_t> visit(nT);
	std::iota(visit.begin(), visit.end(), 0u);
	color.assign(nT, 0);
	int best = GreedyColor(idx, nV, visit, color);
	if (best < 0) { return -1; }
	std::vector<uint32_t> trial(nT);
	std::vector<uint32_t> cur = color;
	int curCount = best;
	for (int round = 0; round < 6; round++) {
		std::vector<uint32_t> sizes(curCount, 0);
		for (uint32_t e = 0; e < nT; e++) { sizes[cur[e]]++; }
		std::vector<uint32_t> classOrder(curCount);
		std::iota(classOrder.begin(), classOrder.end(), 0u);
		if (round % 2 == 0) { std::reverse(classOrder.begin(), classOrder.end()); }
		else { std::stable_sort(classOrder.begin(), classOrder.end(), [&](uint32_t a, uint32_t b) { return sizes[a] > sizes[b]; }); }
		std::vector<uint32_t> rank(curCount);
		for (int k = 0; k < curCount; k++) { rank[classOrder[k]] = (uint32_t)k; }
		std::stable_sort(visit.begin(), visit.end(), [&](uint32_t a, uint32_t b) { return rank[cur[a]] < rank[cur[b]]; });
		int n = GreedyColor(idx, nV, visit, trial);
		if (n < 0) { break; }
		cur = trial;
		curCount = n;
		if (n < best) { best = n; color = trial; }
	}
	return best;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Clustered colouring.  The 24 (or more) elements around a vertex are mutually dependent, so a colour sweep is a chain of
// >= 24 element latencies per substep, each paid through L2.  If ONE thread solves a small cluster of elements that share
// vertices (MeshGen: the six Kuhn tets of a cell, 8 vertices) back to back, the dependences inside the cluster never
// leave the thread, and only clusters need colouring (8 classes on the lattice).  As a serial order this is
// (cluster colour, position in cluster, cluster), expressed for every other schedule as an ordinary colouring with
// colour = clusterColour * groupSize + position: two elements of one such colour sit in different clusters of one
// cluster colour, which share no vertex.
// Clusters: consecutive stream elements while they stay connected and touch <= 8 distinct vertices.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t kClusterVerts = 8, kClusterMaxElems = 8;

bool ColorClustered(HostMesh& m) {
	const uint32_t nT = m.nT;
	std::vector<uint32_t> clusterStart; // offsets into the stream
	{
		uint32_t verts[kClusterVerts];
		uint32_t nVerts = 0, count = 0;
		for (uint32_t e = 0; e < nT; e++) {
			const uint32_t* v = &m.idx[4 * (size_t)e];
			uint32_t fresh = 0, shared = 0;
			for (int j = 0; j < 4; j++) {
				bool found = false;
				for (uint32_t k = 0; k < nVerts; k++) { found = found || verts[k] == v[j]; }
				if (found) { shared++; } else { fresh++; }
			}
			if (count == 0 || count >= kClusterMaxElems || shared == 0 || nVerts + fresh > kClusterVerts) {
				clusterStart.push_back(e);
				nVerts = 0;
				count = 0;
			}
			for (int j = 0; j < 4; j++) {
				bool found = false;
				for (uint32_t k = 0; k < nVerts; k++) { found = found || verts[k] == v[j]; }
				if (!found) { verts[nVerts++] = v[j]; }
			}
			count++;
		}
		clusterStart.push_back(nT);
	}
	const uint32_t nClusters = (uint32_t)clusterStart.size() - 1;
	uint32_t G = 0;
	for (uint32_t c = 0; c < nClusters; c++) { G = std::max(G, clusterStart[c + 1] - clusterStart[c]); }
	if (G < 2) { return false; } // nothing to gain
	// first-fit colouring of the clusters
	std::vector<ColorSet> seen(m.nV);
	std::vector<uint32_t> cColor(nClusters);
	uint32_t nCC = 0;
	for (uint32_t c = 0; c < nClusters; c++) {
		ColorSet used;
		for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) {
			for (int j = 0; j < 4; j++) {
				const ColorSet& s = seen[m.idx[4 * (size_t)e + j]];
				for (int wd = 0; wd < kMaxColors / 64; wd++) { used.bits[wd] |= s.bits[wd]; }
			}
		}
		int col = -1;
		for (int wd = 0; wd < kMaxColors / 64 && col < 0; wd++) { if (~used.bits[wd]) { col = wd * 64 + __builtin_ctzll(~used.bits[wd]); } }
		if (col < 0) { return false; }
		cColor[c] = (uint32_t)col;
		nCC = std::max(nCC, (uint32_t)col + 1);
		for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) {
			for (int j = 0; j < 4; j++) { seen[m.idx[4 * (size_t)e + j]].Add((uint32_t)col); }
		}
	}
	if ((uint64_t)nCC * G > 254u) { return false; } // stage codes are bytes; ordinary colouring instead
	// clusters of one colour: larger first, so that "cluster k" means the same thing at every position t
	std::vector<uint32_t> byColor(nClusters);
	std::iota(byColor.begin(), byColor.end(), 0u);
	std::stable_sort(byColor.begin(), byColor.end(), [&](uint32_t a, uint32_t b) {
		if (cColor[a] != cColor[b]) { return cColor[a] < cColor[b]; }
		return clusterStart[a + 1] - clusterStart[a] > clusterStart[b + 1] - clusterStart[b];
	});
	const uint32_t nColors = nCC * G;
	m.color.assign(nT, 0);
	m.colorStart.assign(nColors + 1, 0);
	for (uint32_t c = 0; c < nClusters; c++) {
		for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) {
			m.color[e] = cColor[c] * G + (e - clusterStart[c]);
			m.colorStart[m.color[e] + 1]++;
		}
	}
	for (uint32_t c = 0; c < nColors; c++) { m.colorStart[c + 1] += m.colorStart[c]; }
	m.order.resize(nT);
	{
		std::vector<uint32_t> cursor(m.colorStart.begin(), m.colorStart.end() - 1);
		for (uint32_t k = 0; k < nClusters; k++) { // clusters in (colour, size-descending) order fill every position in that order
			const uint32_t c = byColor[k];
			for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) { m.order[cursor[m.color[e]]++] = e; }
		}
	}
	// slot / first / last bits
	m.clusterInfo.assign(nT, 0);
	for (uint32_t c = 0; c < nClusters; c++) {
		uint32_t verts[kClusterVerts], lastElem[kClusterVerts];
		uint32_t nVerts = 0;
		for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) {
			for (int j = 0; j < 4; j++) {
				const uint32_t v = m.idx[4 * (size_t)e + j];
				uint32_t slot = nVerts;
				for (uint32_t k = 0; k < nVerts; k++) { if (verts[k] == v) { slot = k; } }
				uint32_t bits = slot;
				if (slot == nVerts) { verts[nVerts++] = v; bits |= 8u; }
				lastElem[slot] = e;
				m.clusterInfo[e] |= bits << (5 * j);
			}
		}
		for (uint32_t e = clusterStart[c]; e < clusterStart[c + 1]; e++) {
			for (int j = 0; j < 4; j++) {
				const uint32_t slot = (m.clusterInfo[e] >> (5 * j)) & 7u;
				if (lastElem[slot] == e) { m.clusterInfo[e] |= 16u << (5 * j); }
			}
		}
	}
	m.groupSize = G;
	return true;
}
}  // namespace

int PrepareMesh(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, float density, bool autoResize,
                const uint32_t* colorHint, uint32_t colorHintCount, HostMesh* out, std::string* err, bool clustered) {
	if (!nodeXYZ || !idxStream || idxCount == 0 || idxCount % 5 != 0 || nodeFloatCount % 3 != 0) {
		*err = "mesh stream must be [4,v0,v1,v2,v3]* (CON_TET records, Connectivity.h:16) with xyz float triples";
		return XF_ERR_INVALID;
	}
	const uint32_t nT = idxCount / 5;
	uint32_t maxVert = 0;
	for (uint32_t e = 0; e < nT; e++) {
		if (idxStream[5 * (size_t)e] != 4) {
			*err = "only CON_TET (4) records are on the tet hot path; record " + std::to_string(e) + " has type " + std::to_string(idxStream[5 * (size_t)e]);
			return XF_ERR_UNSUPPORTED;
		}
		for (int j = 1; j <= 4; j++) { maxVert = std::max(maxVert, idxStream[5 * (size_t)e + j]); }
	}
	const uint32_t nV = maxVert + 1;
	if (nV != nodeFloatCount / 3) { // the reference asserts this (Geo.cpp:700)
		*err = "vertex count implied by the index stream (" + std::to_string(nV) + ") != node count (" + std::to_string(nodeFloatCount / 3) + ")";
		return XF_ERR_INVALID;
	}
	HostMesh& m = *out;
	m.nV = nV;
	m.nT = nT;
	m.X0.resize(3 * (size_t)nV);
	m.flags.assign(nV, (uint8_t)XF_VERT_PICKABLE);
	std::vector<float> mass(nV, 0.0f);

	// AABB in fp32, Left/Right flags at <2% / >98% of the x extent, optional resize (Geo.cpp:711-729)
	float lo[3] = { 1.0e24f, 1.0e24f, 1.0e24f }, hi[3] = { -1.0e24f, -1.0e24f, -1.0e24f };
	for (uint32_t i = 0; i < nV; i++) {
		for (int k = 0; k < 3; k++) {
			const float f = nodeXYZ[3 * (size_t)i + k];
			m.X0[3 * (size_t)i + k] = (double)f;
			lo[k] = fminf(lo[k], f);
			hi[k] = fmaxf(hi[k], f);
		}
	}
	const float ext[3] = { hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] };
	const double maxDim = (double)fmaxf(ext[0], fmaxf(ext[1], ext[2]));
	for (uint32_t i = 0; i < nV; i++) {
		const float rel = ((float)m.X0[3 * (size_t)i] - lo[0]) / (hi[0] - lo[0]);
		if (rel < 0.02f) { m.flags[i] |= XF_VERT_LEFT; }
		if (rel > 0.98f) { m.flags[i] |= XF_VERT_RIGHT; }
		if (autoResize) {
			for (int k = 0; k < 3; k++) {
				const double centre = 0.5 * (double)(lo[k] + hi[k]);
				m.X0[3 * (size_t)i + k] = 0.085 * ((m.X0[3 * (size_t)i + k] - centre) / maxDim);
			}
		}
	}

	m.idx.resize(4 * (size_t)nT);
	m.Qi.resize(9 * (size_t)nT);
	m.QQ.resize(3 * (size_t)nT);
	m.QR.resize(3 * (size_t)nT);
	m.volume.resize(nT);
	m.area.resize(nT);
	for (uint32_t e = 0; e < nT; e++) {
		for (int j = 0; j < 4; j++) { m.idx[4 * (size_t)e + j] = idxStream[5 * (size_t)e + 1 + j]; }
		InitTet(m.X0.data(), &m.idx[4 * (size_t)e], density, mass.data(), &m.Qi[9 * (size_t)e], &m.QQ[3 * (size_t)e], &m.QR[3 * (size_t)e],
		        &m.volume[e], &m.area[e]);
	}
	m.w.resize(nV);
	for (uint32_t i = 0; i < nV; i++) { m.w[i] = 1.0f / mass[i]; }

	// colouring
	if (clustered && ColorClustered(m)) { return XF_OK; }
	m.groupSize = 0;
	uint32_t nColors = 0;
	bool hinted = false;
	if (colorHint && colorHintCount == nT && colorHint[0] != 0xffffffffu) {
		if (!ColoringIsValid(m.idx, nV, colorHint, nT, &nColors)) {
			*err = "colorHint is not a conflict-free colouring of the mesh (two elements of one colour share a vertex, or colour >= " +
			       std::to_string(kMaxColors) + ")";
			return XF_ERR_COLORING;
		}
		m.color.assign(colorHint, colorHint + nT);
		hinted = true;
	} else if (colorHint && colorHintCount != 0 && colorHintCount != nT) {
		*err = "colorHintCount must equal the element count";
		return XF_ERR_INVALID;
	}
	if (!hinted) {
		int n = ColorGeneric(m.idx, nV, nT, m.color);
		if (n < 0) {
			*err = "mesh needs more than " + std::to_string(kMaxColors) + " colours (vertex valence too high)";
			return XF_ERR_COLORING;
		}
		nColors = (uint32_t)n;
	}
	// drop empty classes, then order = (colour, stream index)
	{
		std::vector<uint32_t> sizes(nColors, 0);
		for (uint32_t e = 0; e < nT; e++) { sizes[m.color[e]]++; }
		std::vector<uint32_t> remap(nColors, 0);
		uint32_t used = 0;
		for (uint32_t c = 0; c < nColors; c++) { if (sizes[c]) { remap[c] = used++; } }
		for (uint32_t e = 0; e < nT; e++) { m.color[e] = remap[m.color[e]]; }
		nColors = used;
	}
	m.colorStart.assign(nColors + 1, 0);
	for (uint32_t e = 0; e < nT; e++) { m.colorStart[m.color[e] + 1]++; }
	for (uint32_t c = 0; c < nColors; c++) { m.colorStart[c + 1] += m.colorStart[c]; }
	m.order.resize(nT);
	{
		std::vector<uint32_t> cursor(m.colorStart.begin(), m.colorStart.end() - 1);
		for (uint32_t e = 0; e < nT; e++) { m.order[cursor[m.color[e]]++] = e; }
	}
	return XF_OK;
}

// Damping sweeps on the barrier-free schedule (xf_dataflow_general.cu): V records are versioned by a write count.  Rank of every element
// (serial position k) among the elements around each of its corners' vertices, 4 x 8 bits, and per vertex how many of those elements lie
// below each boundary nT*q/8, q = 1..8, of the amortised damping slices (Geo.cpp:794-797), 8 x 8 bits (byte 7 = the valence).
void DampingCodes(uint32_t nT, uint32_t nV, const uint32_t* idxOfPos, std::vector<uint32_t>* rankOut, std::vector<uint64_t>* belowOut) {
	std::vector<uint32_t>& rank = *rankOut;
	std::vector<uint64_t>& below = *belowOut;
	rank.assign(nT, 0);
	below.assign(nV, 0);
	std::vector<uint8_t> seen(nV, 0);
	uint32_t bound[XF_AMORTIZATION_PERIOD + 1];
	for (uint32_t q = 0; q <= XF_AMORTIZATION_PERIOD; q++) { bound[q] = (uint32_t)(((uint64_t)nT * q) / XF_AMORTIZATION_PERIOD); }
	for (uint32_t k = 0, q0 = 1; k < nT; k++) {
		while (q0 < XF_AMORTIZATION_PERIOD && k >= bound[q0]) { q0++; } // first boundary above position k
		for (int j = 0; j < 4; j++) {
			const uint32_t v = idxOfPos[4 * (size_t)k + j];
			rank[k] |= (uint32_t)seen[v]++ << (8 * j);
			for (uint32_t q = q0; q <= XF_AMORTIZATION_PERIOD; q++) { below[v] += 1ull << (8 * (q - 1)); }
		}
	}
}

void StageCodes(const HostMesh& m, const std::vector<uint32_t>& order, std::vector<uint8_t>* pred, std::vector<uint8_t>* last) {
	pred->assign(4 * (size_t)m.nT, 0);
	last->assign(m.nV, 0);
	for (uint32_t k = 0; k < m.nT; k++) { // colour-major: the elements around a vertex are met in colour order
		const uint32_t e = order[k];
		const uint8_t code = (uint8_t)(1u + m.color[e]);
		for (int j = 0; j < 4; j++) {
			uint8_t& l = (*last)[m.idx[4 * (size_t)e + j]];
			(*pred)[4 * (size_t)k + j] = l;
			l = code;
		}
	}
}

// Chained sweep (k_substeps_chain, xf_dataflow.cu).  The thread that runs position j of colour c also runs position j of
// the next colour.  When the element it meets there shares a vertex with the one it just solved, and nobody else writes
// that vertex in between (the vertex's previous-writer code IS the colour just solved), the record need not travel
// through L2: it stays in one of the thread's four private shared-memory slots.  info (one word per device position,
// same layout as HostMesh::clusterInfo): per corner n, bits [5n, 5n+5) = slot | first << 3 | last << 4, first = gather
// from L2 (and wait for the stage tag), last = scatter to L2.  On a MeshGen lattice with the ring-ordered hint a cell's
// six tets cost 9 gathers + 9 scatters instead of 24 + 24.  Returns the number of corner uses served from a slot.
uint64_t ChainInfo(const HostMesh& m, const std::vector<uint32_t>& order, const std::vector<uint8_t>& pred, std::vector<uint32_t>* info) {
	const uint32_t nC = (uint32_t)m.colorStart.size() - 1;
	info->assign(m.nT, 0);
	uint32_t maxCount = 0;
	for (uint32_t c = 0; c < nC; c++) { maxCount = std::max(maxCount, m.colorStart[c + 1] - m.colorStart[c]); }
	uint64_t held = 0;
	for (uint32_t j = 0; j < maxCount; j++) {
		bool havePrev = false;
		uint32_t prevPos = 0, prevColor = 0, prevVerts[4] = { 0, 0, 0, 0 }, prevSlots[4] = { 0, 0, 0, 0 };
		for (uint32_t c = 0; c < nC; c++) {
			if (j >= m.colorStart[c + 1] - m.colorStart[c]) { continue; }
			const uint32_t pos = m.colorStart[c] + j;
			const uint32_t* v = &m.idx[4 * (size_t)order[pos]];
			uint32_t slot[4], usedSlots = 0;
			bool first[4];
			for (int n = 0; n < 4; n++) {
				first[n] = true;
				if (havePrev && pred[4 * (size_t)pos + n] == (uint8_t)(1u + prevColor)) {
					for (int i = 0; i < 4; i++) {
						if (prevVerts[i] != v[n]) { continue; }
						first[n] = false;
						slot[n] = prevSlots[i];
						usedSlots |= 1u << slot[n];
						(*info)[prevPos] &= ~(16u << (5 * i)); // the previous element keeps it in the slot
						held++;
					}
				}
			}
			for (int n = 0; n < 4; n++) {
				if (!first[n]) { continue; }
				uint32_t f = 0;
				while (usedSlots & (1u << f)) { f++; }
				slot[n] = f;
				usedSlots |= 1u << f;
			}
			uint32_t word = 0;
			for (int n = 0; n < 4; n++) { word |= (slot[n] | (first[n] ? 8u : 0u) | 16u) << (5 * n); }
			(*info)[pos] = word;
			havePrev = true;
			prevPos = pos;
			prevColor = c;
			for (int n = 0; n < 4; n++) { prevVerts[n] = v[n]; prevSlots[n] = slot[n]; }
		}
	}
	return held;
}

void PackElements(const HostMesh& m, const std::vector<uint32_t>& elems, const uint32_t* localIdx, PackedElements* out) {
	const size_t n = elems.size();
	out->a.resize(n);
	out->b.resize(n);
	out->c.resize(n);
	out->area.resize(n);
	for (size_t k = 0; k < n; k++) {
		const uint32_t e = elems[k];
		const uint32_t* v = localIdx ? &localIdx[4 * k] : &m.idx[4 * (size_t)e];
		const float* Q = &m.Qi[9 * (size_t)e];
		const float* qq = &m.QQ[3 * (size_t)e];
		const float* qr = &m.QR[3 * (size_t)e];
		out->a[k] = ElemRecA{ { v[0], v[1], v[2], v[3] }, { Q[0], Q[1], Q[2], Q[3] } };
		out->b[k] = ElemRecB{ { Q[4], Q[5], Q[6], Q[7], Q[8] }, m.volume[e], { qq[0], qq[1] } };
		out->c[k] = float4{ qq[2], qr[0], qr[1], qr[2] };
		out->area[k] = m.area[e];
	}
}

// ------------------------------------------------------------------------------------------------
// Per-call constants.
// ------------------------------------------------------------------------------------------------
int FillSubstepParams(const xf_settings* st, const xf_manipulator* manip, float dt, const HostMesh& mesh, SubstepParams* p, std::string* err) {
	uint32_t energy = (st->flags >> XF_SETTINGS_ENERGY_BIT) & XF_SETTINGS_ENERGY_MASK;
	switch (energy) {
	case XF_ENERGY_MIXED: case XF_ENERGY_MIXED_SEL: case XF_ENERGY_YEOH_SKIN: case XF_ENERGY_YEOH_SKIN_FAST: break;
	case 0: case 1: case 2: case 6: case 8: case 9: case 10: case 11: case 12:
		*err = "energy id " + std::to_string(energy) + " (Pixar*/YeohSkinSel/Continuous*/Cube*) is outside the tet hot path this library implements";
		return XF_ERR_UNSUPPORTED;
	default: energy = XF_ENERGY_MIXED_SEL; break; // TSolveElement's default label, Fem.cpp:884-887
	}
	if (!(dt > 0.0f)) { *err = "dt must be positive"; return XF_ERR_INVALID; }
	p->dt = dt;
	p->dt2 = dt * dt;
	p->invDt = 1.0f / dt;
	p->gdtX = st->gravity[0] * dt;
	p->gdtY = st->gravity[1] * dt;
	p->keep = 1.0f - st->timeCorrectedDrag;
	const float mu = 1.0f / st->compliance;
	const float lambda = (2.0f * mu * st->poissonsRatio) / (1.0f - 2.0f * st->poissonsRatio);
	p->a = 1.0f + mu / lambda;
	p->invMu = 1.0f / mu;
	p->invLambda = 1.0f / lambda;
	p->damping = st->damping;
	p->compliance = st->compliance;
	p->energy = energy;
	p->simultaneous = (st->flags >> XF_SETTINGS_XPBD_SOLVE_BIT) & 1u;
	p->rayleigh = (st->flags >> XF_SETTINGS_RAYLEIGH_TYPE_BIT) & XF_SETTINGS_RAYLEIGH_TYPE_MASK;
	p->lockLeft = (st->flags & XF_SETTINGS_LOCK_LEFT) ? 1u : 0u;
	p->lockRight = (st->flags & XF_SETTINGS_LOCK_RIGHT) ? 1u : 0u;
	p->volumePasses = st->volumePasses;
	p->tickId = st->tickId;
	// Damp(): amortised settings scale damping by the period and swap the PBD constant (Geo.cpp:346-355)
	const bool amortized = p->rayleigh == XF_RAYLEIGH_POST_AMORTIZED;
	p->dampDamping = amortized ? st->damping * (float)XF_AMORTIZATION_PERIOD : st->damping;
	p->pbdDamping = amortized ? st->amortizedVolumeAndTimeCorrectedPbdDamping : st->volumeAndTimeCorrectedPbdDamping;
	p->doDamp = (p->rayleigh >= XF_RAYLEIGH_POST && p->dampDamping > 0.0f) ? 1u : 0u;
	p->doPbdDamp = st->pbdDamping > 0.0f ? 1u : 0u;
	memcpy(p->lockT, st->lockedRightTransform3d, sizeof(p->lockT));
	memcpy(p->origin, mesh.origin, sizeof(p->origin));
	p->c18 = 1.8f / (dt * dt);
	p->manipOn = 0;
	p->manipIdx = 0;
	p->manipTarget[0] = p->manipTarget[1] = p->manipTarget[2] = 0.0f;
	if (manip && manip->picked) {
		if (manip->pickedPointIdx >= mesh.nV) { *err = "manipulator pickedPointIdx out of range"; return XF_ERR_INVALID; }
		// target = pos + t * pickDirTarget, t = n.(pick0 - pos) / n.pickDirTarget   (Geo.cpp:336-337)
		const F3 n = { manip->manipPlaneNormal[0], manip->manipPlaneNormal[1], manip->manipPlaneNormal[2] };
		const F3 d = { manip->pick0[0] - manip->pos[0], manip->pick0[1] - manip->pos[1], manip->pick0[2] - manip->pos[2] };
		const F3 dir = { manip->pickDirTarget[0], manip->pickDirTarget[1], manip->pickDirTarget[2] };
		const float t = Dot(n, d) / Dot(n, dir);
		for (int k = 0; k < 3; k++) { p->manipTarget[k] = manip->pos[k] + t * dir[k]; }
		p->manipOn = 1;
		p->manipIdx = manip->pickedPointIdx;
	}
	p->nColors = (uint32_t)mesh.colorStart.size() - 1;
	for (size_t c = 0; c < mesh.colorStart.size(); c++) { p->colorStart[c] = mesh.colorStart[c]; }
	p->vary = nullptr;
	return XF_OK;
}

}  // namespace xf

// ------------------------------------------------------------------------------------------------
// C ABI: mesh producer
// ------------------------------------------------------------------------------------------------
extern "C" int xf_generate_tet_block(uint32_t width, uint32_t height, uint32_t depth, float sx, float sy, float sz, uint32_t pattern,
                                     float wonkiness, float* nodes, uint32_t* idxStream, uint32_t* colorHint) {
	if (!nodes || !idxStream || width == 0 || height == 0 || depth == 0) { return XF_ERR_INVALID; }
	const uint32_t vx = width + 1, vy = height + 1;
	const float scale[3] = { sx, sy, sz };
	const float corner[3] = { sx * (-0.5f * (float)width), sy * (-0.5f * (float)height), sz * (-0.5f * (float)depth) };
	uint32_t rng = 1;
	for (uint32_t z = 0; z <= depth; z++) {
		for (uint32_t y = 0; y <= height; y++) {
			for (uint32_t x = 0; x <= width; x++) {
				const bool onFace[3] = { x == 0 || x == width, y == 0 || y == height, z == 0 || z == depth };
				// jitter draws: the reference's vec3(RandF(), RandF(), RandF()) is evaluated z, y, x by g++
				float jitter[3];
				for (int k = 2; k >= 0; k--) { jitter[k] = onFace[k] ? 0.0f : (float)xf::Lehmer(rng) * 9.3132258e-10f - 1.0f; }
				const float lattice[3] = { (float)x, (float)y, (float)z };
				float* out = nodes + 3 * (size_t)(x + y * vx + z * vx * vy);
				for (int k = 0; k < 3; k++) { out[k] = corner[k] + scale[k] * (lattice[k] + 0.5f * (wonkiness * jitter[k])); }
			}
		}
	}
	// six Kuhn tets around the 0-7 diagonal of every hex (MeshGen.cpp:235-240)
	static const uint8_t kTets[6][4] = { { 0, 1, 5, 7 }, { 0, 7, 3, 1 }, { 0, 2, 3, 7 }, { 0, 7, 6, 2 }, { 0, 4, 6, 7 }, { 0, 7, 5, 4 } };
	size_t cell = 0;
	for (uint32_t z = 0; z < depth; z++) {
		for (uint32_t y = 0; y < height; y++) {
			for (uint32_t x = 0; x < width; x++, cell++) {
				uint32_t corners[8];
				for (uint32_t bits = 0; bits < 8; bits++) {
					uint32_t i = bits & 1, j = (bits >> 1) & 1, k = bits >> 2;
					if (pattern == XF_PATTERN_UNIFORM) {
						corners[bits] = (x + i) + (y + j) * vx + (z + k) * vx * vy;
					} else { // mirrored: flip the local frame in odd cells (MeshGen.cpp:196-205)
						uint32_t fi = (x & 1) ? 1 - i : i, fj = (y & 1) ? 1 - j : j, fk = (z & 1) ? 1 - k : k;
						uint32_t slot = ((x + y + z) & 1) ? 7 - bits : bits;
						corners[slot] = (x + fi) + (y + fj) * vx + (z + fk) * vx * vy;
					}
				}
				for (uint32_t t = 0; t < 6; t++) {
					uint32_t* rec = idxStream + 30 * cell + 5 * t;
					rec[0] = 4;
					for (int c = 0; c < 4; c++) { rec[1 + c] = corners[kTets[t][c]]; }
					if (colorHint) {
						// Uniform Kuhn lattice: tets of one type conflict only across cell offsets whose
						// (dx+dy+dz) mod 4 != 0, so (type, (x+y+z)%4) is a perfect, balanced 24-colouring.  The classes are
						// numbered cell-class-major: colours 6k .. 6k+5 walk the ring of six tets around the 0-7 diagonal of
						// the cells of class k, consecutive tets sharing a face.  Position j of each of those six colours is
						// the same cell, which is what the chained sweep (ChainInfo below, k_substeps_chain) feeds on.
						colorHint[6 * cell + t] = pattern == XF_PATTERN_UNIFORM ? 6 * ((x + y + z) & 3) + t : 0xffffffffu;
					}
				}
			}
		}
	}
	return XF_OK;
}
