// Host-side plan for stepping ONE mesh on several GPUs (BASELINE config 4).
//
// Elements are split into `nRanks` contiguous chunks of the order "rest-pose centroid x, then stream index"
// (x-slabs for MeshGen blocks, a 1-D geometric partition for anything else).  A rank keeps a local copy of every
// vertex its elements touch; vertices touched by several ranks are *shared*.  The colouring is global, so in a
// colour phase a shared vertex is modified by at most one element in the whole job; after the phase its new
// position is pushed to the other copies (in-kernel peer stores on the device, xf_part.cu; explicit send/recv
// lists for a host-driven transport or a CPU emulation).  Vertex phases (predict / post) are computed
// redundantly on every copy from identical inputs, so no velocity or mass traffic is needed, and the whole
// run stays bit-identical to the single-GPU schedule, i.e. to the serial sweep in xf_get_order's order.
//
// Every rank computes the same plan from the same inputs; nothing here communicates.
#include <algorithm>
#include <numeric>

#include "xf_partition.h"

namespace xf {

namespace {

// XF_PARTITION_GRAPH: greedy graph growing.  Elements are nodes, two elements are adjacent when they share a vertex.  Part q
// starts from the unassigned element that is farthest (in breadth-first hops) from everything assigned so far - the first one
// from a peripheral element found by two sweeps - and grows breadth-first until it holds its share of the elements; what a
// part cannot reach (disconnected leftovers) goes to the part that owns a neighbour, else to the smallest part.  Deterministic:
// every rank computes the same assignment.
void GraphGrow(const HostMesh& m, uint32_t nRanks, std::vector<uint8_t>* elemRank) {
	const uint32_t nT = m.nT, kNone = 0xffu;
	std::vector<uint32_t> vStart(m.nV + 1, 0), vElems(4 * (size_t)nT);
	for (size_t k = 0; k < 4 * (size_t)nT; k++) { vStart[m.idx[k] + 1]++; }
	for (uint32_t v = 0; v < m.nV; v++) { vStart[v + 1] += vStart[v]; }
	{
		std::vector<uint32_t> fill(vStart.begin(), vStart.end() - 1);
		for (uint32_t e = 0; e < nT; e++) { for (int j = 0; j < 4; j++) { vElems[fill[m.idx[4 * (size_t)e + j]]++] = e; } }
	}
	std::vector<uint8_t>& owner = *elemRank;
	owner.assign(nT, (uint8_t)kNone);
	std::vector<uint32_t> dist(nT), queue;
	queue.reserve(nT);
	// breadth-first distances from `sources` over the elements for which `pass(e)` holds; returns the last element reached
	auto sweep = [&](const std::vector<uint32_t>& sources, bool unassignedOnly) {
		std::fill(dist.begin(), dist.end(), 0xffffffffu);
		queue.clear();
		for (uint32_t s : sources) { dist[s] = 0; queue.push_back(s); }
		uint32_t last = sources.empty() ? 0u : sources[0];
		for (size_t h = 0; h < queue.size(); h++) {
			const uint32_t e = queue[h];
			if (!unassignedOnly || owner[e] == kNone) { last = e; }
			for (int j = 0; j < 4; j++) {
				const uint32_t v = m.idx[4 * (size_t)e + j];
				for (uint32_t k = vStart[v]; k < vStart[v + 1]; k++) {
					const uint32_t f = vElems[k];
					if (dist[f] == 0xffffffffu) { dist[f] = dist[e] + 1; queue.push_back(f); }
				}
			}
		}
		return last;
	};
	uint32_t assigned = 0;
	std::vector<uint32_t> partSize(nRanks, 0), frontier;
	for (uint32_t q = 0; q < nRanks; q++) {
		const uint32_t want = (uint32_t)(((uint64_t)nT * (q + 1)) / nRanks) - assigned;
		// seed: farthest unassigned element from the assigned ones (q == 0: a peripheral element of the mesh)
		uint32_t seed = 0;
		if (q == 0) {
			seed = sweep({ sweep({ 0u }, false) }, false);
		} else {
			std::vector<uint32_t> src;
			for (uint32_t e = 0; e < nT; e++) { if (owner[e] != kNone) { src.push_back(e); } }
			seed = sweep(src, true);
			if (owner[seed] != kNone) { for (uint32_t e = 0; e < nT; e++) { if (owner[e] == kNone) { seed = e; break; } } }
		}
		// grow
		frontier.assign(1, seed);
		owner[seed] = (uint8_t)q;
		uint32_t got = 1;
		for (size_t h = 0; h < frontier.size() && got < want; h++) {
			const uint32_t e = frontier[h];
			for (int j = 0; j < 4 && got < want; j++) {
				const uint32_t v = m.idx[4 * (size_t)e + j];
				for (uint32_t k = vStart[v]; k < vStart[v + 1] && got < want; k++) {
					const uint32_t f = vElems[k];
					if (owner[f] == kNone) { owner[f] = (uint8_t)q; frontier.push_back(f); got++; }
				}
			}
		}
		partSize[q] = got;
		assigned += got;
	}
	// leftovers (a part ran out of reachable elements): join a neighbour's part, else the smallest part
	for (bool progress = true; progress && assigned < nT;) {
		progress = false;
		for (uint32_t e = 0; e < nT; e++) {
			if (owner[e] != kNone) { continue; }
			uint32_t best = kNone;
			for (int j = 0; j < 4; j++) {
				const uint32_t v = m.idx[4 * (size_t)e + j];
				for (uint32_t k = vStart[v]; k < vStart[v + 1]; k++) {
					const uint32_t o = owner[vElems[k]];
					if (o != kNone && (best == kNone || partSize[o] < partSize[best])) { best = o; }
				}
			}
			if (best != kNone) { owner[e] = (uint8_t)best; partSize[best]++; assigned++; progress = true; }
		}
	}
	for (uint32_t e = 0; e < nT; e++) {
		if (owner[e] == kNone) {
			const uint32_t q = (uint32_t)(std::min_element(partSize.begin(), partSize.end()) - partSize.begin());
			owner[e] = (uint8_t)q;
			partSize[q]++;
		}
	}
}

}  // namespace

int BuildPartition(const HostMesh& m, uint32_t nRanks, uint32_t rank, PartPlan* out, std::string* err, uint32_t method) {
	if (nRanks == 0 || nRanks > 64 || rank >= nRanks) { *err = "nRanks must be 1..64 and rank < nRanks"; return XF_ERR_INVALID; }
	if (m.nT < nRanks) { *err = "fewer elements than ranks"; return XF_ERR_INVALID; }
	PartPlan& p = *out;
	p.nRanks = nRanks;
	p.rank = rank;
	const uint32_t nColors = (uint32_t)m.colorStart.size() - 1;

	// 1. element -> rank
	std::vector<float> cx(m.nT);
	for (uint32_t e = 0; e < m.nT; e++) {
		const uint32_t* v = &m.idx[4 * (size_t)e];
		cx[e] = (float)(m.X0[3 * (size_t)v[0]] + m.X0[3 * (size_t)v[1]] + m.X0[3 * (size_t)v[2]] + m.X0[3 * (size_t)v[3]]);
	}
	std::vector<uint32_t> byX(m.nT);
	std::iota(byX.begin(), byX.end(), 0u);
	std::stable_sort(byX.begin(), byX.end(), [&](uint32_t a, uint32_t b) { return cx[a] < cx[b]; });
	p.elemRank.assign(m.nT, 0);
	for (uint32_t k = 0; k < m.nT; k++) { p.elemRank[byX[k]] = (uint8_t)(((uint64_t)k * nRanks) / m.nT); }
	if (method == XF_PARTITION_GRAPH) { GraphGrow(m, nRanks, &p.elemRank); }

	// 2. which ranks touch each vertex
	std::vector<uint64_t> mask(m.nV, 0);
	for (uint32_t e = 0; e < m.nT; e++) {
		for (int j = 0; j < 4; j++) { mask[m.idx[4 * (size_t)e + j]] |= (uint64_t)1 << p.elemRank[e]; }
	}
	const uint64_t me = (uint64_t)1 << rank;

	// 3. local vertices (ascending global id), local index of every shared vertex on every rank that has it
	std::vector<uint32_t> globalToLocal(m.nV, 0xffffffffu);
	std::vector<uint32_t> counters(nRanks, 0);
	uint64_t peerMask = 0;
	p.verts.clear();
	p.shareStart.assign(1, 0);
	p.sharePeerRank.clear();
	p.shareRemoteIdx.clear();
	for (uint32_t v = 0; v < m.nV; v++) {
		const uint64_t mk = mask[v];
		if (mk & me) {
			globalToLocal[v] = (uint32_t)p.verts.size();
			p.verts.push_back(v);
			for (uint32_t q = 0; q < nRanks; q++) {
				if (q != rank && (mk >> q & 1)) {
					p.sharePeerRank.push_back(q);
					p.shareRemoteIdx.push_back(counters[q]);
					peerMask |= (uint64_t)1 << q;
				}
			}
			p.shareStart.push_back((uint32_t)p.sharePeerRank.size());
		}
		for (uint32_t q = 0; q < nRanks; q++) { if (mk >> q & 1) { counters[q]++; } }
	}
	p.peers.clear();
	for (uint32_t q = 0; q < nRanks; q++) { if (peerMask >> q & 1) { p.peers.push_back(q); } }
	std::vector<uint32_t> slotOfRank(nRanks, 0xffffffffu);
	for (size_t s = 0; s < p.peers.size(); s++) { slotOfRank[p.peers[s]] = (uint32_t)s; }

	// 4. local elements, colour-major; inside a colour the interface elements (touching a shared vertex) first
	p.colorStart.assign(nColors + 1, 0);
	p.ifaceEnd.assign(nColors, 0);
	p.elems.clear();
	for (uint32_t c = 0; c < nColors; c++) {
		p.colorStart[c] = (uint32_t)p.elems.size();
		for (int pass = 0; pass < 2; pass++) {
			for (uint32_t k = m.colorStart[c]; k < m.colorStart[c + 1]; k++) {
				const uint32_t e = m.order[k];
				if (p.elemRank[e] != rank) { continue; }
				bool iface = false;
				for (int j = 0; j < 4; j++) { iface = iface || (mask[m.idx[4 * (size_t)e + j]] != me); }
				if (iface == (pass == 0)) { p.elems.push_back(e); }
			}
			if (pass == 0) { p.ifaceEnd[c] = (uint32_t)p.elems.size(); }
		}
	}
	p.colorStart[nColors] = (uint32_t)p.elems.size();
	p.localIdx.resize(4 * p.elems.size());
	for (size_t k = 0; k < p.elems.size(); k++) {
		for (int j = 0; j < 4; j++) { p.localIdx[4 * k + j] = globalToLocal[m.idx[4 * (size_t)p.elems[k] + j]]; }
	}

	// 5. explicit halo lists per (colour, peer slot): what this rank sends after the phase, what it receives;
	//    both ascending in global vertex id, so sender and receiver enumerate the same vertices in the same order.
	const size_t nPeers = p.peers.size();
	std::vector<std::vector<uint32_t>> send(nColors * nPeers), recv(nColors * nPeers);
	for (uint32_t c = 0; c < nColors; c++) {
		for (uint32_t k = m.colorStart[c]; k < m.colorStart[c + 1]; k++) {
			const uint32_t e = m.order[k];
			const uint32_t owner = p.elemRank[e];
			for (int j = 0; j < 4; j++) {
				const uint32_t v = m.idx[4 * (size_t)e + j];
				const uint64_t mk = mask[v];
				if (owner == rank) {
					for (size_t s = 0; s < nPeers; s++) { if (mk >> p.peers[s] & 1) { send[c * nPeers + s].push_back(v); } }
				} else if ((mk & me) && slotOfRank[owner] != 0xffffffffu) {
					recv[c * nPeers + slotOfRank[owner]].push_back(v);
				}
			}
		}
	}
	// 6. stage codes of the barrier-free schedule, from the global colour-major order
	{
		std::vector<uint8_t> last(m.nV, 0), pred(4 * (size_t)m.nT, 0);
		for (uint32_t k = 0; k < m.nT; k++) {
			const uint32_t e = m.order[k];
			const uint8_t code = (uint8_t)std::min<uint32_t>(1u + m.color[e], 254u);
			for (int j = 0; j < 4; j++) {
				uint8_t& l = last[m.idx[4 * (size_t)e + j]];
				pred[4 * (size_t)e + j] = l;
				l = code;
			}
		}
		p.predCode.resize(4 * p.elems.size());
		for (size_t k = 0; k < p.elems.size(); k++) {
			for (int j = 0; j < 4; j++) {
				const uint32_t v = m.idx[4 * (size_t)p.elems[k] + j];
				uint8_t c = pred[4 * (size_t)p.elems[k] + j];
				if (c == 0 && mask[v] != me) { c = 255; }
				p.predCode[4 * k + j] = c;
			}
		}
		p.lastCode.resize(p.verts.size());
		for (size_t i = 0; i < p.verts.size(); i++) { p.lastCode[i] = last[p.verts[i]]; }
		// a property of the whole job (every rank must reach the same verdict): at most two copies of any vertex
		p.dataflowOk = nColors <= 253;
		for (uint32_t v = 0; v < m.nV; v++) { if (__builtin_popcountll(mask[v]) > 2) { p.dataflowOk = false; } }
		for (uint32_t q = 0; q < nRanks; q++) { if (counters[q] > 0x01000000u) { p.dataflowOk = false; } }
	}
	p.sendStart.assign(1, 0);
	p.recvStart.assign(1, 0);
	p.sendVerts.clear();
	p.recvVerts.clear();
	for (size_t k = 0; k < send.size(); k++) {
		std::sort(send[k].begin(), send[k].end());
		std::sort(recv[k].begin(), recv[k].end());
		for (uint32_t v : send[k]) { p.sendVerts.push_back(globalToLocal[v]); }
		for (uint32_t v : recv[k]) { p.recvVerts.push_back(globalToLocal[v]); }
		p.sendStart.push_back((uint32_t)p.sendVerts.size());
		p.recvStart.push_back((uint32_t)p.recvVerts.size());
	}
	return XF_OK;
}

}  // namespace xf
