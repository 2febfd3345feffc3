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

int BuildPartition(const HostMesh& m, uint32_t nRanks, uint32_t rank, PartPlan* out, std::string* err) {
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
