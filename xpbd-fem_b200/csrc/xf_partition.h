// Partition plan for single-mesh multi-GPU stepping (see xf_partition.cpp).
#pragma once

#include "xf_scene.h"

namespace xf {

struct PartPlan {
	uint32_t nRanks = 1, rank = 0;
	std::vector<uint8_t> elemRank;        // owner of every element of the full mesh
	std::vector<uint32_t> elems;          // global ids of this rank's elements: colour-major, interface first
	std::vector<uint32_t> colorStart;     // nColors + 1 offsets into elems
	std::vector<uint32_t> ifaceEnd;       // per colour: end of the interface sub-range
	std::vector<uint32_t> localIdx;       // 4 local vertex ids per local element
	std::vector<uint32_t> verts;          // local -> global vertex id (ascending)
	std::vector<uint32_t> peers;          // ranks sharing at least one vertex (ascending)
	std::vector<uint32_t> shareStart;     // CSR over local vertices
	std::vector<uint32_t> sharePeerRank;  //   rank holding another copy
	std::vector<uint32_t> shareRemoteIdx; //   local index of the vertex on that rank
	std::vector<uint32_t> sendStart, sendVerts; // [(colour * nPeers + slot)] -> local vertex ids to send after the phase
	std::vector<uint32_t> recvStart, recvVerts; // same shape: local vertex ids overwritten by that peer's phase
	// barrier-free schedule (xf_part.cu: k_part_dataflow): who wrote a vertex last, in the GLOBAL serial order.
	// predCode: 4 per local element - 0 = this substep's vertex phase, 255 = vertex phase of a SHARED vertex (the element
	// must also see the peer's acknowledgement that its copy has been through the vertex phase), else 1 + colour of the
	// previous element around the vertex (on whichever rank).  lastCode: per local vertex, 1 + colour of its last element
	// (0 = none).  dataflowOk: every shared vertex has exactly one other copy and the codes fit a byte.
	std::vector<uint8_t> predCode, lastCode;
	bool dataflowOk = false;
};

int BuildPartition(const HostMesh& full, uint32_t nRanks, uint32_t rank, PartPlan* out, std::string* err, uint32_t method = XF_PARTITION_SLABS);

}  // namespace xf
