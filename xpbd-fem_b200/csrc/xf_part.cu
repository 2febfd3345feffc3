// One mesh on several GPUs: each rank owns a slab of elements plus local copies of the vertices they touch
// (plan: xf_partition.cpp).  Per phase (vertex phase, colour 0, colour 1, ...) every rank launches one kernel:
//   prologue  thread 0 of each CTA waits until every peer has finished the previous phase (flag >= epoch - 1,
//             ld.acquire.sys on flags the peers write into this GPU's memory over NVLink);
//   body      the rank's elements of the colour; elements touching a shared vertex ALSO store the new position
//             into the peers' copies (plain stores to cudaIpc-mapped peer memory: the halo exchange is fused
//             into the sweep, no pack/send/unpack kernels and no NCCL on the data path);
//   epilogue  the last CTA to finish issues a system-scope fence and writes `epoch` into each peer's flag slot.
// A rank can therefore never run more than one phase ahead of a neighbour, which is exactly the dependence
// the global colour order needs.  NCCL/torch.distributed (or anything else) is only needed to move the 128-byte
// IPC handles at start-up.  Spin loops carry a bail-out so a protocol error reports XF_ERR_CUDA instead of
// hanging the GPU.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "xf_dispatch.cuh"
#include "xf_element.cuh"
#include "xf_partition.h"
#include "xf_phase.cuh"

namespace xf {

constexpr int kMaxPeers = 16;

struct PartDevice {
	DeviceScene local;            // local sub-mesh (local vertex numbering)
	uint32_t nPeers;
	uint32_t myRank;
	VertexRec* peerXw[kMaxPeers]; // peers' vertex arrays (IPC-mapped)
	double4* peerV[kMaxPeers];    // peers' velocity arrays (same IPC mapping: V follows Xw in one allocation)
	unsigned long long* peerFlags[kMaxPeers]; // peers' flag arrays, indexed by sender rank
	unsigned long long* myFlags;  // indexed by sender rank
	uint32_t peerRank[kMaxPeers];
	const uint32_t* shareStart;   // CSR over local vertices
	const uint32_t* shareSlot;    //   peer slot
	const uint32_t* shareRemoteIdx;
	unsigned int* doneCounter;
	unsigned int* errorFlag;
	const uint32_t* colorStart;   // device copies for the persistent kernel
	const uint32_t* ifaceEnd;
	uint32_t nColors;
	// barrier-free schedule: acknowledgement words, one per local vertex, written by the peer that holds the other copy
	uint32_t* myAck;
	uint32_t* peerAck[kMaxPeers];
	// ... and its two warp pools: `nIfaceWarps` warps do nothing but the interface elements and the shared vertices (the
	// work whose dependence chain crosses NVLink), the others the interior elements and the private vertices
	const uint32_t* sharedList; // local ids of the shared vertices
	const uint32_t* privList;   // local ids of the others
	uint32_t nSharedList, nPrivList, nIfaceWarps;
	// barrier-free schedule: 1 = a system-scope fence between the peer store and the local store of a shared vertex (the hand-off
	// to a same-rank successor is then a release/acquire chain by construction); 0 = program order of two relaxed stores only
	uint32_t releaseStores;
	// ... and a gap (ns) between the peer store and the local store of a shared vertex: widens the margin of hazard (a) (a
	// same-rank successor overtaking our peer store) by that much, at the price of the same delay on the interface chain
	uint32_t storeGapNs;
};

__device__ __forceinline__ unsigned long long LoadAcquireSys(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void StoreReleaseSys(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void PhaseWait(const PartDevice& pd, unsigned long long epoch) {
	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < pd.nPeers; s++) {
			const unsigned long long* f = pd.myFlags + pd.peerRank[s];
			unsigned long long spins = 0;
			while (LoadAcquireSys(f) + 1 < epoch) {
				if (++spins > (1ull << 27)) { atomicExch(pd.errorFlag, 1u); break; } // ~seconds: protocol error, do not hang
			}
		}
	}
	__syncthreads();
}
__device__ __forceinline__ void PhaseSignal(const PartDevice& pd, unsigned long long epoch) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		const unsigned int prev = atomicAdd(pd.doneCounter, 1u);
		if (prev == gridDim.x - 1) {
			*pd.doneCounter = 0;
			__threadfence_system();
			for (uint32_t s = 0; s < pd.nPeers; s++) { StoreReleaseSys(pd.peerFlags[s] + pd.myRank, epoch); }
		}
	}
}

// Global store that mirrors position updates of shared vertices into the peers' copies.
struct MirroredStore {
	GlobalStore base;
	const PartDevice* pd;
	__device__ __forceinline__ VertexRegs LoadX(uint32_t i) const { return base.LoadX(i); }
	__device__ __forceinline__ void StoreX(uint32_t i, const VertexRegs& v) const {
		base.StoreX(i, v);
		const uint32_t b = __ldg(pd->shareStart + i), e = __ldg(pd->shareStart + i + 1);
		for (uint32_t k = b; k < e; k++) { StoreVertex(pd->peerXw[__ldg(pd->shareSlot + k)], __ldg(pd->shareRemoteIdx + k), v); }
	}
	__device__ __forceinline__ void LoadO(uint32_t i, double* o) const { base.LoadO(i, o); }
	__device__ __forceinline__ void LoadV(uint32_t i, double* o) const { base.LoadV(i, o); }
	// damping sweeps: velocities of shared vertices are mirrored like positions
	__device__ __forceinline__ void StoreV(uint32_t i, const double* v) const {
		base.StoreV(i, v);
		const uint32_t b = __ldg(pd->shareStart + i), e = __ldg(pd->shareStart + i + 1);
		for (uint32_t k = b; k < e; k++) { StoreD3(pd->peerV[__ldg(pd->shareSlot + k)], __ldg(pd->shareRemoteIdx + k), v); }
	}
};

// KIND 0: main constraint solve   1: volume-only pass   2: Rayleigh damp (V)   3: PBD damp (V)   (as in xf_kernels.cu)
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED, typename VS>
__device__ __forceinline__ void PartSweepOne(const VS& vs, const DeviceScene& sc, const SubstepParams& p, uint32_t e) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	ElemRec rec;
	if (KIND == 3) {
		rec.idx = LoadElementIdx(sc, e);
		PbdDampElement<EXACT>(vs, p, __ldg(sc.eArea + e), rec.idx);
		return;
	}
	LoadElement<(KIND != 1) && kPrefactored, EXACT>(sc, e, rec);
	if (KIND == 0) { SolveElement<ENERGY, SIMUL, EXACT, DAMPED>(vs, p, rec); }
	if (KIND == 1) { SolveVolumeOnly<EXACT>(vs, p, rec); }
	if (KIND == 2) { DampElement<ENERGY, SIMUL, EXACT>(vs, p, rec); }
}

// One phase of the flag protocol: the rank's elements [begin, end) of one colour (interface elements first, up to ifaceEnd).
// Damping sweeps (KIND >= 2) act on the elements whose GLOBAL serial position lies in [lo, hi) (amortised slices, Geo.cpp:794-797).
template <int KIND, int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__global__ void __launch_bounds__(256) k_part_sweep(const __grid_constant__ PartDevice pd, const __grid_constant__ SubstepParams p, uint32_t begin,
                                                    uint32_t ifaceEnd, uint32_t end, uint32_t lo, uint32_t hi, unsigned long long epoch) {
	PhaseWait(pd, epoch);
	const uint32_t e = begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (e < end && (KIND < 2 || InSlice(pd.local, e, lo, hi))) {
		if (e < ifaceEnd) {
			const MirroredStore vs{ StoreOf(pd.local), &pd };
			PartSweepOne<KIND, ENERGY, SIMUL, EXACT, DAMPED>(vs, pd.local, p, e);
		} else {
			const GlobalStore vs = StoreOf(pd.local);
			PartSweepOne<KIND, ENERGY, SIMUL, EXACT, DAMPED>(vs, pd.local, p, e);
		}
	}
	PhaseSignal(pd, epoch);
}

// An empty phase: orders this rank's earlier launches (a barrier-free launch of the previous call) before the peers' next phase.
__global__ void __launch_bounds__(32) k_part_sync_phase(const __grid_constant__ PartDevice pd, unsigned long long epoch) {
	PhaseWait(pd, epoch);
	PhaseSignal(pd, epoch);
}

// vertex phase of the partitioned run: identical arithmetic to k_vertex_phase, on every local copy
template <bool EXACT>
__device__ __forceinline__ void PartVertexPhaseOne(const DeviceScene& sc, const SubstepParams& p, uint32_t i, bool doPost, bool doPredict) {
	typedef Op<EXACT> O;
	VertexRegs v = LoadVertex(sc.Xw, i);
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
		if (p.lockLeft && (v.flags & XF_VERT_LEFT)) { v.x[0] = o[0]; v.x[1] = o[1]; v.x[2] = o[2]; v.w = 0.0f; }
		if (p.lockRight && (v.flags & XF_VERT_RIGHT)) {
			double x0d[3];
			LoadD3(sc.X0, i, x0d);
			float x0[3] = { __double2float_rn(x0d[0]), __double2float_rn(x0d[1]), __double2float_rn(x0d[2]) };
#pragma unroll
			for (int r = 0; r < 3; r++) {
				float t = O::dot(p.lockT[0 + r], p.lockT[4 + r], p.lockT[8 + r], x0[0], x0[1], x0[2]);
				double q = (double)O::add(p.origin[r], t);
				v.x[r] = q;
				o[r] = q;
			}
			v.w = 0.0f;
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
	StoreVertex(sc.Xw, i, v);
	StoreD3(sc.O, i, o);
	StoreD3(sc.V, i, vel);
}

// vertex phase of the partitioned run: identical arithmetic to k_vertex_phase, on every local copy
template <bool EXACT>
__global__ void __launch_bounds__(256) k_part_vertex_phase(const __grid_constant__ PartDevice pd, const __grid_constant__ SubstepParams p, int doPost,
                                                           int doPredict, unsigned long long epoch) {
	PhaseWait(pd, epoch);
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < pd.local.nV) { PartVertexPhaseOne<EXACT>(pd.local, p, i, doPost != 0, doPredict != 0); }
	PhaseSignal(pd, epoch);
}

// ---- persistent variant: one cooperative launch per xf_part_substep call --------------------------------------
// Per phase: wait for the peers' previous phase -> interface elements (their new positions are also stored into
// the peers' copies) -> the last CTA to finish them signals the peers -> interior elements run while the signal
// crosses NVLink -> local grid barrier.  A peer that sees the signal may already write this rank's shared vertices
// for the next phase: safe, because only interior elements (private vertices) are still running here.
__device__ __forceinline__ void SignalPeers(const PartDevice& pd, unsigned long long epoch) {
	__threadfence_system();
	for (uint32_t s = 0; s < pd.nPeers; s++) { StoreReleaseSys(pd.peerFlags[s] + pd.myRank, epoch); }
}
// Peer wait by ONE thread of the grid, placed right before it joins the local grid barrier: the barrier's release then
// also means "every peer has finished the interface part of this phase", nobody else polls, and the NVLink latency
// overlaps the interior elements and the barrier itself.
__device__ __forceinline__ void PeerWaitSingle(const PartDevice& pd, unsigned long long epoch) {
	for (uint32_t s = 0; s < pd.nPeers; s++) {
		const unsigned long long* f = pd.myFlags + pd.peerRank[s];
		unsigned long long spins = 0, v;
		do {
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
			if (++spins > (1ull << 27)) { atomicExch(pd.errorFlag, 1u); break; }
		} while (v < epoch);
	}
	asm volatile("fence.acq_rel.sys;" ::: "memory");
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
__global__ void __launch_bounds__(256, 2) k_part_persistent(const __grid_constant__ PartDevice pd, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                            unsigned long long epoch) {
	namespace cg = cooperative_groups;
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	cg::grid_group grid = cg::this_grid();
	const DeviceScene& sc = pd.local;
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	const uint32_t slot = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u);
	const MirroredStore mirrored{ StoreOf(sc), &pd };
	const GlobalStore plain = StoreOf(sc);
	const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
	// the previous launch ended with a vertex phase whose epoch the peers may not have reached yet
	if (leader) { PeerWaitSingle(pd, epoch); }
	grid.sync();
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		// vertex phase (post of substep s-1, predict of substep s); the last iteration is the closing post only.
		// It touches every local vertex, so the peers are told only after the whole grid has finished it.
		++epoch;
		const bool doPost = s > 0, doPredict = s < nSubsteps;
		for (uint32_t i = gtid; i < sc.nV; i += gsize) { PartVertexPhaseOne<EXACT>(sc, p, i, doPost, doPredict); }
		grid.sync();
		if (leader) { SignalPeers(pd, epoch); PeerWaitSingle(pd, epoch); }
		grid.sync();
		if (s == nSubsteps) { break; }
		for (uint32_t c = 0; c < pd.nColors; c++) {
			++epoch;
			const uint32_t b = __ldg(pd.colorStart + c), mid = __ldg(pd.ifaceEnd + c), end = __ldg(pd.colorStart + c + 1);
			bool wroteRemote = false;
			for (uint32_t e = b + slot; e < mid; e += gsize) {
				ElemRec rec;
				LoadElement<kPrefactored, EXACT>(sc, e, rec);
				SolveElement<ENERGY, SIMUL, EXACT, false>(mirrored, p, rec);
				wroteRemote = true;
			}
			if (__syncthreads_or(wroteRemote ? 1 : 0)) { if (threadIdx.x == 0) { __threadfence_system(); } }
			if (threadIdx.x == 0) {
				if (atomicAdd(pd.doneCounter, 1u) == gridDim.x - 1) {
					*pd.doneCounter = 0;
					SignalPeers(pd, epoch);
				}
			}
			for (uint32_t e = mid + slot; e < end; e += gsize) {
				ElemRec rec;
				LoadElement<kPrefactored, EXACT>(sc, e, rec);
				SolveElement<ENERGY, SIMUL, EXACT, false>(plain, p, rec);
			}
			if (leader) { PeerWaitSingle(pd, epoch); }
			grid.sync();
		}
	}
}

// ---- barrier-free variant (XF_SCHEDULE_DATAFLOW): versioned records across GPUs -----------------------------------
// Same idea as xf_dataflow.cu: a vertex record carries the stage that wrote it and every writer waits until the record
// carries its predecessor's stage.  Here a shared vertex has a copy on two ranks; an element writes BOTH copies (its
// local one and, over NVLink, the peer's) with one 32-byte store each, and everybody polls only its LOCAL copy.  Around
// one vertex the writers are totally ordered and each waits for its predecessor, so the stores to one copy are causally
// ordered even though they come from two GPUs: no flags, no fences, no epochs on the element path.  The vertex phase is
// computed on both copies (identical inputs -> identical bits, as in the other partitioned schedules); the first element
// around a shared vertex in a substep (code 255) additionally waits for the peer's acknowledgement word, so that its
// remote store cannot land before the peer has pushed its copy through the vertex phase.
__device__ __forceinline__ VertexRegs LoadVertexSys(const VertexRec* Xw, uint32_t i) {
	VertexRegs v;
	double packed;
	asm volatile("ld.relaxed.sys.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x[0]), "=d"(v.x[1]), "=d"(v.x[2]), "=d"(packed) : "l"(Xw + i) : "memory");
	const long long bits = __double_as_longlong(packed);
	v.w = __int_as_float((int)(bits & 0xffffffffll));
	v.flags = (uint32_t)((unsigned long long)bits >> 32);
	return v;
}
__device__ __forceinline__ void StoreVertexSys(VertexRec* Xw, uint32_t i, const VertexRegs& v) {
	const long long bits = (long long)(((unsigned long long)v.flags << 32) | (unsigned long long)(uint32_t)__float_as_int(v.w));
	asm volatile("st.relaxed.sys.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(Xw + i), "d"(v.x[0]), "d"(v.x[1]), "d"(v.x[2]), "d"(__longlong_as_double(bits)) : "memory");
}
__device__ __forceinline__ uint32_t LoadAckSys(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// Order of the two stores of a shared vertex: the PEER's copy first, then the local one.  The next writer around the
// vertex may sit on this same rank; it starts when it sees the local record and will then store to the same peer
// address from another SM.  Its store must not overtake ours on the way to the peer, so ours has to be on its way
// before the local record can be seen - and its address is resolved before the solve (`remote`), not between the
// two stores.  (Measured: with local-first stores and the address looked up in between, >= 4M-tet meshes lost
// updates under load; a lost update is fail-stop - the tag never matches and the kernel traps - never a wrong result.)
struct VersionedMirrorStore {
	GlobalStore base;
	uint32_t vid[4];
	VertexRec* remote[4]; // the peer's copy of corner n, or nullptr
	bool release;         // fence between the peer store and the local store (PartDevice::releaseStores)
	uint32_t gapNs;       // PartDevice::storeGapNs
	__device__ __forceinline__ VertexRegs LoadX(uint32_t i) const { return LoadVertexSys(base.Xw, i); }
	__device__ __forceinline__ void StoreX(uint32_t i, const VertexRegs& v) const {
		VertexRec* r = i == vid[0] ? remote[0] : (i == vid[1] ? remote[1] : (i == vid[2] ? remote[2] : remote[3]));
		if (r) {
			StoreVertexSys(r, 0, v);
			// With the fence the peer store is PERFORMED before the local record can be seen: a same-rank successor that reads the
			// local record and then stores to the same peer address is ordered after us by the memory model (fence + relaxed store
			// = release; its poll + the fence it issues before ITS peer store = acquire), not just by the order the stores left.
			if (release) { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
			if (gapNs) { __nanosleep(gapNs); }
		}
		StoreVertexSys(base.Xw, i, v);
	}
	__device__ __forceinline__ void LoadO(uint32_t i, double* o) const { base.LoadO(i, o); }
	__device__ __forceinline__ void LoadV(uint32_t i, double* o) const { base.LoadV(i, o); }
	__device__ __forceinline__ void StoreV(uint32_t i, const double* v) const { base.StoreV(i, v); }
};

constexpr uint32_t kPartSpinLimit = 1u << 26; // polls before a rank gives up on a peer (tens of seconds)
constexpr uint32_t kTagMask = 0xffffff00u;

template <int ENERGY, bool SIMUL, bool EXACT>
__device__ __forceinline__ bool PartDataflowElement(const PartDevice& pd, const SubstepParams& p, const ElemRec& rec, unsigned mask, bool iface,
                                                    uint32_t stageBase, uint32_t c) {
	const uint32_t raw[4] = { rec.idx.x, rec.idx.y, rec.idx.z, rec.idx.w };
	uint32_t vid[4], expectTag[4];
	bool needAck[4];
	VersionedMirrorStore vs;
	vs.base = StoreOf(pd.local);
	vs.release = pd.releaseStores != 0;
	vs.gapNs = pd.storeGapNs;
#pragma unroll
	for (int n = 0; n < 4; n++) {
		vid[n] = raw[n] & 0x00ffffffu;
		const uint32_t code = raw[n] >> 24;
		needAck[n] = code == 255u;
		expectTag[n] = (stageBase + (needAck[n] ? 0u : code)) << 8;
		vs.vid[n] = vid[n];
		vs.remote[n] = nullptr;
		if (iface) { // at most one other copy (PartPlan::dataflowOk)
			const uint32_t b = __ldg(pd.shareStart + vid[n]), e = __ldg(pd.shareStart + vid[n] + 1);
			if (b < e) { vs.remote[n] = pd.peerXw[__ldg(pd.shareSlot + b)] + __ldg(pd.shareRemoteIdx + b); }
		}
	}
	VertexRegs v[4];
	bool ackOk[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
		v[n] = vs.LoadX(vid[n]);
		ackOk[n] = !needAck[n];
	}
	const uint32_t ackWant = stageBase & 0x00ffffffu;
	const ElemCompliance ec = ElemCompliance{ 0.0f, 0.0f, rec.alpha0, rec.alpha1 }; // DeviceScene::eAlpha (undamped: comp unused)
	for (uint32_t spins = 0;; spins++) {
		bool ok[4];
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if (!ackOk[n]) { ackOk[n] = LoadAckSys(pd.myAck + vid[n]) == ackWant; }
			ok[n] = ackOk[n] && (v[n].flags & kTagMask) == expectTag[n];
		}
		if (__all_sync(mask, ok[0] && ok[1] && ok[2] && ok[3])) { break; }
		if (SpinGiveUp(pd.errorFlag, nullptr, mask, spins, kPartSpinLimit)) { return false; } // report + drain, never trap
#pragma unroll
		for (int n = 0; n < 4; n++) {
			if ((v[n].flags & kTagMask) != expectTag[n]) { v[n] = vs.LoadX(vid[n]); }
		}
	}
	const uint32_t newTag = (stageBase + 1u + c) << 8;
#pragma unroll
	for (int n = 0; n < 4; n++) { v[n].flags = (v[n].flags & 0xffu) | newTag; }
	ElemRec r = rec;
	r.idx = make_uint4(vid[0], vid[1], vid[2], vid[3]);
	SolveElementGathered<ENERGY, SIMUL, EXACT, false>(vs, p, r, v, ec);
	return true;
}

template <int ENERGY, bool SIMUL, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_part_dataflow(const __grid_constant__ PartDevice pd, const __grid_constant__ SubstepParams p, uint32_t nSubsteps,
                                                          uint32_t verBase) {
	constexpr bool kPrefactored = (ENERGY == XF_ENERGY_MIXED_SEL || ENERGY == XF_ENERGY_YEOH_SKIN_FAST);
	const DeviceScene& sc = pd.local;
	const uint32_t lane = threadIdx.x & 31u;
	// Two warp pools (global warp id = round-robin over the CTAs).  The cross-GPU dependence chain runs through the
	// interface elements and the shared vertices: stage c on this rank waits for stage c-1 on the peer plus one NVLink
	// flight.  If the warps that carry it also had interior work queued in front of it, every stage of that chain would
	// wait for a whole local stage (measured: 8M tets on 2 GPUs no faster than on 1).  So a few warps carry nothing else.
	const uint32_t gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
	const uint32_t nW = gridDim.x * (blockDim.x >> 5);
	const uint32_t nI = min(pd.nIfaceWarps, nW / 2u);
	const bool ifacePool = gw < nI;
	const uint32_t poolSlot = (ifacePool ? gw : gw - nI) * 32u, poolStride = (ifacePool ? nI : nW - nI) * 32u;
	const uint32_t* vertList = ifacePool ? pd.sharedList : pd.privList;
	const uint32_t nVertList = ifacePool ? pd.nSharedList : pd.nPrivList;
	const uint32_t nC = pd.nColors;
	const uint32_t stride = nC + 1u;
	bool dead = false; // gave up on a stalled record: touch nothing any more, run out of the loops (the host reports XF_ERR_CUDA)
	for (uint32_t s = 0; s <= nSubsteps; s++) {
		const bool closing = s == nSubsteps;
		const uint32_t stageBase = verBase + s * stride;
		// vertex phase on every local copy
		for (uint32_t k0 = poolSlot; k0 < nVertList; k0 += poolStride) {
			const bool has = k0 + lane < nVertList && !dead;
			const unsigned mask = __ballot_sync(0xffffffffu, has);
			if (has) {
				const uint32_t i = __ldg(vertList + k0 + lane);
				VertexRegs v = LoadVertexSys(sc.Xw, i);
				if (s > 0) {
					const uint32_t expectTag = (stageBase - stride + (uint32_t)__ldg(sc.lastCode + i)) << 8;
					for (uint32_t spins = 0;; spins++) {
						const bool ok = (v.flags & kTagMask) == expectTag;
						if (__all_sync(mask, ok)) { break; }
						if (SpinGiveUp(pd.errorFlag, nullptr, mask, spins, kPartSpinLimit)) { dead = true; break; }
						if (!ok) { v = LoadVertexSys(sc.Xw, i); }
					}
				}
				if (dead) { continue; }
				VertexPhaseBody<EXACT>(sc, p, i, v, s > 0, !closing);
				v.flags = (v.flags & 0xffu) | (stageBase << 8);
				StoreVertexSys(sc.Xw, i, v);
				const uint32_t b = __ldg(pd.shareStart + i), e = __ldg(pd.shareStart + i + 1);
				if (b < e) { // tell the holder of the other copy that this copy is through the vertex phase of this stage
					__threadfence_system(); // the record above must be in place before the peer's element may overwrite it
					for (uint32_t k = b; k < e; k++) {
						uint32_t* dst = pd.peerAck[__ldg(pd.shareSlot + k)] + __ldg(pd.shareRemoteIdx + k);
						asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(stageBase & 0x00ffffffu) : "memory");
					}
				}
			}
		}
		if (closing) { break; }
		for (uint32_t c = 0; c < nC; c++) {
			const uint32_t mid = __ldg(pd.ifaceEnd + c);
			const uint32_t begin = ifacePool ? __ldg(pd.colorStart + c) : mid, end = ifacePool ? mid : __ldg(pd.colorStart + c + 1);
			for (uint32_t e0 = begin + poolSlot; e0 < end; e0 += poolStride) {
				const uint32_t e = e0 + lane;
				const bool has = e < end && !dead;
				const unsigned mask = __ballot_sync(0xffffffffu, has);
				if (has) {
					ElemRec rec;
					LoadElementFrom<kPrefactored, EXACT>(sc.eAd, sc, e, rec);
					const float2 al = __ldg(sc.eAlpha + e); // alpha of this call's settings, k_element_alpha
					rec.alpha0 = al.x;
					rec.alpha1 = al.y;
					dead = !PartDataflowElement<ENERGY, SIMUL, EXACT>(pd, p, rec, mask, ifacePool, stageBase, c);
				}
			}
		}
	}
}

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct PartDataflowRunner {
	static cudaError_t Run(const PartDevice& pd, const SubstepParams& p, uint32_t nSubsteps, uint32_t verBase, int smCount, cudaStream_t st,
	                       uint64_t* launches) {
		if (DAMPED) { return cudaErrorNotSupported; }
		auto fn = k_part_dataflow<ENERGY, SIMUL, EXACT>;
		int perSm = 0;
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 256, 0);
		if (e != cudaSuccess) { return e; }
		if (perSm < 1) { return cudaErrorLaunchOutOfResources; }
		const dim3 grid((unsigned)(std::min(perSm, 2) * smCount));
		void* args[] = { (void*)&pd, (void*)&p, (void*)&nSubsteps, (void*)&verBase };
		e = cudaLaunchCooperativeKernel((const void*)fn, grid, dim3(256), args, 0, st); // cooperative = co-resident CTAs
		++*launches;
		return e;
	}
};

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct PartPersistentRunner {
	static cudaError_t Run(const PartDevice& pd, const SubstepParams& p, uint32_t nSubsteps, unsigned long long* epoch, int smCount, cudaStream_t st,
	                       uint64_t* launches) {
		if (DAMPED) { return cudaErrorNotSupported; }
		auto fn = k_part_persistent<ENERGY, SIMUL, EXACT, false>;
		int perSm = 0;
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, 256, 0);
		if (e != cudaSuccess) { return e; }
		if (perSm < 1) { return cudaErrorLaunchOutOfResources; }
		const dim3 grid((unsigned)(std::min(perSm, 2) * smCount));
		unsigned long long epoch0 = *epoch;
		void* args[] = { (void*)&pd, (void*)&p, (void*)&nSubsteps, (void*)&epoch0 };
		e = cudaLaunchCooperativeKernel((const void*)fn, grid, dim3(256), args, 0, st);
		*epoch += (unsigned long long)nSubsteps * (pd.nColors + 1) + 1; // one epoch per vertex phase and per colour
		++*launches;
		return e;
	}
};

template <int ENERGY, bool SIMUL, bool EXACT, bool DAMPED>
struct PartRunner {
	// `globalColorStart`: colour ranges of the FULL mesh's serial order (identical on every rank): a damping sweep visits the colours
	// that meet the slice, and every rank must run the same sequence of phases.
	static cudaError_t Run(const PartDevice& pd, const SubstepParams& p, const std::vector<uint32_t>& colorStart, const std::vector<uint32_t>& ifaceEnd,
	                       const std::vector<uint32_t>& globalColorStart, uint32_t nTGlobal, uint32_t nSubsteps, unsigned long long* epoch,
	                       cudaStream_t st, uint64_t* launches) {
		const uint32_t nV = pd.local.nV;
		const dim3 vgrid((nV + 255) / 256);
		const uint32_t nC = (uint32_t)colorStart.size() - 1;
		const bool anyDamp = p.doDamp || p.doPbdDamp;
		auto blocksOf = [&](uint32_t c) { return dim3(std::max<uint32_t>(1u, (colorStart[c + 1] - colorStart[c] + 255) / 256)); }; // an empty colour still takes part
		k_part_sync_phase<<<1, 32, 0, st>>>(pd, ++*epoch);
		++*launches;
		for (uint32_t s = 0; s < nSubsteps; s++) {
			k_part_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(pd, p, (s > 0 && !anyDamp) ? 1 : 0, 1, ++*epoch);
			++*launches;
			for (uint32_t c = 0; c < nC; c++) {
				k_part_sweep<0, ENERGY, SIMUL, EXACT, DAMPED><<<blocksOf(c), 256, 0, st>>>(pd, p, colorStart[c], ifaceEnd[c], colorStart[c + 1], 0, 0, ++*epoch);
				++*launches;
			}
			for (uint32_t pass = 0; pass < p.volumePasses; pass++) {
				for (uint32_t c = 0; c < nC; c++) {
					k_part_sweep<1, ENERGY, SIMUL, EXACT, false><<<blocksOf(c), 256, 0, st>>>(pd, p, colorStart[c], ifaceEnd[c], colorStart[c + 1], 0, 0, ++*epoch);
					++*launches;
				}
			}
			if (anyDamp) {
				k_part_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(pd, p, 1, 0, ++*epoch);
				++*launches;
				uint32_t lo = 0, hi = nTGlobal;
				if (p.rayleigh == XF_RAYLEIGH_POST_AMORTIZED) { // Geo.cpp:794-797
					const uint32_t k = (p.tickId + s) % XF_AMORTIZATION_PERIOD;
					lo = (uint32_t)(((uint64_t)nTGlobal * k) / XF_AMORTIZATION_PERIOD);
					hi = (uint32_t)(((uint64_t)nTGlobal * (k + 1)) / XF_AMORTIZATION_PERIOD);
				}
				for (int sweep = 0; sweep < 2; sweep++) {
					if (sweep == 0 ? !p.doDamp : !p.doPbdDamp) { continue; }
					for (uint32_t c = 0; c < nC; c++) {
						if (!(globalColorStart[c] < hi && globalColorStart[c + 1] > lo)) { continue; } // the same verdict on every rank
						if (sweep == 0) {
							k_part_sweep<2, ENERGY, SIMUL, EXACT, false><<<blocksOf(c), 256, 0, st>>>(pd, p, colorStart[c], ifaceEnd[c], colorStart[c + 1], lo, hi, ++*epoch);
						} else {
							k_part_sweep<3, ENERGY, SIMUL, EXACT, false><<<blocksOf(c), 256, 0, st>>>(pd, p, colorStart[c], ifaceEnd[c], colorStart[c + 1], lo, hi, ++*epoch);
						}
						++*launches;
					}
				}
			}
		}
		if (!anyDamp) {
			k_part_vertex_phase<EXACT><<<vgrid, 256, 0, st>>>(pd, p, 1, 0, ++*epoch);
			++*launches;
		}
		return cudaGetLastError();
	}
};

}  // namespace xf

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace xf;

struct xf_partition {
	HostMesh mesh;    // the FULL mesh (every rank prepares it identically)
	PartPlan plan;
	PartDevice dev;
	int device = -1;
	int precision = XF_PRECISION_EXACT;
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	bool connected = false;
	unsigned long long epoch = 0;
	uint32_t verBase = 1; // first stage tag of the next barrier-free launch (24 bits, wraps; identical on all ranks)
	float alphaKey[3] = { 0.0f, 0.0f, 0.0f }; // (invMu, invLambda, dt2) DeviceScene::eAlpha was computed for
	bool alphaValid = false;
	uint64_t launches = 0;
	uint32_t groundOn = 0;
	float groundY = 0.0f, groundFriction = 0.0f;
	int schedule = XF_SCHEDULE_AUTO;
	int smCount = 0;
	uint32_t* dColorStart = nullptr;
	uint32_t* dIfaceEnd = nullptr;
	uint32_t* dShareStart = nullptr;
	uint32_t* dShareSlot = nullptr;
	uint32_t* dShareRemote = nullptr;
	uint32_t* dSharedList = nullptr;
	uint32_t* dPrivList = nullptr;
	double* dPackX = nullptr;
	double* dPackV = nullptr;
	float* dPackW = nullptr;
	std::vector<void*> openedPeers;
};

#define XFP_CUDA(call)                                         \
	do {                                                       \
		cudaError_t _e = (call);                               \
		if (_e != cudaSuccess) { return FailCuda(_e, #call); } \
	} while (0)

namespace {
template <typename T>
cudaError_t UploadVecP(T** dst, const std::vector<T>& src) {
	cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * std::max<size_t>(src.size(), 1));
	if (e != cudaSuccess) { return e; }
	return cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
}
struct IpcBlob { // 128 bytes moved between ranks by the host
	cudaIpcMemHandle_t xw;
	cudaIpcMemHandle_t flags;
};
static_assert(sizeof(IpcBlob) == 128, "IPC blob must be 128 bytes");

int UploadPart(xf_partition* P) {
	const HostMesh& m = P->mesh;
	const PartPlan& pl = P->plan;
	DeviceScene& d = P->dev.local;
	const uint32_t nV = (uint32_t)pl.verts.size(), nT = (uint32_t)pl.elems.size();
	d.nV = nV;
	d.nT = nT;
	d.nColors = (uint32_t)pl.colorStart.size() - 1;
	std::vector<VertexRec> xw(nV);
	std::vector<double4> x0(nV), zero(nV, double4{ 0, 0, 0, 0 });
	for (uint32_t i = 0; i < nV; i++) {
		const uint32_t g = pl.verts[i];
		xw[i] = VertexRec{ m.X0[3 * (size_t)g], m.X0[3 * (size_t)g + 1], m.X0[3 * (size_t)g + 2], m.w[g], (uint32_t)m.flags[g] };
		x0[i] = double4{ xw[i].x, xw[i].y, xw[i].z, 0.0 };
	}
	PackedElements pk;
	PackElements(m, pl.elems, pl.localIdx.data(), &pk);
	std::vector<uint32_t> slotOfRank(pl.nRanks, 0);
	for (size_t s = 0; s < pl.peers.size(); s++) { slotOfRank[pl.peers[s]] = (uint32_t)s; }
	std::vector<uint32_t> shareSlot(pl.sharePeerRank.size());
	for (size_t k = 0; k < shareSlot.size(); k++) { shareSlot[k] = slotOfRank[pl.sharePeerRank[k]]; }
	// positions and velocities in ONE allocation (V follows Xw): one IPC handle gives a peer both arrays
	XFP_CUDA(cudaMalloc((void**)&d.Xw, sizeof(VertexRec) * 2 * std::max(nV, 1u)));
	XFP_CUDA(cudaMemcpy(d.Xw, xw.data(), sizeof(VertexRec) * nV, cudaMemcpyHostToDevice));
	d.V = reinterpret_cast<double4*>(d.Xw + std::max(nV, 1u));
	XFP_CUDA(cudaMemset(d.V, 0, sizeof(double4) * std::max(nV, 1u)));
	XFP_CUDA(UploadVecP(&d.O, x0));
	XFP_CUDA(UploadVecP(&d.X0, x0));
	XFP_CUDA(UploadVecP(&d.eA, pk.a));
	XFP_CUDA(UploadVecP(&d.eB, pk.b));
	XFP_CUDA(UploadVecP(&d.eC, pk.c));
	XFP_CUDA(UploadVecP(&d.eArea, pk.area));
	{ // global serial position of every local element: the damping slices are ranges of the FULL mesh's serial order
		std::vector<uint32_t> serialPos(m.nT), canon(nT);
		for (uint32_t k = 0; k < m.nT; k++) { serialPos[m.order[k]] = k; }
		for (uint32_t k = 0; k < nT; k++) { canon[k] = serialPos[pl.elems[k]]; }
		XFP_CUDA(UploadVecP(&d.canonPos, canon));
	}
	XFP_CUDA(UploadVecP(&P->dShareStart, pl.shareStart));
	XFP_CUDA(UploadVecP(&P->dShareSlot, shareSlot));
	XFP_CUDA(UploadVecP(&P->dShareRemote, pl.shareRemoteIdx));
	P->dev.shareStart = P->dShareStart;
	P->dev.shareSlot = P->dShareSlot;
	P->dev.shareRemoteIdx = P->dShareRemote;
	// flag words (64 x u64) followed by one acknowledgement word per local vertex: one allocation, one IPC handle
	const size_t flagBytes = sizeof(unsigned long long) * 64 + sizeof(uint32_t) * std::max(nV, 1u);
	XFP_CUDA(cudaMalloc((void**)&P->dev.myFlags, flagBytes));
	XFP_CUDA(cudaMemset(P->dev.myFlags, 0, flagBytes));
	{ // slot 63 (no rank writes it: nRanks <= 63 flag slots are indexed by sender rank): this rank's local vertex count, so that a
	  // peer can find V behind Xw in the mapped allocation
		const unsigned long long nv = std::max(nV, 1u);
		XFP_CUDA(cudaMemcpy(P->dev.myFlags + 63, &nv, sizeof(nv), cudaMemcpyHostToDevice));
	}
	P->dev.myAck = reinterpret_cast<uint32_t*>(P->dev.myFlags + 64);
	if (pl.dataflowOk) {
		std::vector<uint32_t> sharedList, privList;
		for (uint32_t i = 0; i < nV; i++) { (pl.shareStart[i + 1] > pl.shareStart[i] ? sharedList : privList).push_back(i); }
		uint32_t maxIface = 0;
		for (size_t c = 0; c + 1 < pl.colorStart.size(); c++) { maxIface = std::max(maxIface, pl.ifaceEnd[c] - pl.colorStart[c]); }
		XFP_CUDA(UploadVecP(&P->dSharedList, sharedList));
		XFP_CUDA(UploadVecP(&P->dPrivList, privList));
		P->dev.sharedList = P->dSharedList;
		P->dev.privList = P->dPrivList;
		P->dev.nSharedList = (uint32_t)sharedList.size();
		P->dev.nPrivList = (uint32_t)privList.size();
		// one warp per interface chunk of the largest colour, and enough of them for the shared vertices' phase
		P->dev.nIfaceWarps = pl.peers.empty() ? 0u : std::max((maxIface + 31u) / 32u, ((uint32_t)sharedList.size() + 63u) / 64u);
		if (const char* env = getenv("XF_PART_IFACE_WARPS")) { P->dev.nIfaceWarps = (uint32_t)atoi(env); }
		// Measured on 4 x B200 (profiles/r2_part_release_ab_4gpu.log): the fence BREAKS the schedule - every mesh stalls (reported, never
		// wrong).  While this rank waits in the fence, the peer's successor has already seen our record in ITS copy, solved, and
		// stored its newer record into OUR copy - and our delayed local store then overwrites it.  The order "peer store, then the
		// local store a few cycles later" is what keeps both hazards closed (a same-rank successor cannot overtake a peer store issued
		// >= one L2 round trip + one solve earlier; the peer's answer needs two NVLink flights + a solve, the local store a few cycles):
		// a property of the timing, not of the memory model, and fail-stop when violated.  Off by default; the knob documents it.
		P->dev.releaseStores = 0;
		if (const char* env = getenv("XF_PART_RELEASE")) { P->dev.releaseStores = (uint32_t)atoi(env); }
		P->dev.storeGapNs = 0;
		if (const char* env = getenv("XF_PART_STORE_GAP_NS")) { P->dev.storeGapNs = (uint32_t)atoi(env); }
		if (!sharedList.empty()) { P->dev.nIfaceWarps = std::max(P->dev.nIfaceWarps, 1u); } // somebody must carry the interface
		std::vector<ElemRecA> ad = pk.a;
		for (size_t k = 0; k < ad.size(); k++) {
			for (int j = 0; j < 4; j++) { ad[k].idx[j] |= (uint32_t)pl.predCode[4 * k + j] << 24; }
		}
		XFP_CUDA(UploadVecP(&d.eAd, ad));
		XFP_CUDA(UploadVecP(&d.lastCode, pl.lastCode));
		XFP_CUDA(cudaMalloc((void**)&d.eAlpha, sizeof(float2) * std::max(nT, 1u)));
	}
	XFP_CUDA(cudaMalloc((void**)&P->dev.doneCounter, 2 * sizeof(unsigned int)));
	XFP_CUDA(cudaMemset(P->dev.doneCounter, 0, 2 * sizeof(unsigned int)));
	P->dev.errorFlag = P->dev.doneCounter + 1;
	XFP_CUDA(cudaMalloc((void**)&P->dPackX, sizeof(double) * 3 * std::max(nV, 1u)));
	XFP_CUDA(cudaMalloc((void**)&P->dPackV, sizeof(double) * 3 * std::max(nV, 1u)));
	XFP_CUDA(cudaMalloc((void**)&P->dPackW, sizeof(float) * std::max(nV, 1u)));
	XFP_CUDA(UploadVecP(&P->dColorStart, pl.colorStart));
	XFP_CUDA(UploadVecP(&P->dIfaceEnd, pl.ifaceEnd));
	P->dev.colorStart = P->dColorStart;
	P->dev.ifaceEnd = P->dIfaceEnd;
	P->dev.nColors = (uint32_t)pl.colorStart.size() - 1;
	P->dev.nPeers = (uint32_t)pl.peers.size();
	P->dev.myRank = pl.rank;
	for (size_t s = 0; s < pl.peers.size(); s++) { P->dev.peerRank[s] = pl.peers[s]; }
	return XF_OK;
}
}  // namespace

extern "C" {

int xf_part_create(const xf_create_params* params, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount,
                   uint32_t nRanks, uint32_t rank, xf_partition** outPart) {
	if (!params || !outPart) { return Fail(XF_ERR_INVALID, "null params/outPart"); }
	*outPart = nullptr;
	if (params->abiVersion != XF_ABI_VERSION) { return Fail(XF_ERR_INVALID, "xf_create_params.abiVersion mismatch"); }
	xf_partition* P = new (std::nothrow) xf_partition();
	if (!P) { return Fail(XF_ERR_NOMEM, "out of host memory"); }
	std::string err;
	int rc = PrepareMesh(nodeXYZ, nodeFloatCount, idxStream, idxCount, params->density, params->autoResize != 0, params->colorHint,
	                     params->colorHintCount, &P->mesh, &err);
	if (rc == XF_OK) { rc = BuildPartition(P->mesh, nRanks, rank, &P->plan, &err, params->partition); }
	if (rc != XF_OK) { delete P; return Fail(rc, err); }
	if (P->plan.peers.size() > (size_t)kMaxPeers) { delete P; return Fail(XF_ERR_UNSUPPORTED, "a rank may share vertices with at most 16 other ranks"); }
	memset(&P->dev, 0, sizeof(P->dev));
	P->device = params->device;
	P->precision = params->precision;
	if (P->device >= 0) {
		cudaError_t e = cudaSetDevice(P->device);
		if (e != cudaSuccess) { delete P; return FailCuda(e, "cudaSetDevice"); }
		cudaDeviceProp prop;
		e = cudaGetDeviceProperties(&prop, P->device);
		if (e != cudaSuccess) { delete P; return FailCuda(e, "cudaGetDeviceProperties"); }
		P->smCount = prop.multiProcessorCount;
		P->schedule = params->schedule;
		// Measured on 2 x B200 (998 250 tets): launch-per-phase 312 us/substep, persistent 390-426 us/substep; both are
		// bound by the cross-GPU release/acquire chain (system fence round trip + flag flight, ~10 us per phase), and the
		// back-to-back launches overlap it slightly better.  AUTO therefore means one launch per phase here.
		// The barrier-free schedule (versioned records mirrored by peer stores) removes that chain from the element path.
		if (P->schedule == XF_SCHEDULE_AUTO || P->schedule == XF_SCHEDULE_BRICKS) {
			// graph partitions: measured stalls of the barrier-free schedule at 8M tets on 4 GPUs (irregular interfaces under load,
			// profiles/r2_part_release_ab_4gpu.log) - AUTO keeps them on the flag protocol, which is ordered by construction
			const bool slabs = params->partition == XF_PARTITION_SLABS;
			P->schedule = (P->plan.dataflowOk && prop.cooperativeLaunch && slabs) ? XF_SCHEDULE_DATAFLOW : XF_SCHEDULE_LAUNCH_PER_COLOR;
		}
		if (P->schedule == XF_SCHEDULE_DATAFLOW && !P->plan.dataflowOk) {
			delete P;
			return Fail(XF_ERR_UNSUPPORTED, "XF_SCHEDULE_DATAFLOW on a partitioned mesh needs at most two copies of every vertex, <= 253 colours and <= 2^24 local vertices");
		}
		if (P->schedule >= XF_SCHEDULE_PERSISTENT && !prop.cooperativeLaunch) { P->schedule = XF_SCHEDULE_LAUNCH_PER_COLOR; }
		if (params->stream) { P->stream = (cudaStream_t)params->stream; }
		else {
			e = cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking);
			if (e != cudaSuccess) { delete P; return FailCuda(e, "cudaStreamCreate"); }
			P->ownStream = true;
		}
		rc = UploadPart(P);
		if (rc != XF_OK) { delete P; return rc; }
		if (P->plan.peers.empty()) { P->connected = true; }
	}
	*outPart = P;
	return XF_OK;
}

int xf_part_destroy(xf_partition* P) {
	if (!P) { return XF_OK; }
	if (P->device >= 0) {
		cudaSetDevice(P->device);
		if (P->stream) { cudaStreamSynchronize(P->stream); }
		for (void* p : P->openedPeers) { cudaIpcCloseMemHandle(p); }
		DeviceScene& d = P->dev.local;
		void* ptrs[] = { d.Xw, d.O, d.X0, d.eA, d.eB, d.eC, d.eArea, d.canonPos, d.eAd, d.lastCode, d.eAlpha, P->dSharedList, P->dPrivList, P->dShareStart, P->dShareSlot, P->dShareRemote, P->dColorStart, P->dIfaceEnd, P->dev.myFlags,
			             P->dev.doneCounter, P->dPackX, P->dPackV, P->dPackW };
		for (void* p : ptrs) { if (p) { cudaFree(p); } }
		if (P->ownStream && P->stream) { cudaStreamDestroy(P->stream); }
	}
	delete P;
	return XF_OK;
}

uint32_t xf_part_local_vert_count(const xf_partition* P) { return P ? (uint32_t)P->plan.verts.size() : 0; }
uint32_t xf_part_local_element_count(const xf_partition* P) { return P ? (uint32_t)P->plan.elems.size() : 0; }
uint32_t xf_part_peer_count(const xf_partition* P) { return P ? (uint32_t)P->plan.peers.size() : 0; }
uint32_t xf_part_color_count(const xf_partition* P) { return P ? (uint32_t)P->plan.colorStart.size() - 1 : 0; }
uint32_t xf_part_global_vert_count(const xf_partition* P) { return P ? P->mesh.nV : 0; }
uint32_t xf_part_global_element_count(const xf_partition* P) { return P ? P->mesh.nT : 0; }

int xf_part_get_local_verts(const xf_partition* P, uint32_t* localToGlobal) {
	if (!P || !localToGlobal) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(localToGlobal, P->plan.verts.data(), sizeof(uint32_t) * P->plan.verts.size());
	return XF_OK;
}
int xf_part_get_local_elements(const xf_partition* P, uint32_t* globalElementIds, uint32_t* colorStart) {
	if (!P) { return Fail(XF_ERR_INVALID, "null argument"); }
	if (globalElementIds) { memcpy(globalElementIds, P->plan.elems.data(), sizeof(uint32_t) * P->plan.elems.size()); }
	if (colorStart) { memcpy(colorStart, P->plan.colorStart.data(), sizeof(uint32_t) * P->plan.colorStart.size()); }
	return XF_OK;
}
int xf_part_get_peers(const xf_partition* P, uint32_t* peerRanks) {
	if (!P || !peerRanks) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(peerRanks, P->plan.peers.data(), sizeof(uint32_t) * P->plan.peers.size());
	return XF_OK;
}
int xf_part_get_halo(const xf_partition* P, uint32_t color, uint32_t peerSlot, int send, uint32_t* count, uint32_t* localVerts) {
	if (!P || !count) { return Fail(XF_ERR_INVALID, "null argument"); }
	const size_t nPeers = P->plan.peers.size();
	if (color + 1 >= P->plan.colorStart.size() || peerSlot >= nPeers) { return Fail(XF_ERR_INVALID, "colour / peer slot out of range"); }
	const std::vector<uint32_t>& start = send ? P->plan.sendStart : P->plan.recvStart;
	const std::vector<uint32_t>& verts = send ? P->plan.sendVerts : P->plan.recvVerts;
	const size_t k = color * nPeers + peerSlot;
	*count = start[k + 1] - start[k];
	if (localVerts) { memcpy(localVerts, verts.data() + start[k], sizeof(uint32_t) * *count); }
	return XF_OK;
}
int xf_part_get_order(const xf_partition* P, uint32_t* fullOrder) {
	if (!P || !fullOrder) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(fullOrder, P->mesh.order.data(), sizeof(uint32_t) * P->mesh.nT);
	return XF_OK;
}
int xf_part_get_global_color_start(const xf_partition* P, uint32_t* colorStart) {
	if (!P || !colorStart) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(colorStart, P->mesh.colorStart.data(), sizeof(uint32_t) * P->mesh.colorStart.size());
	return XF_OK;
}
int xf_part_get_initial(const xf_partition* P, float* w, uint8_t* flags) { // local vertices: global inverse masses / flags
	if (!P) { return Fail(XF_ERR_INVALID, "null argument"); }
	for (size_t i = 0; i < P->plan.verts.size(); i++) {
		if (w) { w[i] = P->mesh.w[P->plan.verts[i]]; }
		if (flags) { flags[i] = P->mesh.flags[P->plan.verts[i]]; }
	}
	return XF_OK;
}

// Stage codes of the barrier-free schedule (xf_partition.h): 4 per local element, 1 per local vertex; returns whether the
// job qualifies for it (the same verdict on every rank).
int xf_part_get_dataflow_codes(const xf_partition* P, uint8_t* predCode4, uint8_t* lastCode, int* outOk) {
	if (!P) { return Fail(XF_ERR_INVALID, "null argument"); }
	if (predCode4) { memcpy(predCode4, P->plan.predCode.data(), P->plan.predCode.size()); }
	if (lastCode) { memcpy(lastCode, P->plan.lastCode.data(), P->plan.lastCode.size()); }
	if (outOk) { *outOk = P->plan.dataflowOk ? 1 : 0; }
	return XF_OK;
}

int xf_part_ipc_export(xf_partition* P, void* out128) {
	if (!P || !out128) { return Fail(XF_ERR_INVALID, "null argument"); }
	if (P->device < 0) { return Fail(XF_ERR_CUDA, "host-only partition has no device memory"); }
	XFP_CUDA(cudaSetDevice(P->device));
	IpcBlob blob;
	XFP_CUDA(cudaIpcGetMemHandle(&blob.xw, P->dev.local.Xw));
	XFP_CUDA(cudaIpcGetMemHandle(&blob.flags, P->dev.myFlags));
	memcpy(out128, &blob, sizeof(blob));
	return XF_OK;
}

// allRanks: nRanks consecutive 128-byte blobs, blob r from rank r's xf_part_ipc_export.
int xf_part_ipc_connect(xf_partition* P, const void* allRanks) {
	if (!P || !allRanks) { return Fail(XF_ERR_INVALID, "null argument"); }
	if (P->device < 0) { return Fail(XF_ERR_CUDA, "host-only partition has no device memory"); }
	XFP_CUDA(cudaSetDevice(P->device));
	const IpcBlob* blobs = (const IpcBlob*)allRanks;
	for (size_t s = 0; s < P->plan.peers.size(); s++) {
		const IpcBlob& b = blobs[P->plan.peers[s]];
		void* xw = nullptr;
		void* fl = nullptr;
		XFP_CUDA(cudaIpcOpenMemHandle(&xw, b.xw, cudaIpcMemLazyEnablePeerAccess));
		XFP_CUDA(cudaIpcOpenMemHandle(&fl, b.flags, cudaIpcMemLazyEnablePeerAccess));
		P->openedPeers.push_back(xw);
		P->openedPeers.push_back(fl);
		P->dev.peerXw[s] = (VertexRec*)xw;
		P->dev.peerFlags[s] = (unsigned long long*)fl;
		unsigned long long peerNV = 0; // the peer's local vertex count (flag slot 63): its V array follows its Xw array
		XFP_CUDA(cudaMemcpy(&peerNV, (unsigned long long*)fl + 63, sizeof(peerNV), cudaMemcpyDeviceToHost));
		P->dev.peerV[s] = reinterpret_cast<double4*>((VertexRec*)xw + peerNV);
		P->dev.peerAck[s] = reinterpret_cast<uint32_t*>((unsigned long long*)fl + 64);
	}
	P->connected = true;
	return XF_OK;
}

int xf_part_set_ground(xf_partition* P, int enabled, float y0, float friction) {
	if (!P) { return Fail(XF_ERR_INVALID, "null partition"); }
	P->groundOn = enabled ? 1u : 0u;
	P->groundY = y0;
	P->groundFriction = friction;
	return XF_OK;
}

int xf_part_substep(xf_partition* P, const xf_settings* st, float dt, uint32_t n) {
	if (!P || !st) { return Fail(XF_ERR_INVALID, "null argument"); }
	if (P->device < 0) { return Fail(XF_ERR_CUDA, "host-only partition: there is no CPU compute path"); }
	if (!P->connected) { return Fail(XF_ERR_INVALID, "xf_part_ipc_connect has not been called"); }
	if (n == 0) { return XF_OK; }
	XFP_CUDA(cudaSetDevice(P->device));
	SubstepParams p;
	std::string err;
	int rc = FillSubstepParams(st, nullptr, dt, P->mesh, &p, &err);
	if (rc != XF_OK) { return Fail(rc, err); }
	// damping (in-constraint or sweeps) and volume passes run on the flag protocol, whatever the schedule of the plain sweep:
	// one launch per phase, velocities of shared vertices mirrored like positions
	const bool inConstraint = p.damping > 0.0f && p.rayleigh < XF_RAYLEIGH_POST;
	const bool general = inConstraint || p.doDamp || p.doPbdDamp || p.volumePasses;
	if (P->plan.nRanks > 63) { return Fail(XF_ERR_UNSUPPORTED, "at most 63 ranks"); }
	p.groundOn = P->groundOn;
	p.groundY = P->groundY;
	p.groundKeep = 1.0f - P->groundFriction;
	p.handleCount = 0;
	if (general) {
		XFP_CUDA(DispatchConfig<PartRunner>(p.energy, p.simultaneous != 0, P->precision == XF_PRECISION_EXACT, inConstraint, P->dev, p, P->plan.colorStart,
		                                    P->plan.ifaceEnd, P->mesh.colorStart, P->mesh.nT, n, &P->epoch, P->stream, &P->launches));
	} else if (P->schedule == XF_SCHEDULE_DATAFLOW) {
		if (!(P->alphaValid && P->alphaKey[0] == p.invMu && P->alphaKey[1] == p.invLambda && P->alphaKey[2] == p.dt2)) {
			XFP_CUDA(LaunchElementAlpha(P->dev.local, p, P->precision == XF_PRECISION_EXACT, P->stream, &P->launches));
			P->alphaKey[0] = p.invMu; P->alphaKey[1] = p.invLambda; P->alphaKey[2] = p.dt2;
			P->alphaValid = true;
		}
		const uint32_t stride = P->dev.nColors + 1u;
		const uint32_t maxPerLaunch = (0x00ffffffu - 2u) / stride;
		for (uint32_t done = 0; done < n;) {
			const uint32_t m = std::min(n - done, maxPerLaunch);
			XFP_CUDA(DispatchConfig<PartDataflowRunner>(p.energy, p.simultaneous != 0, P->precision == XF_PRECISION_EXACT, false, P->dev, p, m, P->verBase,
			                                            P->smCount, P->stream, &P->launches));
			P->verBase = (P->verBase + m * stride + 1u) & 0x00ffffffu;
			done += m;
		}
	} else if (P->schedule == XF_SCHEDULE_PERSISTENT) {
		XFP_CUDA(DispatchConfig<PartPersistentRunner>(p.energy, p.simultaneous != 0, P->precision == XF_PRECISION_EXACT, false, P->dev, p, n, &P->epoch,
		                                              P->smCount, P->stream, &P->launches));
	} else {
		XFP_CUDA(DispatchConfig<PartRunner>(p.energy, p.simultaneous != 0, P->precision == XF_PRECISION_EXACT, false, P->dev, p, P->plan.colorStart,
		                                    P->plan.ifaceEnd, P->mesh.colorStart, P->mesh.nT, n, &P->epoch, P->stream, &P->launches));
	}
	return XF_OK;
}

int xf_part_sync(xf_partition* P) {
	if (!P || P->device < 0) { return Fail(XF_ERR_INVALID, "null or host-only partition"); }
	XFP_CUDA(cudaSetDevice(P->device));
	XFP_CUDA(cudaStreamSynchronize(P->stream));
	unsigned int flag = 0;
	XFP_CUDA(cudaMemcpy(&flag, P->dev.errorFlag, sizeof(flag), cudaMemcpyDeviceToHost));
	if (flag) { return Fail(XF_ERR_CUDA, "a peer did not reach the expected phase in time (halo protocol error)"); }
	return XF_OK;
}

int xf_part_get_state(xf_partition* P, double* X, double* V, float* w) {
	if (!P || P->device < 0) { return Fail(XF_ERR_INVALID, "null or host-only partition"); }
	int rc = xf_part_sync(P);
	if (rc != XF_OK) { return rc; }
	const uint32_t nV = (uint32_t)P->plan.verts.size();
	uint64_t dummy = 0;
	XFP_CUDA(LaunchPackState(P->dev.local, X ? P->dPackX : nullptr, V ? P->dPackV : nullptr, w ? P->dPackW : nullptr, P->stream, &dummy));
	if (X) { XFP_CUDA(cudaMemcpyAsync(X, P->dPackX, sizeof(double) * 3 * nV, cudaMemcpyDeviceToHost, P->stream)); }
	if (V) { XFP_CUDA(cudaMemcpyAsync(V, P->dPackV, sizeof(double) * 3 * nV, cudaMemcpyDeviceToHost, P->stream)); }
	if (w) { XFP_CUDA(cudaMemcpyAsync(w, P->dPackW, sizeof(float) * nV, cudaMemcpyDeviceToHost, P->stream)); }
	XFP_CUDA(cudaStreamSynchronize(P->stream));
	return XF_OK;
}

int xf_part_get_info(const xf_partition* P, uint64_t* launches, uint64_t* epoch) {
	if (!P) { return Fail(XF_ERR_INVALID, "null partition"); }
	if (launches) { *launches = P->launches; }
	if (epoch) { *epoch = P->epoch; }
	return XF_OK;
}

}  // extern "C"
