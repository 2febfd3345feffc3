// Internal types shared by the host side (xf_api.cpp, xf_prepare.cpp) and the kernels (xf_kernels.cu).
// Nothing here crosses the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/xpbd_fem_b200.h"

namespace xf {

// ------------------------------------------------------------------------------------------------
// HBM layout
//
// Vertices (AoS, one 32-byte L2 sector per vertex so an element gather costs exactly 4 sectors):
//   VertexRec  Xw[nV]   { double x, y, z; float w; uint32 flags }   position + inverse mass + lock flags
//   double4    O[nV], V[nV], X0[nV]                                  (.w unused)
// Elements: three planes sorted by colour then by original index.  The sweep is bound by L2 *requests*, so a
// record is fetched with two 256-bit loads (+ one 128-bit load when the reference's own prefactored
// coefficients are needed), each lane reading whole 32-byte sectors:
//   ElemRecA eA[nT]   32 B   idx[4]; Qi[0][0] Qi[0][1] Qi[0][2] Qi[1][0]      (Qi column-major like mat3)
//   ElemRecB eB[nT]   32 B   Qi[1][1] Qi[1][2] Qi[2][0] Qi[2][1] Qi[2][2]; volume; QQ0 QQ1
//   float4   eC[nT]   16 B   QQ2 QR0 QR1 QR2         (XF_PRECISION_EXACT + prefactored energies only)
//   float    eArea[nT]       surfaceArea (PbdDamp only)
// => 64 B (fast / non-prefactored) or 80 B (exact prefactored) streamed per element per sweep.
// ------------------------------------------------------------------------------------------------
struct alignas(32) VertexRec {
	double x, y, z;
	float w;
	uint32_t flags;
};

struct alignas(32) ElemRecA {
	uint32_t idx[4];
	float q[4];
};
struct alignas(32) ElemRecB {
	float q[5];
	float volume;
	float qq01[2];
};

struct DeviceScene {
	uint32_t nV = 0, nT = 0, nColors = 0;
	VertexRec* Xw = nullptr;
	double4* O = nullptr;
	double4* V = nullptr;
	double4* X0 = nullptr;
	ElemRecA* eA = nullptr;
	ElemRecB* eB = nullptr;
	float4* eC = nullptr;
	float* eArea = nullptr;
	float* eScratch = nullptr;      // nT floats (per-element volume terms)
	double* statScratch = nullptr;  // reduction outputs
	uint32_t* streamToSorted = nullptr; // nT: position of stream element s in the colour-sorted planes
	uint32_t* canonPos = nullptr;       // nT: position in the serial order (xf_get_order) of the element at device position k (identity on one GPU; global position on a partition)
	unsigned int* barrier = nullptr; // grid-barrier counter
	// Stall report of the barrier-free schedules.  A record that never reaches the expected stage (a broken schedule, or a caller
	// that rewrote the state under a running launch) must not hang the device and must not kill the CUDA context either: the warp
	// that gives up sets both words, every other waiting warp sees errDev within 1024 polls, all of them stop working and the
	// kernel drains.  errHost is pinned host memory (device alias here), so the host reads the verdict after a plain stream sync.
	unsigned int* errDev = nullptr;
	unsigned int* errHost = nullptr;
	uint32_t spinLimit = 1u << 24; // polls before a waiting warp gives up: seconds (a healthy wait is a few polls)
	// dataflow schedule (XF_SCHEDULE_DATAFLOW, xf_dataflow.cu): eA whose vertex indices carry, in their top byte, the stage
	// code of the previous writer of that vertex (0 = vertex phase, 1 + colour otherwise); per vertex the code of its last writer
	ElemRecA* eAd = nullptr;
	uint8_t* lastCode = nullptr;
	// device vertex id -> caller's vertex id (nullptr = identity); used where state crosses the ABI
	uint32_t* extOfInt = nullptr;
	// clustered dataflow (k_substeps_cluster) and chained dataflow (k_substeps_chain): per element the slot / first / last
	// bits of its corners
	uint32_t* eK = nullptr;
	uint32_t groupSize = 0;
	uint32_t chained = 0;       // eK holds ChainInfo words (groupSize <= 1)
	uint32_t maxColorSize = 0;  // the chained kernel needs every colour to fit one wave of the co-resident grid
	uint32_t dataflowBlock = 0; // threads per CTA of the barrier-free kernels (0 = DataflowBlockThreads picks)
	// damping sweeps on the barrier-free schedule (k_substeps_dataflow_general): V records are versioned by a write COUNT.
	// eRank: per element 4 x 8 bits, the rank of the element among the elements around each corner's vertex (serial order);
	// vSlice: per vertex 8 x 8 bits, the number of elements around it whose serial position is below nT*q/8, q = 1..8
	// (the amortised damping slices of Geo.cpp:794-797; byte 7 = the valence)
	uint32_t* eRank = nullptr;
	uint2* vSlice = nullptr;
	// alpha = {1/mu/volume/dt^2, 1/lambda/volume/dt^2} per element (Fem.cpp:449, Xpbd.h:154) for the settings of the running call:
	// two chained IEEE divisions per element and substep (~25 of ~700 instructions) that depend on (compliance, nu, dt) only, so
	// k_element_alpha evaluates them once per change of those three - same operations, same bits - and the barrier-free kernels
	// load 8 bytes instead
	float2* eAlpha = nullptr;
};

// Threads per CTA of the barrier-free kernels.  Work is dealt one element per thread and colour, so only
// ceil(maxColorSize / 32) warps of the grid ever run an element; the others only take part in the vertex phase and poll for
// it the rest of the time (L2 traffic and issue slots next to the warps on the dependence chain).  Measured on B200
// (tools/ab_check.sh, profiles/r1_results.md): with no spare warps 384k tets run 9.7 % faster (64 instead of 256 threads:
// 500 of 2368 warps had work), 24k tets 0.7 % faster; at 1M tets (1300 of 2368 warps busy) 160 threads are 0.8 % SLOWER than
// 256 - the spare warps shorten the vertex phase by one round.  So: no spare warps while less than half of a 256-thread
// grid would have work, else 256.  The grid keeps its CTA count (co-resident, one wave per colour when it fits).
inline int DataflowBlockThreads(const DeviceScene& sc, int gridBlocks) {
	if (sc.dataflowBlock >= 32 && sc.dataflowBlock <= 256) { return (int)(sc.dataflowBlock & ~31u); }
	if (gridBlocks <= 0) { return 256; }
	const uint64_t perCta = ((uint64_t)sc.maxColorSize + (uint64_t)gridBlocks - 1) / (uint64_t)gridBlocks;
	const uint64_t threads = ((perCta + 31) / 32) * 32;
	return (int)(threads < 64 ? 64 : (threads > 128 ? 256 : threads));
}

constexpr int kMaxHandles = 64;
constexpr int kMaxColors = 256;

// Everything one xf_substep call needs, passed to the kernels by value (__grid_constant__).
// All derived fp32 constants are computed on the host with the reference's expressions
// (file compiled with -ffp-contract=off) so the device sees the very same bits.
struct SubstepParams {
	float dt, dt2, invDt;      // dt, dt*dt, 1.0f/dt
	float gdtX, gdtY;          // gravity.x*dt, gravity.y*dt                       Geo.cpp:308
	float keep;                // 1.0f - timeCorrectedDrag                         Geo.cpp:309
	float invMu, invLambda, a; // 1.0f/mu, 1.0f/lambda, 1.0f + mu/lambda           Fem.cpp:445-449
	float damping;             // settings.damping                                 Fem.cpp:450
	float compliance;          // settings.compliance (volume passes, Fem.cpp:844)
	float pbdDamping;          // volumeAndTimeCorrectedPbdDamping to use this call
	float dampDamping;         // damping used by the post-solve Damp sweep (x8 when amortised, Geo.cpp:349)
	uint32_t energy;           // XF_ENERGY_*
	uint32_t simultaneous;     // Settings_XpbdSolveBit
	uint32_t rayleigh;         // XF_RAYLEIGH_*
	uint32_t lockLeft, lockRight;
	uint32_t volumePasses;
	uint32_t tickId;           // of the first substep of the call
	uint32_t doDamp, doPbdDamp;
	float lockT[12];           // lockedRightTransform3d, padded columns
	float origin[3];
	uint32_t groundOn;
	float groundY, groundKeep; // y0, 1.0f - friction
	uint32_t manipOn, manipIdx;
	float manipTarget[3];
	float c18;                 // 1.8f / (dt*dt)                                   Geo.cpp:338
	uint32_t handleCount;
	uint32_t handleIdx[kMaxHandles];
	float handleTarget[kMaxHandles][3];
	uint32_t colorStart[kMaxColors + 1];
	uint32_t nColors;
	// Parameters that change from substep to substep inside one call (xf_substep_varying: an animated right-side lock, a moving
	// manipulator ray - Sim::Update's per-substep work, Demo.cpp:67-90): kVaryFloats floats per substep in device memory,
	// {lockT[12], manipTarget[3], pad}; nullptr = lockT / manipTarget above hold for the whole call.
	const float* vary;
};
constexpr int kVaryFloats = 16;

// Per-scene constants of a batched call (xf_batch.cu): the scalar part of SubstepParams.  The element math is
// templated on the parameter type and only touches fields that exist in both.
struct SceneConsts {
	float dt, dt2, invDt;
	float gdtX, gdtY;
	float keep;
	float invMu, invLambda, a;
	float damping;
	float compliance;
	float pbdDamping;
	float dampDamping;
	uint32_t rayleigh;
	uint32_t lockLeft, lockRight;
	uint32_t volumePasses;
	uint32_t tickId;
	uint32_t doDamp, doPbdDamp;
	float lockT[12];
	float origin[3];
	uint32_t groundOn;
	float groundY, groundKeep;
};

struct LaunchShape {
	int blockThreads = 256;
	int gridBlocks = 0;     // persistent grid (co-resident)
	int smCount = 0;
};

// ---- host-side prepared mesh (xf_prepare.cpp) ----
struct HostMesh {
	uint32_t nV = 0, nT = 0;
	std::vector<double> X0;       // 3*nV, after autoResize
	std::vector<float> w;         // inverse lumped masses
	std::vector<uint8_t> flags;
	std::vector<uint32_t> idx;    // 4*nT, stream order
	std::vector<float> Qi, QQ, QR, volume, area; // 9/3/3/1/1 per element, stream order
	std::vector<uint32_t> color;  // per element (stream order)
	std::vector<uint32_t> order;  // equivalent serial order (colour-major, index-minor)
	std::vector<uint32_t> colorStart; // nColors + 1 offsets into `order`
	float origin[3] = { 0.0f, 0.0f, 0.0f };
	// clustered colouring (PrepareMesh with clusterVerts > 0): colours come in groups of `groupSize`; element k of colour
	// C*groupSize + t is the t-th element of cluster k of cluster-colour C, and a cluster's elements touch at most 8
	// distinct vertices.  clusterInfo (stream order): per corner n, bits [5n, 5n+5) = slot (3 bits) | first use in the
	// cluster (bit 3) | last use in the cluster (bit 4).
	uint32_t groupSize = 0;
	std::vector<uint32_t> clusterInfo;
};

// Returns 0 or an xf_status; on failure `err` holds the message.
int PrepareMesh(const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount, float density,
                bool autoResize, const uint32_t* colorHint, uint32_t colorHintCount, HostMesh* out, std::string* err, bool clustered = false);
// Element planes for the elements `elems` (global ids, in device order); `localIdx` (4 per element) replaces the
// mesh's vertex ids when the device uses a local vertex numbering (partitioned meshes), else nullptr.
struct PackedElements {
	std::vector<ElemRecA> a;
	std::vector<ElemRecB> b;
	std::vector<float4> c;
	std::vector<float> area;
};
// Stage codes of the barrier-free schedule for the elements in `order` (colour-major): pred = 4 per element, the code of the
// previous writer of each corner (0 = the substep's vertex phase, else 1 + colour); last = per vertex (caller's numbering)
// the code of its last writer.
void StageCodes(const HostMesh& mesh, const std::vector<uint32_t>& order, std::vector<uint8_t>* pred, std::vector<uint8_t>* last);
// Slot / first / last bits for the chained sweep (xf_prepare.cpp); `pred` as produced by StageCodes for the same `order`.
uint64_t ChainInfo(const HostMesh& mesh, const std::vector<uint32_t>& order, const std::vector<uint8_t>& pred, std::vector<uint32_t>* info);
void PackElements(const HostMesh& mesh, const std::vector<uint32_t>& elems, const uint32_t* localIdx, PackedElements* out);
int FillSubstepParams(const xf_settings* st, const xf_manipulator* manip, float dt, const HostMesh& mesh, SubstepParams* p,
                      std::string* err);

// Sim::Update over several scenes (xf_frame.cpp); xf_frame_update and xf_sim_update are thin wrappers
int FrameUpdateGeos(xf_scene* const* scenes, uint32_t count, int pickedGeo, xf_settings* settings, xf_manipulator* manip, float dt,
                    float medianFrameTime, xf_frame_state* state, uint32_t* outSubsteps);

// error reporting shared by the ABI translation units (xf_api.cpp)
int Fail(int status, const std::string& msg);
int FailCuda(cudaError_t e, const char* what);

// ---- kernel launchers (xf_kernels.cu) ----
cudaError_t QueryLaunchShape(int device, uint32_t energy, bool exact, LaunchShape* shape);
cudaError_t LaunchSubstepsPerColor(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, cudaStream_t stream,
                                   uint64_t* launchCount);
cudaError_t LaunchSubstepsDataflow(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                   uint32_t sleepNs, cudaStream_t stream, uint64_t* launchCount);
// Falls back to LaunchSubstepsDataflow when a colour does not fit one wave of the grid.
cudaError_t LaunchSubstepsChain(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                uint32_t tuning, cudaStream_t stream, uint64_t* launchCount);
// Volume passes and post-solve damping sweeps (Rayleigh_Post / PostAmortized, PbdDamp) without barriers; `vEpoch` numbers the
// substeps of the scene's life (V-record tags), `stride` = stages per substep (the caller advances verBase by n * stride + 1).
uint32_t DataflowGeneralStride(const SubstepParams& p);
cudaError_t LaunchSubstepsDataflowGeneral(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                          uint64_t vEpoch, uint32_t tuning, cudaStream_t stream, uint64_t* launchCount);
cudaError_t LaunchSubstepsCluster(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, int smCount, uint32_t verBase,
                                  uint32_t tuning, cudaStream_t stream, uint64_t* launchCount);
cudaError_t LaunchSubstepsPersistent(const DeviceScene& sc, const SubstepParams& p, bool exact, uint32_t nSubsteps, const LaunchShape& shape,
                                     cudaStream_t stream, uint64_t* launchCount);
cudaError_t LaunchElementAlpha(const DeviceScene& sc, const SubstepParams& p, bool exact, cudaStream_t stream, uint64_t* launchCount);  // -> sc.eAlpha
cudaError_t LaunchElementVolumes(const DeviceScene& sc, cudaStream_t stream, uint64_t* launchCount);  // -> sc.eScratch, stream order
cudaError_t LaunchTransform(const DeviceScene& sc, const float* m9, cudaStream_t stream, uint64_t* launchCount);
cudaError_t LaunchStats(const DeviceScene& sc, const SubstepParams& p, double gx, double gy, int smCount, cudaStream_t stream,
                        uint64_t* launchCount);  // -> sc.statScratch[6]
cudaError_t LaunchPackState(const DeviceScene& sc, double* dX, double* dV, float* dW, cudaStream_t stream, uint64_t* launchCount);
cudaError_t LaunchUnpackState(const DeviceScene& sc, const double* dX, const double* dV, const float* dW, cudaStream_t stream,
                              uint64_t* launchCount);

}  // namespace xf
