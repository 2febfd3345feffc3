// C ABI of libxpbd_fem_b200.so (include/xpbd_fem_b200.h): scene lifetime, uploads, stepping, state access.
// Host C++ only; every device operation goes through the launchers in xf_kernels.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>

#include "xf_scene.h"

using namespace xf;

namespace {
thread_local std::string g_lastError;
}

namespace xf {
int Fail(int status, const std::string& msg) {
	g_lastError = msg;
	return status;
}
int FailCuda(cudaError_t e, const char* what) {
	g_lastError = std::string(what) + ": " + cudaGetErrorString(e);
	return XF_ERR_CUDA;
}
}  // namespace xf

namespace {
#define XF_CUDA(call)                                                \
	do {                                                             \
		cudaError_t _e = (call);                                     \
		if (_e != cudaSuccess) { return FailCuda(_e, #call); }       \
	} while (0)

// XF_GROUPING_AUTO on the barrier-free schedule: chained sweep or plain one element per thread
constexpr bool kChainByDefault = false;
constexpr uint32_t kChainMinPermille = 100;

template <typename T>
cudaError_t Upload(T** dst, const std::vector<T>& src) {
	cudaError_t e = cudaMalloc((void**)dst, sizeof(T) * std::max<size_t>(src.size(), 1));
	if (e != cudaSuccess) { return e; }
	return cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
}

}  // namespace

struct xf_scene {
	HostMesh mesh;
	DeviceScene dev;
	int device = -1;          // < 0: host-only scene (introspection of init/colouring; cannot step)
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	int precision = XF_PRECISION_EXACT;
	int schedule = XF_SCHEDULE_AUTO;
	bool cooperative = false;
	std::map<uint32_t, LaunchShape> shapes; // per energy
	bool dataflowOk = false;  // stage codes fit the vertex-index top byte / the 24-bit record tag
	uint32_t verBase = 1;     // first stage tag of the next dataflow launch (24-bit, wraps)
	uint32_t spinSleepNs = 400; // back-off of the vertex-phase spin (XF_DATAFLOW_SLEEP_NS)
	bool wantChain = false;   // XF_GROUPING_CHAINS (or AUTO resolved to it)
	std::vector<uint32_t> chainInfo; // ChainInfo words in serial-order positions (empty: not chained)
	uint32_t chainedPermille = 0;
	std::vector<uint32_t> intOfExt, extOfInt; // caller's vertex id <-> device vertex id
	int smCount = 0;
	size_t l2Bytes = 0;
	uint64_t launches = 0;
	uint32_t lastKernel = 0;  // xf_kernel_id of the last stepping launch
	uint64_t vEpoch = 1;      // substeps stepped by k_substeps_dataflow_general so far (V-record tags)
	float alphaKey[3] = { 0.0f, 0.0f, 0.0f }; // (invMu, invLambda, dt2) the alpha plane (DeviceScene::eAlpha) was computed for
	bool alphaValid = false;
	float* dVary = nullptr;   // per-substep parameter rows of xf_substep_varying (kVaryFloats floats each)
	size_t dVaryRows = 0;
	// extensions
	uint32_t groundOn = 0;
	float groundY = 0.0f, groundFriction = 0.0f;
	uint32_t handleCount = 0;
	uint32_t handleIdx[kMaxHandles];
	float handleTarget[kMaxHandles][3];
	// staging for packed state transfers
	double* dPackX = nullptr;
	double* dPackV = nullptr;
	float* dPackW = nullptr;
	std::vector<float> hostScratch;
	volatile unsigned int* stallWord = nullptr; // pinned host word a barrier-free kernel sets when it gives up on a record
};

namespace xf {
// xf_prepare.cpp: write-count codes of the damping sweeps on the barrier-free schedule (DeviceScene::eRank / vSlice); `idxOfPos` = the
// four vertex ids of the element at every serial position
void DampingCodes(uint32_t nT, uint32_t nV, const uint32_t* idxOfPos, std::vector<uint32_t>* rank, std::vector<uint64_t>* below);
}

namespace {

void FreeDevice(xf_scene* s) {
	if (s->device < 0) { return; }
	cudaSetDevice(s->device);
	DeviceScene& d = s->dev;
	void* ptrs[] = { d.Xw, d.O, d.V, d.X0, d.eA, d.eB, d.eC, d.eArea, d.eAd, d.lastCode, d.eRank, d.vSlice, d.eAlpha, d.eK, d.extOfInt, d.canonPos, d.eScratch, d.statScratch, d.streamToSorted,
		             d.barrier, s->dPackX, s->dPackV, s->dPackW, s->dVary };
	for (void* p : ptrs) { if (p) { cudaFree(p); } }
	if (s->stallWord) { cudaFreeHost((void*)s->stallWord); }
	if (s->ownStream && s->stream) { cudaStreamDestroy(s->stream); }
}

int UploadScene(xf_scene* s) {
	const HostMesh& m = s->mesh;
	DeviceScene& d = s->dev;
	d.nV = m.nV;
	d.nT = m.nT;
	d.nColors = (uint32_t)m.colorStart.size() - 1;
	// element planes: colour-major, inside a colour in stream order = the serial order of xf_get_order
	const std::vector<uint32_t>& deviceOrder = m.order;
	// Device vertex numbering.  A warp handles 32 consecutive elements of one colour and its lanes gather/scatter corner n
	// of their element with one 256-bit access each; every distinct 128-byte line costs one L1TEX wavefront, and the sweep's
	// per-element cost IS those wavefronts (~276 per warp-element when every lane hits its own line).  Numbering the vertices
	// in first-touch order (colour-major, corner-major, element-minor) makes consecutive lanes touch consecutive records
	// wherever the mesh allows it (exactly so on MeshGen lattices: four interleaved row-ordered classes), i.e. four lanes
	// per line.  The caller's numbering is restored at the ABI (pack/unpack kernels, handle and manipulator indices).
	s->intOfExt.assign(m.nV, 0xffffffffu);
	s->extOfInt.clear();
	s->extOfInt.reserve(m.nV);
	if (!getenv("XF_NO_RENUMBER")) {
		for (uint32_t c = 0; c < d.nColors; c++) {
			for (int j = 0; j < 4; j++) {
				for (uint32_t k = m.colorStart[c]; k < m.colorStart[c + 1]; k++) {
					const uint32_t v = m.idx[4 * (size_t)deviceOrder[k] + j];
					if (s->intOfExt[v] == 0xffffffffu) { s->intOfExt[v] = (uint32_t)s->extOfInt.size(); s->extOfInt.push_back(v); }
				}
			}
		}
	}
	for (uint32_t v = 0; v < m.nV; v++) { // vertices no element touches (or all of them: identity numbering)
		if (s->intOfExt[v] == 0xffffffffu) { s->intOfExt[v] = (uint32_t)s->extOfInt.size(); s->extOfInt.push_back(v); }
	}
	std::vector<VertexRec> xw(m.nV);
	std::vector<double4> x0(m.nV), zero(m.nV, double4{ 0.0, 0.0, 0.0, 0.0 });
	for (uint32_t i = 0; i < m.nV; i++) {
		const uint32_t v = s->extOfInt[i];
		xw[i] = VertexRec{ m.X0[3 * (size_t)v], m.X0[3 * (size_t)v + 1], m.X0[3 * (size_t)v + 2], m.w[v], (uint32_t)m.flags[v] };
		x0[i] = double4{ xw[i].x, xw[i].y, xw[i].z, 0.0 };
	}
	XF_CUDA(Upload(&d.Xw, xw));
	XF_CUDA(Upload(&d.O, x0));
	XF_CUDA(Upload(&d.X0, x0));
	XF_CUDA(Upload(&d.V, zero));
	XF_CUDA(Upload(&d.extOfInt, s->extOfInt));
	std::vector<uint32_t> streamToSorted(m.nT), canonPos(m.nT), serialPos(m.nT);
	for (uint32_t pos = 0; pos < m.nT; pos++) { serialPos[m.order[pos]] = pos; }
	for (uint32_t pos = 0; pos < m.nT; pos++) { streamToSorted[deviceOrder[pos]] = pos; canonPos[pos] = serialPos[deviceOrder[pos]]; }
	std::vector<uint32_t> devIdx(4 * (size_t)m.nT);
	for (uint32_t k = 0; k < m.nT; k++) {
		for (int j = 0; j < 4; j++) { devIdx[4 * (size_t)k + j] = s->intOfExt[m.idx[4 * (size_t)deviceOrder[k] + j]]; }
	}
	PackedElements pk;
	PackElements(m, deviceOrder, devIdx.data(), &pk);
	XF_CUDA(Upload(&d.eA, pk.a));
	XF_CUDA(Upload(&d.eB, pk.b));
	XF_CUDA(Upload(&d.eC, pk.c));
	XF_CUDA(Upload(&d.eArea, pk.area));
	// dataflow schedule: previous-writer stage code per (element, vertex) in the top byte of the index, last-writer code per vertex
	s->dataflowOk = m.nV <= 0x01000000u && d.nColors <= 254u;
	if (s->dataflowOk) {
		std::vector<uint8_t> pred, lastExt, lastCode(m.nV, 0);
		StageCodes(m, deviceOrder, &pred, &lastExt);
		std::vector<ElemRecA> ad = pk.a;
		for (uint32_t k = 0; k < m.nT; k++) {
			for (int j = 0; j < 4; j++) { ad[k].idx[j] = pk.a[k].idx[j] | ((uint32_t)pred[4 * (size_t)k + j] << 24); }
		}
		for (uint32_t v = 0; v < m.nV; v++) { lastCode[s->intOfExt[v]] = lastExt[v]; }
		XF_CUDA(Upload(&d.eAd, ad));
		XF_CUDA(Upload(&d.lastCode, lastCode));
		XF_CUDA(cudaMalloc((void**)&d.eAlpha, sizeof(float2) * std::max<size_t>(m.nT, 1)));
		if (m.groupSize > 1) { // clustered colouring: slot / first / last bits of every corner, device order
			std::vector<uint32_t> ek(m.nT);
			for (uint32_t k = 0; k < m.nT; k++) { ek[k] = m.clusterInfo[deviceOrder[k]]; }
			XF_CUDA(Upload(&d.eK, ek));
			d.groupSize = m.groupSize;
		} else if (!s->chainInfo.empty()) { // chained sweep: device order == serial order
			XF_CUDA(Upload(&d.eK, s->chainInfo));
			d.chained = 1;
		}
		if (m.groupSize <= 1) {
			// damping sweeps on the barrier-free schedule (xf_dataflow_general.cu): V records are versioned by a write count.
			// Rank of every element among the elements around each of its corners' vertices (serial order), and per vertex how
			// many of those elements lie below each boundary nT*q/8 of the amortised damping slices (Geo.cpp:794-797).
			std::vector<uint32_t> rank;
			std::vector<uint64_t> below;
			DampingCodes(m.nT, m.nV, devIdx.data(), &rank, &below);
			std::vector<uint2> slice(m.nV);
			for (uint32_t v = 0; v < m.nV; v++) { slice[v] = make_uint2((uint32_t)below[v], (uint32_t)(below[v] >> 32)); }
			XF_CUDA(Upload(&d.eRank, rank));
			XF_CUDA(Upload(&d.vSlice, slice));
		}
		for (uint32_t c = 0; c < d.nColors; c++) { d.maxColorSize = std::max(d.maxColorSize, m.colorStart[c + 1] - m.colorStart[c]); }
		if (const char* env = getenv("XF_DATAFLOW_BLOCK")) { d.dataflowBlock = (uint32_t)atoi(env); }
	}
	XF_CUDA(Upload(&d.canonPos, canonPos));
	XF_CUDA(Upload(&d.streamToSorted, streamToSorted));
	XF_CUDA(cudaMalloc((void**)&d.eScratch, sizeof(float) * m.nT));
	XF_CUDA(cudaMalloc((void**)&d.statScratch, sizeof(double) * 8));
	XF_CUDA(cudaMalloc((void**)&d.barrier, sizeof(unsigned int) * 32));
	XF_CUDA(cudaMemset(d.barrier, 0, sizeof(unsigned int) * 32));
	d.errDev = d.barrier + 16;
	{
		void* host = nullptr;
		XF_CUDA(cudaHostAlloc(&host, sizeof(unsigned int), cudaHostAllocMapped));
		s->stallWord = (volatile unsigned int*)host;
		*s->stallWord = 0u;
		XF_CUDA(cudaHostGetDevicePointer((void**)&d.errHost, host, 0));
	}
	XF_CUDA(cudaMalloc((void**)&s->dPackX, sizeof(double) * 3 * m.nV));
	XF_CUDA(cudaMalloc((void**)&s->dPackV, sizeof(double) * 3 * m.nV));
	XF_CUDA(cudaMalloc((void**)&s->dPackW, sizeof(float) * m.nV));
	return XF_OK;
}

int NeedDevice(const xf_scene* s) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	if (s->device < 0) { return Fail(XF_ERR_CUDA, "scene was created host-only (device < 0): there is no CPU compute path"); }
	cudaError_t e = cudaSetDevice(s->device);
	if (e != cudaSuccess) { return FailCuda(e, "cudaSetDevice"); }
	return XF_OK;
}

// After a stream sync: did a barrier-free launch give up on a record (SpinGiveUp)?  The state is then undefined; the scene keeps
// failing until xf_set_state installs a new one.  The CUDA context is intact, other scenes are unaffected.
int CheckStall(const xf_scene* s) {
	if (s->stallWord && *s->stallWord != 0u) {
		return Fail(XF_ERR_CUDA, "barrier-free schedule stalled: a vertex record never reached its expected stage (state rewritten under a running "
		                         "launch, or a broken schedule); the state is undefined until xf_set_state");
	}
	return XF_OK;
}

int BuildParams(xf_scene* s, const xf_settings* st, const xf_manipulator* manip, float dt, SubstepParams* p) {
	std::string err;
	int rc = FillSubstepParams(st, manip, dt, s->mesh, p, &err);
	if (rc != XF_OK) { return Fail(rc, err); }
	p->groundOn = s->groundOn;
	p->groundY = s->groundY;
	p->groundKeep = 1.0f - s->groundFriction;
	p->handleCount = s->handleCount;
	memcpy(p->handleIdx, s->handleIdx, sizeof(p->handleIdx));
	if (!s->intOfExt.empty()) { // device vertex numbering
		if (p->manipOn) { p->manipIdx = s->intOfExt[p->manipIdx]; }
		for (uint32_t h = 0; h < p->handleCount; h++) { p->handleIdx[h] = s->intOfExt[p->handleIdx[h]]; }
	}
	memcpy(p->handleTarget, s->handleTarget, sizeof(p->handleTarget));
	return XF_OK;
}

}  // namespace

extern "C" {

const char* xf_last_error(void) { return g_lastError.c_str(); }

int xf_device_count(int* outCount) {
	if (!outCount) { return Fail(XF_ERR_INVALID, "null outCount"); }
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { *outCount = 0; return FailCuda(e, "cudaGetDeviceCount"); }
	*outCount = n;
	return XF_OK;
}

void xf_default_create_params(xf_create_params* p) {
	if (!p) { return; }
	memset(p, 0, sizeof(*p));
	p->abiVersion = XF_ABI_VERSION;
	p->device = 0;
	p->density = 1.0f;
	p->autoResize = 0;
	p->precision = XF_PRECISION_EXACT;
	p->schedule = XF_SCHEDULE_AUTO;
}

int xf_create(const xf_create_params* params, const float* nodeXYZ, uint32_t nodeFloatCount, const uint32_t* idxStream, uint32_t idxCount,
              xf_scene** outScene) {
	if (!params || !outScene) { return Fail(XF_ERR_INVALID, "null params/outScene"); }
	*outScene = nullptr;
	if (params->abiVersion != XF_ABI_VERSION) { return Fail(XF_ERR_INVALID, "xf_create_params.abiVersion mismatch"); }
	if (params->precision != XF_PRECISION_EXACT && params->precision != XF_PRECISION_FAST) { return Fail(XF_ERR_INVALID, "bad precision"); }
	if (params->schedule < XF_SCHEDULE_AUTO || params->schedule > XF_SCHEDULE_DATAFLOW) { return Fail(XF_ERR_INVALID, "bad schedule"); }
	xf_scene* s = new (std::nothrow) xf_scene();
	if (!s) { return Fail(XF_ERR_NOMEM, "out of host memory"); }
	std::string err;
	// clustered colouring (one thread per cluster of elements, xf_dataflow.cu): opt-in.  Measured on B200: 83 us per substep at
	// 1M tets against 53 us with one thread per element (the chain is ~1.2 us of dependent arithmetic per element, not the L2
	// hand-off, so serialising six elements in a thread lengthens it), 99 against 113 us at 2M tets.
	const bool clustered = params->grouping == XF_GROUPING_CLUSTERS || (params->grouping == XF_GROUPING_AUTO && getenv("XF_CLUSTERS") != nullptr);
	int rc = PrepareMesh(nodeXYZ, nodeFloatCount, idxStream, idxCount, params->density, params->autoResize != 0, params->colorHint,
	                     params->colorHintCount, &s->mesh, &err, clustered);
	if (rc != XF_OK) { delete s; return Fail(rc, err); }
	s->precision = params->precision;
	s->schedule = params->schedule;
	s->device = params->device;
	// chained sweep (XF_GROUPING_CHAINS): AUTO follows kChainByDefault unless XF_CHAIN=0/1 says otherwise
	if (params->grouping > XF_GROUPING_CHAINS) { delete s; return Fail(XF_ERR_INVALID, "bad grouping"); }
	{
		const char* env = getenv("XF_CHAIN");
		const bool autoChain = env ? atoi(env) != 0 : kChainByDefault;
		s->wantChain = params->grouping == XF_GROUPING_CHAINS || (params->grouping == XF_GROUPING_AUTO && autoChain);
	}
	if (s->wantChain && s->mesh.groupSize <= 1 && s->mesh.nT > 0 && s->mesh.nV <= 0x01000000u && s->mesh.colorStart.size() <= 255u) {
		std::vector<uint8_t> pred, last;
		StageCodes(s->mesh, s->mesh.order, &pred, &last);
		const uint64_t held = ChainInfo(s->mesh, s->mesh.order, pred, &s->chainInfo);
		s->chainedPermille = (uint32_t)(held * 1000u / (4u * (uint64_t)s->mesh.nT));
		// nothing to keep (e.g. a generic colouring whose positions do not line up): the plain kernel does the same work cheaper
		if (s->chainedPermille < kChainMinPermille) { s->chainInfo.clear(); s->chainedPermille = 0; }
	}
	if (s->device >= 0) {
		cudaError_t e = cudaSetDevice(s->device);
		if (e != cudaSuccess) { delete s; return FailCuda(e, "cudaSetDevice"); }
		cudaDeviceProp prop;
		e = cudaGetDeviceProperties(&prop, s->device);
		if (e != cudaSuccess) { delete s; return FailCuda(e, "cudaGetDeviceProperties"); }
		s->smCount = prop.multiProcessorCount;
		s->l2Bytes = (size_t)prop.l2CacheSize;
		s->cooperative = prop.cooperativeLaunch != 0;
		if (params->stream) {
			s->stream = (cudaStream_t)params->stream;
		} else {
			e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
			if (e != cudaSuccess) { delete s; return FailCuda(e, "cudaStreamCreate"); }
			s->ownStream = true;
		}
		rc = UploadScene(s);
		if (rc != XF_OK) { FreeDevice(s); delete s; return rc; }
		if (const char* env = getenv("XF_DATAFLOW_SLEEP_NS")) { s->spinSleepNs = (uint32_t)atoi(env) & 0x7fffu; }
		if (getenv("XF_DATAFLOW_NO_PREFETCH")) { s->spinSleepNs |= 0x8000u; }
		if (const char* env = getenv("XF_DATAFLOW_ESLEEP_NS")) { s->spinSleepNs |= ((uint32_t)atoi(env) & 0xffffu) << 16; }
		// XF_SCHEDULE_BRICKS (vertices private to a CTA's Morton brick kept in shared memory, round 1) measured slower than the
		// plain grid-barrier kernel (99.6 vs 92.4 us per substep at 1M tets, profiles/r1_results.md) and was retired in round 2:
		// the value is still accepted and runs as XF_SCHEDULE_PERSISTENT (same serial order, same bits).
		if (s->schedule == XF_SCHEDULE_BRICKS) { s->schedule = XF_SCHEDULE_PERSISTENT; }
		if (s->schedule == XF_SCHEDULE_AUTO) {
			s->schedule = !s->cooperative ? XF_SCHEDULE_LAUNCH_PER_COLOR : (s->dataflowOk ? XF_SCHEDULE_DATAFLOW : XF_SCHEDULE_PERSISTENT);
		}
		if (s->schedule == XF_SCHEDULE_DATAFLOW && !s->dataflowOk) {
			FreeDevice(s); delete s;
			return Fail(XF_ERR_UNSUPPORTED, "XF_SCHEDULE_DATAFLOW needs <= 2^24 vertices and <= 254 colours");
		}
		if (s->schedule >= XF_SCHEDULE_PERSISTENT && !s->cooperative) {
			FreeDevice(s); delete s;
			return Fail(XF_ERR_UNSUPPORTED, "device does not support cooperative launches (needed by XF_SCHEDULE_PERSISTENT / DATAFLOW)");
		}
	}
	*outScene = s;
	return XF_OK;
}

int xf_destroy(xf_scene* s) {
	if (!s) { return XF_OK; }
	if (s->device >= 0) { cudaSetDevice(s->device); if (s->stream) { cudaStreamSynchronize(s->stream); } }
	FreeDevice(s);
	delete s;
	return XF_OK;
}

uint32_t xf_vert_count(const xf_scene* s) { return s ? s->mesh.nV : 0; }
uint32_t xf_element_count(const xf_scene* s) { return s ? s->mesh.nT : 0; }
uint32_t xf_color_count(const xf_scene* s) { return s ? (uint32_t)s->mesh.colorStart.size() - 1 : 0; }

int xf_get_order(const xf_scene* s, uint32_t* order) {
	if (!s || !order) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(order, s->mesh.order.data(), sizeof(uint32_t) * s->mesh.nT);
	return XF_OK;
}
int xf_get_colors(const xf_scene* s, uint32_t* colorOfElement) {
	if (!s || !colorOfElement) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(colorOfElement, s->mesh.color.data(), sizeof(uint32_t) * s->mesh.nT);
	return XF_OK;
}
int xf_get_stage_codes(const xf_scene* s, uint8_t* predCode4, uint8_t* lastCode) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	if (s->mesh.colorStart.size() - 1 > 254u) { return Fail(XF_ERR_UNSUPPORTED, "more than 254 colours: stage codes do not fit a byte"); }
	std::vector<uint8_t> pred, last;
	StageCodes(s->mesh, s->mesh.order, &pred, &last);
	if (predCode4) { memcpy(predCode4, pred.data(), pred.size()); }
	if (lastCode) { memcpy(lastCode, last.data(), last.size()); }
	return XF_OK;
}

// Codes of the count-versioned velocity records (k_substeps_dataflow_general), in the caller's vertex numbering and the serial order
// of xf_get_order: rank4[4*k + j] = rank of the element at serial position k among the elements around its corner j;
// below8[8*v + q] = elements around vertex v whose serial position is below nT*(q+1)/8.  Host-only scenes answer too.
int xf_get_damping_codes(const xf_scene* s, uint8_t* rank4, uint8_t* below8) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	const HostMesh& m = s->mesh;
	std::vector<uint32_t> idxOfPos(4 * (size_t)m.nT), rank;
	std::vector<uint64_t> below;
	for (uint32_t k = 0; k < m.nT; k++) { for (int j = 0; j < 4; j++) { idxOfPos[4 * (size_t)k + j] = m.idx[4 * (size_t)m.order[k] + j]; } }
	DampingCodes(m.nT, m.nV, idxOfPos.data(), &rank, &below);
	if (rank4) { for (uint32_t k = 0; k < m.nT; k++) { for (int j = 0; j < 4; j++) { rank4[4 * (size_t)k + j] = (uint8_t)(rank[k] >> (8 * j)); } } }
	if (below8) { for (uint32_t v = 0; v < m.nV; v++) { for (int q = 0; q < 8; q++) { below8[8 * (size_t)v + q] = (uint8_t)(below[v] >> (8 * q)); } } }
	return XF_OK;
}

int xf_get_chain_info(const xf_scene* s, uint32_t* info, uint32_t* outPermille) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	if (outPermille) { *outPermille = s->chainedPermille; }
	if (info) {
		if (s->chainInfo.empty()) { memset(info, 0, sizeof(uint32_t) * s->mesh.nT); }
		else { memcpy(info, s->chainInfo.data(), sizeof(uint32_t) * s->mesh.nT); }
	}
	return XF_OK;
}

int xf_get_elements(const xf_scene* s, uint32_t* idx4, float* Qi9, float* QQ3, float* QR3, float* volume, float* surfaceArea) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	const HostMesh& m = s->mesh;
	if (idx4) { memcpy(idx4, m.idx.data(), sizeof(uint32_t) * m.idx.size()); }
	if (Qi9) { memcpy(Qi9, m.Qi.data(), sizeof(float) * m.Qi.size()); }
	if (QQ3) { memcpy(QQ3, m.QQ.data(), sizeof(float) * m.QQ.size()); }
	if (QR3) { memcpy(QR3, m.QR.data(), sizeof(float) * m.QR.size()); }
	if (volume) { memcpy(volume, m.volume.data(), sizeof(float) * m.volume.size()); }
	if (surfaceArea) { memcpy(surfaceArea, m.area.data(), sizeof(float) * m.area.size()); }
	return XF_OK;
}

// Launches n substeps with the prepared parameters (p.vary, if set, holds n rows; a call that is split into several launches
// advances it).
static int LaunchSubsteps(xf_scene* s, SubstepParams& p, uint32_t firstTick, uint32_t n) {
	const bool exact = s->precision == XF_PRECISION_EXACT;
	// the barrier-free schedule covers everything but in-constraint Rayleigh damping (Paper / Limit read O of other threads' vertices)
	const bool inConstraintDamping = p.damping > 0.0f && p.rayleigh < XF_RAYLEIGH_POST;
	const bool plainSweep = !inConstraintDamping && !p.doDamp && !p.doPbdDamp && p.volumePasses == 0;
	const bool generalOk = !inConstraintDamping && s->dev.eRank && !s->dev.chained && s->dev.groupSize <= 1 && !getenv("XF_NO_DATAFLOW_GENERAL");
	const float* vary0 = p.vary;
	if (s->schedule == XF_SCHEDULE_DATAFLOW && (plainSweep || generalOk)) {
		if (!(s->alphaValid && s->alphaKey[0] == p.invMu && s->alphaKey[1] == p.invLambda && s->alphaKey[2] == p.dt2)) {
			XF_CUDA(LaunchElementAlpha(s->dev, p, exact, s->stream, &s->launches)); // stream-ordered before the launch that reads it
			s->alphaKey[0] = p.invMu; s->alphaKey[1] = p.invLambda; s->alphaKey[2] = p.dt2;
			s->alphaValid = true;
		}
		const uint32_t stride = plainSweep ? p.nColors + 1u : DataflowGeneralStride(p);
		const uint32_t maxPerLaunch = std::max(1u, (0x00ffffffu - 2u) / stride); // tags of one launch must not wrap onto the stale ones
		for (uint32_t done = 0; done < n;) {
			const uint32_t m = std::min(n - done, maxPerLaunch);
			p.tickId = firstTick + done;
			p.vary = vary0 ? vary0 + (size_t)kVaryFloats * done : nullptr;
			s->lastKernel = !plainSweep ? XF_KERNEL_DATAFLOW_GENERAL
			                            : (s->dev.groupSize > 1 ? XF_KERNEL_CLUSTER : (s->dev.chained ? XF_KERNEL_CHAIN : XF_KERNEL_DATAFLOW));
			if (!plainSweep) {
				XF_CUDA(LaunchSubstepsDataflowGeneral(s->dev, p, exact, m, s->smCount, s->verBase, s->vEpoch, s->spinSleepNs, s->stream, &s->launches));
				s->vEpoch += m;
			} else if (s->dev.groupSize > 1) {
				XF_CUDA(LaunchSubstepsCluster(s->dev, p, exact, m, s->smCount, s->verBase, s->spinSleepNs, s->stream, &s->launches));
			} else if (s->dev.chained) {
				XF_CUDA(LaunchSubstepsChain(s->dev, p, exact, m, s->smCount, s->verBase, s->spinSleepNs, s->stream, &s->launches));
			} else {
				XF_CUDA(LaunchSubstepsDataflow(s->dev, p, exact, m, s->smCount, s->verBase, s->spinSleepNs, s->stream, &s->launches));
			}
			s->verBase = (s->verBase + m * stride + 1u) & 0x00ffffffu;
			done += m;
		}
	} else if (s->schedule == XF_SCHEDULE_PERSISTENT || s->schedule == XF_SCHEDULE_DATAFLOW) {
		auto it = s->shapes.find(p.energy);
		if (it == s->shapes.end()) {
			LaunchShape shape;
			XF_CUDA(QueryLaunchShape(s->device, p.energy, exact, &shape));
			it = s->shapes.emplace(p.energy, shape).first;
		}
		s->lastKernel = XF_KERNEL_PERSISTENT;
		XF_CUDA(LaunchSubstepsPersistent(s->dev, p, exact, n, it->second, s->stream, &s->launches));
	} else {
		s->lastKernel = XF_KERNEL_PER_COLOR;
		XF_CUDA(LaunchSubstepsPerColor(s->dev, p, exact, n, s->stream, &s->launches));
	}
	return XF_OK;
}

int xf_substep(xf_scene* s, const xf_settings* st, const xf_manipulator* manip, float dt, uint32_t n) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	if (!st) { return Fail(XF_ERR_INVALID, "null settings"); }
	if (n == 0) { return XF_OK; }
	SubstepParams p;
	rc = BuildParams(s, st, manip, dt, &p);
	if (rc != XF_OK) { return rc; }
	return LaunchSubsteps(s, p, st->tickId, n);
}

// n substeps in ONE launch while the right-side lock transform and / or the manipulator ray change from substep to substep
// (what Sim::Update does between its Substep calls, Demo.cpp:67-90).  lockT3d: n x 12 floats (Settings::lockedRightTransform3d
// of every substep) or NULL; pickDirTarget: n x 3 floats (Manipulator::pickDirTarget of every substep) or NULL.  Everything
// else comes from *st / *manip; tickId advances by one per substep as in xf_substep.  Bit-identical to n calls of
// xf_substep(.., 1) with the same values.
int xf_substep_varying(xf_scene* s, const xf_settings* st, const xf_manipulator* manip, float dt, uint32_t n, const float* lockT3d,
                       const float* pickDirTarget) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	if (!st) { return Fail(XF_ERR_INVALID, "null settings"); }
	if (n == 0) { return XF_OK; }
	const bool picked = manip && manip->picked;
	const bool supported = s->schedule == XF_SCHEDULE_DATAFLOW || s->schedule == XF_SCHEDULE_PERSISTENT;
	if ((!lockT3d && !(pickDirTarget && picked)) || !supported) {
		if (!lockT3d && !(pickDirTarget && picked)) { return xf_substep(s, st, manip, dt, n); }
		// schedules without per-substep rows: one call per substep (same values, same bits)
		xf_settings stK = *st;
		xf_manipulator mK;
		if (manip) { mK = *manip; }
		for (uint32_t k = 0; k < n; k++) {
			if (lockT3d) { memcpy(stK.lockedRightTransform3d, lockT3d + 12 * (size_t)k, sizeof(float) * 12); }
			if (manip && pickDirTarget) { memcpy(mK.pickDirTarget, pickDirTarget + 3 * (size_t)k, sizeof(float) * 3); }
			stK.tickId = st->tickId + k;
			rc = xf_substep(s, &stK, manip ? &mK : nullptr, dt, 1);
			if (rc != XF_OK) { return rc; }
		}
		return XF_OK;
	}
	SubstepParams p;
	rc = BuildParams(s, st, manip, dt, &p);
	if (rc != XF_OK) { return rc; }
	// per-substep rows: the same host arithmetic as BuildParams, once per substep
	std::vector<float> rows((size_t)kVaryFloats * n, 0.0f);
	xf_settings stK = *st;
	xf_manipulator mK;
	if (manip) { mK = *manip; }
	for (uint32_t k = 0; k < n; k++) {
		if (lockT3d) { memcpy(stK.lockedRightTransform3d, lockT3d + 12 * (size_t)k, sizeof(float) * 12); }
		if (manip && pickDirTarget) { memcpy(mK.pickDirTarget, pickDirTarget + 3 * (size_t)k, sizeof(float) * 3); }
		SubstepParams pk;
		std::string err;
		rc = FillSubstepParams(&stK, manip ? &mK : nullptr, dt, s->mesh, &pk, &err);
		if (rc != XF_OK) { return Fail(rc, err); }
		memcpy(&rows[(size_t)kVaryFloats * k], pk.lockT, sizeof(float) * 12);
		memcpy(&rows[(size_t)kVaryFloats * k + 12], pk.manipTarget, sizeof(float) * 3);
	}
	if (s->dVaryRows < n) {
		XF_CUDA(cudaStreamSynchronize(s->stream)); // a launch in flight may still read the old rows
		if (s->dVary) { cudaFree(s->dVary); s->dVary = nullptr; }
		s->dVaryRows = std::max<size_t>(n, 512);
		XF_CUDA(cudaMalloc((void**)&s->dVary, sizeof(float) * kVaryFloats * s->dVaryRows));
	}
	// stream-ordered after the previous launch; the source is pageable memory, so the runtime has copied it out when the call returns
	XF_CUDA(cudaMemcpyAsync(s->dVary, rows.data(), sizeof(float) * rows.size(), cudaMemcpyHostToDevice, s->stream));
	p.vary = s->dVary;
	return LaunchSubsteps(s, p, st->tickId, n);
}

int xf_sync(xf_scene* s) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	XF_CUDA(cudaStreamSynchronize(s->stream));
	return CheckStall(s);
}

int xf_set_ground(xf_scene* s, int enabled, float y0, float friction) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	s->groundOn = enabled ? 1u : 0u;
	s->groundY = y0;
	s->groundFriction = friction;
	return XF_OK;
}

int xf_set_handles(xf_scene* s, uint32_t count, const uint32_t* vertIdx, const float* targetXYZ) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	if (count > (uint32_t)kMaxHandles) { return Fail(XF_ERR_INVALID, "at most 64 drag handles"); }
	if (count && (!vertIdx || !targetXYZ)) { return Fail(XF_ERR_INVALID, "null handle arrays"); }
	for (uint32_t k = 0; k < count; k++) {
		if (vertIdx[k] >= s->mesh.nV) { return Fail(XF_ERR_INVALID, "handle vertex index out of range"); }
	}
	s->handleCount = count;
	for (uint32_t k = 0; k < count; k++) {
		s->handleIdx[k] = vertIdx[k];
		for (int j = 0; j < 3; j++) { s->handleTarget[k][j] = targetXYZ[3 * (size_t)k + j]; }
	}
	return XF_OK;
}

int xf_get_state_async(xf_scene* s, double* X, double* V) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	XF_CUDA(LaunchPackState(s->dev, X ? s->dPackX : nullptr, V ? s->dPackV : nullptr, nullptr, s->stream, &s->launches));
	const size_t bytes = sizeof(double) * 3 * s->mesh.nV;
	if (X) { XF_CUDA(cudaMemcpyAsync(X, s->dPackX, bytes, cudaMemcpyDeviceToHost, s->stream)); }
	if (V) { XF_CUDA(cudaMemcpyAsync(V, s->dPackV, bytes, cudaMemcpyDeviceToHost, s->stream)); }
	return XF_OK;
}

int xf_set_state_async(xf_scene* s, const double* X, const double* V) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	const size_t bytes = sizeof(double) * 3 * s->mesh.nV;
	if (X) { XF_CUDA(cudaMemcpyAsync(s->dPackX, X, bytes, cudaMemcpyHostToDevice, s->stream)); }
	if (V) { XF_CUDA(cudaMemcpyAsync(s->dPackV, V, bytes, cudaMemcpyHostToDevice, s->stream)); }
	XF_CUDA(LaunchUnpackState(s->dev, X ? s->dPackX : nullptr, V ? s->dPackV : nullptr, nullptr, s->stream, &s->launches));
	return XF_OK;
}

int xf_get_state(xf_scene* s, double* X, double* V, float* w) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	if (s->device < 0) { // host-only scene: the initial state
		const HostMesh& m = s->mesh;
		if (X) { memcpy(X, m.X0.data(), sizeof(double) * 3 * m.nV); }
		if (V) { memset(V, 0, sizeof(double) * 3 * m.nV); }
		if (w) { memcpy(w, m.w.data(), sizeof(float) * m.nV); }
		return XF_OK;
	}
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	XF_CUDA(LaunchPackState(s->dev, X ? s->dPackX : nullptr, V ? s->dPackV : nullptr, w ? s->dPackW : nullptr, s->stream, &s->launches));
	const size_t bytes = sizeof(double) * 3 * s->mesh.nV;
	if (X) { XF_CUDA(cudaMemcpyAsync(X, s->dPackX, bytes, cudaMemcpyDeviceToHost, s->stream)); }
	if (V) { XF_CUDA(cudaMemcpyAsync(V, s->dPackV, bytes, cudaMemcpyDeviceToHost, s->stream)); }
	if (w) { XF_CUDA(cudaMemcpyAsync(w, s->dPackW, sizeof(float) * s->mesh.nV, cudaMemcpyDeviceToHost, s->stream)); }
	XF_CUDA(cudaStreamSynchronize(s->stream));
	return CheckStall(s);
}

int xf_set_state(xf_scene* s, const double* X, const double* V, const float* w) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	const size_t bytes = sizeof(double) * 3 * s->mesh.nV;
	if (X) { XF_CUDA(cudaMemcpyAsync(s->dPackX, X, bytes, cudaMemcpyHostToDevice, s->stream)); }
	if (V) { XF_CUDA(cudaMemcpyAsync(s->dPackV, V, bytes, cudaMemcpyHostToDevice, s->stream)); }
	if (w) { XF_CUDA(cudaMemcpyAsync(s->dPackW, w, sizeof(float) * s->mesh.nV, cudaMemcpyHostToDevice, s->stream)); }
	XF_CUDA(LaunchUnpackState(s->dev, X ? s->dPackX : nullptr, V ? s->dPackV : nullptr, w ? s->dPackW : nullptr, s->stream, &s->launches));
	XF_CUDA(cudaStreamSynchronize(s->stream));
	if (s->stallWord && *s->stallWord != 0u) { // a new state ends a reported stall
		XF_CUDA(cudaMemset(s->dev.errDev, 0, sizeof(unsigned int)));
		*s->stallWord = 0u;
	}
	return XF_OK;
}

int xf_get_rest(xf_scene* s, double* X0, double* O, uint8_t* flags) {
	if (!s) { return Fail(XF_ERR_INVALID, "null scene"); }
	const HostMesh& m = s->mesh;
	if (X0) { memcpy(X0, m.X0.data(), sizeof(double) * 3 * m.nV); }
	if (flags) { memcpy(flags, m.flags.data(), m.nV); }
	if (O) {
		if (s->device < 0) { memcpy(O, m.X0.data(), sizeof(double) * 3 * m.nV); return XF_OK; }
		int rc = NeedDevice(s);
		if (rc != XF_OK) { return rc; }
		std::vector<double4> tmp(m.nV);
		XF_CUDA(cudaStreamSynchronize(s->stream));
		XF_CUDA(cudaMemcpy(tmp.data(), s->dev.O, sizeof(double4) * m.nV, cudaMemcpyDeviceToHost));
		for (uint32_t i = 0; i < m.nV; i++) {
			const uint32_t v = s->extOfInt[i];
			O[3 * (size_t)v] = tmp[i].x; O[3 * (size_t)v + 1] = tmp[i].y; O[3 * (size_t)v + 2] = tmp[i].z;
		}
	}
	return XF_OK;
}

int xf_get_origin(const xf_scene* s, float* origin3) {
	if (!s || !origin3) { return Fail(XF_ERR_INVALID, "null argument"); }
	memcpy(origin3, s->mesh.origin, sizeof(float) * 3);
	return XF_OK;
}

int xf_transform(xf_scene* s, const float* m9) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	if (!m9) { return Fail(XF_ERR_INVALID, "null matrix"); }
	XF_CUDA(LaunchTransform(s->dev, m9, s->stream, &s->launches));
	// origin = (t4 * vec4(origin, 1)).xyz, Geo.cpp:363 (fp32, left-assoc dot of each row with (x,y,z,1))
	const float ox = s->mesh.origin[0], oy = s->mesh.origin[1], oz = s->mesh.origin[2];
	s->mesh.origin[0] = m9[0] * ox + m9[3] * oy + 0.0f * oz + m9[6] * 1.0f;
	s->mesh.origin[1] = m9[1] * ox + m9[4] * oy + 0.0f * oz + m9[7] * 1.0f;
	s->mesh.origin[2] = m9[2] * ox + m9[5] * oy + 1.0f * oz + 0.0f * 1.0f;
	return XF_OK;
}

int xf_volume(xf_scene* s, float* outVolume) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	if (!outVolume) { return Fail(XF_ERR_INVALID, "null outVolume"); }
	XF_CUDA(LaunchElementVolumes(s->dev, s->stream, &s->launches));
	s->hostScratch.resize(s->mesh.nT);
	XF_CUDA(cudaMemcpyAsync(s->hostScratch.data(), s->dev.eScratch, sizeof(float) * s->mesh.nT, cudaMemcpyDeviceToHost, s->stream));
	XF_CUDA(cudaStreamSynchronize(s->stream));
	rc = CheckStall(s);
	if (rc != XF_OK) { return rc; }
	float volume = 0.0f; // fp32 running sum in element order, Geo.cpp:828-829
	for (uint32_t e = 0; e < s->mesh.nT; e++) { volume += s->hostScratch[e]; }
	*outVolume = volume;
	return XF_OK;
}

int xf_stats(xf_scene* s, const xf_settings* st, double* out6) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	if (!st || !out6) { return Fail(XF_ERR_INVALID, "null argument"); }
	SubstepParams p;
	rc = BuildParams(s, st, nullptr, 1.0f, &p);
	if (rc != XF_OK) { return rc; }
	XF_CUDA(LaunchStats(s->dev, p, (double)st->gravity[0], (double)st->gravity[1], s->smCount, s->stream, &s->launches));
	XF_CUDA(cudaMemcpyAsync(out6, s->dev.statScratch, sizeof(double) * 6, cudaMemcpyDeviceToHost, s->stream));
	XF_CUDA(cudaStreamSynchronize(s->stream));
	return CheckStall(s);
}

/* Test / measurement hook (declared in the header's debug section).  knob 0: first stage tag of the next barrier-free launch
 * (24 bits; lets a test cross the tag wrap-around without ~13 000 launches); 1: polls before a waiting warp gives up;
 * 2: overwrite the last-writer code of device vertex 0 (breaks the schedule on purpose: the next launch must report a stall,
 * not hang and not kill the context). */
int xf_debug_scene_knob(xf_scene* s, int knob, uint32_t value) {
	int rc = NeedDevice(s);
	if (rc != XF_OK) { return rc; }
	switch (knob) {
	case 0: s->verBase = value & 0x00ffffffu; return XF_OK;
	case 1: s->dev.spinLimit = value; return XF_OK;
	case 2: {
		if (!s->dev.lastCode) { return Fail(XF_ERR_UNSUPPORTED, "scene has no dataflow codes"); }
		const uint8_t code = (uint8_t)value;
		XF_CUDA(cudaStreamSynchronize(s->stream));
		XF_CUDA(cudaMemcpy(s->dev.lastCode, &code, 1, cudaMemcpyHostToDevice));
		return XF_OK;
	}
	default: return Fail(XF_ERR_INVALID, "unknown knob");
	}
}

int xf_get_info(const xf_scene* s, xf_info* out) {
	if (!s || !out) { return Fail(XF_ERR_INVALID, "null argument"); }
	memset(out, 0, sizeof(*out));
	const HostMesh& m = s->mesh;
	out->vertCount = m.nV;
	out->elementCount = m.nT;
	out->colorCount = (uint32_t)m.colorStart.size() - 1;
	uint32_t mn = 0xffffffffu, mx = 0;
	for (size_t c = 0; c + 1 < m.colorStart.size(); c++) {
		uint32_t n = m.colorStart[c + 1] - m.colorStart[c];
		mn = std::min(mn, n);
		mx = std::max(mx, n);
	}
	out->minColorSize = mn;
	out->maxColorSize = mx;
	out->smCount = (uint32_t)s->smCount;
	out->chainedPermille = s->device < 0 || (s->dev.chained && s->schedule == XF_SCHEDULE_DATAFLOW) ? s->chainedPermille : 0;
	if (s->schedule == XF_SCHEDULE_DATAFLOW) {
		out->gridBlocks = (uint32_t)(2 * s->smCount);
		out->blockThreads = s->dev.groupSize > 1 ? 256u : (uint32_t)DataflowBlockThreads(s->dev, 2 * s->smCount);
	} else if (!s->shapes.empty()) {
		out->gridBlocks = (uint32_t)s->shapes.begin()->second.gridBlocks;
		out->blockThreads = (uint32_t)s->shapes.begin()->second.blockThreads;
	}
	out->elementRecordBytes = s->precision == XF_PRECISION_EXACT ? 80u : 64u; // +16 B only for prefactored energies in EXACT
	out->schedule = (uint32_t)s->schedule;
	out->lastKernel = s->lastKernel;
	out->kernelLaunches = s->launches;
	out->l2Bytes = s->l2Bytes;
	return XF_OK;
}

}  // extern "C"
