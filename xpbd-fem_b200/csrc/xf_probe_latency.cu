// Latency probes behind DESIGN section 6's decomposition of one stage of the barrier-free sweep (not part of the stepping path):
//   mode 0  the element solve of a lone warp, records in registers: cycles from gathered records to updated records
//   mode 1  the hand-off: one warp stores a versioned 32-byte record, a warp on ANOTHER SM polls it and answers through a
//           second record (ping-pong): cycles per one-way hand-off (store -> L2 -> a polling load sees it)
//   mode 2  a dependent chain of 256-bit L2 loads (pointer chase through records): cycles per load
// Together they say how much of the ~2400-cycle stage of a latency-bound mesh is arithmetic (what a warp-cooperative,
// several-lanes-per-element evaluation could shorten) and how much is the L2 round trip (which it cannot).
#include "xf_element.cuh"

namespace xf {
namespace {

__global__ void __launch_bounds__(32) k_probe_element(ElemRec rec, SubstepParams p, double4* io, uint32_t iters, long long* outCycles) {
	VertexRegs v[4];
	for (int n = 0; n < 4; n++) {
		const double4 r = io[4 * (blockIdx.x * 32 + threadIdx.x) + n];
		v[n].x[0] = r.x; v[n].x[1] = r.y; v[n].x[2] = r.z;
		v[n].w = 400000.0f;
		v[n].flags = 0;
	}
	const ElemCompliance ec = ComplianceOf<true>(p, rec.volume);
	const long long t0 = clock64();
	for (uint32_t k = 0; k < iters; k++) { SolveElementGathered<XF_ENERGY_YEOH_SKIN_FAST, true, true, false>(NoStore{}, p, rec, v, ec); }
	const long long t1 = clock64();
	for (int n = 0; n < 4; n++) { io[4 * (blockIdx.x * 32 + threadIdx.x) + n] = make_double4(v[n].x[0], v[n].x[1], v[n].x[2], 0.0); }
	if (threadIdx.x == 0 && blockIdx.x == 0) { *outCycles = t1 - t0; }
}

// CTA 0 and CTA 1 land on different SMs (one CTA of 32 threads each, grid = 2 <= SM count).  Lane 0 of each plays.
__global__ void __launch_bounds__(32) k_probe_handoff(VertexRec* recs, uint32_t rounds, long long* outCycles) {
	if (threadIdx.x != 0) { return; }
	const uint32_t me = blockIdx.x, other = 1u - me;
	const long long t0 = clock64();
	for (uint32_t k = 1; k <= rounds; k++) {
		if (me == 0) {
			VertexRegs v; v.x[0] = k; v.x[1] = 0; v.x[2] = 0; v.w = 0; v.flags = k << 8;
			StoreVertex(recs, 0, v);
			VertexRegs r = LoadVertex(recs, 1);
			while ((r.flags >> 8) != k) { r = LoadVertex(recs, 1); }
		} else {
			VertexRegs r = LoadVertex(recs, 0);
			while ((r.flags >> 8) != k) { r = LoadVertex(recs, 0); }
			r.flags = k << 8;
			StoreVertex(recs, 1, r);
		}
	}
	const long long t1 = clock64();
	if (me == 0) { *outCycles = t1 - t0; }
	(void)other;
}

__global__ void __launch_bounds__(32) k_probe_chase(const VertexRec* recs, uint32_t n, uint32_t steps, long long* outCycles, uint32_t* sink) {
	if (threadIdx.x != 0) { return; }
	uint32_t i = 0;
	const long long t0 = clock64();
	for (uint32_t k = 0; k < steps; k++) {
		const VertexRegs r = LoadVertex(recs, i);
		i = r.flags % n;
	}
	const long long t1 = clock64();
	*outCycles = t1 - t0;
	*sink = i;
}

}  // namespace
}  // namespace xf

// mode 0: cycles per element solve (lone warp); mode 1: cycles per one-way record hand-off between two SMs; mode 2: cycles per
// dependent 256-bit L2 load.
extern "C" int xf_debug_stage_latency(int device, int mode, uint32_t iterations, double* outCycles) {
	using namespace xf;
	if (!outCycles || iterations == 0) { return XF_ERR_INVALID; }
	if (cudaSetDevice(device) != cudaSuccess) { return XF_ERR_CUDA; }
	long long* dCycles = nullptr;
	if (cudaMalloc(&dCycles, sizeof(long long)) != cudaSuccess) { return XF_ERR_NOMEM; }
	int rc = XF_OK;
	if (mode == 0) {
		// a slightly sheared rest tet (Qi = identity / edge length) so that the solve does real work every iteration
		ElemRec rec;
		memset(&rec, 0, sizeof(rec));
		const float h = 0.002f;
		for (int c = 0; c < 3; c++) { rec.Qi[c][c] = 1.0f / h; }
		rec.volume = h * h * h / 6.0f;
		for (int i = 0; i < 3; i++) { rec.QQ[i] = 1.0f / (h * h); rec.QR[i] = 0.0f; }
		SubstepParams p;
		memset(&p, 0, sizeof(p));
		p.dt = 1.0f / 3000.0f; p.dt2 = p.dt * p.dt; p.invDt = 3000.0f; p.invMu = 1.0f; p.invLambda = 0.0f; p.a = 1.0f;
		std::vector<double4> host(4 * 32);
		for (int t = 0; t < 32; t++) {
			host[4 * t + 0] = make_double4(h * 1.01, 0.0, 0.0, 0.0);
			host[4 * t + 1] = make_double4(0.0, h * 0.99, 1e-5 * t, 0.0);
			host[4 * t + 2] = make_double4(1e-5, 0.0, h * 1.02, 0.0);
			host[4 * t + 3] = make_double4(0.0, 0.0, 0.0, 0.0);
		}
		double4* io = nullptr;
		cudaMalloc(&io, sizeof(double4) * host.size());
		cudaMemcpy(io, host.data(), sizeof(double4) * host.size(), cudaMemcpyHostToDevice);
		k_probe_element<<<1, 32>>>(rec, p, io, iterations, dCycles);
		if (cudaDeviceSynchronize() != cudaSuccess) { rc = XF_ERR_CUDA; }
		cudaFree(io);
	} else if (mode == 1) {
		VertexRec* recs = nullptr;
		cudaMalloc(&recs, sizeof(VertexRec) * 64);
		cudaMemset(recs, 0, sizeof(VertexRec) * 64);
		k_probe_handoff<<<2, 32>>>(recs, iterations, dCycles);
		if (cudaDeviceSynchronize() != cudaSuccess) { rc = XF_ERR_CUDA; }
		cudaFree(recs);
	} else {
		const uint32_t n = 1u << 16;
		std::vector<VertexRec> host(n);
		uint32_t x = 12345u;
		for (uint32_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; host[i] = VertexRec{ 0.0, 0.0, 0.0, 0.0f, x >> 8 }; }
		VertexRec* recs = nullptr;
		uint32_t* sink = nullptr;
		cudaMalloc(&recs, sizeof(VertexRec) * n);
		cudaMalloc(&sink, 4);
		cudaMemcpy(recs, host.data(), sizeof(VertexRec) * n, cudaMemcpyHostToDevice);
		k_probe_chase<<<1, 32>>>(recs, n, iterations, dCycles, sink);
		if (cudaDeviceSynchronize() != cudaSuccess) { rc = XF_ERR_CUDA; }
		cudaFree(recs);
		cudaFree(sink);
	}
	long long cycles = 0;
	cudaMemcpy(&cycles, dCycles, sizeof(cycles), cudaMemcpyDeviceToHost);
	cudaFree(dCycles);
	*outCycles = (double)cycles / (double)iterations / (mode == 1 ? 2.0 : 1.0);
	return rc;
}
