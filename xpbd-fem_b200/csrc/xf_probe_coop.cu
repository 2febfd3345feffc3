// Measurement of the warp-cooperative element solve (xf_element_coop.cuh: four lanes per element) against the one-thread solve
// the stepping kernels use (SolveElementGathered -> SolvePrefactoredSimulPacked), on the SAME elements and vertex records
// (not part of the stepping path):
//   parity      both variants run `iterations` chained solves of every element; the final positions are compared bit for bit
//   latency     a lone warp (32 elements one-thread, 8 elements four-lane): cycles per solve = the arithmetic part of a stage of the
//               barrier-free sweep on a latency-bound mesh (solve + hand-off + bookkeeping)
//   throughput  `warpsPerSm` warps on every SM: element solves per second = what the issue slots of the chip sustain, the bound
//               of a stage at 1M tets
#include "xf_element.cuh"
#include "xf_element_coop.cuh"

#include <string.h>

#include <vector>

namespace xf {
namespace {

struct ProbeElem { // 64 bytes
	float Qi[9]; // [col][row], as xf_get_elements returns it
	float volume;
	float QQ[3], QR[3];
};

__device__ __forceinline__ ElemRec LoadProbeElem(const ProbeElem* elems, uint32_t i) {
	const ProbeElem pe = elems[i];
	ElemRec r;
	r.idx = make_uint4(0u, 1u, 2u, 3u);
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { r.Qi[c][k] = pe.Qi[3 * c + k]; }
	}
	r.volume = pe.volume;
#pragma unroll
	for (int k = 0; k < 3; k++) { r.QQ[k] = pe.QQ[k]; r.QR[k] = pe.QR[k]; }
	r.alpha0 = r.alpha1 = 0.0f;
	return r;
}

// The one-thread solve WITHOUT the two-wide arithmetic: the scalar branch of SolveElementGathered (xf_element.cuh), same bits.
template <int ENERGY>
__device__ __forceinline__ void SolveScalar(const SubstepParams& p, const ElemRec& e, VertexRegs (&v)[4], const ElemCompliance& ec) {
	float P[3][3], F[3][3], g0[4][3], g1[4][3];
	float U0;
	bool haveF;
	Edges<true>(v, P);
	DeviatoricTerm<ENERGY, true>(e, P, U0, g0, F, haveF);
	DeformationGradient<true>(e, P, F);
	const float U1 = VolumetricFromF<true>(e, F, p.a, g1);
	ConstrainBoth<true, false>(NoStore{}, p, e.idx, v, U0, U1, g0, g1, ec.comp0, ec.comp1, ec.alpha0, ec.alpha1, p.damping);
}

// one thread per element (thread t -> element t mod nElems); PACKED: the two-wide arithmetic the stepping kernels run
template <int ENERGY, bool PACKED>
__global__ void __launch_bounds__(256) k_probe_single(const ProbeElem* elems, uint32_t nElems, const double* X, const float* W, SubstepParams p,
                                                     uint32_t iters, double* out, long long* outCycles) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t ei = t % nElems;
	const ElemRec rec = LoadProbeElem(elems, ei);
	VertexRegs v[4];
#pragma unroll
	for (int n = 0; n < 4; n++) {
#pragma unroll
		for (int k = 0; k < 3; k++) { v[n].x[k] = X[12 * (size_t)ei + 3 * n + k]; }
		v[n].w = W[4 * (size_t)ei + n];
		v[n].flags = 0;
	}
	const ElemCompliance ec = ComplianceOf<true>(p, rec.volume);
	const long long t0 = clock64();
	for (uint32_t k = 0; k < iters; k++) {
		if (PACKED) {
			SolveElementGathered<ENERGY, true, true, false>(NoStore{}, p, rec, v, ec);
		} else {
			SolveScalar<ENERGY>(p, rec, v, ec);
		}
	}
	const long long t1 = clock64();
	if (out && t < nElems) {
#pragma unroll
		for (int n = 0; n < 4; n++) {
#pragma unroll
			for (int k = 0; k < 3; k++) { out[12 * (size_t)t + 3 * n + k] = v[n].x[k]; }
		}
	}
	if (outCycles && t == 0) { *outCycles = t1 - t0; }
}

// four lanes per element (thread t -> element (t / 4) mod nElems, vertex t mod 4)
// X3G: every lane also gathers vertex 3's position itself (xf_element_coop.cuh).  The probe gathers once, so in a chain of solves
// lanes 0..2 keep the FIRST x3: the work and the dependence chain through a lane's own record are those of a sweep, where every
// element gathers fresh records, but only the first solve of the chain is comparable with the one-thread variant.
template <int ENERGY, bool X3G>
__global__ void __launch_bounds__(256) k_probe_coop4(const ProbeElem* elems, uint32_t nElems, const double* X, const float* W, SubstepParams p,
                                                    uint32_t iters, double* out, long long* outCycles) {
	__shared__ __align__(16) float sm[(256 / 4) * kCoopQuadFloats];
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t quad = t >> 2;
	const int lane4 = (int)(t & 3u);
	const uint32_t ei = quad % nElems;
	const ElemRec rec = LoadProbeElem(elems, ei);
	double x[3];
#pragma unroll
	for (int k = 0; k < 3; k++) { x[k] = X[12 * (size_t)ei + 3 * lane4 + k]; }
	double x3[3] = { 0.0, 0.0, 0.0 };
	if (X3G) {
#pragma unroll
		for (int k = 0; k < 3; k++) { x3[k] = X[12 * (size_t)ei + 9 + k]; }
	}
	const float w = W[4 * (size_t)ei + lane4];
	const ElemCompliance ec = ComplianceOf<true>(p, rec.volume);
	const DevLane4 ln{ 0xffffffffu, lane4, sm + (threadIdx.x >> 2) * kCoopQuadFloats };
	const long long t0 = clock64();
	for (uint32_t k = 0; k < iters; k++) { SolvePrefactoredSimulCoop4<ENERGY, X3G>(ln, p.a, rec, ec.alpha0, ec.alpha1, x, w, x3); }
	const long long t1 = clock64();
	if (out && quad < nElems) {
#pragma unroll
		for (int k = 0; k < 3; k++) { out[12 * (size_t)quad + 3 * lane4 + k] = x[k]; }
	}
	if (outCycles && t == 0) { *outCycles = t1 - t0; }
}

template <int ENERGY>
void Launch(int coop, int blocks, int threads, const ProbeElem* elems, uint32_t nElems, const double* X, const float* W, const SubstepParams& p,
            uint32_t iters, double* out, long long* cyc) {
	if (coop == 2) {
		k_probe_coop4<ENERGY, true><<<blocks, threads>>>(elems, nElems, X, W, p, iters, out, cyc);
	} else if (coop == 1) {
		k_probe_coop4<ENERGY, false><<<blocks, threads>>>(elems, nElems, X, W, p, iters, out, cyc);
	} else if (coop == 0) {
		k_probe_single<ENERGY, true><<<blocks, threads>>>(elems, nElems, X, W, p, iters, out, cyc);
	} else {
		k_probe_single<ENERGY, false><<<blocks, threads>>>(elems, nElems, X, W, p, iters, out, cyc);
	}
}
void LaunchE(int energy, int coop, int blocks, int threads, const ProbeElem* elems, uint32_t nElems, const double* X, const float* W,
             const SubstepParams& p, uint32_t iters, double* out, long long* cyc) {
	if (energy == (int)XF_ENERGY_YEOH_SKIN_FAST) {
		Launch<XF_ENERGY_YEOH_SKIN_FAST>(coop, blocks, threads, elems, nElems, X, W, p, iters, out, cyc);
	} else {
		Launch<XF_ENERGY_MIXED_SEL>(coop, blocks, threads, elems, nElems, X, W, p, iters, out, cyc);
	}
}

}  // namespace
}  // namespace xf

// elemConsts: nElems x 16 floats {Qi[9] ([col][row]), volume, QQ[3], QR[3]}; X: nElems x 12 doubles (the four gathered vertex
// positions of every element); w: nElems x 4; params4 = {a = 1 + mu/lambda, 1/mu, 1/lambda, dt^2}.
// out8 = {cycles per solve of a lone warp: one-thread, four-lane; element solves per second with warpsPerSm warps on every SM:
//         one-thread, four-lane; doubles that differ between the variants after `iterations` chained solves, doubles compared;
//         SM count, SM clock in kHz}.  outXSingle / outXCoop (optional): nElems x 12 doubles, the final positions of each variant.
// variant bit 0: 0 = lane 3 broadcasts vertex 3, 1 = every lane gathered vertex 3 itself (parity is then meaningful for
// iterations == 1 only, see k_probe_coop4); bit 1: the one-thread side runs the SCALAR arithmetic instead of the two-wide one.
extern "C" int xf_debug_coop_element(int device, int energy, int variant, const float* elemConsts, const double* X, const float* w,
                                     uint32_t nElems, const float* params4, uint32_t iterations, int warpsPerSm, double* out8,
                                     double* outXSingle, double* outXCoop) {
	using namespace xf;
	if (!elemConsts || !X || !w || !params4 || !out8 || nElems == 0 || iterations == 0 || warpsPerSm < 1 || warpsPerSm > 64 || variant < 0 ||
	    variant > 3) {
		return XF_ERR_INVALID;
	}
	if (warpsPerSm > 8 && warpsPerSm % 4 != 0) { return XF_ERR_INVALID; } // one CTA of <= 8 warps, or CTAs of 4 warps
	const int coopKind = 1 + (variant & 1);
	const int singleKind = (variant & 2) ? -1 : 0;
	if (energy != (int)XF_ENERGY_YEOH_SKIN_FAST && energy != (int)XF_ENERGY_MIXED_SEL) { return XF_ERR_UNSUPPORTED; }
	if (cudaSetDevice(device) != cudaSuccess) { return XF_ERR_CUDA; }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { return XF_ERR_CUDA; }
	int clockKHz = 0;
	cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, device);
	SubstepParams p;
	memset(&p, 0, sizeof(p));
	p.a = params4[0]; p.invMu = params4[1]; p.invLambda = params4[2]; p.dt2 = params4[3];
	p.dt = 1.0f; p.invDt = 1.0f;

	ProbeElem* dE = nullptr;
	double *dX = nullptr, *dOutS = nullptr, *dOutC = nullptr;
	float* dW = nullptr;
	long long* dCyc = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	int rc = XF_OK;
	auto ok = [&](cudaError_t e) { if (e != cudaSuccess && rc == XF_OK) { rc = XF_ERR_CUDA; } return e == cudaSuccess; };
	const size_t nX = 12 * (size_t)nElems;
	ok(cudaMalloc(&dE, sizeof(ProbeElem) * nElems));
	ok(cudaMalloc(&dX, sizeof(double) * nX));
	ok(cudaMalloc(&dOutS, sizeof(double) * nX));
	ok(cudaMalloc(&dOutC, sizeof(double) * nX));
	ok(cudaMalloc(&dW, sizeof(float) * 4 * (size_t)nElems));
	ok(cudaMalloc(&dCyc, sizeof(long long) * 2));
	ok(cudaEventCreate(&e0));
	ok(cudaEventCreate(&e1));
	if (rc == XF_OK) {
		ok(cudaMemcpy(dE, elemConsts, sizeof(ProbeElem) * nElems, cudaMemcpyHostToDevice));
		ok(cudaMemcpy(dX, X, sizeof(double) * nX, cudaMemcpyHostToDevice));
		ok(cudaMemcpy(dW, w, sizeof(float) * 4 * (size_t)nElems, cudaMemcpyHostToDevice));
		ok(cudaMemset(dCyc, 0, sizeof(long long) * 2));
	}
	std::vector<double> hS(nX), hC(nX);
	if (rc == XF_OK) {
		// ---- parity: every element, `iterations` chained solves, both variants
		LaunchE(energy, singleKind, (int)((nElems + 255) / 256), 256, dE, nElems, dX, dW, p, iterations, dOutS, nullptr);
		LaunchE(energy, coopKind, (int)((4 * (size_t)nElems + 255) / 256), 256, dE, nElems, dX, dW, p, iterations, dOutC, nullptr);
		ok(cudaDeviceSynchronize());
		ok(cudaMemcpy(hS.data(), dOutS, sizeof(double) * nX, cudaMemcpyDeviceToHost));
		ok(cudaMemcpy(hC.data(), dOutC, sizeof(double) * nX, cudaMemcpyDeviceToHost));
	}
	if (rc == XF_OK) {
		uint64_t bad = 0;
		for (size_t i = 0; i < nX; i++) { bad += memcmp(&hS[i], &hC[i], sizeof(double)) != 0 ? 1u : 0u; }
		out8[4] = (double)bad;
		out8[5] = (double)nX;
		if (outXSingle) { memcpy(outXSingle, hS.data(), sizeof(double) * nX); }
		if (outXCoop) { memcpy(outXCoop, hC.data(), sizeof(double) * nX); }
		// ---- latency: a lone warp
		LaunchE(energy, singleKind, 1, 32, dE, nElems, dX, dW, p, iterations, nullptr, dCyc);
		LaunchE(energy, coopKind, 1, 32, dE, nElems, dX, dW, p, iterations, nullptr, dCyc + 1);
		ok(cudaDeviceSynchronize());
		long long cyc[2] = { 0, 0 };
		ok(cudaMemcpy(cyc, dCyc, sizeof(cyc), cudaMemcpyDeviceToHost));
		out8[0] = (double)cyc[0] / (double)iterations;
		out8[1] = (double)cyc[1] / (double)iterations;
	}
	if (rc == XF_OK) {
		// ---- throughput: warpsPerSm warps on every SM: CTAs of 4 warps (one per scheduler) when warpsPerSm is a multiple of 4,
		// else one CTA of warpsPerSm warps per SM
		const bool quads = warpsPerSm % 4 == 0;
		const int threads = quads ? 128 : warpsPerSm * 32;
		const int blocks = prop.multiProcessorCount * (quads ? warpsPerSm / 4 : 1);
		for (int v = 0; v < 2 && rc == XF_OK; v++) {
			const bool coop = v == 1;
			LaunchE(energy, coop ? coopKind : singleKind, blocks, threads, dE, nElems, dX, dW, p, iterations, nullptr, nullptr); // warm-up
			float best = 0.0f;
			for (int rep = 0; rep < 3; rep++) {
				ok(cudaEventRecord(e0));
				LaunchE(energy, coop ? coopKind : singleKind, blocks, threads, dE, nElems, dX, dW, p, iterations, nullptr, nullptr);
				ok(cudaEventRecord(e1));
				ok(cudaEventSynchronize(e1));
				float ms = 0.0f;
				ok(cudaEventElapsedTime(&ms, e0, e1));
				if (rep == 0 || ms < best) { best = ms; }
			}
			const double solves = (double)blocks * (double)threads / (coop ? 4.0 : 1.0) * (double)iterations;
			out8[2 + v] = best > 0.0f ? solves / ((double)best * 1e-3) : 0.0;
		}
		out8[6] = (double)prop.multiProcessorCount;
		out8[7] = (double)clockKHz;
	}
	if (cudaGetLastError() != cudaSuccess && rc == XF_OK) { rc = XF_ERR_CUDA; }
	cudaFree(dE); cudaFree(dX); cudaFree(dOutS); cudaFree(dOutC); cudaFree(dW); cudaFree(dCyc);
	if (e0) { cudaEventDestroy(e0); }
	if (e1) { cudaEventDestroy(e1); }
	return rc;
}
