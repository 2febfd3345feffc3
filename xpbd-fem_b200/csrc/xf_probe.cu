// Measurement probes (not part of the stepping path): the L2 bandwidth the roofline of an L2-resident mesh is quoted
// against ("L2 peak is not in MEASURED_PEAKS.json; builder must measure it with a resident-buffer copy"), and
// a torn-record stress test for the one hardware property the barrier-free schedules rely on: a 32-byte-aligned 256-bit
// access is ONE transaction at L2 (and across NVLink), so a reader never sees half of a record.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/xpbd_fem_b200.h"

namespace {

__device__ __forceinline__ void Ld32(const void* p, double& a, double& b, double& c, double& d) {
	asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void St32(void* p, double a, double b, double c, double d) {
	asm volatile("st.relaxed.gpu.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void Ld32Sys(const void* p, double& a, double& b, double& c, double& d) {
	asm volatile("ld.relaxed.sys.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void St32Sys(void* p, double a, double b, double c, double d) {
	asm volatile("st.relaxed.sys.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// MODE 0: copy src -> dst (one 256-bit load + one 256-bit store per 32 bytes; the sweep's gather + scatter mix)
// MODE 1: read only (the loads are kept alive through a never-true store)
// The buffers stay resident in L2 (the caller sizes them); every pass of the grid walks them once, fully coalesced.
template <int MODE>
__global__ void __launch_bounds__(256) k_l2_stream(const double4* __restrict__ src, double4* __restrict__ dst, uint32_t n, uint32_t passes) {
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	double acc = 0.0;
	for (uint32_t p = 0; p < passes; p++) {
		for (uint32_t i = gtid; i < n; i += gsize) {
			double a, b, c, d;
			Ld32(src + i, a, b, c, d);
			if (MODE == 0) {
				St32(dst + i, a, b, c, d);
			} else {
				acc += a + b + c + d;
			}
		}
	}
	if (MODE == 1 && acc == 123.456) { dst[gtid].x = acc; }
}

// Torn-record stress.  A record is four 8-byte words that all encode ONE counter value (word k = counter * 4 + k, as doubles'
// bit patterns).  Writer CTAs rewrite `nRecords` records over and over with increasing counters, one 256-bit store each;
// reader CTAs (other SMs, or - with remote = 1 - this GPU reading/writing a peer's memory through an IPC/P2P mapping) load
// them with one 256-bit load and count records whose four words do not belong to the same counter.
__global__ void __launch_bounds__(256) k_torn_records(double4* rec, uint32_t nRecords, uint32_t writerBlocks, uint32_t rounds, int sysScope,
                                                      unsigned long long* outReads, unsigned long long* outTorn) {
	const bool writer = blockIdx.x < writerBlocks;
	if (writer) {
		const uint32_t wtid = blockIdx.x * blockDim.x + threadIdx.x, wsize = writerBlocks * blockDim.x;
		// stores are fire-and-forget while the readers wait for every load: the writers run 8x the rounds so that they are
		// still rewriting records when the last reader finishes
		for (uint32_t r = 1; r <= 8u * rounds; r++) {
			for (uint32_t i = wtid; i < nRecords; i += wsize) {
				const long long base = (long long)r * 4;
				const double a = __longlong_as_double(base), b = __longlong_as_double(base + 1), c = __longlong_as_double(base + 2),
				             d = __longlong_as_double(base + 3);
				if (sysScope) { St32Sys(rec + i, a, b, c, d); } else { St32(rec + i, a, b, c, d); }
			}
		}
	} else {
		const uint32_t rtid = (blockIdx.x - writerBlocks) * blockDim.x + threadIdx.x, rsize = (gridDim.x - writerBlocks) * blockDim.x;
		unsigned long long reads = 0, torn = 0;
		for (uint32_t r = 0; r < rounds; r++) {
			for (uint32_t i = rtid; i < nRecords; i += rsize) {
				double a, b, c, d;
				if (sysScope) { Ld32Sys(rec + i, a, b, c, d); } else { Ld32(rec + i, a, b, c, d); }
				const long long wa = __double_as_longlong(a), wb = __double_as_longlong(b), wc = __double_as_longlong(c), wd = __double_as_longlong(d);
				reads++;
				if (!(wb == wa + 1 && wc == wa + 2 && wd == wa + 3 && (wa & 3) == 0)) { torn++; }
			}
		}
		atomicAdd(outReads, reads);
		atomicAdd(outTorn, torn);
	}
}

}  // namespace

// Bandwidth of an L2-resident streaming copy (mode 0: bytes read + bytes written per second) or read (mode 1), in GB/s,
// best of `reps` launches.  `bytes` = size of EACH buffer (two buffers in copy mode).
extern "C" int xf_debug_l2_bandwidth(int device, int mode, uint64_t bytes, uint32_t passes, int reps, int blocksPerSm, double* outGBs) {
	if (!outGBs || bytes < 4096 || passes == 0) { return XF_ERR_INVALID; }
	if (cudaSetDevice(device) != cudaSuccess) { return XF_ERR_CUDA; }
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { return XF_ERR_CUDA; }
	const uint32_t n = (uint32_t)(bytes / sizeof(double4));
	double4 *src = nullptr, *dst = nullptr;
	if (cudaMalloc(&src, n * sizeof(double4)) != cudaSuccess || cudaMalloc(&dst, n * sizeof(double4)) != cudaSuccess) { cudaFree(src); return XF_ERR_NOMEM; }
	cudaMemset(src, 0, n * sizeof(double4));
	cudaMemset(dst, 0, n * sizeof(double4));
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	const int grid = prop.multiProcessorCount * (blocksPerSm > 0 ? blocksPerSm : 8);
	double best = 0.0;
	int rc = XF_OK;
	for (int r = 0; r < reps + 1; r++) { // first launch warms the buffers into L2
		cudaEventRecord(a);
		if (mode == 0) { k_l2_stream<0><<<grid, 256>>>(src, dst, n, passes); } else { k_l2_stream<1><<<grid, 256>>>(src, dst, n, passes); }
		cudaEventRecord(b);
		if (cudaEventSynchronize(b) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = XF_ERR_CUDA; break; }
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, a, b);
		const double moved = (double)n * sizeof(double4) * (double)passes * (mode == 0 ? 2.0 : 1.0);
		const double gbs = moved / (ms * 1e-3) / 1e9;
		if (r > 0 && gbs > best) { best = gbs; }
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	cudaFree(src);
	cudaFree(dst);
	*outGBs = best;
	return rc;
}

// Torn-record stress on `device`.  remoteDevice < 0: records live on `device`, writers and readers are CTAs of one launch on
// different SMs (gpu-scope accesses).  remoteDevice >= 0: records live on `remoteDevice` (peer access enabled here); `device`
// runs the WRITERS with sys-scope stores over NVLink while `remoteDevice` runs the readers on its local memory - the pattern of
// the partitioned schedule (peer stores, local polls).  Returns the number of reads and of torn records seen.
extern "C" int xf_debug_torn_records(int device, int remoteDevice, uint32_t nRecords, uint32_t rounds, uint64_t* outReads, uint64_t* outTorn) {
	if (!outReads || !outTorn || nRecords == 0) { return XF_ERR_INVALID; }
	const int home = remoteDevice >= 0 ? remoteDevice : device;
	if (cudaSetDevice(home) != cudaSuccess) { return XF_ERR_CUDA; }
	double4* rec = nullptr;
	unsigned long long* counters = nullptr;
	if (cudaMalloc(&rec, nRecords * sizeof(double4)) != cudaSuccess || cudaMalloc(&counters, 2 * sizeof(unsigned long long)) != cudaSuccess) { return XF_ERR_NOMEM; }
	{ // counter 0 everywhere: words 0,1,2,3
		long long* host = (long long*)malloc(nRecords * sizeof(double4));
		for (uint32_t i = 0; i < nRecords; i++) {
			for (int k = 0; k < 4; k++) { host[4 * (size_t)i + k] = k; }
		}
		cudaMemcpy(rec, host, nRecords * sizeof(double4), cudaMemcpyHostToDevice);
		free(host);
	}
	cudaMemset(counters, 0, 2 * sizeof(unsigned long long));
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, home);
	int rc = XF_OK;
	if (remoteDevice < 0) {
		const uint32_t blocks = (uint32_t)prop.multiProcessorCount, writers = blocks / 4;
		k_torn_records<<<blocks, 256>>>(rec, nRecords, writers, rounds, 0, counters, counters + 1);
		if (cudaDeviceSynchronize() != cudaSuccess) { rc = XF_ERR_CUDA; }
	} else {
		int can = 0;
		cudaDeviceCanAccessPeer(&can, device, remoteDevice);
		if (!can) { cudaFree(rec); cudaFree(counters); return XF_ERR_UNSUPPORTED; }
		cudaSetDevice(device);
		cudaError_t pe = cudaDeviceEnablePeerAccess(remoteDevice, 0);
		if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { rc = XF_ERR_CUDA; }
		cudaGetLastError();
		unsigned long long* dummy = nullptr;
		cudaMalloc(&dummy, 2 * sizeof(unsigned long long));
		cudaMemset(dummy, 0, 2 * sizeof(unsigned long long));
		cudaStream_t sw, sr;
		cudaStreamCreateWithFlags(&sw, cudaStreamNonBlocking);
		cudaSetDevice(remoteDevice);
		cudaStreamCreateWithFlags(&sr, cudaStreamNonBlocking);
		const uint32_t blocks = (uint32_t)prop.multiProcessorCount;
		// readers on the home GPU (all blocks read: writerBlocks = 0), writers on `device` over NVLink (all blocks write)
		k_torn_records<<<blocks, 256, 0, sr>>>(rec, nRecords, 0, rounds, 1, counters, counters + 1);
		cudaSetDevice(device);
		k_torn_records<<<blocks, 256, 0, sw>>>(rec, nRecords, blocks, rounds, 1, dummy, dummy + 1);
		if (cudaStreamSynchronize(sw) != cudaSuccess) { rc = XF_ERR_CUDA; }
		cudaSetDevice(remoteDevice);
		if (cudaStreamSynchronize(sr) != cudaSuccess) { rc = XF_ERR_CUDA; }
		cudaStreamDestroy(sr);
		cudaSetDevice(device);
		cudaStreamDestroy(sw);
		cudaFree(dummy);
		cudaSetDevice(home);
	}
	unsigned long long host[2] = { 0, 0 };
	cudaMemcpy(host, counters, sizeof(host), cudaMemcpyDeviceToHost);
	cudaFree(rec);
	cudaFree(counters);
	*outReads = host[0];
	*outTorn = host[1];
	return rc;
}
