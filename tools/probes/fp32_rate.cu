// Probe: issue rate of scalar vs packed (f32x2) round-to-nearest fp32 instructions on sm_100a, per SM sub-partition.
// Decides whether the bit-exact element body (no FMA contraction allowed: separate FMUL / FADD) can be packed two-wide.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp32_rate tools/probes/fp32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int VARIANT, int ILP>
__global__ void __launch_bounds__(256) k_rate(float* out, int iters, float seed) {
	float a[ILP], b[ILP];
	float2 p[ILP], q[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) {
		a[i] = seed + i + threadIdx.x; b[i] = seed * 0.999f + 1e-3f * (float)(i + (int)blockIdx.x % 3);
		p[i] = make_float2(a[i], a[i] + 0.5f); q[i] = make_float2(b[i], b[i]);
	}
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) {
			if (VARIANT == 0) { a[i] = __fmul_rn(a[i], b[i]); }
			if (VARIANT == 1) { a[i] = __fadd_rn(a[i], b[i]); }
			if (VARIANT == 2) { p[i] = __fmul2_rn(p[i], q[i]); }
			if (VARIANT == 3) { p[i] = __fadd2_rn(p[i], q[i]); }
			if (VARIANT == 4) { a[i] = __fmaf_rn(a[i], b[i], b[i]); }
			if (VARIANT == 5) { p[i] = __ffma2_rn(p[i], q[i], q[i]); }
			if (VARIANT == 6) { a[i] = __fadd_rn(__fmul_rn(a[i], b[i]), b[i]); }               // the exact kernel's pattern: mul then add
			if (VARIANT == 7) { p[i] = __fadd2_rn(__fmul2_rn(p[i], q[i]), q[i]); }
		}
	}
	float s = 0.0f;
#pragma unroll
	for (int i = 0; i < ILP; i++) { s += a[i] + p[i].x + p[i].y; }
	if (s == 123.456f) { out[0] = s; }
}

template <int VARIANT, int ILP>
void run(const char* name, int opsPerIter, int warpsPerSm) {
	int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
	float* out; cudaMalloc(&out, 4);
	const int iters = 20000;
	const int threads = 256, blocksPerSm = warpsPerSm * 32 / threads > 0 ? warpsPerSm * 32 / threads : 1;
	const int th = warpsPerSm * 32 < threads ? warpsPerSm * 32 : threads;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_rate<VARIANT, ILP><<<prop.multiProcessorCount * blocksPerSm, th>>>(out, 100, 1.0f);
	cudaEventRecord(e0);
	k_rate<VARIANT, ILP><<<prop.multiProcessorCount * blocksPerSm, th>>>(out, iters, 1.0f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
	const double cycles = ms * 1e-3 * clk * 1e3;
	const double warpInstr = (double)iters * ILP * opsPerIter * warpsPerSm; // per SM
	printf("%-28s ILP %d warps/SM %2d: %.3f warp-instr/clk/SM  (%.3f per SMSP), flop-lanes/clk/SM %.1f\n", name, ILP, warpsPerSm, warpInstr / cycles,
	       warpInstr / cycles / 4.0, warpInstr / cycles * 32.0 * ((VARIANT == 2 || VARIANT == 3 || VARIANT == 5 || VARIANT == 7) ? 2.0 : 1.0));
	cudaFree(out);
}

int main() {
	for (int w : { 4, 8, 16, 32 }) {
		if (w == 4) {
			run<0, 8>("FMUL", 1, 4); run<1, 8>("FADD", 1, 4); run<2, 8>("FMUL2", 1, 4); run<3, 8>("FADD2", 1, 4); run<4, 8>("FFMA", 1, 4);
			run<5, 8>("FFMA2", 1, 4); run<6, 8>("FMUL+FADD", 2, 4); run<7, 8>("FMUL2+FADD2", 2, 4);
		} else if (w == 8) {
			run<0, 8>("FMUL", 1, 8); run<1, 8>("FADD", 1, 8); run<2, 8>("FMUL2", 1, 8); run<3, 8>("FADD2", 1, 8); run<4, 8>("FFMA", 1, 8);
			run<5, 8>("FFMA2", 1, 8); run<6, 8>("FMUL+FADD", 2, 8); run<7, 8>("FMUL2+FADD2", 2, 8);
		} else if (w == 16) {
			run<0, 8>("FMUL", 1, 16); run<1, 8>("FADD", 1, 16); run<2, 8>("FMUL2", 1, 16); run<3, 8>("FADD2", 1, 16); run<4, 8>("FFMA", 1, 16);
			run<5, 8>("FFMA2", 1, 16); run<6, 8>("FMUL+FADD", 2, 16); run<7, 8>("FMUL2+FADD2", 2, 16);
		} else {
			run<0, 8>("FMUL", 1, 32); run<1, 8>("FADD", 1, 32); run<2, 8>("FMUL2", 1, 32); run<3, 8>("FADD2", 1, 32); run<4, 8>("FFMA", 1, 32);
			run<5, 8>("FFMA2", 1, 32); run<6, 8>("FMUL+FADD", 2, 32); run<7, 8>("FMUL2+FADD2", 2, 32);
		}
	}
	// latency: ILP 1, one warp per SMSP
	run<0, 1>("FMUL lat", 1, 4); run<2, 1>("FMUL2 lat", 1, 4); run<1, 1>("FADD lat", 1, 4); run<3, 1>("FADD2 lat", 1, 4);
	return 0;
}
