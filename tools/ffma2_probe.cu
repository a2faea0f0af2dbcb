// Throughput of packed fp32 FFMA2 vs scalar FFMA on B200 (issue-slot economics of the NCC inner loop).
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
	float x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
	unsigned long long p[4]; for (int i = 0; i < 4; ++i) p[i] = ((unsigned long long)__float_as_uint(x[2 * i + 1]) << 32) | __float_as_uint(x[2 * i]);
	unsigned long long A = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a), B = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
		} else {
#pragma unroll
			for (int i = 0; i < 4; ++i) asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(A), "l"(B));
		}
	}
	float s = 0; for (int i = 0; i < 8; ++i) s += x[i]; for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
	float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
		const int iters = 20000; float ms;
		cudaEventRecord(e0);
		if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
		cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
		double fmas = 148.0 * 8 * 256 * iters * 8;
		if (rep == 2) printf("%s: %.3f ms, %.2f TFMA/s (%.1f fp32 FMA lanes/clk/SM at 1.92 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / 1.92e9);
	}
	return 0;
}
