// Quad-footprint probe: fetch throughput of fp32 bilinear TEX as a function of how the 4 lanes of a
// quad are spread (dx, dy pattern), with quad base positions scattered inside an L1-resident window.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
struct Pat { float dx[4], dy[4]; };
__global__ void k_bench(cudaTextureObject_t tex, float *out, int iters, Pat p, float scale, int region) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int quad = t >> 2, l = t & 3;
	unsigned h = quad * 2654435761u;
	float bx = 64.0f + (float)(h % region) + 0.37f, by = 64.0f + (float)((h >> 12) % region) + 0.61f;
	float x = bx + p.dx[l] * scale, y = by + p.dy[l] * scale;
	float acc = 0.f;
#pragma unroll 4
	for (int i = 0; i < iters; ++i) {
		acc += tex2D<float>(tex, x, y);
		x += 1.37f; y += 0.73f;
		if (x > 64.0f + region) x -= region; if (y > 64.0f + region) y -= region;
	}
	out[t] = acc;
}
int main() {
	const int W = 2048, H = 2048;
	std::vector<float> img((size_t)W * H, 1.0f);
	cudaChannelFormatDesc d = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	cudaArray_t arr; CK(cudaMallocArray(&arr, &d, W, H));
	CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	cudaResourceDesc r; memset(&r, 0, sizeof(r)); r.resType = cudaResourceTypeArray; r.res.array.array = arr;
	cudaTextureDesc t; memset(&t, 0, sizeof(t)); t.addressMode[0] = t.addressMode[1] = cudaAddressModeClamp; t.filterMode = cudaFilterModeLinear; t.readMode = cudaReadModeElementType;
	cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &r, &t, nullptr));
	float *dacc; const int threads = 148 * 2048 * 4; CK(cudaMalloc(&dacc, threads * 4));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	struct { const char *name; Pat p; } pats[] = {
		{"same point          ", {{0, 0, 0, 0}, {0, 0, 0, 0}}},
		{"row dx=1            ", {{0, 1, 2, 3}, {0, 0, 0, 0}}},
		{"row dx=2 (current)  ", {{0, 2, 4, 6}, {0, 0, 0, 0}}},
		{"col dy=1            ", {{0, 0, 0, 0}, {0, 1, 2, 3}}},
		{"col dy=2            ", {{0, 0, 0, 0}, {0, 2, 4, 6}}},
		{"2x2 d=1             ", {{0, 1, 0, 1}, {0, 0, 1, 1}}},
		{"2x2 d=2             ", {{0, 2, 0, 2}, {0, 0, 2, 2}}},
		{"diamond (checker)   ", {{0, 2, 1, 3}, {0, 0, 1, 1}}},
		{"2x2 dx=2 dy=1       ", {{0, 2, 0, 2}, {0, 0, 1, 1}}},
		{"2x2 dx=1 dy=2       ", {{0, 1, 0, 1}, {0, 0, 2, 2}}},
		{"row dx=4            ", {{0, 4, 8, 12}, {0, 0, 0, 0}}},
		{"scatter 40px        ", {{0, 40, 7, 33}, {0, 13, 38, 25}}},
	};
	for (int region : {96, 1500}) for (auto &pp : pats) {
		float ms = 0; const int iters = 256;
		for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); k_bench<<<threads / 256, 256>>>(tex, dacc, iters, pp.p, 1.0f, region); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
		printf("region %4d  %s %8.3f ms %8.1f Gfetch/s\n", region, pp.name, ms, (double)threads * iters / ms * 1e-6);
	}
	return 0;
}
