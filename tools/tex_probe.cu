// Texture-unit probe for B200 (run under gpurun): (1) dumps bilinear samples so the exact hardware
// filter arithmetic can be modelled offline, (2) measures fetch throughput of the sampling modes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_sample(cudaTextureObject_t lin, cudaTextureObject_t gat, const float2 *xy, float *out, float4 *g, int n) {
	int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
	out[i] = tex2D<float>(lin, xy[i].x, xy[i].y);
	g[i] = tex2Dgather<float4>(gat, xy[i].x, xy[i].y, 0);
}

template <int MODE>
__global__ void k_bench(cudaTextureObject_t tex, float *out, int iters, float step, int W) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	// 16x2 pixel footprint per warp, walking diagonally: similar locality to the NCC kernel
	float x = (float)((t & 15) + ((t >> 5) * 16) % (W - 64)) + 0.37f;
	float y = (float)(((t >> 4) & 1) + ((t >> 5) * 16 / (W - 64)) * 2 % (W - 64)) + 0.61f;
	float acc = 0.f;
#pragma unroll 4
	for (int i = 0; i < iters; ++i) {
		if (MODE == 0) acc += tex2D<float>(tex, x, y);
		if (MODE == 1) { float4 g = tex2Dgather<float4>(tex, x, y, 0); acc += g.x + g.y + g.z + g.w; }
		if (MODE == 2) acc += tex2DLayered<float>(tex, x, y, i & 7);
		x += step; y += 0.31f * step;
		if (x > W - 8) { x -= (W - 16); } if (y > W - 8) { y -= (W - 16); }
	}
	out[t] = acc;
}

int main() {
	const int W = 1024, H = 1024;
	std::vector<float> img((size_t)W * H);
	srand(7);
	for (auto &v : img) v = (float)(rand() % 256000) / 1000.0f;
	cudaChannelFormatDesc d = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	cudaArray_t arr, arrg, arrl;
	CK(cudaMallocArray(&arr, &d, W, H));
	CK(cudaMallocArray(&arrg, &d, W, H, cudaArrayTextureGather));
	CK(cudaMalloc3DArray(&arrl, &d, make_cudaExtent(W, H, 8), cudaArrayLayered));
	CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	CK(cudaMemcpy2DToArray(arrg, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
	for (int l = 0; l < 8; ++l) { cudaMemcpy3DParms p; memset(&p, 0, sizeof(p)); p.srcPtr = make_cudaPitchedPtr(img.data(), W * 4, W, H);
		p.dstArray = arrl; p.dstPos = make_cudaPos(0, 0, l); p.extent = make_cudaExtent(W, H, 1); p.kind = cudaMemcpyHostToDevice; CK(cudaMemcpy3D(&p)); }
	auto mk = [&](cudaArray_t a, cudaTextureFilterMode fm) { cudaResourceDesc r; memset(&r, 0, sizeof(r)); r.resType = cudaResourceTypeArray; r.res.array.array = a;
		cudaTextureDesc t; memset(&t, 0, sizeof(t)); t.addressMode[0] = t.addressMode[1] = t.addressMode[2] = cudaAddressModeClamp; t.filterMode = fm; t.readMode = cudaReadModeElementType;
		cudaTextureObject_t o; CK(cudaCreateTextureObject(&o, &r, &t, nullptr)); return o; };
	cudaTextureObject_t lin = mk(arr, cudaFilterModeLinear), pnt = mk(arr, cudaFilterModePoint), gat = mk(arrg, cudaFilterModePoint), lay = mk(arrl, cudaFilterModeLinear);

	// ---- (1) sample dump
	const int n = 1 << 18;
	std::vector<float2> xy(n);
	for (int i = 0; i < n; ++i) {
		float x = (float)(rand() % (W * 1000)) / 1000.0f, y = (float)(rand() % (H * 1000)) / 1000.0f;
		if (i % 7 == 0) x = (float)(rand() % W) + (float)(rand() % 513) / 512.0f;      // exact 1/512 steps
		if (i % 11 == 0) { x = (float)(rand() % 40) - 20.0f + 0.123f * (rand() % 9); }    // border clamp
		if (i % 13 == 0) { y = (float)H - 3.0f + 0.37f * (rand() % 17); }
		xy[i] = make_float2(x, y);
	}
	float2 *dxy; float *dout; float4 *dg;
	CK(cudaMalloc(&dxy, n * 8)); CK(cudaMalloc(&dout, n * 4)); CK(cudaMalloc(&dg, n * 16));
	CK(cudaMemcpy(dxy, xy.data(), n * 8, cudaMemcpyHostToDevice));
	k_sample<<<n / 256, 256>>>(lin, gat, dxy, dout, dg, n);
	CK(cudaDeviceSynchronize());
	std::vector<float> out(n); std::vector<float4> g(n);
	CK(cudaMemcpy(out.data(), dout, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g.data(), dg, n * 16, cudaMemcpyDeviceToHost));
	system("mkdir -p gpurun_out");
	FILE *f = fopen("gpurun_out/tex_probe.bin", "wb");
	int hdr[4] = {W, H, n, 0}; fwrite(hdr, 4, 4, f);
	fwrite(img.data(), 4, img.size(), f); fwrite(xy.data(), 8, n, f); fwrite(out.data(), 4, n, f); fwrite(g.data(), 16, n, f);
	fclose(f);

	// ---- (2) throughput
	float *dacc; const int threads = 148 * 2048 * 4; CK(cudaMalloc(&dacc, threads * 4));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (float step : {0.05f, 1.0f, 3.7f}) {
		for (int mode = 0; mode < 4; ++mode) {
			const int iters = 512; float ms = 0;
			for (int rep = 0; rep < 3; ++rep) {
				cudaEventRecord(e0);
				if (mode == 0) k_bench<0><<<threads / 256, 256>>>(lin, dacc, iters, step, W);
				if (mode == 1) k_bench<1><<<threads / 256, 256>>>(gat, dacc, iters, step, W);
				if (mode == 2) k_bench<2><<<threads / 256, 256>>>(lay, dacc, iters, step, W);
				if (mode == 3) k_bench<0><<<threads / 256, 256>>>(pnt, dacc, iters, step, W);
				cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
			}
			const char *nm[] = {"linear fp32", "gather4 fp32", "layered linear fp32", "point fp32"};
			printf("step %.2f %-20s %8.3f ms  %8.2f Gfetch/s\n", step, nm[mode], ms, (double)threads * iters / ms * 1e-6);
		}
	}
	return 0;
}
