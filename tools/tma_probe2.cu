#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// variant 0: 1-D bulk copy (no tensor map); variant 1: tensor 2D with map in param; 
template <int V>
__global__ void k(const __grid_constant__ CUtensorMap tmap, const float *src, float *out, int cx, int cy) {
	extern __shared__ __align__(1024) unsigned char smem[];
	float *tile = (float *)smem;
	__shared__ __align__(8) uint64_t mbar_s;
	uint64_t *mbar = &mbar_s;
	const int tid = threadIdx.x;
	const uint32_t bytes = 32 * 16 * 4;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
		if (V == 0)
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)), "l"(src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
		else if (V == 1)
			asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
			             ::"r"(smem_u32(tile)), "l"(&tmap), "r"(cx), "r"(cy), "r"(smem_u32(mbar)) : "memory");
		else
			asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
			             ::"r"(smem_u32(tile)), "l"(&tmap), "r"(cx), "r"(cy), "r"(smem_u32(mbar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
	for (int i = tid; i < 512; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv) {
	const int variant = argc > 1 ? atoi(argv[1]) : 0;
	const int pitch = 256, rows = 64;
	std::vector<float> h(pitch * rows); for (int i = 0; i < pitch * rows; ++i) h[i] = (float)i;
	float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 4096); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
	void *fn = nullptr; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	CUtensorMap m;
	cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
	cuuint32_t box[2] = {32, 16}; cuuint32_t es[2] = {1, 1};
	CUresult r = ((PFN_encodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("variant %d encode %d\n", variant, (int)r);
	const int cx = argc > 2 ? atoi(argv[2]) : 0, cy = argc > 3 ? atoi(argv[3]) : 0;
	if (variant == 0) k<0><<<1, 128, 4096>>>(m, d, o, cx, cy); else if (variant == 1) k<1><<<1, 128, 4096>>>(m, d, o, cx, cy); else k<2><<<1, 128, 4096>>>(m, d, o, cx, cy);
	cudaError_t e = cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
	std::vector<float> res(512); cudaMemcpy(res.data(), o, 2048, cudaMemcpyDeviceToHost);
	printf("tile[0]=%g tile[33]=%g (1-D expects 33, 2-D expects %d)\n", res[0], res[33], pitch + 1);
	return 0;
}
