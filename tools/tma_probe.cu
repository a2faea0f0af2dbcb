#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <vector>
#include <cstdlib>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int PW = 56, PH = 18;
__global__ void k(const __grid_constant__ CUtensorMap tmap, float *out, int cx, int cy) {
	extern __shared__ __align__(128) unsigned char smem[];
	float *tile = (float *)smem;
	uint64_t *mbar = (uint64_t *)(tile + PW * PH);
	const int tid = threadIdx.x;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(PW * PH * 4) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		             ::"r"(smem_u32(tile)), "l"(&tmap), "r"(cx), "r"(cy), "r"(smem_u32(mbar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
	for (int i = tid; i < PW * PH; i += blockDim.x) out[i] = tile[i];
}
__global__ void k2(const CUtensorMap *tmapp, float *out, int cx, int cy, int bytes) {
	extern __shared__ __align__(128) unsigned char smem[];
	float *tile = (float *)smem;
	uint64_t *mbar = (uint64_t *)(smem + 8192);
	const int tid = threadIdx.x;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(tmapp) : "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		             ::"r"(smem_u32(tile)), "l"(tmapp), "r"(cx), "r"(cy), "r"(smem_u32(mbar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
	for (int i = tid; i < bytes / 4; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv) {
	const int variant = argc > 1 ? atoi(argv[1]) : 0;
	const int pitch = 272, rows = 272;
	std::vector<float> h(pitch * rows); for (int i = 0; i < pitch * rows; ++i) h[i] = (float)i;
	float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, PW * PH * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
	void *fn = nullptr; cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	printf("entry: %d %p q=%d\n", (int)e, fn, (int)q);
	CUtensorMap m;
	cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
	cuuint32_t box[2] = {PW, PH}; cuuint32_t es[2] = {1, 1};
	CUresult r = ((PFN_encodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode: %d\n", (int)r);
	if (variant == 0) {
	k<<<1, 128, PW * PH * 4 + 16>>>(m, o, 3, 3);
	e = cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
	std::vector<float> res(PW * PH); cudaMemcpy(res.data(), o, PW * PH * 4, cudaMemcpyDeviceToHost);
	printf("tile[0]=%g (expect %d) tile[PW+1]=%g (expect %d)\n", res[0], 3 * pitch + 3, res[PW + 1], 4 * pitch + 4);
	k<<<1, 128, PW * PH * 4 + 16>>>(m, o, 250, 260);
	e = cudaDeviceSynchronize(); printf("sync2: %s\n", cudaGetErrorString(e));
	cudaMemcpy(res.data(), o, PW * PH * 4, cudaMemcpyDeviceToHost);
	printf("oob: tile[0]=%g (expect %d) tile[30]=%g (expect 0)\n", res[0], 260 * pitch + 250, res[30]);
	return 0; }
	for (int bw : {variant == 1 ? 32 : 56}) {
		cuuint32_t box2[2] = {(cuuint32_t)bw, 16};
		CUtensorMap m2; r = ((PFN_encodeTiled)fn)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		CUtensorMap *dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &m2, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
		float *o2; cudaMalloc(&o2, 8192);
		k2<<<1, 128, 8192 + 16>>>(dm, o2, 3, 3, bw * 16 * 4);
		e = cudaDeviceSynchronize(); printf("k2 box %d: encode %d sync: %s\n", bw, (int)r, cudaGetErrorString(e));
		std::vector<float> r2(bw * 16); cudaMemcpy(r2.data(), o2, bw * 16 * 4, cudaMemcpyDeviceToHost);
		printf("  tile[0]=%g (expect %d) tile[bw+1]=%g (expect %d)\n", r2[0], 3 * pitch + 3, r2[bw + 1], 4 * pitch + 4);
	}
	return 0;
}
