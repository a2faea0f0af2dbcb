// GPU depth-map fusion: the C-ABI of include/apd_fusion.h. Follows RunFusion (APD.cpp:826-977) with the raster-order
// mask dependency resolved exactly by rounds of claims (see the header). This TU is compiled with IEEE arithmetic
// (no fast-math, no FMA contraction): the reference's fusion is host code, every float operation rounded on its own.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/apd_fusion.h"

namespace {

struct FCam { float K[9], R[9], t[3]; };
struct FView { const uchar3 *bgr; const float *depth; const float *normal; const uint8_t *states; const uint8_t *block; uint8_t *mask; int *claim; };

// Get3DPointonWorld, APD.cpp:776-800
__device__ __forceinline__ float3 point_on_world(int x, int y, float depth, const FCam &c) {
	float3 p, t;
	p.x = depth * (x - c.K[2]) / c.K[0];
	p.y = depth * (y - c.K[5]) / c.K[4];
	p.z = depth;
	t.x = c.R[0] * p.x + c.R[3] * p.y + c.R[6] * p.z;
	t.y = c.R[1] * p.x + c.R[4] * p.y + c.R[7] * p.z;
	t.z = c.R[2] * p.x + c.R[5] * p.y + c.R[8] * p.z;
	float3 C;
	C.x = -(c.R[0] * c.t[0] + c.R[3] * c.t[1] + c.R[6] * c.t[2]);
	C.y = -(c.R[1] * c.t[0] + c.R[4] * c.t[1] + c.R[7] * c.t[2]);
	C.z = -(c.R[2] * c.t[0] + c.R[5] * c.t[1] + c.R[8] * c.t[2]);
	return make_float3(t.x + C.x, t.y + C.y, t.z + C.z);
}
// ProjectCamera, APD.cpp:802-812
__device__ __forceinline__ void project_camera(const float3 X, const FCam &c, float2 &pt, float &depth) {
	float3 t;
	t.x = c.R[0] * X.x + c.R[1] * X.y + c.R[2] * X.z + c.t[0];
	t.y = c.R[3] * X.x + c.R[4] * X.y + c.R[5] * X.z + c.t[1];
	t.z = c.R[6] * X.x + c.R[7] * X.y + c.R[8] * X.z + c.t[2];
	depth = c.K[6] * t.x + c.K[7] * t.y + c.K[8] * t.z;
	pt.x = (c.K[0] * t.x + c.K[1] * t.y + c.K[2] * t.z) / depth;
	pt.y = (c.K[3] * t.x + c.K[4] * t.y + c.K[5] * t.z) / depth;
}
// GetAngle, APD.cpp:814-823
__device__ __forceinline__ float get_angle(const float *a, const float *b) {
	const float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
	const float angle = acosf(dot);
	return (angle != angle) ? 0.0f : angle;
}

constexpr int kUndecided = 0, kRejected = 1, kAccepted = 2;
constexpr int kMaxSrc = APD_MAX_IMAGES;

struct FuseArgs {
	int W, H, n_src;
	FView ref;
	FCam ref_cam;
	const FView *src;          // [n_src] device
	const FCam *src_cam;       // [n_src] device
	int *cand;                 // [n_src][W*H] source pixel index or -1
	float *term;               // [n_src][W*H] exp(-tmp_index)
	uint8_t *status;           // [W*H]
	float3 *pt_xyz, *pt_col;   // [W*H] accepted point of the pixel
	int *undecided;            // counter
};

// Everything of the inner loop APD.cpp:910-947 that does not depend on marks set while this view is processed.
__global__ void k_fuse_candidates(const FuseArgs a) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
	if (c >= a.W || r >= a.H) return;
	const size_t n = (size_t)a.W * a.H; const size_t p = (size_t)r * a.W + c;
	uint8_t st = kUndecided;
	if (a.ref.block && a.ref.block[p] < 128) st = kRejected;
	else if (a.ref.mask[p] == 1) st = kRejected;
	const float ref_depth = a.ref.depth[p];
	if (ref_depth <= 0.0) st = kRejected;
	bool any = false;
	if (st == kUndecided) {
		const float *ref_normal = a.ref.normal + 3 * p;
		const float3 X = point_on_world(c, r, ref_depth, a.ref_cam);
		for (int j = 0; j < a.n_src; ++j) {
			int cand = -1; float term = 0.0f;
			const FView &sv = a.src[j]; const FCam &sc = a.src_cam[j];
			float2 pt; float proj_depth;
			project_camera(X, sc, pt, proj_depth);
			const int src_r = int(pt.y + 0.5f), src_c = int(pt.x + 0.5f);
			if (src_c >= 0 && src_c < a.W && src_r >= 0 && src_r < a.H) {
				const size_t q = (size_t)src_r * a.W + src_c;
				const float src_depth = sv.depth[q];
				if (sv.mask[q] != 1 && !(src_depth <= 0.0)) {
					const float3 tX = point_on_world(src_c, src_r, src_depth, sc);
					float2 tpt;
					project_camera(tX, a.ref_cam, tpt, proj_depth);
					const double ex = (double)(c - tpt.x), ey = (double)(r - tpt.y);
					const float reproj_error = (float)sqrt(ex * ex + ey * ey);
					const float relative_depth_diff = fabsf(proj_depth - ref_depth) / ref_depth;
					const float angle = get_angle(ref_normal, sv.normal + 3 * q);
					if (reproj_error < 2.0f && relative_depth_diff < 0.01f && angle < 0.174533f) {
						const float tmp_index = reproj_error + 200 * relative_depth_diff + angle * 10;
						cand = (int)q; term = expf(-tmp_index);
						any = true;
					}
				}
			}
			a.cand[(size_t)j * n + p] = cand; a.term[(size_t)j * n + p] = term;
		}
		if (!any) st = kRejected;            // num_consistent == 0
	}
	a.status[p] = st;
	if (st == kUndecided) atomicAdd(a.undecided, 1);
}

// Round, step 1: drop candidates that an accepted earlier pixel has marked; claim the rest with the raster index.
__global__ void k_fuse_claim(const FuseArgs a) {
	const size_t n = (size_t)a.W * a.H;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || a.status[p] != kUndecided) return;
	for (int j = 0; j < a.n_src; ++j) {
		const int q = a.cand[(size_t)j * n + p];
		if (q < 0) continue;
		if (a.src[j].mask[q] == 1) a.cand[(size_t)j * n + p] = -1;
		else atomicMin(&a.src[j].claim[q], (int)p);
	}
}

// Round, step 2: a pixel that holds every claim it made is the earliest undecided contender for all its source
// pixels: its outcome is final (APD.cpp:949-972).
__global__ void k_fuse_decide(const FuseArgs a) {
	const size_t n = (size_t)a.W * a.H;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || a.status[p] != kUndecided) return;
	for (int j = 0; j < a.n_src; ++j) {
		const int q = a.cand[(size_t)j * n + p];
		if (q >= 0 && a.src[j].claim[q] != (int)p) return;        // wait for the earlier pixel
	}
	int num_consistent = 0; float dynamic_consistency = 0.0f;
	for (int j = 0; j < a.n_src; ++j)
		if (a.cand[(size_t)j * n + p] >= 0) { dynamic_consistency += a.term[(size_t)j * n + p]; num_consistent++; }
	const float factor = (a.ref.states[p] == APD_WEAK ? 0.45f : 0.3f);
	uint8_t st = kRejected;
	if (num_consistent >= 1 && (dynamic_consistency > factor * num_consistent)) {
		const int c = (int)(p % a.W), r = (int)(p / a.W);
		const float3 X = point_on_world(c, r, a.ref.depth[p], a.ref_cam);
		const uchar3 rc = a.ref.bgr[p];
		float col[3] = {(float)rc.x, (float)rc.y, (float)rc.z};
		for (int j = 0; j < a.n_src; ++j) {
			const int q = a.cand[(size_t)j * n + p];
			if (q < 0) continue;
			a.src[j].mask[q] = 1;
			const uchar3 sc = a.src[j].bgr[q];
			col[0] += sc.x; col[1] += sc.y; col[2] += sc.z;
		}
		col[0] /= (num_consistent + 1); col[1] /= (num_consistent + 1); col[2] /= (num_consistent + 1);
		a.pt_xyz[p] = X; a.pt_col[p] = make_float3(col[0], col[1], col[2]);
		st = kAccepted;
	}
	a.status[p] = st;
	atomicSub(a.undecided, 1);
}

__global__ void k_fuse_reset_claims(const FuseArgs a) {
	const size_t n = (size_t)a.W * a.H;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || a.status[p] != kUndecided) return;
	for (int j = 0; j < a.n_src; ++j) { const int q = a.cand[(size_t)j * n + p]; if (q >= 0) a.src[j].claim[q] = INT_MAX; }
}

__global__ void k_fuse_flags(const uint8_t *status, int *flags, size_t n) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) flags[p] = status[p] == kAccepted ? 1 : 0;
}
__global__ void k_fuse_compact(const uint8_t *status, const int *offs, const float3 *xyz, const float3 *col, float3 *out_xyz, float3 *out_col, size_t n) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n && status[p] == kAccepted) { out_xyz[offs[p]] = xyz[p]; out_col[offs[p]] = col[p]; }
}
__global__ void k_planes_to_normals(const float4 *planes, float *normals, size_t n) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n) { const float4 v = planes[p]; normals[3 * p] = v.x; normals[3 * p + 1] = v.y; normals[3 * p + 2] = v.z; }
}
__global__ void k_fill_int(int *a, int v, size_t n) { const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (p < n) a[p] = v; }


// ---- Tanks-and-Temples variants: RunFusion_TAT_Intermediate (APD.cpp:979-1147) and RunFusion_TAT_advanced (:1149-1296) ----------
// These loops have a different sequential dependency than the ETH variant. Accepted pixels mark only THEIR OWN view's mask
// (masks[ref], :1139 / :1288), which is read when that view is a SOURCE of a later problem - so inside one view all pixels
// see the same marks. But the per-source measurements live in ONE `std::vector<CostData> diff` per view that is NOT reset
// between pixels (:1059 / :1230): when a pixel's projection falls outside the source image, onto a marked pixel or onto
// an invalid depth, diff[j] keeps the values of the LAST EARLIER pixel (raster order) that did measure source j - including
// its source coordinates, which the Intermediate variant uses for the colour (:1123-1128). Kept exactly: k_tat_measure writes,
// per source j, the pixel's own index where it measured and -1 elsewhere; an inclusive max-scan over the raster order turns
// that into "index of the last measuring pixel <= p"; k_tat_decide reads the measurements from there.
struct TatArgs {
	int W, H, n_src, variant;            // variant 1 = Intermediate, 2 = advanced
	FView ref; FCam ref_cam;
	const FView *src; const FCam *src_cam;
	int *last;                           // [n_src][W*H] own index where source j was measured / after the scan: last measuring pixel
	int *srcq;                           // [n_src][W*H] source pixel of the measurement
	float *dist, *depth, *angle;         // [n_src][W*H]
	uint8_t *status; float3 *pt_xyz, *pt_col;
};
__global__ void k_tat_measure(const TatArgs a) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
	if (c >= a.W || r >= a.H) return;
	const size_t n = (size_t)a.W * a.H; const size_t p = (size_t)r * a.W + c;
	bool processed = !(a.ref.block && a.ref.block[p] < 128);
	const float ref_depth = a.ref.depth[p];
	if (ref_depth <= 0.0) processed = false;
	a.status[p] = processed ? kUndecided : kRejected;
	float3 X = make_float3(0.f, 0.f, 0.f);
	if (processed) X = point_on_world(c, r, ref_depth, a.ref_cam);
	const float *ref_normal = a.ref.normal + 3 * p;
	for (int j = 0; j < a.n_src; ++j) {
		int last = -1;
		if (processed) {
			const FView &sv = a.src[j]; const FCam &sc = a.src_cam[j];
			float2 pt; float proj_depth;
			project_camera(X, sc, pt, proj_depth);
			const int src_r = int(pt.y + 0.5f), src_c = int(pt.x + 0.5f);
			if (src_c >= 0 && src_c < a.W && src_r >= 0 && src_r < a.H) {
				const size_t q = (size_t)src_r * a.W + src_c;
				const float src_depth = sv.depth[q];
				if (sv.mask[q] != 1 && !(src_depth <= 0.0)) {
					const float3 tX = point_on_world(src_c, src_r, src_depth, sc);
					float2 tpt;
					project_camera(tX, a.ref_cam, tpt, proj_depth);
					const double ex = (double)(c - tpt.x), ey = (double)(r - tpt.y);
					a.dist[(size_t)j * n + p] = (float)sqrt(ex * ex + ey * ey);
					a.depth[(size_t)j * n + p] = fabsf(proj_depth - ref_depth) / ref_depth;
					a.angle[(size_t)j * n + p] = get_angle(ref_normal, sv.normal + 3 * q);
					a.srcq[(size_t)j * n + p] = (int)q;
					last = (int)p;
				}
			}
		}
		a.last[(size_t)j * n + p] = last;
	}
}
__global__ void k_tat_decide(const TatArgs a) {
	const size_t n = (size_t)a.W * a.H;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || a.status[p] != kUndecided) return;
	const float dist_base = 0.25f;
	const float depth_base = (a.variant == 1) ? 1.0f / 3500.0f : 1.0f / 3000.0f;
	const float angle_base = 0.06981317007977318f, angle_grad = 0.05235987755982988f;
	uint8_t st = kRejected;
	for (int k = 2; k <= a.n_src; ++k) {
		int count = 0; unsigned use = 0u;
		for (int j = 0; j < a.n_src; ++j) {
			const int q = a.last[(size_t)j * n + p];
			if (q < 0) continue;                                   // never measured so far: CostData() = FLT_MAX
			const size_t o = (size_t)j * n + q;
			bool ok = a.dist[o] < k * dist_base && a.depth[o] < k * depth_base;
			if (a.variant == 1) ok = ok && a.angle[o] < (k * angle_grad + angle_base);
			if (ok) { count++; use |= 1u << j; }
		}
		if (count >= k) {
			const int c = (int)(p % a.W), r = (int)(p / a.W);
			const uchar3 rc = a.ref.bgr[p];
			float col[3] = {(float)rc.x, (float)rc.y, (float)rc.z};
			if (a.variant == 1) {
				for (int j = 0; j < a.n_src; ++j) {
					if (!((use >> j) & 1u)) continue;
					const uchar3 sc = a.src[j].bgr[a.srcq[(size_t)j * n + a.last[(size_t)j * n + p]]];
					col[0] += (float)sc.x; col[1] += (float)sc.y; col[2] += (float)sc.z;
				}
				col[0] /= (count + 1.0f); col[1] /= (count + 1.0f); col[2] /= (count + 1.0f);
			}
			a.pt_xyz[p] = point_on_world(c, r, a.ref.depth[p], a.ref_cam);
			a.pt_col[p] = make_float3(col[0], col[1], col[2]);
			a.ref.mask[p] = 1;
			st = kAccepted;
			break;
		}
	}
	a.status[p] = st;
}
struct MaxOp { __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; } };

struct HostView { uchar3 *bgr = nullptr; float *depth = nullptr, *normal = nullptr; uint8_t *states = nullptr, *block = nullptr, *mask = nullptr; int *claim = nullptr; bool set = false, has_block = false; FCam cam; };
struct HostProblem { int ref; std::vector<int> srcs; };

}  // namespace

struct apd_fusion {
	int device = 0, n_views = 0, W = 0, H = 0;
	cudaStream_t stream = nullptr;
	std::vector<HostView> views;
	std::vector<HostProblem> problems;
	FView *d_src = nullptr; FCam *d_src_cam = nullptr;
	size_t cand_src = 0;       // source views the cand/term buffers were allocated for
	size_t tat_src = 0; int *tat_srcq = nullptr; float *tat_depth = nullptr, *tat_angle = nullptr;     // T&T variants: extra per-(source, pixel) arrays
	void *tat_tmp = nullptr; size_t tat_tmp_bytes = 0;
	int *cand = nullptr; float *term = nullptr; uint8_t *status = nullptr; float3 *pt_xyz = nullptr, *pt_col = nullptr;
	int *undecided = nullptr, *flags = nullptr, *offs = nullptr; void *cub_tmp = nullptr; size_t cub_bytes = 0;
	float4 *tmp_planes = nullptr;
	float3 *out_xyz = nullptr, *out_col = nullptr; size_t out_cap = 0, out_n = 0;
	double gpu_ms = 0.0; int max_rounds = 0;
	std::string err;
};

static thread_local std::string g_fusion_null = "null fusion handle";
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { f->err = std::string(#call) + ": " + cudaGetErrorString(e_); return APD_E_CUDA; } } while (0)
static int ffail(apd_fusion_handle f, int code, const char *msg) { if (f) f->err = msg; return code; }

extern "C" int apd_fusion_create(apd_fusion_handle *out, int device, int n_views, int width, int height) {
	if (!out) return APD_E_ARG;
	*out = nullptr;
	if (n_views < 2 || width < 1 || height < 1 || (size_t)width * height > (size_t)INT_MAX) return APD_E_LIMIT;
	apd_fusion *f = new apd_fusion();
	f->device = device; f->n_views = n_views; f->W = width; f->H = height;
	auto bail = [&](int code) { apd_fusion_destroy(f); return code; };
	if (cudaSetDevice(device) != cudaSuccess) return bail(APD_E_CUDA);
	if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(APD_E_CUDA);
	const size_t n = (size_t)width * height;
	f->views.resize(n_views);
#define FALLOC(ptr, bytes) if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) return bail(APD_E_CUDA)
	for (auto &v : f->views) {
		FALLOC(v.bgr, n * 3); FALLOC(v.depth, n * 4); FALLOC(v.normal, n * 12); FALLOC(v.states, n); FALLOC(v.block, n); FALLOC(v.mask, n); FALLOC(v.claim, n * 4);
	}
	FALLOC(f->d_src, sizeof(FView) * kMaxSrc); FALLOC(f->d_src_cam, sizeof(FCam) * kMaxSrc);
	FALLOC(f->status, n); FALLOC(f->pt_xyz, n * 12); FALLOC(f->pt_col, n * 12);
	FALLOC(f->undecided, 4); FALLOC(f->flags, n * 4); FALLOC(f->offs, n * 4); FALLOC(f->tmp_planes, n * 16);
	cub::DeviceScan::ExclusiveSum(nullptr, f->cub_bytes, f->flags, f->offs, (int)n, f->stream);
	FALLOC(f->cub_tmp, f->cub_bytes ? f->cub_bytes : 16);
#undef FALLOC
	*out = f;
	return APD_OK;
}

extern "C" void apd_fusion_destroy(apd_fusion_handle f) {
	if (!f) return;
	cudaSetDevice(f->device);
	if (f->stream) cudaStreamSynchronize(f->stream);
	for (auto &v : f->views) { void *p[] = {v.bgr, v.depth, v.normal, v.states, v.block, v.mask, v.claim}; for (void *q : p) if (q) cudaFree(q); }
	void *p[] = {f->d_src, f->d_src_cam, f->cand, f->term, f->status, f->pt_xyz, f->pt_col, f->undecided, f->flags, f->offs, f->cub_tmp, f->tmp_planes, f->out_xyz, f->out_col, f->tat_srcq, f->tat_depth, f->tat_angle, f->tat_tmp};
	for (void *q : p) if (q) cudaFree(q);
	if (f->stream) cudaStreamDestroy(f->stream);
	delete f;
}

extern "C" const char *apd_fusion_last_error(apd_fusion_handle f) { return f ? f->err.c_str() : g_fusion_null.c_str(); }

static int set_view(apd_fusion_handle f, int view, const uint8_t *bgr, const apd_camera *cam, const float *depth, const float *normal,
                    const float *planes, const uint8_t *states, const uint8_t *block) {
	if (!f || view < 0 || view >= f->n_views || !bgr || !cam || !depth || !(normal || planes) || !states) return APD_E_ARG;
	CKF(cudaSetDevice(f->device));
	const size_t n = (size_t)f->W * f->H;
	HostView &v = f->views[view];
	CKF(cudaMemcpyAsync(v.bgr, bgr, n * 3, cudaMemcpyDefault, f->stream));
	CKF(cudaMemcpyAsync(v.depth, depth, n * 4, cudaMemcpyDefault, f->stream));
	if (normal) CKF(cudaMemcpyAsync(v.normal, normal, n * 12, cudaMemcpyDefault, f->stream));
	else {
		CKF(cudaMemcpyAsync(f->tmp_planes, planes, n * 16, cudaMemcpyDefault, f->stream));
		k_planes_to_normals<<<(unsigned)((n + 255) / 256), 256, 0, f->stream>>>(f->tmp_planes, v.normal, n);
	}
	CKF(cudaMemcpyAsync(v.states, states, n, cudaMemcpyDefault, f->stream));
	v.has_block = block != nullptr;
	if (block) CKF(cudaMemcpyAsync(v.block, block, n, cudaMemcpyDefault, f->stream));
	CKF(cudaGetLastError());
	CKF(cudaStreamSynchronize(f->stream));
	memcpy(v.cam.K, cam->K, sizeof(v.cam.K)); memcpy(v.cam.R, cam->R, sizeof(v.cam.R)); memcpy(v.cam.t, cam->t, sizeof(v.cam.t));
	v.set = true;
	return APD_OK;
}
extern "C" int apd_fusion_set_view(apd_fusion_handle f, int view, const uint8_t *bgr, const apd_camera *cam, const float *depth,
                                   const float *normal_xyz, const uint8_t *states, const uint8_t *block) {
	return set_view(f, view, bgr, cam, depth, normal_xyz, nullptr, states, block);
}
extern "C" int apd_fusion_set_view_planes(apd_fusion_handle f, int view, const uint8_t *bgr, const apd_camera *cam, const float *depth,
                                          const float *planes_xyzw, const uint8_t *states, const uint8_t *block) {
	return set_view(f, view, bgr, cam, depth, nullptr, planes_xyzw, states, block);
}

extern "C" int apd_fusion_add_problem(apd_fusion_handle f, int ref_view, const int *src_views, int n_src) {
	if (!f || ref_view < 0 || ref_view >= f->n_views || n_src < 0 || (n_src > 0 && !src_views)) return APD_E_ARG;
	if (n_src > kMaxSrc) return ffail(f, APD_E_LIMIT, "too many source views");
	HostProblem p; p.ref = ref_view;
	for (int i = 0; i < n_src; ++i) {
		if (src_views[i] < 0 || src_views[i] >= f->n_views) return ffail(f, APD_E_ARG, "source view out of range");
		if (src_views[i] == ref_view) return ffail(f, APD_E_ARG, "a view cannot be its own source");
		p.srcs.push_back(src_views[i]);
	}
	f->problems.push_back(p);
	return APD_OK;
}

// raster-order compaction of the accepted points of the view just decided (f->status), appended to the cloud
static int append_accepted(apd_fusion_handle f) {
	const size_t n = (size_t)f->W * f->H;
	cudaStream_t st = f->stream;
	const unsigned g1 = (unsigned)((n + 255) / 256);
	k_fuse_flags<<<g1, 256, 0, st>>>(f->status, f->flags, n);
	cub::DeviceScan::ExclusiveSum(f->cub_tmp, f->cub_bytes, f->flags, f->offs, (int)n, st);
	int last_off = 0, last_flag = 0;
	CKF(cudaMemcpyAsync(&last_off, f->offs + (n - 1), 4, cudaMemcpyDeviceToHost, st));
	CKF(cudaMemcpyAsync(&last_flag, f->flags + (n - 1), 4, cudaMemcpyDeviceToHost, st));
	CKF(cudaStreamSynchronize(st));
	const size_t cnt = (size_t)last_off + last_flag;
	if (f->out_n + cnt > f->out_cap) {
		size_t cap = f->out_cap ? f->out_cap : n;
		while (cap < f->out_n + cnt) cap *= 2;
		float3 *nx = nullptr, *nc = nullptr;
		CKF(cudaMalloc((void **)&nx, cap * 12)); CKF(cudaMalloc((void **)&nc, cap * 12));
		if (f->out_n) { CKF(cudaMemcpyAsync(nx, f->out_xyz, f->out_n * 12, cudaMemcpyDeviceToDevice, st)); CKF(cudaMemcpyAsync(nc, f->out_col, f->out_n * 12, cudaMemcpyDeviceToDevice, st)); }
		CKF(cudaStreamSynchronize(st));
		if (f->out_xyz) cudaFree(f->out_xyz);
		if (f->out_col) cudaFree(f->out_col);
		f->out_xyz = nx; f->out_col = nc; f->out_cap = cap;
	}
	if (cnt) k_fuse_compact<<<g1, 256, 0, st>>>(f->status, f->offs, f->pt_xyz, f->pt_col, f->out_xyz + f->out_n, f->out_col + f->out_n, n);
	f->out_n += cnt;
	CKF(cudaGetLastError());
	return APD_OK;
}

static int ensure_pair_buffers(apd_fusion_handle f, size_t max_src) {
	const size_t n = (size_t)f->W * f->H;
	if (max_src > f->cand_src) {       // problems may be added between runs: grow with the largest source count seen
		if (f->cand) { cudaFree(f->cand); f->cand = nullptr; }
		if (f->term) { cudaFree(f->term); f->term = nullptr; }
		f->cand_src = 0;
		CKF(cudaMalloc((void **)&f->cand, max_src * n * 4)); CKF(cudaMalloc((void **)&f->term, max_src * n * 4));
		f->cand_src = max_src;
	}
	return APD_OK;
}

extern "C" int apd_fusion_run(apd_fusion_handle f) {
	if (!f) return APD_E_ARG;
	for (auto &v : f->views) if (!v.set) return ffail(f, APD_E_STATE, "every view needs apd_fusion_set_view first");
	CKF(cudaSetDevice(f->device));
	const size_t n = (size_t)f->W * f->H;
	size_t max_src = 1;
	for (auto &p : f->problems) if (p.srcs.size() > max_src) max_src = p.srcs.size();
	{ int rc = ensure_pair_buffers(f, max_src); if (rc != APD_OK) return rc; }
	cudaStream_t st = f->stream;
	cudaEvent_t e0, e1; CKF(cudaEventCreate(&e0)); CKF(cudaEventCreate(&e1));
	CKF(cudaEventRecord(e0, st));
	const unsigned g1 = (unsigned)((n + 255) / 256);
	for (auto &v : f->views) { CKF(cudaMemsetAsync(v.mask, 0, n, st)); k_fill_int<<<g1, 256, 0, st>>>(v.claim, INT_MAX, n); }
	f->out_n = 0; f->max_rounds = 0;
	for (const HostProblem &pb : f->problems) {
		FuseArgs a; memset(&a, 0, sizeof(a));
		a.W = f->W; a.H = f->H; a.n_src = (int)pb.srcs.size();
		auto fv = [&](const HostView &v) { FView o; o.bgr = v.bgr; o.depth = v.depth; o.normal = v.normal; o.states = v.states; o.block = v.has_block ? v.block : nullptr; o.mask = v.mask; o.claim = v.claim; return o; };
		a.ref = fv(f->views[pb.ref]); a.ref_cam = f->views[pb.ref].cam;
		std::vector<FView> sv; std::vector<FCam> sc;
		for (int s : pb.srcs) { sv.push_back(fv(f->views[s])); sc.push_back(f->views[s].cam); }
		if (a.n_src) {
			CKF(cudaMemcpyAsync(f->d_src, sv.data(), sizeof(FView) * sv.size(), cudaMemcpyHostToDevice, st));
			CKF(cudaMemcpyAsync(f->d_src_cam, sc.data(), sizeof(FCam) * sc.size(), cudaMemcpyHostToDevice, st));
		}
		a.src = f->d_src; a.src_cam = f->d_src_cam; a.cand = f->cand; a.term = f->term; a.status = f->status;
		a.pt_xyz = f->pt_xyz; a.pt_col = f->pt_col; a.undecided = f->undecided;
		CKF(cudaMemsetAsync(f->undecided, 0, 4, st));
		const dim3 b2(32, 8), g2((f->W + 31) / 32, (f->H + 7) / 8);
		k_fuse_candidates<<<g2, b2, 0, st>>>(a);
		int undecided = 0, rounds = 0;
		CKF(cudaMemcpyAsync(&undecided, f->undecided, 4, cudaMemcpyDeviceToHost, st));
		CKF(cudaStreamSynchronize(st));          // also keeps sv / sc alive until the copies are done
		while (undecided > 0) {
			k_fuse_claim<<<g1, 256, 0, st>>>(a);
			k_fuse_decide<<<g1, 256, 0, st>>>(a);
			k_fuse_reset_claims<<<g1, 256, 0, st>>>(a);
			const int before = undecided;
			CKF(cudaMemcpyAsync(&undecided, f->undecided, 4, cudaMemcpyDeviceToHost, st));
			CKF(cudaStreamSynchronize(st));
			++rounds;
			if (undecided >= before) return ffail(f, APD_E_STATE, "fusion rounds made no progress");
		}
		// the pixels k_fuse_decide resolved in the last round still hold claims: clear them for the next view
		for (int s : pb.srcs) k_fill_int<<<g1, 256, 0, st>>>(f->views[s].claim, INT_MAX, n);
		if (rounds > f->max_rounds) f->max_rounds = rounds;
		{ int rc = append_accepted(f); if (rc != APD_OK) return rc; }
	}
	CKF(cudaEventRecord(e1, st));
	CKF(cudaStreamSynchronize(st));
	float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); f->gpu_ms = ms;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return APD_OK;
}

extern "C" int apd_fusion_run_tat(apd_fusion_handle f, int variant) {
	if (!f) return APD_E_ARG;
	if (variant != APD_FUSION_TAT_INTERMEDIATE && variant != APD_FUSION_TAT_ADVANCED) return ffail(f, APD_E_ARG, "variant must be APD_FUSION_TAT_INTERMEDIATE or APD_FUSION_TAT_ADVANCED");
	for (auto &v : f->views) if (!v.set) return ffail(f, APD_E_STATE, "every view needs apd_fusion_set_view first");
	CKF(cudaSetDevice(f->device));
	const size_t n = (size_t)f->W * f->H;
	size_t max_src = 1;
	for (auto &p : f->problems) if (p.srcs.size() > max_src) max_src = p.srcs.size();
	{ int rc = ensure_pair_buffers(f, max_src); if (rc != APD_OK) return rc; }
	if (max_src > f->tat_src) {
		for (void *q : {(void *)f->tat_srcq, (void *)f->tat_depth, (void *)f->tat_angle}) if (q) cudaFree(q);
		f->tat_srcq = nullptr; f->tat_depth = f->tat_angle = nullptr; f->tat_src = 0;
		CKF(cudaMalloc((void **)&f->tat_srcq, max_src * n * 4)); CKF(cudaMalloc((void **)&f->tat_depth, max_src * n * 4)); CKF(cudaMalloc((void **)&f->tat_angle, max_src * n * 4));
		f->tat_src = max_src;
	}
	cudaStream_t st = f->stream;
	if (!f->tat_tmp) {
		cub::DeviceScan::InclusiveScan(nullptr, f->tat_tmp_bytes, f->cand, f->cand, MaxOp(), (int)n, st);
		CKF(cudaMalloc(&f->tat_tmp, f->tat_tmp_bytes ? f->tat_tmp_bytes : 16));
	}
	cudaEvent_t e0, e1; CKF(cudaEventCreate(&e0)); CKF(cudaEventCreate(&e1));
	CKF(cudaEventRecord(e0, st));
	const unsigned g1 = (unsigned)((n + 255) / 256);
	for (auto &v : f->views) CKF(cudaMemsetAsync(v.mask, 0, n, st));
	f->out_n = 0; f->max_rounds = 1;
	for (const HostProblem &pb : f->problems) {
		TatArgs a; memset(&a, 0, sizeof(a));
		a.W = f->W; a.H = f->H; a.n_src = (int)pb.srcs.size(); a.variant = variant;
		auto fv = [&](const HostView &v) { FView o; o.bgr = v.bgr; o.depth = v.depth; o.normal = v.normal; o.states = v.states; o.block = v.has_block ? v.block : nullptr; o.mask = v.mask; o.claim = v.claim; return o; };
		a.ref = fv(f->views[pb.ref]); a.ref_cam = f->views[pb.ref].cam;
		std::vector<FView> sv; std::vector<FCam> sc;
		for (int s : pb.srcs) { sv.push_back(fv(f->views[s])); sc.push_back(f->views[s].cam); }
		if (a.n_src) {
			CKF(cudaMemcpyAsync(f->d_src, sv.data(), sizeof(FView) * sv.size(), cudaMemcpyHostToDevice, st));
			CKF(cudaMemcpyAsync(f->d_src_cam, sc.data(), sizeof(FCam) * sc.size(), cudaMemcpyHostToDevice, st));
			CKF(cudaStreamSynchronize(st));          // sv / sc are locals
		}
		a.src = f->d_src; a.src_cam = f->d_src_cam; a.last = f->cand; a.srcq = f->tat_srcq; a.dist = f->term; a.depth = f->tat_depth; a.angle = f->tat_angle;
		a.status = f->status; a.pt_xyz = f->pt_xyz; a.pt_col = f->pt_col;
		const dim3 b2(32, 8), g2((f->W + 31) / 32, (f->H + 7) / 8);
		k_tat_measure<<<g2, b2, 0, st>>>(a);
		for (int j = 0; j < a.n_src; ++j)          // "last earlier pixel that measured source j", in raster order
			cub::DeviceScan::InclusiveScan(f->tat_tmp, f->tat_tmp_bytes, a.last + (size_t)j * n, a.last + (size_t)j * n, MaxOp(), (int)n, st);
		k_tat_decide<<<g1, 256, 0, st>>>(a);
		CKF(cudaGetLastError());
		{ int rc = append_accepted(f); if (rc != APD_OK) return rc; }
	}
	CKF(cudaEventRecord(e1, st));
	CKF(cudaStreamSynchronize(st));
	float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); f->gpu_ms = ms;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return APD_OK;
}

extern "C" long long apd_fusion_num_points(apd_fusion_handle f) { return f ? (long long)f->out_n : 0; }

extern "C" int apd_fusion_get_points(apd_fusion_handle f, float *xyz, float *color) {
	if (!f) return APD_E_ARG;
	CKF(cudaSetDevice(f->device));
	if (f->out_n) {
		if (xyz) CKF(cudaMemcpyAsync(xyz, f->out_xyz, f->out_n * 12, cudaMemcpyDeviceToHost, f->stream));
		if (color) CKF(cudaMemcpyAsync(color, f->out_col, f->out_n * 12, cudaMemcpyDeviceToHost, f->stream));
		CKF(cudaStreamSynchronize(f->stream));
	}
	return APD_OK;
}

extern "C" int apd_fusion_write_ply(apd_fusion_handle f, const char *path) {
	if (!f || !path) return APD_E_ARG;
	std::vector<float> xyz(f->out_n * 3), col(f->out_n * 3);
	int rc = apd_fusion_get_points(f, xyz.data(), col.data());
	if (rc != APD_OK) return rc;
	FILE *out = fopen(path, "wb");
	if (!out) return ffail(f, APD_E_STATE, "cannot open the PLY file");
	fprintf(out, "ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
	             "property uchar diffuse_blue\nproperty uchar diffuse_green\nproperty uchar diffuse_red\nend_header\n", (int)f->out_n);
	for (size_t i = 0; i < f->out_n; ++i) {
		fwrite(&xyz[3 * i], 4, 3, out);
		const unsigned char px[3] = {(unsigned char)col[3 * i], (unsigned char)col[3 * i + 1], (unsigned char)col[3 * i + 2]};
		fwrite(px, 1, 3, out);
	}
	fclose(out);
	return APD_OK;
}

extern "C" int apd_fusion_get_timing(apd_fusion_handle f, double *gpu_ms, int *max_rounds) {
	if (!f) return APD_E_ARG;
	if (gpu_ms) *gpu_ms = f->gpu_ms;
	if (max_rounds) *max_rounds = f->max_rounds;
	return APD_OK;
}
