// Pass scheduler with a persistent per-scene device cache: the C-ABI of include/apd_scene.h.
// Replaces the host loop main.cpp:168-217 + ProcessProblem (main.cpp:91-138) + the per-pass input preparation of
// APD::InuputInitialization (APD.cpp:399-583), with every intermediate kept in HBM instead of JPEG/.dmb files.
// The PatchMatch itself is the engine of apd_engine.cu, one handle per round (image pyramid level).
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "apd_engine_internal.h"
#include "../../include/apd_scene.h"

namespace {

// ---- device kernels ---------------------------------------------------------------------------------

// cv::resize(INTER_LINEAR) on CV_32FC1 as OpenCV's generic C++ path computes it (imgproc/resize.cpp: the x/y
// offset+weight tables are made on the host exactly as there; horizontal pass S[sx]*a0 + S[sx+1]*a1 into two row
// buffers, then the vertical pass R0*b0 + R1*b1, every product and sum rounded separately).
__global__ void k_resize_linear(const float *__restrict__ src, int sw, int sh, float *__restrict__ dst, int dw, int dh,
                                const int *__restrict__ xofs, const float *__restrict__ xw, const int *__restrict__ yofs,
                                const float *__restrict__ yw) {
	const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
	if (dx >= dw || dy >= dh) return;
	const int sx = xofs[dx], sx1 = min(sx + 1, sw - 1);
	const float a1 = xw[dx], a0 = __fsub_rn(1.0f, a1);
	const int sy = yofs[dy];
	const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
	const float b1 = yw[dy], b0 = __fsub_rn(1.0f, b1);
	const float *r0 = src + (size_t)y0 * sw, *r1 = src + (size_t)y1 * sw;
	const float h0 = __fadd_rn(__fmul_rn(r0[sx], a0), __fmul_rn(r0[sx1], a1));
	const float h1 = __fadd_rn(__fmul_rn(r1[sx], a0), __fmul_rn(r1[sx1], a1));
	dst[(size_t)dy * dw + dx] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
}

// cv::resize switches INTER_LINEAR to the 2x2 box average when both scales are exactly 2 (resize.cpp, "is_area_fast")
__global__ void k_resize_half(const float *__restrict__ src, int sw, float *__restrict__ dst, int dw, int dh) {
	const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
	if (dx >= dw || dy >= dh) return;
	const float *r0 = src + (size_t)(2 * dy) * sw + 2 * dx, *r1 = r0 + sw;
	dst[(size_t)dy * dw + dx] = __fmul_rn(__fadd_rn(__fadd_rn(r0[0], r0[1]), __fadd_rn(r1[0], r1[1])), 0.25f);
}

// RescaleMatToTargetSize<T>, APD.cpp:752-774: nearest sampling; the reference divides the ROW by the x scale and the
// COLUMN by the y scale, which is kept. Elements it leaves unwritten (source index out of range) become 0 here.
template <typename T>
__global__ void k_rescale_nearest(const T *__restrict__ src, int sw, int sh, T *__restrict__ dst, int dw, int dh, float scale_x, float scale_y) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
	if (c >= dw || r >= dh) return;
	const int o_r = (int)__fdiv_rn((float)r, scale_x);
	const int o_c = (int)__fdiv_rn((float)c, scale_y);
	T v; memset(&v, 0, sizeof(T));
	if (o_r >= 0 && o_c >= 0 && o_r < sh && o_c < sw) v = src[(size_t)o_r * sw + o_c];
	dst[(size_t)r * dw + c] = v;
}

// Result hand-over of ProcessProblem, main.cpp:101-124: depth = plane.w, out-of-range depth -> 0 and UNKNOWN.
__global__ void k_collect(const float4 *__restrict__ planes, const uint8_t *__restrict__ states, const uint32_t *__restrict__ views, size_t n,
                          float dmin, float dmax, float4 *__restrict__ out_planes, float *__restrict__ out_depth,
                          uint8_t *__restrict__ out_states, uint32_t *__restrict__ out_views) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 p = planes[i];
	uint8_t st = states[i];
	if (p.w < dmin || p.w > dmax) { p.w = 0.0f; st = APD_UNKNOWN; }
	out_planes[i] = p; out_depth[i] = p.w; out_states[i] = st; out_views[i] = views[i];
}

struct ViewResult {
	float4 *planes = nullptr;   // (normal xyz, depth w) as written to normals.dmb / depths.dmb
	float *depth = nullptr;
	uint8_t *states = nullptr;
	uint32_t *views = nullptr;
	int W = 0, H = 0;
};
struct Problem { int ref; std::vector<int> srcs; };

}  // namespace

struct apd_scene {
	int device = 0, n_views = 0, W = 0, H = 0;
	int round_limit = 1000;      // main.cpp:81: halve until the larger image side is <= 1000
	uint64_t seed = 0;
	cudaStream_t stream = nullptr;
	std::vector<float *> full, scaled;        // device images, full size and current round's size
	std::vector<char> have_view;
	std::vector<apd_camera> cams;
	std::vector<Problem> problems;
	std::vector<ViewResult> res;
	int cur_round = -1, rw = 0, rh = 0;
	apd_handle eng = nullptr;
	// staging for re-sampled priors / depth maps at the round's size
	float4 *tmp_planes = nullptr; uint32_t *tmp_views = nullptr; uint8_t *tmp_states = nullptr;
	std::vector<float *> tmp_depth;           // APD_MAX_IMAGES maps
	int *d_xofs = nullptr, *d_yofs = nullptr; float *d_xw = nullptr, *d_yw = nullptr;
	double pm_ms = 0.0, wall_ms = 0.0; long long launches = 0;
	std::string err;
};

static thread_local std::string g_scene_null = "null scene handle";
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { s->err = std::string(#call) + ": " + cudaGetErrorString(e_); return APD_E_CUDA; } } while (0)
static int sfail(apd_scene_handle s, int code, const std::string &msg) { if (s) s->err = msg; return code; }

static int round_num_for(int W, int H, int limit) {   // ComputeRoundNum, main.cpp:72-88 (limit = 1000 there)
	int max_size = W > H ? W : H, rounds = 1;
	while (max_size > limit) { max_size /= 2; rounds++; }
	return rounds;
}
static int scale_size_for(int rounds, int round) { return 1 << (rounds - 1 - round); }   // main.cpp:188
static void scaled_size(int W, int H, int scale_size, int *w, int *h) {                     // APD.cpp:465-468
	if (scale_size == 1) { *w = W; *h = H; return; }
	const float factor = 1.0f / (float)scale_size;
	*w = (int)std::round(W * factor); *h = (int)std::round(H * factor);
}

extern "C" int apd_scene_create(apd_scene_handle *out, int device, int n_views, int width, int height, uint64_t seed) {
	if (!out) return APD_E_ARG;
	*out = nullptr;
	if (n_views < 2 || width < 16 || height < 16 || width > 32767 || height > 32767) return APD_E_LIMIT;
	apd_scene *s = new apd_scene();
	s->device = device; s->n_views = n_views; s->W = width; s->H = height; s->seed = seed;
	auto bail = [&](int code) { apd_scene_destroy(s); return code; };
	if (cudaSetDevice(device) != cudaSuccess) return bail(APD_E_CUDA);
	if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(APD_E_CUDA);
	const size_t n = (size_t)width * height;
	s->full.assign(n_views, nullptr); s->scaled.assign(n_views, nullptr); s->have_view.assign(n_views, 0);
	s->cams.resize(n_views); s->res.resize(n_views); s->tmp_depth.assign(APD_MAX_IMAGES, nullptr);
#define SALLOC(ptr, bytes) if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) return bail(APD_E_CUDA)
	for (int v = 0; v < n_views; ++v) {
		SALLOC(s->full[v], n * 4);
		SALLOC(s->res[v].planes, n * 16); SALLOC(s->res[v].depth, n * 4); SALLOC(s->res[v].states, n); SALLOC(s->res[v].views, n * 4);
	}
	SALLOC(s->tmp_planes, n * 16); SALLOC(s->tmp_views, n * 4); SALLOC(s->tmp_states, n);
	SALLOC(s->d_xofs, (size_t)width * 4); SALLOC(s->d_xw, (size_t)width * 4); SALLOC(s->d_yofs, (size_t)height * 4); SALLOC(s->d_yw, (size_t)height * 4);
#undef SALLOC
	*out = s;
	return APD_OK;
}

extern "C" void apd_scene_destroy(apd_scene_handle s) {
	if (!s) return;
	cudaSetDevice(s->device);
	if (s->stream) cudaStreamSynchronize(s->stream);
	if (s->eng) apd_destroy(s->eng);
	for (int v = 0; v < (int)s->full.size(); ++v) {
		if (s->scaled[v] && s->scaled[v] != s->full[v]) cudaFree(s->scaled[v]);
		if (s->full[v]) cudaFree(s->full[v]);
	}
	for (auto &r : s->res) { if (r.planes) cudaFree(r.planes); if (r.depth) cudaFree(r.depth); if (r.states) cudaFree(r.states); if (r.views) cudaFree(r.views); }
	for (float *p : s->tmp_depth) if (p) cudaFree(p);
	void *ptrs[] = {s->tmp_planes, s->tmp_views, s->tmp_states, s->d_xofs, s->d_xw, s->d_yofs, s->d_yw};
	for (void *p : ptrs) if (p) cudaFree(p);
	if (s->stream) cudaStreamDestroy(s->stream);
	delete s;
}

extern "C" const char *apd_scene_last_error(apd_scene_handle s) { return s ? s->err.c_str() : g_scene_null.c_str(); }

extern "C" int apd_scene_set_view(apd_scene_handle s, int view, const float *image, size_t pitch_bytes, const apd_camera *cam) {
	if (!s || !image || !cam || view < 0 || view >= s->n_views || pitch_bytes < (size_t)s->W * 4) return APD_E_ARG;
	CKS(cudaSetDevice(s->device));
	CKS(cudaMemcpy2DAsync(s->full[view], (size_t)s->W * 4, image, pitch_bytes, (size_t)s->W * 4, s->H, cudaMemcpyDefault, s->stream));
	CKS(cudaStreamSynchronize(s->stream));
	s->cams[view] = *cam;
	s->cams[view].width = s->W; s->cams[view].height = s->H;       // APD.cpp:439-440
	s->have_view[view] = 1;
	s->cur_round = -1;                                             // cached pyramid level is stale
	return APD_OK;
}

extern "C" int apd_scene_add_problem(apd_scene_handle s, int ref_view, const int *src_views, int n_src) {
	if (!s || ref_view < 0 || ref_view >= s->n_views || n_src < 1 || !src_views) return APD_E_ARG;
	if (n_src + 1 > APD_MAX_IMAGES) return sfail(s, APD_E_LIMIT, "too many images (APD.cpp:428-431)");
	Problem p; p.ref = ref_view;
	for (int i = 0; i < n_src; ++i) {
		if (src_views[i] < 0 || src_views[i] >= s->n_views) return sfail(s, APD_E_ARG, "source view out of range");
		p.srcs.push_back(src_views[i]);
	}
	s->problems.push_back(p);
	return APD_OK;
}

extern "C" int apd_scene_set_round_limit(apd_scene_handle s, int max_size) {
	if (!s) return APD_E_ARG;
	if (max_size < 16) return sfail(s, APD_E_ARG, "round limit must be >= 16");
	if (s->eng) return sfail(s, APD_E_STATE, "the schedule is fixed once a pass has run");
	s->round_limit = max_size;
	return APD_OK;
}

extern "C" int apd_scene_num_rounds(apd_scene_handle s) { return s ? round_num_for(s->W, s->H, s->round_limit) : 0; }

extern "C" int apd_scene_round_size(apd_scene_handle s, int round, int *width, int *height) {
	if (!s || !width || !height) return APD_E_ARG;
	const int rounds = round_num_for(s->W, s->H, s->round_limit);
	if (round < 0 || round >= rounds) return APD_E_ARG;
	scaled_size(s->W, s->H, scale_size_for(rounds, round), width, height);
	return APD_OK;
}

extern "C" int apd_scene_pass_params(apd_scene_handle s, int round, int pass, apd_params *p) {
	if (!s || !p || pass < 0 || pass > 3) return APD_E_ARG;
	const int rounds = round_num_for(s->W, s->H, s->round_limit);
	if (round < 0 || round >= rounds) return APD_E_ARG;
	apd_default_params(p);
	const int i = round;
	p->max_iterations = 3;
	if (i == 0) p->use_APD = 0;
	else {
		p->use_APD = 1;
		p->ransac_threshold = (float)(0.01 - i * 0.00125);                               // main.cpp:181,203
		int rt = (int)std::pow(2, i); p->rotate_time = rt < 4 ? rt : 4;                  // main.cpp:182,204
	}
	if (pass == 0) {                                                                         // main.cpp:171-186
		p->state = (i == 0) ? APD_FIRST_INIT : APD_REFINE_INIT;
		p->geom_consistency = 0;
		p->weak_peak_radius = 6;
	} else {                                                                                 // main.cpp:195-208
		const int j = pass - 1;
		p->state = APD_REFINE_ITER;
		p->geom_consistency = 1;
		const int r = 4 - 2 * j; p->weak_peak_radius = r > 2 ? r : 2;
	}
	return APD_OK;
}

// Build the round's pyramid level and engine (called when the round changes).
static int enter_round(apd_scene_handle s, int round) {
	const int rounds = round_num_for(s->W, s->H, s->round_limit);
	const int scale = scale_size_for(rounds, round);
	int rw, rh; scaled_size(s->W, s->H, scale, &rw, &rh);
	for (int v = 0; v < s->n_views; ++v) if (!s->have_view[v]) return sfail(s, APD_E_STATE, "every view needs apd_scene_set_view first");
	if (s->eng) { apd_destroy(s->eng); s->eng = nullptr; }
	for (int v = 0; v < s->n_views; ++v) { if (s->scaled[v] && s->scaled[v] != s->full[v]) cudaFree(s->scaled[v]); s->scaled[v] = nullptr; }
	for (float *&p : s->tmp_depth) { if (p) cudaFree(p); p = nullptr; }
	if (scale == 1) {
		for (int v = 0; v < s->n_views; ++v) s->scaled[v] = s->full[v];
	} else {
		const dim3 blk(32, 8), grd((rw + 31) / 32, (rh + 7) / 8);
		const bool half = (s->W == 2 * rw) && (s->H == 2 * rh);           // both scales exactly 2 -> box average
		if (!half) {
			// offset / weight tables, imgproc/resize.cpp: fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx;
			// sx < 0 -> (0, 0);  sx >= ssize - 1 -> (ssize - 1, 0). The vertical table keeps its fraction at the border.
			auto table = [](int dn, int sn, bool horizontal, std::vector<int> &ofs, std::vector<float> &w) {
				const double inv_scale = (double)dn / sn, sc = 1.0 / inv_scale;
				ofs.resize(dn); w.resize(dn);
				for (int d = 0; d < dn; ++d) {
					float f = (float)((d + 0.5) * sc - 0.5);
					int o = (int)std::floor(f);
					f -= o;
					if (horizontal) {
						if (o < 0) { f = 0.f; o = 0; }
						if (o >= sn - 1) { f = 0.f; o = sn - 1; }
					}
					ofs[d] = o; w[d] = f;
				}
			};
			std::vector<int> xo, yo; std::vector<float> xw, yw;
			table(rw, s->W, true, xo, xw); table(rh, s->H, false, yo, yw);
			CKS(cudaMemcpyAsync(s->d_xofs, xo.data(), rw * 4, cudaMemcpyHostToDevice, s->stream));
			CKS(cudaMemcpyAsync(s->d_xw, xw.data(), rw * 4, cudaMemcpyHostToDevice, s->stream));
			CKS(cudaMemcpyAsync(s->d_yofs, yo.data(), rh * 4, cudaMemcpyHostToDevice, s->stream));
			CKS(cudaMemcpyAsync(s->d_yw, yw.data(), rh * 4, cudaMemcpyHostToDevice, s->stream));
			CKS(cudaStreamSynchronize(s->stream));      // the host vectors go out of scope below
		}
		for (int v = 0; v < s->n_views; ++v) {
			CKS(cudaMalloc((void **)&s->scaled[v], (size_t)rw * rh * 4));
			if (half) k_resize_half<<<grd, blk, 0, s->stream>>>(s->full[v], s->W, s->scaled[v], rw, rh);
			else k_resize_linear<<<grd, blk, 0, s->stream>>>(s->full[v], s->W, s->H, s->scaled[v], rw, rh, s->d_xofs, s->d_xw, s->d_yofs, s->d_yw);
			s->launches++;
		}
		CKS(cudaGetLastError());
		CKS(cudaStreamSynchronize(s->stream));
	}
	size_t max_images = 2;
	for (const auto &p : s->problems) if (p.srcs.size() + 1 > max_images) max_images = p.srcs.size() + 1;
	apd_params prm; apd_default_params(&prm);
	int rc = apd_create(&s->eng, s->device, rw, rh, (int)max_images, &prm, s->seed);
	if (rc != APD_OK) return sfail(s, rc, "apd_create failed for the round's engine");
	s->cur_round = round; s->rw = rw; s->rh = rh;
	return APD_OK;
}

template <typename T>
static void rescale(apd_scene_handle s, const T *src, int sw, int sh, T *dst, int dw, int dh) {
	const float scale_x = dw / static_cast<float>(sw), scale_y = dh / static_cast<float>(sh);       // APD.cpp:757-758
	const dim3 blk(32, 8), grd((dw + 31) / 32, (dh + 7) / 8);
	k_rescale_nearest<T><<<grd, blk, 0, s->stream>>>(src, sw, sh, dst, dw, dh, scale_x, scale_y);
	s->launches++;
}

extern "C" int apd_scene_run_problem(apd_scene_handle s, int round, int pass, int problem) {
	if (!s || problem < 0 || problem >= (int)s->problems.size()) return APD_E_ARG;
	apd_params prm;
	int rc = apd_scene_pass_params(s, round, pass, &prm);
	if (rc != APD_OK) return sfail(s, rc, "bad round / pass");
	CKS(cudaSetDevice(s->device));
	if (round != s->cur_round) { rc = enter_round(s, round); if (rc != APD_OK) return rc; }
	const Problem &pb = s->problems[problem];
	const int n = 1 + (int)pb.srcs.size();
	const int rw = s->rw, rh = s->rh;
	const size_t npx = (size_t)rw * rh;
	apd_handle e = s->eng;
#define CKE(call) do { int rc_ = (call); if (rc_ != APD_OK) return sfail(s, rc_, std::string(#call) + ": " + apd_last_error(e)); } while (0)
	CKE(apd_set_num_images(e, n));
	// ---- cameras: intrinsics follow the image scale (APD.cpp:470-488); depth range from the reference camera (:454-455)
	std::vector<int> ids(n); ids[0] = pb.ref; for (int i = 1; i < n; ++i) ids[i] = pb.srcs[i - 1];
	std::vector<apd_camera> cams(n);
	const float scale_x = rw / static_cast<float>(s->W), scale_y = rh / static_cast<float>(s->H);
	for (int i = 0; i < n; ++i) {
		cams[i] = s->cams[ids[i]];
		if (rw != s->W || rh != s->H) {
			cams[i].K[0] *= scale_x; cams[i].K[2] *= scale_x; cams[i].K[4] *= scale_y; cams[i].K[5] *= scale_y;
		}
		cams[i].width = rw; cams[i].height = rh;
	}
	prm.depth_min = cams[0].depth_min * 0.6f; prm.depth_max = cams[0].depth_max * 1.2f;
	CKE(apd_set_cameras(e, cams.data()));
	// ---- images of this problem: device-to-device from the round's cache
	std::vector<const float *> ptrs(n);
	for (int i = 0; i < n; ++i) ptrs[i] = s->scaled[ids[i]];
	CKE(apd_set_images(e, ptrs.data(), (size_t)rw * 4));
	// ---- depth maps for the geometric consistency term (APD.cpp:492-510): latest result of each view
	if (prm.geom_consistency) {
		for (int i = 0; i < n; ++i) {
			const ViewResult &r = s->res[ids[i]];
			if (r.W == 0) return sfail(s, APD_E_STATE, "geometric pass before every view has a depth map (pass order)");
			if (r.W == rw && r.H == rh) ptrs[i] = r.depth;
			else {
				if (!s->tmp_depth[i]) CKS(cudaMalloc((void **)&s->tmp_depth[i], npx * 4));
				rescale<float>(s, r.depth, r.W, r.H, s->tmp_depth[i], rw, rh);
				ptrs[i] = s->tmp_depth[i];
			}
		}
		CKS(cudaGetLastError());
		CKS(cudaStreamSynchronize(s->stream));
		CKE(apd_set_depths(e, ptrs.data(), (size_t)rw * 4));
	}
	// ---- priors from the view's previous result (APD.cpp:513-581)
	{
		const ViewResult &r = s->res[pb.ref];
		const bool need_planes = prm.state != APD_FIRST_INIT, need_states = prm.use_APD != 0;
		if ((need_planes || need_states) && r.W == 0) return sfail(s, APD_E_STATE, "refinement pass before the view has a result (pass order)");
		const float4 *pl = nullptr; const uint32_t *vw = nullptr; const uint8_t *st = nullptr;
		const bool same = r.W == rw && r.H == rh;
		if (need_planes) {
			if (same) { pl = r.planes; vw = r.views; }
			else {
				rescale<float4>(s, r.planes, r.W, r.H, s->tmp_planes, rw, rh);
				rescale<uint32_t>(s, r.views, r.W, r.H, s->tmp_views, rw, rh);
				pl = s->tmp_planes; vw = s->tmp_views;
			}
		}
		if (need_states) {
			if (same) st = r.states;
			else { rescale<uint8_t>(s, r.states, r.W, r.H, s->tmp_states, rw, rh); st = s->tmp_states; }
		}
		CKS(cudaGetLastError());
		CKS(cudaStreamSynchronize(s->stream));
		if (pl || st) CKE(apd_set_priors(e, (const float *)pl, vw, st));
	}
	CKE(apd_set_params(e, &prm));
	CKE(apd_set_seed(e, s->seed + (uint64_t)(round * 4 + pass) * 65536ull + (uint64_t)problem));
	CKE(apd_run(e));
#undef CKE
	{
		float ms[10 + 5 * 64];
		const int ns = apd_get_stage_ms(e, ms, (int)(sizeof(ms) / sizeof(ms[0])));
		for (int i = 0; i < ns; ++i) s->pm_ms += ms[i];
		s->launches += apd_get_launch_count(e);
	}
	// ---- results stay on the device (main.cpp:101-124)
	ViewResult &out = s->res[pb.ref];
	k_collect<<<(unsigned)((npx + 255) / 256), 256, 0, s->stream>>>(e->planes, e->states, e->sel_views, npx, prm.depth_min, prm.depth_max,
	                                                                 out.planes, out.depth, out.states, out.views);
	s->launches++;
	CKS(cudaGetLastError());
	CKS(cudaStreamSynchronize(s->stream));
	out.W = rw; out.H = rh;
	return APD_OK;
}

static void timing_begin(apd_scene_handle s) { s->pm_ms = 0.0; s->wall_ms = 0.0; s->launches = 0; }

static int run_pass(apd_scene_handle s, int round, int pass) {
	for (int k = 0; k < (int)s->problems.size(); ++k) {
		int rc = apd_scene_run_problem(s, round, pass, k);
		if (rc != APD_OK) return rc;
	}
	return APD_OK;
}

extern "C" int apd_scene_run_pass(apd_scene_handle s, int round, int pass) {
	if (!s) return APD_E_ARG;
	if (s->problems.empty()) return sfail(s, APD_E_STATE, "no problems");
	timing_begin(s);
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = run_pass(s, round, pass);
	s->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return rc;
}

extern "C" int apd_scene_run(apd_scene_handle s) {
	if (!s) return APD_E_ARG;
	if (s->problems.empty()) return sfail(s, APD_E_STATE, "no problems");
	timing_begin(s);
	const auto t0 = std::chrono::steady_clock::now();
	const int rounds = round_num_for(s->W, s->H, s->round_limit);
	int rc = APD_OK;
	for (int i = 0; i < rounds && rc == APD_OK; ++i)
		for (int pass = 0; pass < 4 && rc == APD_OK; ++pass) rc = run_pass(s, i, pass);
	s->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return rc;
}

extern "C" int apd_scene_result_size(apd_scene_handle s, int view, int *width, int *height) {
	if (!s || view < 0 || view >= s->n_views || !width || !height) return APD_E_ARG;
	*width = s->res[view].W; *height = s->res[view].H;
	return APD_OK;
}

static int get_result(apd_scene_handle s, int view, void *out, const void *src, size_t bytes_per_px) {
	if (!s || !out || view < 0 || view >= s->n_views) return APD_E_ARG;
	const ViewResult &r = s->res[view];
	if (r.W == 0) return sfail(s, APD_E_STATE, "view has no result yet");
	CKS(cudaSetDevice(s->device));
	CKS(cudaMemcpyAsync(out, src, (size_t)r.W * r.H * bytes_per_px, cudaMemcpyDeviceToHost, s->stream));
	CKS(cudaStreamSynchronize(s->stream));
	return APD_OK;
}
extern "C" int apd_scene_get_depth(apd_scene_handle s, int view, float *depth) {
	if (!s || view < 0 || view >= s->n_views) return APD_E_ARG;
	return get_result(s, view, depth, s->res[view].depth, 4);
}
extern "C" int apd_scene_get_states(apd_scene_handle s, int view, uint8_t *states) {
	if (!s || view < 0 || view >= s->n_views) return APD_E_ARG;
	return get_result(s, view, states, s->res[view].states, 1);
}
extern "C" int apd_scene_get_views(apd_scene_handle s, int view, uint32_t *views) {
	if (!s || view < 0 || view >= s->n_views) return APD_E_ARG;
	return get_result(s, view, views, s->res[view].views, 4);
}
extern "C" int apd_scene_get_normal(apd_scene_handle s, int view, float *normal_xyz) {
	if (!s || !normal_xyz || view < 0 || view >= s->n_views) return APD_E_ARG;
	const ViewResult &r = s->res[view];
	if (r.W == 0) return sfail(s, APD_E_STATE, "view has no result yet");
	const size_t n = (size_t)r.W * r.H;
	std::vector<float> tmp(n * 4);
	int rc = get_result(s, view, tmp.data(), r.planes, 16);
	if (rc != APD_OK) return rc;
	for (size_t i = 0; i < n; ++i) { normal_xyz[3 * i] = tmp[4 * i]; normal_xyz[3 * i + 1] = tmp[4 * i + 1]; normal_xyz[3 * i + 2] = tmp[4 * i + 2]; }
	return APD_OK;
}

__global__ void k_pack_planes(const float *__restrict__ normal, const float *__restrict__ depth, float4 *__restrict__ planes, size_t n) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) planes[i] = make_float4(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2], depth[i]);
}

extern "C" int apd_scene_set_result(apd_scene_handle s, int view, int width, int height, const float *depth, const float *normal_xyz,
                                    const uint8_t *states, const uint32_t *selected_views) {
	if (!s || view < 0 || view >= s->n_views || !depth || !normal_xyz || !states || !selected_views) return APD_E_ARG;
	if (width < 1 || height < 1 || (size_t)width * height > (size_t)s->W * s->H) return sfail(s, APD_E_ARG, "result larger than the full-resolution buffers");
	CKS(cudaSetDevice(s->device));
	const size_t n = (size_t)width * height;
	ViewResult &r = s->res[view];
	// the normals travel through tmp_planes (3 floats per pixel fit into its 4)
	CKS(cudaMemcpyAsync(r.depth, depth, n * 4, cudaMemcpyDefault, s->stream));
	CKS(cudaMemcpyAsync(s->tmp_planes, normal_xyz, n * 12, cudaMemcpyDefault, s->stream));
	k_pack_planes<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(reinterpret_cast<const float *>(s->tmp_planes), r.depth, r.planes, n);
	CKS(cudaMemcpyAsync(r.states, states, n, cudaMemcpyDefault, s->stream));
	CKS(cudaMemcpyAsync(r.views, selected_views, n * 4, cudaMemcpyDefault, s->stream));
	CKS(cudaGetLastError());
	CKS(cudaStreamSynchronize(s->stream));
	r.W = width; r.H = height;
	return APD_OK;
}

extern "C" int apd_scene_result_device(apd_scene_handle s, int view, void **planes, void **depth, void **states, void **views) {
	if (!s || view < 0 || view >= s->n_views) return APD_E_ARG;
	const ViewResult &r = s->res[view];
	if (planes) *planes = r.planes;
	if (depth) *depth = r.depth;
	if (states) *states = r.states;
	if (views) *views = r.views;
	return APD_OK;
}
extern "C" int apd_scene_mark_result(apd_scene_handle s, int view, int width, int height) {
	if (!s || view < 0 || view >= s->n_views) return APD_E_ARG;
	if (width < 1 || height < 1 || (size_t)width * height > (size_t)s->W * s->H) return sfail(s, APD_E_ARG, "result larger than the full-resolution buffers");
	s->res[view].W = width; s->res[view].H = height;
	return APD_OK;
}

extern "C" int apd_scene_get_scaled_image(apd_scene_handle s, int round, int view, float *image) {
	if (!s || !image || view < 0 || view >= s->n_views) return APD_E_ARG;
	CKS(cudaSetDevice(s->device));
	if (round != s->cur_round) { int rc = enter_round(s, round); if (rc != APD_OK) return rc; }
	CKS(cudaMemcpyAsync(image, s->scaled[view], (size_t)s->rw * s->rh * 4, cudaMemcpyDeviceToHost, s->stream));
	CKS(cudaStreamSynchronize(s->stream));
	return APD_OK;
}

extern "C" int apd_scene_get_timing(apd_scene_handle s, double *patchmatch_ms, double *wall_ms, long long *launches) {
	if (!s) return APD_E_ARG;
	if (patchmatch_ms) *patchmatch_ms = s->pm_ms;
	if (wall_ms) *wall_ms = s->wall_ms;
	if (launches) *launches = s->launches;
	return APD_OK;
}
