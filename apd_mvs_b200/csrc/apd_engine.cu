// Host side of libapd_b200.so: the C-ABI of include/apd_b200.h on top of the sm_100a kernels.
// Replaces APD::CudaSpaceInitialization / SetDataPassHelperInCuda / RunPatchMatch / getters /
// ~APD (APD.cpp:361-397, :585-727; APD.cu:2386-2495). One handle = one device + one stream;
// all buffers are allocated once in apd_create and reused by every run (the reference
// re-allocates ~25 buffers + 2N cudaArrays per (view, pass)). No host synchronisation between
// the launches of a run; stage timing is taken with events on the handle's stream.
#include <cuda_runtime.h>
#include <cuda.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include "apd_device.cuh"
#include "apd_engine_internal.h"

namespace apd {
void launch_setup_views(cudaStream_t, const apd_camera *, int, ViewConst *, RefConst *, float *);
void launch_pad_ref(cudaStream_t, const float *, int, int, int, float *, int, int);
void launch_rng_seed(cudaStream_t, const Args &, unsigned long long);
cudaError_t launch_init_planes(cudaStream_t, const Args &);
cudaError_t launch_strong(cudaStream_t, const Args &, int iter, int color, const CUtensorMap *);
int make_tensor_maps(const float *, int, int, CUtensorMap *, CUtensorMap *);
void launch_depth_normal(cudaStream_t, const Args &);
void launch_median(cudaStream_t, const Args &, int color);
cudaError_t launch_sweep(cudaStream_t, const Args &, int mode, const CUtensorMap *);
// deformation path (apd_kernels_weak.cu)
cudaError_t launch_nearest_strong(cudaStream_t, const Args &);
cudaError_t launch_gen_anchors(cudaStream_t, const Args &, void *anchor_consts);
size_t anchor_consts_bytes();
cudaError_t launch_demote_unreliable(cudaStream_t, const Args &);
cudaError_t launch_fit_plane(cudaStream_t, const Args &);
cudaError_t launch_weak(cudaStream_t, const Args &, int iter, int color);
// quad-per-pixel WEAK propagation over compacted lists (apd_kernels_weakq.cu)
cudaError_t launch_weak_lists(cudaStream_t, const Args &, bool split);
cudaError_t launch_weak_q(cudaStream_t, const Args &, int iter, int color, int work_slot, int num_sms);
cudaError_t launch_sweep_q(cudaStream_t, const Args &, int mode, int num_sms);
}  // namespace apd

using namespace apd;


static thread_local std::string g_null_err = "null handle";

#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return APD_E_CUDA; } } while (0)

static int fail(apd_handle h, int code, const char *msg) { if (h) h->err = msg; return code; }

extern "C" void apd_default_params(apd_params *p) {
	if (!p) return;
	memset(p, 0, sizeof(*p));
	p->max_iterations = 3; p->num_images = 5; p->sigma_spatial = 5.0f; p->sigma_color = 3.0f; p->top_k = 4;
	p->depth_min = 0.0f; p->depth_max = 1.0f; p->geom_consistency = 0;
	p->strong_radius = 5; p->strong_increment = 2; p->weak_radius = 5; p->weak_increment = 5;
	p->use_APD = 1; p->weak_peak_radius = 2; p->rotate_time = 4; p->ransac_threshold = 0.005f; p->geom_factor = 0.2f;
	p->state = APD_FIRST_INIT;
}

static int check_params(apd_handle h, const apd_params *p) {
	// the kernels are specialised for the window geometry main.cpp never changes (main.h:83-86)
	if (p->strong_radius != 5 || p->strong_increment != 2 || p->weak_radius != 5 || p->weak_increment != 5)
		return fail(h, APD_E_LIMIT, "only strong 5/2 and weak 5/5 windows (the reference defaults) are built");
	if (p->max_iterations < 0 || p->max_iterations > 64) return fail(h, APD_E_ARG, "max_iterations out of range");
	if (p->state < APD_FIRST_INIT || p->state > APD_REFINE_ITER) return fail(h, APD_E_ARG, "bad state");
	if (p->rotate_time < 1 || p->rotate_time > 4) return fail(h, APD_E_ARG, "rotate_time must be 1..4 (APD.cu:1790)");
	if (p->top_k < 1) return fail(h, APD_E_ARG, "top_k must be >= 1");
	// k_sweep evaluates profile entries 1..59 around a centre window of at most 29 steps; the reference's peak rules
	// (APD.cu:2092-2143) match it for radii up to 28 (main.cpp uses 2..6)
	if (p->weak_peak_radius < 0 || p->weak_peak_radius > 28) return fail(h, APD_E_LIMIT, "weak_peak_radius must be 0..28");
	return APD_OK;
}

static int make_layered(apd_handle h, cudaArray_t *arr, cudaTextureObject_t *tex) {
	cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	// h->capacity layers, not the current image count: a later problem of the same round may have more source views
	CKH(cudaMalloc3DArray(arr, &desc, make_cudaExtent(h->W, h->H, h->capacity), cudaArrayLayered));
	cudaResourceDesc res; memset(&res, 0, sizeof(res));
	res.resType = cudaResourceTypeArray; res.res.array.array = *arr;
	cudaTextureDesc td; memset(&td, 0, sizeof(td));
	// The reference asks for Wrap with unnormalised coordinates (APD.cpp:598-602), which CUDA
	// executes as clamp-to-edge; Clamp is requested explicitly here. Linear filter, element reads.
	td.addressMode[0] = cudaAddressModeClamp; td.addressMode[1] = cudaAddressModeClamp; td.addressMode[2] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	CKH(cudaCreateTextureObject(tex, &res, &td, nullptr));
	return APD_OK;
}

extern "C" int apd_create(apd_handle *out, int device, int width, int height, int num_images, const apd_params *params, uint64_t seed) {
	if (!out || !params) return APD_E_ARG;
	*out = nullptr;
	if (width < 16 || height < 16 || width > 32767 || height > 32767) return APD_E_LIMIT;   // anchors are short2 (APD.h:48)
	if (num_images < 2 || num_images > APD_MAX_IMAGES) return APD_E_LIMIT;                   // APD.cpp:428-431
	apd_engine *h = new apd_engine();
	h->device = device; h->W = width; h->H = height; h->N = num_images; h->S = num_images - 1; h->capacity = num_images;
	h->npx = (size_t)width * height; h->seed = seed; h->params = *params; h->params.num_images = num_images;
	int rc = check_params(h, params);
	if (rc != APD_OK) { delete h; return rc; }
	auto bail = [&](int code) { apd_destroy(h); return code; };
	if (cudaSetDevice(device) != cudaSuccess) return bail(APD_E_CUDA);
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(APD_E_CUDA);
	if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(APD_E_CUDA);
	if (cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_late, cudaEventDisableTiming) != cudaSuccess) return bail(APD_E_CUDA);
	const size_t n = h->npx;
	h->ref_pitch = ((width + 2 * kRefPad + 3) / 4) * 4;
	h->ref_rows = height + 2 * kRefPad;
#define ALLOC(ptr, bytes) if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) return bail(APD_E_CUDA)
	ALLOC(h->ref_lin, n * 4);
	ALLOC(h->ref_pad, (size_t)h->ref_pitch * h->ref_rows * 4);
	ALLOC(h->d_cams, sizeof(apd_camera) * num_images);
	ALLOC(h->d_views, sizeof(ViewConst) * APD_MAX_IMAGES);
	ALLOC(h->d_ref, sizeof(RefConst));
	ALLOC(h->d_invw, 16);
	ALLOC(h->planes, n * 16); ALLOC(h->fit_planes, n * 16); ALLOC(h->prior_planes, n * 16);
	ALLOC(h->costs, n * 4);
	ALLOC(h->sel_views, n * 4); ALLOC(h->prior_views, n * 4);
	ALLOC(h->states, n); ALLOC(h->prior_states, n); ALLOC(h->reliable, n);
	ALLOC(h->rng, n * 24); ALLOC(h->view_w, n * 16);
	ALLOC(h->anchors, n * APD_NEIGHBOUR_NUM * sizeof(short2)); ALLOC(h->nearest, n * sizeof(short2));
	{   // slab pool (apd_device.cuh): one slab per resident block, big enough for k_strong / k_weak's [9*S][128] cost
		// matrices and k_sweep's [61][128] profile
		const int rows = 9 * h->S > 61 ? 9 * h->S : 61;
		h->slab_stride = rows * 128;
		ALLOC(h->scratch, (size_t)kSlabSMs * kSlabPerSM * h->slab_stride * 4);
		ALLOC(h->slab_slots, (size_t)kSlabSMs * kSlabPerSM * 4);
	}
	h->wlist_stride = (int)(((n + 1) / 2 + 63) / 64 * 64);
	ALLOC(h->wlist, (size_t)2 * h->wlist_stride * sizeof(int));
	ALLOC(h->wctrl, (size_t)(kWorkBase + kWorkSlots) * sizeof(int));
	ALLOC(h->anchor_consts, anchor_consts_bytes());
	{
		int sms = 0;
		if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms < 1) return bail(APD_E_CUDA);
		h->num_sms = sms;
		const char *w = getenv("APD_WEAK_IMPL");
		h->weak_impl = (w && !strcmp(w, "old")) ? 0 : 1;
		const char *sw = getenv("APD_SWEEP_IMPL");
		h->sweep_impl = (sw && !strcmp(sw, "old")) ? 0 : 1;
	}
#undef ALLOC
	if (make_layered(h, &h->img_arr, &h->img_tex) != APD_OK) return bail(APD_E_CUDA);
	if (make_tensor_maps(h->ref_pad, h->ref_pitch, h->ref_rows, &h->tmap_strong, &h->tmap_sweep) != 0) { h->err = "cuTensorMapEncodeTiled failed"; return bail(APD_E_CUDA); }
	cudaMemsetAsync(h->slab_slots, 0, (size_t)kSlabSMs * kSlabPerSM * 4, h->stream);
	cudaMemsetAsync(h->costs, 0, n * 4, h->stream);
	cudaMemsetAsync(h->view_w, 0, n * 16, h->stream);
	cudaMemsetAsync(h->planes, 0, n * 16, h->stream);
	cudaMemsetAsync(h->fit_planes, 0, n * 16, h->stream);
	cudaMemsetAsync(h->reliable, 0, n, h->stream);
	if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(APD_E_CUDA);
	*out = h;
	return APD_OK;
}

extern "C" void apd_destroy(apd_handle h) {
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
	if (h->ev_early) cudaEventDestroy(h->ev_early);
	if (h->ev_late) cudaEventDestroy(h->ev_late);
	for (auto e : h->events) cudaEventDestroy(e);
	if (h->img_tex) cudaDestroyTextureObject(h->img_tex);
	if (h->depth_tex) cudaDestroyTextureObject(h->depth_tex);
	if (h->img_arr) cudaFreeArray(h->img_arr);
	if (h->depth_arr) cudaFreeArray(h->depth_arr);
	void *ptrs[] = {h->ref_lin, h->ref_pad, h->d_cams, h->d_views, h->d_ref, h->d_invw, h->planes, h->fit_planes, h->prior_planes,
	                h->costs, h->sel_views, h->prior_views, h->states, h->prior_states, h->reliable, h->rng, h->view_w, h->anchors, h->nearest, h->scratch, h->slab_slots, h->wlist, h->wctrl, h->anchor_consts};
	for (void *p : ptrs) if (p) cudaFree(p);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
}

extern "C" const char *apd_last_error(apd_handle h) { return h ? h->err.c_str() : g_null_err.c_str(); }

extern "C" int apd_set_params(apd_handle h, const apd_params *p) {
	if (!h || !p) return APD_E_ARG;
	int rc = check_params(h, p); if (rc != APD_OK) return rc;
	h->params = *p; h->params.num_images = h->N;
	return APD_OK;
}
extern "C" int apd_set_seed(apd_handle h, uint64_t seed) { if (!h) return APD_E_ARG; h->seed = seed; return APD_OK; }

static void drain_uploads(apd_handle h);
extern "C" int apd_set_num_images(apd_handle h, int num_images) {
	if (!h) return APD_E_ARG;
	if (num_images < 2 || num_images > h->capacity) return fail(h, APD_E_LIMIT, "num_images must be 2..the count given to apd_create");
	if (num_images != h->N) { drain_uploads(h); h->have_images = h->have_cams = h->have_depths = false; }
	h->N = num_images; h->S = num_images - 1; h->params.num_images = num_images;
	return APD_OK;
}

extern "C" int apd_reset_inputs(apd_handle h) {
	if (!h) return APD_E_ARG;
	drain_uploads(h);
	h->have_images = h->have_cams = h->have_depths = h->have_planes = h->have_states = false;
	return APD_OK;
}
extern "C" int apd_get_capacity(apd_handle h) { return h ? h->capacity : 0; }

extern "C" int apd_set_upload_mode(apd_handle h, int asynchronous) {
	if (!h) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	if (h->async_upload && !asynchronous) { CKH(cudaStreamSynchronize(h->copy_stream)); h->pending_early = h->pending_late = false; }
	h->async_upload = asynchronous != 0;
	return APD_OK;
}
// Stream a host upload goes to, and what ends it: a wait (default) or an event the run waits for (asynchronous mode).
static inline cudaStream_t upload_stream(apd_handle h, bool host) { return (h->async_upload && host) ? h->copy_stream : h->stream; }
static int finish_upload(apd_handle h, bool host, bool late) {
	if (h->async_upload && host) {
		CKH(cudaEventRecord(late ? h->ev_late : h->ev_early, h->copy_stream));
		(late ? h->pending_late : h->pending_early) = true;
		return APD_OK;
	}
	CKH(cudaStreamSynchronize(h->stream));
	return APD_OK;
}

// Asynchronous uploads still in flight are completed before anything that ends their contract early (inputs forgotten, a run refused):
// the caller may release its host buffers as soon as such a call returns.
static void drain_uploads(apd_handle h) {
	if (h->pending_early || h->pending_late) { cudaStreamSynchronize(h->copy_stream); h->pending_early = h->pending_late = false; }
}

extern "C" int apd_set_cameras(apd_handle h, const apd_camera *cams) {
	if (!h || !cams) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	for (int i = 0; i < h->N; ++i)
		if (cams[i].width != h->W || cams[i].height != h->H) return fail(h, APD_E_ARG, "camera width/height must equal the image size (APD.cpp:439-449)");
	if (h->async_upload) {   // the caller's array may be a temporary (112 B per camera): keep our own copy until the run
		h->cams_host.assign(cams, cams + h->N);
		CKH(cudaMemcpyAsync(h->d_cams, h->cams_host.data(), sizeof(apd_camera) * h->N, cudaMemcpyHostToDevice, h->copy_stream));
	} else CKH(cudaMemcpyAsync(h->d_cams, cams, sizeof(apd_camera) * h->N, cudaMemcpyHostToDevice, h->stream));
	{ int rc = finish_upload(h, true, false); if (rc) return rc; }
	h->have_cams = true;
	return APD_OK;
}

static int copy_stack(apd_handle h, cudaArray_t arr, const float *const *host_imgs, const float *dev_stack, size_t pitch, size_t stride) {
	cudaStream_t st = upload_stream(h, host_imgs != nullptr);
	for (int i = 0; i < h->N; ++i) {
		cudaMemcpy3DParms p; memset(&p, 0, sizeof(p));
		const float *src = host_imgs ? host_imgs[i] : (const float *)((const char *)dev_stack + (size_t)i * stride);
		p.srcPtr = make_cudaPitchedPtr((void *)src, pitch, h->W, h->H);
		p.dstArray = arr; p.dstPos = make_cudaPos(0, 0, i);
		p.extent = make_cudaExtent(h->W, h->H, 1);
		p.kind = cudaMemcpyDefault;      // host or device pointers (unified addressing)
		CKH(cudaMemcpy3DAsync(&p, st));
	}
	return APD_OK;
}

static int finish_images(apd_handle h, const float *img0, size_t pitch, bool host) {
	cudaStream_t st = upload_stream(h, host);
	CKH(cudaMemcpy2DAsync(h->ref_lin, (size_t)h->W * 4, img0, pitch, (size_t)h->W * 4, h->H, cudaMemcpyDefault, st));
	launch_pad_ref(st, h->ref_lin, h->W, h->H, h->W, h->ref_pad, h->ref_pitch, h->ref_rows);
	CKH(cudaGetLastError());
	{ int rc = finish_upload(h, host, true); if (rc) return rc; }
	h->have_images = true;
	return APD_OK;
}

extern "C" int apd_set_images(apd_handle h, const float *const *images, size_t pitch_bytes) {
	if (!h || !images || pitch_bytes < (size_t)h->W * 4) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	int rc = copy_stack(h, h->img_arr, images, nullptr, pitch_bytes, 0); if (rc) return rc;
	return finish_images(h, images[0], pitch_bytes, true);
}
extern "C" int apd_set_images_device(apd_handle h, const float *dev_stack, size_t pitch_bytes, size_t image_stride_bytes) {
	if (!h || !dev_stack || pitch_bytes < (size_t)h->W * 4) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	int rc = copy_stack(h, h->img_arr, nullptr, dev_stack, pitch_bytes, image_stride_bytes); if (rc) return rc;
	return finish_images(h, dev_stack, pitch_bytes, false);
}

static int ensure_depth_array(apd_handle h) {
	if (h->depth_arr) return APD_OK;
	return make_layered(h, &h->depth_arr, &h->depth_tex);
}
extern "C" int apd_set_depths(apd_handle h, const float *const *depths, size_t pitch_bytes) {
	if (!h || !depths || pitch_bytes < (size_t)h->W * 4) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	int rc = ensure_depth_array(h); if (rc) return rc;
	rc = copy_stack(h, h->depth_arr, depths, nullptr, pitch_bytes, 0); if (rc) return rc;
	rc = finish_upload(h, true, true); if (rc) return rc;
	h->have_depths = true;
	return APD_OK;
}
extern "C" int apd_set_depths_device(apd_handle h, const float *dev_stack, size_t pitch_bytes, size_t image_stride_bytes) {
	if (!h || !dev_stack || pitch_bytes < (size_t)h->W * 4) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	int rc = ensure_depth_array(h); if (rc) return rc;
	rc = copy_stack(h, h->depth_arr, nullptr, dev_stack, pitch_bytes, image_stride_bytes); if (rc) return rc;
	CKH(cudaStreamSynchronize(h->stream));
	h->have_depths = true;
	return APD_OK;
}

extern "C" int apd_set_priors(apd_handle h, const float *planes, const uint32_t *views, const uint8_t *states) {
	if (!h) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	const size_t n = h->npx;
	// host or device pointers: only host memory takes the copy stream in asynchronous mode
	auto is_host = [](const void *p) { cudaPointerAttributes at; return cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeUnregistered; };
	const bool host = (planes ? is_host(planes) : true) && (states ? is_host(states) : true);
	cudaGetLastError();
	cudaStream_t st = upload_stream(h, host);
	if (planes) {
		if (!views) return fail(h, APD_E_ARG, "planes need views (APD.cpp:552-581)");
		CKH(cudaMemcpyAsync(h->prior_planes, planes, n * 16, cudaMemcpyDefault, st));
		CKH(cudaMemcpyAsync(h->prior_views, views, n * 4, cudaMemcpyDefault, st));
		h->have_planes = true;
	}
	if (states) {
		CKH(cudaMemcpyAsync(h->prior_states, states, n, cudaMemcpyDefault, st));
		h->have_states = true;
	}
	return finish_upload(h, host, false);
}

extern "C" int apd_num_stages(apd_handle h) { return h ? 10 + 5 * h->params.max_iterations : 0; }

static Args make_args(apd_handle h) {
	Args a; memset(&a, 0, sizeof(a));
	const apd_params &p = h->params;
	a.W = h->W; a.H = h->H; a.S = h->S;
	a.half_rows = 32 * ((h->H / 2 + 15) / 16);     // rows the reference's half launch reaches (APD.cu:2400-2403)
	a.ref_pitch = h->ref_pitch;
	a.depth_min = p.depth_min; a.depth_max = p.depth_max;
	a.top_k = p.top_k; a.state = p.state; a.geom = p.geom_consistency ? 1 : 0;
	a.weak_peak_radius = p.weak_peak_radius; a.rotate_time = p.rotate_time;
	a.geom_factor = p.geom_factor; a.ransac_threshold = p.ransac_threshold;
	a.inv_w = h->d_invw;
	a.img_tex = h->img_tex; a.depth_tex = h->depth_tex;
	a.ref_pad = h->ref_pad; a.views = h->d_views; a.ref = h->d_ref;
	a.planes = h->planes; a.fit_planes = h->fit_planes; a.costs = h->costs;
	a.sel_views = h->sel_views; a.states = h->states; a.rng = h->rng; a.view_w = h->view_w;
	a.anchors = h->anchors; a.nearest = h->nearest; a.reliable = h->reliable; a.scratch = h->scratch; a.slab_slots = h->slab_slots; a.slab_stride = h->slab_stride;
	a.wlist = h->wlist; a.wlist_stride = h->wlist_stride; a.wctrl = h->wctrl;
	return a;
}

extern "C" int apd_run_until(apd_handle h, int stage_end) {
	if (!h) return APD_E_ARG;
	const apd_params &p = h->params;
	const char *missing = nullptr;
	if (!h->have_images || !h->have_cams) missing = "set cameras and images first";
	else if (p.geom_consistency && !h->have_depths) missing = "geom_consistency needs apd_set_depths (APD.cpp:492-510)";
	else if (p.state != APD_FIRST_INIT && !h->have_planes) missing = "state != FIRST_INIT needs prior planes+views (APD.cpp:552-581)";
	else if (p.use_APD && !h->have_states) missing = "use_APD needs prior pixel states (APD.cpp:513-519)";
	if (missing) { drain_uploads(h); return fail(h, APD_E_STATE, missing); }
	CKH(cudaSetDevice(h->device));
	const int nstages = apd_num_stages(h);
	if (stage_end < 0 || stage_end >= nstages) stage_end = nstages - 1;
	while ((int)h->events.size() < nstages + 1) { cudaEvent_t e; CKH(cudaEventCreate(&e)); h->events.push_back(e); }
	cudaStream_t st = h->stream;
	const size_t n = h->npx;
	const Args a = make_args(h);
	h->launches = 0;
	const bool apd_on = p.use_APD != 0;

	// ---- asynchronous uploads: cameras and priors are read by the very first operations
	if (h->pending_early) { CKH(cudaStreamWaitEvent(st, h->ev_early, 0)); }
	// ---- restore the run's inputs (what CudaSpaceInitialization uploads, APD.cpp:643-661)
	if (apd_on) CKH(cudaMemcpyAsync(h->states, h->prior_states, n, cudaMemcpyDeviceToDevice, st));
	else CKH(cudaMemsetAsync(h->states, APD_STRONG, n, st));                       // APD.cpp:540-548
	if (p.state != APD_FIRST_INIT) {
		CKH(cudaMemcpyAsync(h->planes, h->prior_planes, n * 16, cudaMemcpyDeviceToDevice, st));
		CKH(cudaMemcpyAsync(h->sel_views, h->prior_views, n * 4, cudaMemcpyDeviceToDevice, st));
	} else {
		CKH(cudaMemsetAsync(h->planes, 0, n * 16, st));
		CKH(cudaMemsetAsync(h->sel_views, 0, n * 4, st));
	}
	CKH(cudaMemsetAsync(h->fit_planes, 0, n * 16, st));                            // APD.cpp:651
	CKH(cudaMemsetAsync(h->slab_slots, 0, (size_t)kSlabSMs * kSlabPerSM * 4, st));  // all slabs free
	CKH(cudaMemsetAsync(h->wctrl, 0, (size_t)(kWorkBase + kWorkSlots) * sizeof(int), st));
	// the reference allocates these per run (APD.cpp:645-649): pixels no update reaches must not see the previous run's values
	CKH(cudaMemsetAsync(h->view_w, 0, n * 16, st));
	CKH(cudaMemsetAsync(h->costs, 0, n * 4, st));
	launch_setup_views(st, h->d_cams, h->S, h->d_views, h->d_ref, h->d_invw); h->launches++;
	CKH(cudaGetLastError());

	int stage = 0;
	CKH(cudaEventRecord(h->events[0], st));
#define STAGE_END() do { CKH(cudaGetLastError()); CKH(cudaEventRecord(h->events[stage + 1], st)); if (stage == stage_end) goto done; ++stage; } while (0)
	launch_rng_seed(st, a, h->seed); h->launches++; STAGE_END();                                          // 0  K1
	if (apd_on) { CKH(launch_nearest_strong(st, a)); h->launches += 2; } STAGE_END();                         // 1  K2
	if (apd_on) { CKH(launch_gen_anchors(st, a, h->anchor_consts)); h->launches += 3; } STAGE_END();        // 2  K3
	if (apd_on) {                                                                                          // 3  K4 (+ the WEAK lists K9/K10 walk)
		CKH(launch_demote_unreliable(st, a)); h->launches++;
		if (h->weak_impl == 1) { CKH(launch_weak_lists(st, a, true)); h->launches++; }
	}
	STAGE_END();
	if (h->pending_late) { CKH(cudaStreamWaitEvent(st, h->ev_late, 0)); }     // images, padded reference, depth maps: first read by K5
	CKH(launch_init_planes(st, a)); h->launches++; STAGE_END();                                            // 4  K5
	for (int it = 0; it < p.max_iterations; ++it) {
		CKH(launch_strong(st, a, it, 0, &h->tmap_strong)); h->launches++; STAGE_END();                                      // K6
		CKH(launch_strong(st, a, it, 1, &h->tmap_strong)); h->launches++; STAGE_END();                                      // K7
		if (apd_on) { CKH(launch_fit_plane(st, a)); h->launches++; } STAGE_END();                          // K8
		const int slot = (2 * it) % kWorkSlots;
		if (apd_on) { CKH(h->weak_impl == 1 ? launch_weak_q(st, a, it, 0, slot, h->num_sms) : launch_weak(st, a, it, 0)); h->launches++; } STAGE_END();        // K9
		if (apd_on) { CKH(h->weak_impl == 1 ? launch_weak_q(st, a, it, 1, slot + 1, h->num_sms) : launch_weak(st, a, it, 1)); h->launches++; } STAGE_END();    // K10
	}
	launch_depth_normal(st, a); h->launches++; STAGE_END();                                                // K11
	launch_median(st, a, 0); h->launches++; STAGE_END();                                                   // K12
	launch_median(st, a, 1); h->launches++; STAGE_END();                                                   // K13
	// K14 and K15 are one fused launch (reported in the K14 slot) unless the run stops between them
	if (stage_end == stage) { CKH(h->sweep_impl == 1 ? launch_sweep_q(st, a, 0, h->num_sms) : launch_sweep(st, a, 0, &h->tmap_sweep)); h->launches++; STAGE_END(); }   // K14 alone
	CKH(h->sweep_impl == 1 ? launch_sweep_q(st, a, 2, h->num_sms) : launch_sweep(st, a, 2, &h->tmap_sweep)); h->launches++; STAGE_END();                                 // K14 + K15
	STAGE_END();
#undef STAGE_END
done:
	CKH(cudaStreamSynchronize(st));
	if (h->pending_early || h->pending_late) {   // the host buffers are free again when the run returns, whatever stage it stopped at
		CKH(cudaStreamSynchronize(h->copy_stream)); h->pending_early = h->pending_late = false;
	}
	h->stages_run = stage + 1;
	h->stage_ms.assign(nstages, 0.0f);
	for (int i = 0; i < h->stages_run; ++i) cudaEventElapsedTime(&h->stage_ms[i], h->events[i], h->events[i + 1]);
	return APD_OK;
}

extern "C" int apd_run(apd_handle h) { return apd_run_until(h, -1); }

#define GETTER(name, type, field, bytes_per_px)                                                        \
	extern "C" int name(apd_handle h, type *out) {                                                     \
		if (!h || !out) return APD_E_ARG;                                                              \
		CKH(cudaSetDevice(h->device));                                                                 \
		CKH(cudaMemcpyAsync(out, h->field, h->npx * (bytes_per_px), cudaMemcpyDeviceToHost, h->stream)); \
		CKH(cudaStreamSynchronize(h->stream));                                                         \
		return APD_OK;                                                                                 \
	}
GETTER(apd_get_planes, float, planes, 16)
GETTER(apd_get_states, uint8_t, states, 1)
GETTER(apd_get_views, uint32_t, sel_views, 4)
GETTER(apd_get_costs, float, costs, 4)
GETTER(apd_get_rng, uint32_t, rng, 24)
#undef GETTER

extern "C" int apd_get_view_weights(apd_handle h, uint8_t *out) {
	if (!h || !out) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	std::vector<uint32_t> packed(h->npx * 4);
	CKH(cudaMemcpyAsync(packed.data(), h->view_w, h->npx * 16, cudaMemcpyDeviceToHost, h->stream));
	CKH(cudaStreamSynchronize(h->stream));
	for (size_t i = 0; i < h->npx; ++i)
		for (int v = 0; v < 32; ++v) out[i * 32 + v] = (uint8_t)((packed[i * 4 + (v >> 3)] >> (4 * (v & 7))) & 15u);
	return APD_OK;
}

extern "C" int apd_get_anchors(apd_handle h, int16_t *anchors_xy, int16_t *nearest_xy, uint8_t *reliable, float *fit_planes) {
	if (!h) return APD_E_ARG;
	CKH(cudaSetDevice(h->device));
	const size_t n = h->npx;
	if (anchors_xy) {   // device layout [9][n] -> caller layout [n][9]
		std::vector<short2> tmp(n * APD_NEIGHBOUR_NUM);
		CKH(cudaMemcpyAsync(tmp.data(), h->anchors, n * APD_NEIGHBOUR_NUM * sizeof(short2), cudaMemcpyDeviceToHost, h->stream));
		CKH(cudaStreamSynchronize(h->stream));
		for (size_t i = 0; i < n; ++i)
			for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) {
				anchors_xy[(i * APD_NEIGHBOUR_NUM + k) * 2 + 0] = tmp[(size_t)k * n + i].x;
				anchors_xy[(i * APD_NEIGHBOUR_NUM + k) * 2 + 1] = tmp[(size_t)k * n + i].y;
			}
	}
	if (nearest_xy) CKH(cudaMemcpyAsync(nearest_xy, h->nearest, n * sizeof(short2), cudaMemcpyDeviceToHost, h->stream));
	if (reliable) CKH(cudaMemcpyAsync(reliable, h->reliable, n, cudaMemcpyDeviceToHost, h->stream));
	if (fit_planes) CKH(cudaMemcpyAsync(fit_planes, h->fit_planes, n * 16, cudaMemcpyDeviceToHost, h->stream));
	CKH(cudaStreamSynchronize(h->stream));
	return APD_OK;
}

extern "C" int apd_get_stage_ms(apd_handle h, float *ms, int capacity) {
	if (!h) return 0;
	const int n = (int)h->stage_ms.size();
	for (int i = 0; i < n && i < capacity; ++i) ms[i] = h->stage_ms[i];
	return n;
}
extern "C" int apd_get_launch_count(apd_handle h) { return h ? h->launches : 0; }
extern "C" void *apd_get_stream(apd_handle h) { return h ? (void *)h->stream : nullptr; }
