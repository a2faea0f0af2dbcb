// oracle/_ref/libapd_ref.so — the UNMODIFIED reference PatchMatch (APD.cu of whoiszzj/APD-MVS)
// recompiled for sm_100 behind a small C-ABI. TEST / BASELINE INFRASTRUCTURE ONLY:
// only tests/, __graft_entry__.smoke() and bench.py's reference arm may load it.
//
// The reference source is compiled from where it lies (/root/reference/APD.cu, passed as
// -DAPD_REF_SOURCE=...); nothing of it is copied into this repository. Three preprocessor
// seams make it drivable without editing it:
//   * clock64()              -> a __device__ seed variable   (APD.cu:803 seeds curand with clock64())
//   * cudaDeviceSynchronize  -> a hook that really syncs, then time-stamps / snapshots the
//                               device state after each of the 25 launches (APD.cu:2409-2494)
//   * private                -> public, so this harness can fill the members that
//                               APD::InuputInitialization / CudaSpaceInitialization /
//                               SetDataPassHelperInCuda (APD.cpp:399-699, needs OpenCV+Boost,
//                               not buildable here) would have filled from disk.
// The harness below is our own code: it uploads caller-supplied arrays with the same texture
// descriptors and buffer shapes as APD.cpp:585-699 and then calls the reference's own
// APD::RunPatchMatch().
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <curand_kernel.h>
#include <vector_types.h>
#include <vector>
#include <string>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <map>
#include <memory>
#include <chrono>
#include <iomanip>
#include <unordered_set>
#include <cstdarg>
#include <random>
#include <unordered_map>
#include <opencv2/opencv.hpp>
#include <boost/filesystem.hpp>

// ---- seam 1: deterministic seed instead of clock64() -------------------------------------
__device__ unsigned long long g_apdref_seed = 0ULL;

// ---- seam 2: sync hook ---------------------------------------------------------------------
struct ApdRefCtx;
static ApdRefCtx *g_ctx = nullptr;
static cudaError_t apdref_sync_hook();
static inline cudaError_t apdref_real_sync() { return cudaDeviceSynchronize(); }

#define clock64() (g_apdref_seed)
#define cudaDeviceSynchronize() apdref_sync_hook()
#define private public
#include APD_REF_SOURCE
#undef private
#undef cudaDeviceSynchronize
#undef clock64

static_assert(sizeof(Camera) == 112, "Camera layout (main.h:47-56)");
static_assert(sizeof(PatchMatchParams) == 72, "PatchMatchParams layout (main.h:75-94)");
static_assert(sizeof(curandState) == 48, "curandState XORWOW");

// Members declared in APD.h whose definitions live in APD.cpp (not compiled here).
APD::APD(const Problem &p) { params_host = p.params; this->problem = p; }
APD::~APD() {}
void CudaSafeCall(const cudaError_t error, const std::string &file, const int line) {
	if (error != cudaSuccess) {
		std::cerr << cudaGetErrorString(error) << " in " << file << " at line " << line << std::endl;
		exit(EXIT_FAILURE);  // reference behaviour, APD.cpp:315-323
	}
}

struct Snapshot {
	std::vector<float4> planes;
	std::vector<float> costs;
	std::vector<unsigned> views;
	std::vector<uchar> states;
	std::vector<uchar> view_weights;
	std::vector<curandState> rng;
};

struct ApdRefCtx {
	APD *apd = nullptr;
	int device = 0, W = 0, H = 0, N = 0;
	unsigned long long seed = 0;
	bool uploaded = false;
	std::vector<uchar> weak_host;
	std::vector<int> nmap_host;
	std::vector<unsigned> views_host;
	// run bookkeeping
	int sync_index = 0;
	std::vector<double> stage_ms;
	std::chrono::steady_clock::time_point t_prev;
	std::vector<int> want;
	std::map<int, Snapshot> snaps;
	std::string err;
};

static void take_snapshot(ApdRefCtx *c, Snapshot &s) {
	APD *a = c->apd;
	const size_t n = (size_t)c->W * c->H;
	s.planes.resize(n); s.costs.resize(n); s.views.resize(n); s.states.resize(n);
	s.view_weights.resize(n * MAX_IMAGES); s.rng.resize(n);
	cudaMemcpy(s.planes.data(), a->plane_hypotheses_cuda, n * sizeof(float4), cudaMemcpyDeviceToHost);
	cudaMemcpy(s.costs.data(), a->costs_cuda, n * sizeof(float), cudaMemcpyDeviceToHost);
	cudaMemcpy(s.views.data(), a->selected_views_cuda, n * sizeof(unsigned), cudaMemcpyDeviceToHost);
	cudaMemcpy(s.states.data(), a->weak_info_cuda, n, cudaMemcpyDeviceToHost);
	cudaMemcpy(s.view_weights.data(), a->view_weight_cuda, n * MAX_IMAGES, cudaMemcpyDeviceToHost);
	cudaMemcpy(s.rng.data(), a->rand_states_cuda, n * sizeof(curandState), cudaMemcpyDeviceToHost);
}

static cudaError_t apdref_sync_hook() {
	cudaError_t e = apdref_real_sync();
	ApdRefCtx *c = g_ctx;
	if (!c) return e;
	auto now = std::chrono::steady_clock::now();
	c->stage_ms.push_back(std::chrono::duration<double, std::milli>(now - c->t_prev).count());
	if (std::find(c->want.begin(), c->want.end(), c->sync_index) != c->want.end()) {
		take_snapshot(c, c->snaps[c->sync_index]);
		now = std::chrono::steady_clock::now();
	}
	c->t_prev = now;
	c->sync_index++;
	return e;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -2; } } while (0)

static void make_tex(cudaArray **arr, cudaTextureObject_t *tex, const float *src, size_t pitch, int W, int H) {
	// same descriptor as APD.cpp:589-603 (Wrap + Linear + unnormalised coordinates)
	cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
	cudaMallocArray(arr, &desc, W, H);
	cudaMemcpy2DToArray(*arr, 0, 0, src, pitch, W * sizeof(float), H, cudaMemcpyHostToDevice);
	cudaResourceDesc res; memset(&res, 0, sizeof(res));
	res.resType = cudaResourceTypeArray; res.res.array.array = *arr;
	cudaTextureDesc td; memset(&td, 0, sizeof(td));
	td.addressMode[0] = cudaAddressModeWrap; td.addressMode[1] = cudaAddressModeWrap;
	td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
	cudaCreateTextureObject(tex, &res, &td, NULL);
}

extern "C" {

int apdref_create(void **out, int device, int W, int H, int N, const void *params72, unsigned long long seed) {
	if (!out || !params72 || N < 2 || N > MAX_IMAGES || W <= 0 || H <= 0) return -1;
	ApdRefCtx *c = new ApdRefCtx();
	c->device = device; c->W = W; c->H = H; c->N = N; c->seed = seed;
	if (cudaSetDevice(device) != cudaSuccess) { delete c; return -2; }
	Problem pb; pb.index = 0; pb.ref_image_id = 0; pb.iteration = 0;
	memcpy(&pb.params, params72, sizeof(PatchMatchParams));
	pb.params.num_images = N;
	c->apd = new APD(pb);
	c->apd->num_images = N; c->apd->width = W; c->apd->height = H;
	c->apd->plane_hypotheses_host = new float4[(size_t)W * H];
	*out = c;
	return 0;
}

// One call doing what InuputInitialization's tail + CudaSpaceInitialization + SetDataPassHelperInCuda do.
// images/depths: N host pointers, row pitch in bytes. depths may be NULL unless geom_consistency.
// planes (float4 world normal + depth), views, states may be NULL for FIRST_INIT / !use_APD.
int apdref_upload(void *h, const float *const *images, size_t pitch, const void *cameras,
                  const float *const *depths, size_t dpitch,
                  const float *planes, const unsigned *views, const unsigned char *states) {
	ApdRefCtx *c = (ApdRefCtx *)h; APD *a = c->apd;
	if (c->uploaded) { c->err = "already uploaded"; return -1; }
	CK(cudaSetDevice(c->device));
	const int W = c->W, H = c->H, N = c->N; const size_t n = (size_t)W * H;
	PatchMatchParams &P = a->params_host;
	a->cameras.assign((const Camera *)cameras, (const Camera *)cameras + N);
	for (int i = 0; i < N; ++i) make_tex(&a->cuArray[i], &a->texture_objects_host.images[i], images[i], pitch, W, H);
	CK(cudaMalloc((void **)&a->texture_objects_cuda, sizeof(cudaTextureObjects)));
	CK(cudaMemcpy(a->texture_objects_cuda, &a->texture_objects_host, sizeof(cudaTextureObjects), cudaMemcpyHostToDevice));
	a->texture_depths_cuda = nullptr;
	if (P.geom_consistency) {
		if (!depths) { c->err = "geom_consistency needs depths"; return -1; }
		for (int i = 0; i < N; ++i) make_tex(&a->cuDepthArray[i], &a->texture_depths_host.images[i], depths[i], dpitch, W, H);
		CK(cudaMalloc((void **)&a->texture_depths_cuda, sizeof(cudaTextureObjects)));
		CK(cudaMemcpy(a->texture_depths_cuda, &a->texture_depths_host, sizeof(cudaTextureObjects), cudaMemcpyHostToDevice));
	}
	// pixel states + compact weak index (APD.cpp:513-548)
	c->weak_host.assign(n, (uchar)STRONG);
	c->nmap_host.assign(n, 0);
	a->weak_count = 0;
	if (P.use_APD) {
		if (!states) { c->err = "use_APD needs states"; return -1; }
		for (size_t i = 0; i < n; ++i) {
			c->weak_host[i] = states[i];
			if (states[i] == WEAK) c->nmap_host[i] = a->weak_count++;
		}
	}
	c->views_host.assign(n, 0u);
	if (P.state != FIRST_INIT) {
		if (!planes || !views) { c->err = "state != FIRST_INIT needs planes+views"; return -1; }
		memcpy(a->plane_hypotheses_host, planes, n * sizeof(float4));
		memcpy(c->views_host.data(), views, n * sizeof(unsigned));
	} else {
		memset(a->plane_hypotheses_host, 0, n * sizeof(float4));
	}
	a->weak_info_host.rows = H; a->weak_info_host.cols = W; a->weak_info_host.data = c->weak_host.data(); a->weak_info_host.step = W;
	a->selected_views_host.rows = H; a->selected_views_host.cols = W; a->selected_views_host.data = (unsigned char *)c->views_host.data(); a->selected_views_host.step = W * 4;

	CK(cudaMalloc((void **)&a->cameras_cuda, sizeof(Camera) * N));
	CK(cudaMemcpy(a->cameras_cuda, a->cameras.data(), sizeof(Camera) * N, cudaMemcpyHostToDevice));
	CK(cudaMalloc((void **)&a->costs_cuda, sizeof(float) * n));
	CK(cudaMemset(a->costs_cuda, 0, sizeof(float) * n));
	CK(cudaMalloc((void **)&a->rand_states_cuda, sizeof(curandState) * n));
	CK(cudaMalloc((void **)&a->selected_views_cuda, sizeof(unsigned) * n));
	CK(cudaMemcpy(a->selected_views_cuda, c->views_host.data(), sizeof(unsigned) * n, cudaMemcpyHostToDevice));
	CK(cudaMalloc((void **)&a->view_weight_cuda, n * MAX_IMAGES));
	CK(cudaMemset(a->view_weight_cuda, 0, n * MAX_IMAGES));
	CK(cudaMalloc((void **)&a->plane_hypotheses_cuda, sizeof(float4) * n));
	CK(cudaMemcpy(a->plane_hypotheses_cuda, a->plane_hypotheses_host, sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMalloc((void **)&a->fit_plane_hypotheses_cuda, sizeof(float4) * n));
	CK(cudaMemset(a->fit_plane_hypotheses_cuda, 0, sizeof(float4) * n));
	CK(cudaMalloc((void **)&a->weak_info_cuda, n));
	CK(cudaMemcpy(a->weak_info_cuda, c->weak_host.data(), n, cudaMemcpyHostToDevice));
	CK(cudaMalloc((void **)&a->weak_reliable_cuda, n));
	CK(cudaMemset(a->weak_reliable_cuda, 0, n));
	CK(cudaMalloc((void **)&a->weak_nearest_strong, n * sizeof(short2)));
	CK(cudaMalloc((void **)&a->neigbours_map_cuda, n * sizeof(int)));
	CK(cudaMemcpy(a->neigbours_map_cuda, c->nmap_host.data(), n * sizeof(int), cudaMemcpyHostToDevice));
	CK(cudaMalloc((void **)&a->neighbours_cuda, (size_t)std::max(a->weak_count, 1) * NEIGHBOUR_NUM * sizeof(short2)));
	CK(cudaMalloc((void **)&a->params_cuda, sizeof(PatchMatchParams)));
	CK(cudaMemcpy(a->params_cuda, &P, sizeof(PatchMatchParams), cudaMemcpyHostToDevice));

	DataPassHelper &hp = a->helper_host;
	memset(&hp, 0, sizeof(hp));
	hp.width = W; hp.height = H; hp.ref_index = 0;
	hp.texture_objects_cuda = a->texture_objects_cuda; hp.texture_depths_cuda = a->texture_depths_cuda;
	hp.cameras_cuda = a->cameras_cuda; hp.plane_hypotheses_cuda = a->plane_hypotheses_cuda;
	hp.rand_states_cuda = a->rand_states_cuda; hp.selected_views_cuda = a->selected_views_cuda;
	hp.neighbours_cuda = a->neighbours_cuda; hp.neighbours_map_cuda = a->neigbours_map_cuda;
	hp.weak_info_cuda = a->weak_info_cuda; hp.costs_cuda = a->costs_cuda; hp.params = a->params_cuda;
	hp.debug_point = make_int2(DEBUG_POINT_X, DEBUG_POINT_Y); hp.show_ncc_info = false;
	hp.fit_plane_hypotheses_cuda = a->fit_plane_hypotheses_cuda; hp.weak_reliable_cuda = a->weak_reliable_cuda;
	hp.view_weight_cuda = a->view_weight_cuda; hp.weak_nearest_strong = a->weak_nearest_strong;
	CK(cudaMalloc((void **)&a->helper_cuda, sizeof(DataPassHelper)));
	CK(cudaMemcpy(a->helper_cuda, &hp, sizeof(DataPassHelper), cudaMemcpyHostToDevice));
	c->uploaded = true;
	return 0;
}

// Snapshot the device state after these launch indices (0 = InitRandomStates ... see APD.cu:2409-2471).
int apdref_want_snapshots(void *h, const int *stages, int n) {
	ApdRefCtx *c = (ApdRefCtx *)h;
	c->want.assign(stages, stages + n);
	return 0;
}

// Calls the reference's own APD::RunPatchMatch(). State-restoring inputs are NOT re-uploaded:
// call once per handle when state != FIRST_INIT.
int apdref_run(void *h, int quiet) {
	ApdRefCtx *c = (ApdRefCtx *)h;
	if (!c->uploaded) { c->err = "upload first"; return -1; }
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyToSymbol(g_apdref_seed, &c->seed, sizeof(c->seed)));
	c->sync_index = 0; c->stage_ms.clear(); c->snaps.clear();
	std::streambuf *old = nullptr; std::ostringstream sink;
	if (quiet) old = std::cout.rdbuf(sink.rdbuf());
	g_ctx = c;
	c->t_prev = std::chrono::steady_clock::now();
	c->apd->RunPatchMatch();
	g_ctx = nullptr;
	if (quiet) std::cout.rdbuf(old);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return -2; }
	return 0;
}

int apdref_get_stage_ms(void *h, double *out, int cap) {
	ApdRefCtx *c = (ApdRefCtx *)h;
	int n = std::min<int>(cap, (int)c->stage_ms.size());
	for (int i = 0; i < n; ++i) out[i] = c->stage_ms[i];
	return (int)c->stage_ms.size();
}

// stage < 0: live device state (after the run: the final outputs). Any pointer may be NULL.
int apdref_get(void *h, int stage, float *planes, float *costs, unsigned *views, unsigned char *states,
               unsigned char *view_weights, void *rng48) {
	ApdRefCtx *c = (ApdRefCtx *)h;
	Snapshot live; const Snapshot *s = nullptr;
	if (stage < 0) { CK(cudaSetDevice(c->device)); take_snapshot(c, live); s = &live; }
	else {
		auto it = c->snaps.find(stage);
		if (it == c->snaps.end()) { c->err = "no snapshot for stage"; return -1; }
		s = &it->second;
	}
	const size_t n = (size_t)c->W * c->H;
	if (planes) memcpy(planes, s->planes.data(), n * 16);
	if (costs) memcpy(costs, s->costs.data(), n * 4);
	if (views) memcpy(views, s->views.data(), n * 4);
	if (states) memcpy(states, s->states.data(), n);
	if (view_weights) memcpy(view_weights, s->view_weights.data(), n * MAX_IMAGES);
	if (rng48) memcpy(rng48, s->rng.data(), n * 48);
	return 0;
}

// Outputs exactly as the reference hands them to main.cpp (APD.cu:2490-2492).
int apdref_get_outputs(void *h, float *planes, unsigned char *states, unsigned *views) {
	ApdRefCtx *c = (ApdRefCtx *)h; APD *a = c->apd;
	const size_t n = (size_t)c->W * c->H;
	if (planes) memcpy(planes, a->plane_hypotheses_host, n * 16);
	if (states) memcpy(states, c->weak_host.data(), n);
	if (views) memcpy(views, c->views_host.data(), n * 4);
	return 0;
}

// Deformable anchors of WEAK pixels (neighbours[weak_idx*9 + k]) and helper maps, for kernel-level diffs.
int apdref_get_anchors(void *h, short *neighbours, int *nmap, short *nearest, unsigned char *reliable, float *fit_planes) {
	ApdRefCtx *c = (ApdRefCtx *)h; APD *a = c->apd;
	const size_t n = (size_t)c->W * c->H;
	CK(cudaSetDevice(c->device));
	if (neighbours && a->weak_count > 0) CK(cudaMemcpy(neighbours, a->neighbours_cuda, (size_t)a->weak_count * NEIGHBOUR_NUM * sizeof(short2), cudaMemcpyDeviceToHost));
	if (nmap) memcpy(nmap, c->nmap_host.data(), n * sizeof(int));
	if (nearest) CK(cudaMemcpy(nearest, a->weak_nearest_strong, n * sizeof(short2), cudaMemcpyDeviceToHost));
	if (reliable) CK(cudaMemcpy(reliable, a->weak_reliable_cuda, n, cudaMemcpyDeviceToHost));
	if (fit_planes) CK(cudaMemcpy(fit_planes, a->fit_plane_hypotheses_cuda, n * 16, cudaMemcpyDeviceToHost));
	return a->weak_count;
}

const char *apdref_last_error(void *h) { return h ? ((ApdRefCtx *)h)->err.c_str() : "null handle"; }

void apdref_destroy(void *h) {
	ApdRefCtx *c = (ApdRefCtx *)h; if (!c) return;
	APD *a = c->apd;
	cudaSetDevice(c->device);
	if (c->uploaded) {
		for (int i = 0; i < c->N; ++i) { cudaDestroyTextureObject(a->texture_objects_host.images[i]); cudaFreeArray(a->cuArray[i]); }
		cudaFree(a->texture_objects_cuda);
		if (a->params_host.geom_consistency) {
			for (int i = 0; i < c->N; ++i) { cudaDestroyTextureObject(a->texture_depths_host.images[i]); cudaFreeArray(a->cuDepthArray[i]); }
			cudaFree(a->texture_depths_cuda);
		}
		cudaFree(a->cameras_cuda); cudaFree(a->plane_hypotheses_cuda); cudaFree(a->fit_plane_hypotheses_cuda);
		cudaFree(a->costs_cuda); cudaFree(a->rand_states_cuda); cudaFree(a->selected_views_cuda);
		cudaFree(a->params_cuda); cudaFree(a->helper_cuda); cudaFree(a->neighbours_cuda);
		cudaFree(a->neigbours_map_cuda); cudaFree(a->weak_info_cuda); cudaFree(a->weak_reliable_cuda);
		cudaFree(a->view_weight_cuda); cudaFree(a->weak_nearest_strong);
	}
	delete[] a->plane_hypotheses_host;
	delete a;
	delete c;
}

}  // extern "C"
