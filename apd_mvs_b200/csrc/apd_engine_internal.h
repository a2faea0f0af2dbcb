// The engine handle, shared by apd_engine.cu (the per-run C-ABI) and apd_scene.cu (the pass scheduler that feeds it
// from device-resident per-view caches). Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <string>
#include <vector>
#include "apd_device.cuh"

struct apd_engine {
	int device = 0, W = 0, H = 0, N = 0, S = 0;
	int capacity = 0;            // layers / scratch were allocated for this many images; N <= capacity
	size_t npx = 0;
	apd_params params;
	uint64_t seed = 0;
	cudaStream_t stream = nullptr;
	// asynchronous upload mode (apd_set_upload_mode): host copies go to copy_stream, the run waits for these events
	cudaStream_t copy_stream = nullptr; cudaEvent_t ev_early = nullptr, ev_late = nullptr;   // early: cameras + priors, late: images + depth maps
	bool async_upload = false, pending_early = false, pending_late = false;
	std::vector<apd_camera> cams_host;
	cudaArray_t img_arr = nullptr, depth_arr = nullptr;
	cudaTextureObject_t img_tex = 0, depth_tex = 0;
	float *ref_lin = nullptr, *ref_pad = nullptr;
	int ref_pitch = 0, ref_rows = 0;
	apd_camera *d_cams = nullptr; apd::ViewConst *d_views = nullptr; apd::RefConst *d_ref = nullptr; float *d_invw = nullptr;
	float4 *planes = nullptr, *fit_planes = nullptr, *prior_planes = nullptr;
	float *costs = nullptr;
	uint32_t *sel_views = nullptr, *prior_views = nullptr;
	uint8_t *states = nullptr, *prior_states = nullptr, *reliable = nullptr;
	uint2 *rng = nullptr; uint4 *view_w = nullptr;
	short2 *anchors = nullptr, *nearest = nullptr;
	float *scratch = nullptr; int *slab_slots = nullptr; int slab_stride = 0;
	int *wlist = nullptr, *wctrl = nullptr; int wlist_stride = 0;      // compacted WEAK-pixel lists (apd_kernels_weakq.cu)
	void *anchor_consts = nullptr;                                     // K3's rotation constants, one set per handle
	int num_sms = 0;
	int weak_impl = 1;                                                 // 1 = quad-per-pixel k_weak_q, 0 = first design (APD_WEAK_IMPL=old)
	int sweep_impl = 1;                                                // 1 = quad-per-pixel k_sweep_q, 0 = first design (APD_SWEEP_IMPL=old)
	CUtensorMap tmap_strong, tmap_sweep;
	bool have_images = false, have_cams = false, have_depths = false, have_planes = false, have_states = false;
	std::vector<cudaEvent_t> events;
	std::vector<float> stage_ms;
	int launches = 0, stages_run = 0;
	std::string err;
};
