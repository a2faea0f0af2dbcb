// Drop-in replacement for the device half of the reference's `class APD` (APD.h:67-145).
//
// Build it INSTEAD of APD.cu, next to the reference's unchanged main.cpp / APD.h / main.h and an
// APD.cpp whose device-side member definitions (the destructor APD.cpp:361-397,
// CudaSpaceInitialization :585-671, SetDataPassHelperInCuda :673-699) are compiled out with
// `#ifndef USE_APD_B200`. The host half — InuputInitialization (file reading,
// APD.cpp:399-583) and the getters (:701-727) — stays the reference's own code and keeps filling /
// reading the same private members; this file only moves the data across the C-ABI.
//
// The engine handle is kept in the (otherwise unused) `helper_cuda` pointer member so that APD.h needs
// no change.
#include "APD.h"
#include "apd_b200.h"
#include <chrono>
#include <cstdlib>

static_assert(sizeof(apd_params) == sizeof(PatchMatchParams), "PatchMatchParams layout");
static_assert(sizeof(apd_camera) == sizeof(Camera), "Camera layout");

namespace {
inline apd_handle handle_of(DataPassHelper *p) { return reinterpret_cast<apd_handle>(p); }
void check(int rc, apd_handle h, const char *what) {
	if (rc != APD_OK) {   // reference behaviour on CUDA errors: message + exit (APD.cpp:315-323)
		std::cerr << what << " failed (" << rc << "): " << apd_last_error(h) << std::endl;
		exit(EXIT_FAILURE);
	}
}
uint64_t pick_seed() {
	if (const char *s = getenv("APD_SEED")) return strtoull(s, nullptr, 10);
	return (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();   // reference: clock64(), APD.cu:803
}
}  // namespace

APD::~APD() {
	delete[] plane_hypotheses_host;
	apd_destroy(handle_of(helper_cuda));
}

void APD::CudaSpaceInitialization() {
	int device = 0;
	cudaGetDevice(&device);                       // main() chose it with cudaSetDevice (main.cpp:153)
	apd_handle h = nullptr;
	check(apd_create(&h, device, width, height, num_images, reinterpret_cast<const apd_params *>(&params_host), pick_seed()), h, "apd_create");
	helper_cuda = reinterpret_cast<DataPassHelper *>(h);
	check(apd_set_cameras(h, reinterpret_cast<const apd_camera *>(cameras.data())), h, "apd_set_cameras");
	std::vector<const float *> ptrs(num_images);
	for (int i = 0; i < num_images; ++i) ptrs[i] = images[i].ptr<float>(0);
	check(apd_set_images(h, ptrs.data(), (size_t)images[0].step), h, "apd_set_images");
	if (params_host.geom_consistency) {
		for (int i = 0; i < num_images; ++i) ptrs[i] = depths[i].ptr<float>(0);
		check(apd_set_depths(h, ptrs.data(), (size_t)depths[0].step), h, "apd_set_depths");
	}
	const bool has_planes = params_host.state != FIRST_INIT;
	if (has_planes || params_host.use_APD)
		check(apd_set_priors(h, has_planes ? reinterpret_cast<const float *>(plane_hypotheses_host) : nullptr,
		                     has_planes ? selected_views_host.ptr<unsigned int>(0) : nullptr,
		                     params_host.use_APD ? weak_info_host.ptr<uchar>(0) : nullptr), h, "apd_set_priors");
}

void APD::SetDataPassHelperInCuda() {}            // kernel arguments travel by value; nothing to upload

void APD::RunPatchMatch() {
	apd_handle h = handle_of(helper_cuda);
	check(apd_run(h), h, "apd_run");
	// APD.cu:2490-2492
	check(apd_get_planes(h, reinterpret_cast<float *>(plane_hypotheses_host)), h, "apd_get_planes");
	check(apd_get_states(h, weak_info_host.ptr<uchar>(0)), h, "apd_get_states");
	check(apd_get_views(h, selected_views_host.ptr<unsigned int>(0)), h, "apd_get_views");
}
