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
// no change. Destroyed objects park their handle in a pool (apd_b200_facade.h): ProcessProblem creates one
// APD per (view, pass), the pool turns that into one allocation per image size.
#include "APD.h"
#include "apd_b200.h"
#include "apd_b200_facade.h"
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

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

// ---- engine-handle pool + object registry -------------------------------------------------------------------------
struct Pooled { apd_handle h; int device, W, H; };
struct Live { apd_handle h; const float4 *planes; int W, H; };
struct Pool {
	std::mutex mu;
	std::vector<Pooled> idle;                              // most recently parked last
	std::unordered_map<const APD *, Live> live;
	int created = 0, reused = 0;
	bool enabled = true;
	Pool() { if (const char *e = getenv("APD_B200_POOL")) enabled = strcmp(e, "0") != 0; }
	~Pool() { release(); }
	void release() {
		std::lock_guard<std::mutex> lk(mu);
		for (auto &p : idle) apd_destroy(p.h);
		idle.clear();
	}
	// a parked handle of this size that can hold `n` images, or nullptr
	apd_handle take(int device, int W, int H, int n) {
		std::lock_guard<std::mutex> lk(mu);
		for (size_t i = idle.size(); i-- > 0;) {
			Pooled p = idle[i];
			if (p.device != device || p.W != W || p.H != H) continue;
			idle.erase(idle.begin() + i);
			if (apd_get_capacity(p.h) >= n) { ++reused; return p.h; }
			apd_destroy(p.h);                                  // too few layers for this problem: replaced by a larger one
			break;
		}
		return nullptr;
	}
	void park(apd_handle h, int device, int W, int H) {
		if (!h) return;
		if (!enabled) { apd_destroy(h); return; }
		std::lock_guard<std::mutex> lk(mu);
		idle.push_back({h, device, W, H});
		// the multi-scale schedule works on one image size at a time (main.cpp:168-217); two sizes are kept so that a
		// change of round does not thrash, older ones are freed
		while (idle.size() > 2) { apd_destroy(idle.front().h); idle.erase(idle.begin()); }
	}
};
Pool &pool() { static Pool p; return p; }

// APD_B200_TIMING=1: wall time spent inside the facade's own methods, printed at exit next to the program's "Cost time"
// lines (main.cpp:135-137) - what is left of a ProcessProblem call is the reference's host code (file reading in
// InuputInitialization, per-pixel GetPlaneHypothesis, WriteBinMat, the visualisation images).
struct Timing {
	double init_ms = 0, run_ms = 0, dtor_ms = 0; int calls = 0; bool on = false;
	Timing() { const char *e = getenv("APD_B200_TIMING"); on = e && strcmp(e, "0") != 0; }
	~Timing() {
		if (on) std::cerr << "[apd_b200 facade] " << calls << " APD objects: CudaSpaceInitialization " << init_ms << " ms, RunPatchMatch "
		                  << run_ms << " ms, ~APD " << dtor_ms << " ms" << std::endl;
	}
};
Timing &timing() { static Timing t; return t; }
struct Scope {
	double &acc; std::chrono::steady_clock::time_point t0;
	explicit Scope(double &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
	~Scope() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace

APD::~APD() {
	Scope sc(timing().dtor_ms);
	apd_handle h = handle_of(helper_cuda);
	int device = 0, W = 0, H = 0;
	{
		Pool &P = pool();
		std::lock_guard<std::mutex> lk(P.mu);
		auto it = P.live.find(this);
		if (it != P.live.end()) { W = it->second.W; H = it->second.H; P.live.erase(it); }
	}
	delete[] plane_hypotheses_host;
	if (h) {
		cudaGetDevice(&device);
		if (W > 0) pool().park(h, device, W, H); else apd_destroy(h);
	}
}

void APD::CudaSpaceInitialization() {
	timing();                                     // constructed before the pool: destroyed (and printed) after it
	Scope sc(timing().init_ms);
	++timing().calls;
	int device = 0;
	cudaGetDevice(&device);                       // main() chose it with cudaSetDevice (main.cpp:153)
	const apd_params *prm = reinterpret_cast<const apd_params *>(&params_host);
	apd_handle h = pool().take(device, width, height, num_images);
	if (h) {   // a handle of this size from an earlier APD object: same device buffers, new problem
		check(apd_reset_inputs(h), h, "apd_reset_inputs");
		check(apd_set_num_images(h, num_images), h, "apd_set_num_images");
		check(apd_set_params(h, prm), h, "apd_set_params");
		check(apd_set_seed(h, pick_seed()), h, "apd_set_seed");
	} else {
		// problems of one scene have different numbers of source views (main.cpp:36-46): leave room for a few more
		// layers than this one needs so that the next problems of the round fit the same handle
		int cap = ((num_images + 3) / 4) * 4;
		if (cap > APD_MAX_IMAGES) cap = APD_MAX_IMAGES;
		if (!pool().enabled) cap = num_images;
		check(apd_create(&h, device, width, height, cap, prm, pick_seed()), h, "apd_create");
		check(apd_set_num_images(h, num_images), h, "apd_set_num_images");
		std::lock_guard<std::mutex> lk(pool().mu);
		++pool().created;
	}
	helper_cuda = reinterpret_cast<DataPassHelper *>(h);
	{
		std::lock_guard<std::mutex> lk(pool().mu);
		pool().live[this] = Live{h, plane_hypotheses_host, width, height};
	}
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
	Scope sc(timing().run_ms);
	apd_handle h = handle_of(helper_cuda);
	check(apd_run(h), h, "apd_run");
	// APD.cu:2490-2492
	check(apd_get_planes(h, reinterpret_cast<float *>(plane_hypotheses_host)), h, "apd_get_planes");
	check(apd_get_states(h, weak_info_host.ptr<uchar>(0)), h, "apd_get_states");
	check(apd_get_views(h, selected_views_host.ptr<unsigned int>(0)), h, "apd_get_views");
}

// ---- extras (apd_b200_facade.h) ---------------------------------------------------------------------------------
namespace apd_b200 {
static bool lookup(const APD &apd, Live &out) {
	Pool &P = pool();
	std::lock_guard<std::mutex> lk(P.mu);
	auto it = P.live.find(&apd);
	if (it == P.live.end()) return false;
	out = it->second;
	return true;
}
const float4 *GetPlaneHypotheses(const APD &apd) { Live l; return lookup(apd, l) ? l.planes : nullptr; }
int GetCosts(const APD &apd, float *out) { Live l; return lookup(apd, l) ? apd_get_costs(l.h, out) : APD_E_STATE; }
apd_handle HandleOf(const APD &apd) { Live l; return lookup(apd, l) ? l.h : nullptr; }
void ReleasePool() { pool().release(); }
void PoolStats(int *created, int *reused) {
	Pool &P = pool();
	std::lock_guard<std::mutex> lk(P.mu);
	if (created) *created = P.created;
	if (reused) *reused = P.reused;
}
}  // namespace apd_b200
