/*
 * apd_b200_facade.h — extras of the C++ facade (apd_mvs_b200/facade/APD_b200.cpp) next to the reference's
 * `class APD` (APD.h:67-145), which stays untouched: bulk accessors the reference does not have (SURVEY F1:
 * its public surface is the per-pixel `GetPlaneHypothesis(r, c)`, APD.h:73, and `costs_cuda` never leaves the
 * device, APD.cu:2490-2492) as free functions taking the object, and control of the facade's engine-handle pool.
 *
 * Handle pool. `ProcessProblem` (main.cpp:91-138) constructs and destroys one `APD` per (view, pass); in the
 * reference that is ~25 cudaMalloc + 2N cudaMallocArray and as many frees each time (APD.cpp:361-397, 585-671).
 * The facade keeps the engine handle of a destroyed `APD` in a small process-wide pool keyed by
 * (device, width, height) and hands it to the next `APD::CudaSpaceInitialization` of that size
 * (apd_reset_inputs + apd_set_params + apd_set_seed + apd_set_num_images instead of apd_create), so main.cpp runs
 * unchanged and pays the allocations once per image size. Set APD_B200_POOL=0 to switch the pool off.
 */
#ifndef APD_B200_FACADE_H
#define APD_B200_FACADE_H
#include <vector_types.h>
#include "apd_b200.h"

class APD;

namespace apd_b200 {
/* Host mirror of all plane hypotheses after RunPatchMatch: width*height float4 (world normal xyz, depth w), row-major,
 * i.e. what GetPlaneHypothesis(r, c) (APD.cpp:701-703) indexes. Valid until the object is destroyed. */
const float4 *GetPlaneHypotheses(const APD &apd);
/* Matching costs of the final planes (costs_cuda, APD.h:129): copies width*height floats from the device into `out`.
 * Returns 0 or a negative apd_status. */
int GetCosts(const APD &apd, float *out);
/* The engine handle behind an APD object (valid between CudaSpaceInitialization and destruction), or nullptr. */
apd_handle HandleOf(const APD &apd);
/* Destroys every pooled handle (also done at process exit). */
void ReleasePool();
/* Number of apd_create calls / pool hits so far (instrumentation for the integration test). */
void PoolStats(int *created, int *reused);
}  // namespace apd_b200
#endif /* APD_B200_FACADE_H */
