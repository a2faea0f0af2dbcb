// placeholder, replaced below
#include "apd_device.cuh"
namespace apd {
cudaError_t launch_nearest_strong(cudaStream_t, const Args &) { return cudaErrorNotSupported; }
cudaError_t launch_gen_anchors(cudaStream_t, const Args &) { return cudaErrorNotSupported; }
cudaError_t launch_demote_unreliable(cudaStream_t, const Args &) { return cudaErrorNotSupported; }
cudaError_t launch_fit_plane(cudaStream_t, const Args &) { return cudaErrorNotSupported; }
cudaError_t launch_weak(cudaStream_t, const Args &, int, int) { return cudaErrorNotSupported; }
}
