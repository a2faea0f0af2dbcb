// Rotation constants of the deformable-anchor search (APD.cu:1790-1795). They are double-precision
// libdevice cos/sin/tan results rounded to float, so this one tiny kernel is compiled with the
// reference's own math flags (--use_fast_math, FMA contraction on) instead of the engine's
// contraction-free flags: same library code, same flags, same bits.
#include <cuda_runtime.h>
namespace apd {
struct AnchorConsts { float cos_a, sin_a, thresh; int shift_range; };
#define APD_PI 3.14159265358979323846          /* APD.h:7 */
__global__ void k_anchor_consts(int rotate_time, AnchorConsts *out) {
	const float angle = 45.0f / rotate_time;
	AnchorConsts c;
	c.cos_a = cos(angle * APD_PI / 180.f);
	c.sin_a = sin(angle * APD_PI / 180.f);
	c.thresh = cos((angle / 2.0f) * APD_PI / 180.0f);
	const int s = (int)(tan((angle / 2.0f) * APD_PI / 180.0f) * 20);
	c.shift_range = s < 1 ? 1 : s;
	*out = c;
}
void launch_anchor_consts(cudaStream_t st, int rotate_time, AnchorConsts *out) { k_anchor_consts<<<1, 1, 0, st>>>(rotate_time, out); }
}  // namespace apd
