// sm_100a kernels of the adaptive patch deformation path (pixels whose state is WEAK):
//   k_row_nearest + k_nearest_strong   K2  FindNearestStrongPoint   APD.cu:2234-2270
//   k_gen_anchors                      K3  GenNeighbours            APD.cu:1750-1969
//   k_demote_unreliable                K4  NeigbourUpdate           APD.cu:1971-1987
//   k_fit_plane                        K8  RANSACToGetFitPlane      APD.cu:2272-2384
//   k_weak                             K9/K10 Black/RedPixelUpdateWeak APD.cu:1510-1545 -> :1323-1508
//                                          -> :892-980, deformable NCC :400-528
// Numerics follow the same contract as apd_device.cuh (explicit FMAs where the reference SASS has them).
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <curand_kernel.h>
#include "apd_device.cuh"

namespace apd {

struct AnchorConsts { float cos_a, sin_a, thresh; int shift_range; };
void launch_anchor_consts(cudaStream_t st, int rotate_time, AnchorConsts *out);   // apd_anchor_consts.cu (fast-math TU)

// ------------------------------------------------------------------------------------------------
// K2. The reference scans a 201x201 window per WEAK pixel (40 401 loads) for the nearest STRONG pixel,
// ties broken by scan order (dx ascending, then dy ascending, strict <). Exact two-pass equivalent:
// (1) per pixel, signed dx of the nearest STRONG pixel of its own row within |dx| <= 100, the
// negative one on ties; (2) per WEAK pixel, minimum over the 201 rows of (dx^2+dy^2, dx, dy).
constexpr int kNearR = 100;
__global__ void k_row_nearest(const uint8_t *states, int W, int H, int8_t *rowdx) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= W || y >= H) return;
	const uint8_t *row = states + (size_t)y * W;
	int8_t best = 127;
	for (int d = 0; d <= kNearR; ++d) {
		if (x - d >= 0 && row[x - d] == APD_STRONG) { best = (int8_t)(-d); break; }
		if (x + d < W && row[x + d] == APD_STRONG) { best = (int8_t)d; break; }
	}
	rowdx[(size_t)y * W + x] = best;
}
__global__ void k_nearest_strong(const Args a, const int8_t *rowdx) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= a.W || y >= a.H) return;
	const size_t c = (size_t)y * a.W + x;
	short2 out = make_short2(-1, -1);
	if (a.states[c] == APD_WEAK) {
		int best_d2 = 0x7fffffff, best_dx = 0, best_dy = 0;
		for (int dy = -kNearR; dy <= kNearR; ++dy) {
			const int yy = y + dy;
			if (yy < 0 || yy >= a.H) continue;
			const int dx = rowdx[(size_t)yy * a.W + x];
			if (dx == 127) continue;
			const int d2 = dx * dx + dy * dy;
			if (d2 < best_d2 || (d2 == best_d2 && dx < best_dx)) { best_d2 = d2; best_dx = dx; best_dy = dy; }
		}
		if (best_d2 != 0x7fffffff) out = make_short2((short)(x + best_dx), (short)(y + best_dy));
	}
	a.nearest[c] = out;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void normalize2(float &x, float &y) {   // NormalizeVec2, APD.cu:135-141
	const float r = rsqrtaf(fmaf(x, x, y * y));
	x *= r; y *= r;
}
__device__ __forceinline__ float3 point3(const RefConst &rc, int x, int y, float depth) {   // Get3DPoint, APD.cu:159-171
	float X0, X1; backproject(rc, (float)x, (float)y, depth, X0, X1);
	return make_float3(X0, X1, depth);
}
// PointinTriangle, APD.cu:91-112
__device__ __forceinline__ bool point_in_triangle(short2 A, short2 B, short2 C, int px, int py) {
	const float abx = (float)(B.x - A.x), aby = (float)(B.y - A.y);
	const float bcx = (float)(C.x - B.x), bcy = (float)(C.y - B.y);
	const float cax = (float)(A.x - C.x), cay = (float)(A.y - C.y);
	const float ab = sqrtaf(fmaf(abx, abx, aby * aby)), bc = sqrtaf(fmaf(bcx, bcx, bcy * bcy)), ca = sqrtaf(fmaf(cax, cax, cay * cay));
	if (ab <= 2.0f || bc <= 2.0f || ca <= 2.0f) return false;
	if (!(ab + bc > ca && bc + ca > ab && ab + ca > bc)) return false;
	const float pax = (float)(A.x - px), pay = (float)(A.y - py);
	const float pbx = (float)(B.x - px), pby = (float)(B.y - py);
	const float pcx = (float)(C.x - px), pcy = (float)(C.y - py);
	const float t1 = fmaf(pax, pby, -(pay * pbx));
	const float t2 = fmaf(pbx, pcy, -(pby * pcx));
	const float t3 = fmaf(pcx, pay, -(pcy * pax));
	return t1 * t2 >= 0.0f && t1 * t3 >= 0.0f;
}
// plane through three points, unit normal + offset (APD.cu:1897-1907, 2338-2349); false if degenerate
__device__ __forceinline__ bool plane_from_points(const float3 A, const float3 B, const float3 C, float4 &pl) {
	const float acx = A.x - C.x, acy = A.y - C.y, acz = A.z - C.z;
	const float bcx = B.x - C.x, bcy = B.y - C.y, bcz = B.z - C.z;
	float cx = fmaf(acy, bcz, -(bcy * acz));
	float cy = -fmaf(acx, bcz, -(bcx * acz));
	float cz = fmaf(acx, bcy, -(bcx * acy));
	if ((cx == 0.0f && cy == 0.0f && cz == 0.0f) || isnan(cx) || isnan(cy) || isnan(cz)) return false;
	normalize3(cx, cy, cz);
	const float d = fmaf(cz, A.z, fmaf(cx, A.x, cy * A.y));
	pl = make_float4(cx, cy, cz, -d);
	return true;
}
__device__ __forceinline__ float plane_dist(const float4 pl, const float3 p) {
	return fabsf(fmaf(pl.z, p.z, fmaf(pl.x, p.x, pl.y * p.y)) + pl.w);
}

// ------------------------------------------------------------------------------------------------
// K3. One thread per WEAK pixel of the compacted list (k_weak_lists<false>): deformable-anchor search along 8 base
// directions x rotate_time sub-rotations, then a 50-iteration RANSAC plane through the anchors' 3-D points.
// RANGE = the jitter range `shift_range` (APD.cu:1795-1796: 8 / 3 / 1 for rotate_time 1 / 2 / 4) as a compile-time
// constant, 0 = read it at run time. The kernel is ALU bound (ncu: 76 % of the ALU pipe, profiles/r02_*), so:
//  * `% shift_range` becomes a multiply-shift;
//  * with RANGE == 1 the jitter is always 0, the four tries of a radius test the SAME candidate against read-only
//    data, so it is tested once; the draws the reference consumes (4 on success at the first try, else 16) are only
//    COUNTED during the search - their values are never read - and the generator jumps ahead once, before the RANSAC.
// Advance the generator by n draws whose values nobody reads: cuRAND's own jump-ahead (precomputed matrices of the XORWOW
// recurrence for 4^k steps, curand_kernel.h `skipahead`) - a dozen 160x160-bit matrix-vector products for the ~18 000 draws
// a pixel's search consumes at 6221x4146, instead of 18 000 generator steps.
__device__ __forceinline__ void rng_skipahead(Rng &r, unsigned int n) {
	curandStateXORWOW_t st;
	st.v[0] = r.v0; st.v[1] = r.v1; st.v[2] = r.v2; st.v[3] = r.v3; st.v[4] = r.v4; st.d = r.d;
	skipahead((unsigned long long)n, &st);
	r.v0 = st.v[0]; r.v1 = st.v[1]; r.v2 = st.v[2]; r.v3 = st.v[3]; r.v4 = st.v[4]; r.d = st.d;
}

template <int RANGE>
__global__ void __launch_bounds__(128) k_gen_anchors(const Args a, const AnchorConsts *acp) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= a.wctrl[2]) return;
	const int W = a.W, H = a.H;
	const size_t n = (size_t)W * H;
	const int center = a.wlist[idx];
	const int py = center / W, px = center - py * W;
	const AnchorConsts ac = *acp;
	const RefConst rc = *a.ref;
	for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) a.anchors[(size_t)k * n + center] = make_short2(-1, -1);
	a.anchors[center] = make_short2((short)px, (short)py);
	Rng rng = rng_load(a.rng, center);
	short2 sp[32]; unsigned valid = 0u; int found = 0;
	unsigned int skipped = 0u;           // RANGE == 1: draws consumed by the search so far
	for (int i = 0; i < 32; ++i) sp[i] = make_short2(-1, -1);
	const float pxf = (float)px, pyf = (float)py, Wf = (float)W, Hf = (float)H;
	const unsigned range = RANGE ? (unsigned)RANGE : (unsigned)ac.shift_range;
	// A run-time +0.0f the compiler cannot prove to be zero. When the two loops below are unrolled, normalize2() of the
	// constant base directions would be folded at BUILD time - and a folded rsqrt.approx is the correctly rounded value,
	// whereas the MUFU.RSQ the reference executes is not (1 ulp apart for (+-1, +-1)): 227 of 9.75 M WEAK pixels of the
	// 6221x4146 case then take a different branch somewhere along their search (found with the full-size parity check).
	const float rt_zero = ac.thresh * 0.0f;
	int base = -1;
	for (int ox = -1; ox <= 1; ++ox) for (int oy = -1; oy <= 1; ++oy) {
		if (ox == 0 && oy == 0) continue;
		float dx = (float)ox + rt_zero, dy = (float)oy + rt_zero;
		normalize2(dx, dy);
		++base;
		for (int rot = 0; rot < a.rotate_time; ++rot) {
			const int di = base * 4 + rot;
			for (int radius = 2; radius <= APD_MAX_SEARCH_RADIUS; radius = min(radius * 2, radius + 25)) {
				const float rf = (float)radius;
				const float tx = fmaf(rf, dx, pxf), ty = fmaf(rf, dy, pyf);
				if (tx < 0.0f || ty < 0.0f || tx >= Wf || ty >= Hf) break;
				const int tries = (RANGE == 1) ? 1 : 4;
				for (int t = 0; t < tries; ++t) {
					// (curand()%2==0 ? 1 : -1) * curand() % shift_range, all in unsigned arithmetic (APD.cu:1813-1814)
					uint32_t xs = 0u, ys = 0u;
					if (RANGE != 1) {
						const uint32_t d1 = rng_next(rng), d2 = rng_next(rng), d3 = rng_next(rng), d4 = rng_next(rng);
						xs = (((d1 & 1u) == 0u) ? d2 : (0u - d2)) % range;
						ys = (((d3 & 1u) == 0u) ? d4 : (0u - d4)) % range;
					}
					bool hit = false;
					float ddx = fmaf(dx, 20.0f, (float)xs), ddy = fmaf(dy, 20.0f, (float)ys);
					normalize2(ddx, ddy);
					int nx = (short)(int)fmaf(rf, ddx, pxf), ny = (short)(int)fmaf(rf, ddy, pyf);
					if (!(nx < 6 || ny < 6 || nx >= W - 6 || ny >= H - 6)) {
						int nc = nx + ny * W;
						bool ok = true;
						if (a.states[nc] != APD_STRONG) {
							const short2 s = a.nearest[nc];
							if (s.x == -1 || s.y == -1) ok = false;
							nx = s.x; ny = s.y;
						}
						if (ok) {
							float tdx = (float)(nx - px), tdy = (float)(ny - py);
							normalize2(tdx, tdy);
							const float cosv = fmaf(tdx, dx, tdy * dy);
							hit = cosv > ac.thresh;
						}
					}
					if (RANGE == 1) skipped += hit ? 4u : 16u;      // first try succeeds, or all four fail alike
					if (hit) { sp[di] = make_short2((short)nx, (short)ny); valid |= 1u << di; ++found; break; }
				}
				if ((valid >> di) & 1u) break;
			}
			const float rx = fmaf(dx, ac.cos_a, -(dy * ac.sin_a)), ry = fmaf(dx, ac.sin_a, dy * ac.cos_a);
			dx = rx; dy = ry;
			normalize2(dx, dy);
		}
	}
	if (RANGE == 1) rng_skipahead(rng, skipped);
	if (found <= 3) { a.reliable[center] = 0; rng_store(a.rng, center, rng); return; }

	short2 pts[32]; float3 p3[32]; int vc = 0;
	const float3 c3 = point3(rc, px, py, a.planes[center].w);       // planes[].w still holds the prior depth here
	for (int i = 0; i < 32; ++i) {
		pts[i] = make_short2(-1, -1);
		if ((valid >> i) & 1u) {
			const short2 s = sp[i];
			pts[vc] = s;
			p3[vc] = point3(rc, s.x, s.y, a.planes[s.x + s.y * W].w);
			++vc;
		}
	}
	const float inv_dd = rcpf(a.depth_max - a.depth_min);
	float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
	int ua = -1, ub = -1, uc = -1, max_count = 3; bool has = false; float min_cost = FLT_MAX;
	for (int it = 0; it < 50; ++it) {
		const int ia = (int)(rng_next(rng) % (unsigned)vc), ib = (int)(rng_next(rng) % (unsigned)vc), ic = (int)(rng_next(rng) % (unsigned)vc);
		if (ia == ib || ib == ic || ia == ic) continue;
		if (!point_in_triangle(pts[ia], pts[ib], pts[ic], px, py)) continue;
		float4 pl;
		if (!plane_from_points(p3[ia], p3[ib], p3[ic], pl)) continue;
		int cnt = 0;
		for (int s = 0; s < vc; ++s) if (plane_dist(pl, p3[s]) * inv_dd < a.ransac_threshold) ++cnt;
		if (cnt < 6) continue;
		if (cnt > max_count) {
			max_count = cnt; min_cost = plane_dist(pl, c3); best = pl; has = true; ua = ia; ub = ib; uc = ic;
		} else if (cnt == max_count) {
			const float cd = plane_dist(pl, c3);
			if (cd < min_cost) { min_cost = cd; best = pl; ua = ia; ub = ib; uc = ic; }
		}
	}
	rng_store(a.rng, center, rng);
	if (!has) { a.reliable[center] = 0; return; }
	float wgt[32];
	for (int i = 0; i < vc; ++i) {
		float d = plane_dist(best, p3[i]);
		if (d * inv_dd >= a.ransac_threshold) { pts[i] = make_short2(-1, -1); wgt[i] = FLT_MAX; continue; }
		if (i == ua || i == ub || i == uc) d -= 1.0f;
		wgt[i] = d;
	}
	for (int i = 1; i < vc; ++i) {   // sort_small_weighted, APD.cu:14-27
		const short2 tp = pts[i]; const float tw = wgt[i]; int j = i;
		for (; j >= 1 && tw < wgt[j - 1]; --j) { pts[j] = pts[j - 1]; wgt[j] = wgt[j - 1]; }
		pts[j] = tp; wgt[j] = tw;
	}
	for (int k = 1; k < APD_NEIGHBOUR_NUM; ++k) a.anchors[(size_t)k * n + center] = pts[k - 1];
	a.reliable[center] = 1;
}

// K4
__global__ void k_demote_unreliable(const Args a) {
	const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
	if (px >= a.W || py >= a.H) return;
	const size_t c = (size_t)py * a.W + px;
	if (a.states[c] == APD_WEAK && a.reliable[c] != 1) a.states[c] = APD_UNKNOWN;
}

// ------------------------------------------------------------------------------------------------
// K8
__global__ void __launch_bounds__(128) k_fit_plane(const Args a) {
	const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
	if (px >= a.W || py >= a.H) return;
	const int W = a.W; const size_t n = (size_t)W * a.H;
	const int center = py * W + px;
	if (a.states[center] != APD_WEAK) { a.fit_planes[center] = a.planes[center]; return; }
	const RefConst rc = *a.ref;
	short2 pts[8]; float3 p3[8]; int cnt = 0;
	for (int k = 1; k < APD_NEIGHBOUR_NUM; ++k) {
		const short2 s = a.anchors[(size_t)k * n + center];
		if (s.x == -1 || s.y == -1) continue;
		const float depth = plane_depth(rc, a.planes[s.x + s.y * W], (float)s.x, (float)s.y);
		pts[cnt] = s; p3[cnt] = point3(rc, s.x, s.y, depth); ++cnt;
	}
	if (cnt < 3) { a.fit_planes[center] = a.planes[center]; return; }
	Rng rng = rng_load(a.rng, center);
	float min_cost = FLT_MAX; float4 best = make_float4(0.f, 0.f, 0.f, 0.f); bool has = false;
	for (int it = 0; it < 50; ++it) {
		const int ia = (int)(rng_next(rng) % (unsigned)cnt), ib = (int)(rng_next(rng) % (unsigned)cnt), ic = (int)(rng_next(rng) % (unsigned)cnt);
		if (ia == ib || ib == ic || ia == ic) continue;
		if (!point_in_triangle(pts[ia], pts[ib], pts[ic], px, py)) continue;
		float4 pl;
		if (!plane_from_points(p3[ia], p3[ib], p3[ic], pl)) continue;
		float cost = 0.0f;
		for (int s = 0; s < cnt; ++s) { if (s == ia || s == ib || s == ic) continue; cost += plane_dist(pl, p3[s]); }
		if (cost < min_cost) { min_cost = cost; best = pl; has = true; }
		if (min_cost == 0.0f) break;
	}
	rng_store(a.rng, center, rng);
	if (has) {
		const float xf = (float)px, yf = (float)py;
		const float depth = plane_depth(rc, a.planes[center], xf, yf);
		float X0, X1; backproject(rc, xf, yf, depth, X0, X1);
		float n2 = X1 * X1; n2 = fmaf(X0, X0, n2); n2 = fmaf(depth, depth, n2);
		const float rn = rcpf(sqrtaf(n2));
		float dot = (X1 * rn) * best.y; dot = fmaf(X0 * rn, best.x, dot); dot = fmaf(depth * rn, best.z, dot);
		if (dot > 0.0f) best = make_float4(-best.x, -best.y, -best.z, -best.w);
		a.fit_planes[center] = best;
	} else {
		a.fit_planes[center] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
}

// ------------------------------------------------------------------------------------------------
// ComputeBilateralNCCNew, APD.cu:400-528: patch k = 0 is the pixel's own 6x6 window, patches 1..8 are 3x3
// windows (offsets {-5,0,5}) centred on the deformable anchors, all warped by the same homography.
template <int INC>
__device__ __forceinline__ float ncc_window(const Args &a, int layer, const float *h, int cx, int cy, float inv_w) {
	NccSums t = {0.f, 0.f, 0.f, 0.f, 0.f};
	const float *base = a.ref_pad + (size_t)(cy + kRefPad) * a.ref_pitch + (cx + kRefPad);
#pragma unroll
	for (int i = -5; i <= 5; i += INC) {
		const float xf = (float)(cx + i);
		const float ax = h[0] * xf, ay = h[3] * xf, az = h[6] * xf;
		NccSums r = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = __ldg(base + (ptrdiff_t)j * a.ref_pitch + i);
			const float sp = src_tap(a.img_tex, layer, h, ax, ay, az, (float)(cy + j));
			r.r += rp; r.rr = fmaf(rp, rp, r.rr); r.rs = fmaf(rp, sp, r.rs);
			r.s += sp; r.ss = fmaf(sp, sp, r.ss);
		}
		t.r += r.r; t.rr += r.rr; t.s += r.s; t.ss += r.ss; t.rs += r.rs;
	}
	return ncc_cost(t, inv_w);
}

struct Anchors { short2 p[APD_NEIGHBOUR_NUM]; };

// The reference-image side of the nine windows of a WEAK pixel does not depend on the view or the plane: the 72 tap
// values of the eight anchor windows and every window's sum / sum of squares (accumulated in ncc_window's order) are
// gathered ONCE per pixel into a shared-memory column [kRefTaps + 18][kWeakNT] instead of 72 scattered global loads
// per (hypothesis, view). The own window stays in global memory: neighbouring lanes read neighbouring addresses,
// and shared memory spent on it would come out of the L1 that the scattered texture fetches live on.
constexpr int kRefTaps = 8 * 9;      // anchor windows only: the own 6x6 window is read coalesced from global (L1)
constexpr int kRefCacheRows = kRefTaps + 2 * APD_NEIGHBOUR_NUM;
constexpr int kWeakNT = 128;

template <int INC, bool STORE>
__device__ __forceinline__ void cache_window(const Args &a, int cx, int cy, float *col, float *sums) {
	const float *base = a.ref_pad + (size_t)(cy + kRefPad) * a.ref_pitch + (cx + kRefPad);
	float R = 0.f, RR = 0.f;
	int t = 0;
#pragma unroll
	for (int i = -5; i <= 5; i += INC) {
		float r = 0.f, rr = 0.f;
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = __ldg(base + (ptrdiff_t)j * a.ref_pitch + i);
			if (STORE) col[t * kWeakNT] = rp;
			++t;
			r += rp; rr = fmaf(rp, rp, rr);
		}
		R += r; RR += rr;
	}
	sums[0] = R; sums[kWeakNT] = RR;
}

template <int INC, bool CACHED>
__device__ __forceinline__ float ncc_window_cached(const Args &a, int layer, const float *h, int cx, int cy, float inv_w,
                                                   const float *col, const float *sums) {
	NccSums t = {sums[0], sums[kWeakNT], 0.f, 0.f, 0.f};
	const float *base = a.ref_pad + (size_t)(cy + kRefPad) * a.ref_pitch + (cx + kRefPad);
	int n = 0;
#pragma unroll
	for (int i = -5; i <= 5; i += INC) {
		const float xf = (float)(cx + i);
		const float ax = h[0] * xf, ay = h[3] * xf, az = h[6] * xf;
		float rs = 0.f, ss = 0.f, sm = 0.f;
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = CACHED ? col[n * kWeakNT] : __ldg(base + (ptrdiff_t)j * a.ref_pitch + i);
			++n;
			const float sp = src_tap(a.img_tex, layer, h, ax, ay, az, (float)(cy + j));
			rs = fmaf(rp, sp, rs);
			sm += sp; ss = fmaf(sp, sp, ss);
		}
		t.s += sm; t.ss += ss; t.rs += rs;
	}
	return ncc_cost(t, inv_w);
}

__device__ float ncc_deform(const Args &a, const RefConst &rc, const ViewConst &vc, int v, const float4 pl,
                            const Anchors &an, const float *rcol, int px, int py, float inv36, float inv9) {
	const Homog Hm = make_homography(rc, vc, pl);
	if (!centre_inside(Hm, vc, (float)px, (float)py)) return kCostMax;
	const float *h = Hm.h;
	const float Wf = (float)a.W, Hf = (float)a.H;
	float center_cost = 0.0f, strong_cost = 0.0f; int cnt = 0;
#pragma unroll 1
	for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) {
		const short2 q = an.p[k];
		if (q.x == -1 || q.y == -1) continue;
		const float xf = (float)q.x, yf = (float)q.y;
		const float rz = rcpf(h[8] + fmaf(h[6], xf, h[7] * yf));
		const float sx = (h[2] + fmaf(h[0], xf, h[1] * yf)) * rz;
		const float sy = (h[5] + fmaf(h[3], xf, h[4] * yf)) * rz;
		if (sx < 0.0f || sy < 0.0f || sx >= Wf || sy >= Hf) {
			if (k == 0) return kCostMax;
			if ((a.sel_views[q.x + q.y * a.W] >> v) & 1u) { strong_cost += kCostMax; ++cnt; }
			continue;
		}
		const float *sums = rcol + (kRefTaps + 2 * k) * kWeakNT;
		if (k == 0) center_cost = ncc_window_cached<2, false>(a, v + 1, h, q.x, q.y, inv36, rcol, sums);
		else { strong_cost += ncc_window_cached<5, true>(a, v + 1, h, q.x, q.y, inv9, rcol + 9 * (k - 1) * kWeakNT, sums); ++cnt; }
	}
	if (cnt == 0) return center_cost;
	strong_cost = strong_cost * rcpf((float)cnt);
	strong_cost = (strong_cost > kCostMax) ? kCostMax : strong_cost;      // OpenCV MIN(strong_cost, cost_max)
	return (float)fma((double)center_cost, 0.25, (double)strong_cost * 0.75);
}

// weighted cost of one plane for a WEAK pixel over the sampled views:
//   sum_v w_v * (ncc_deform + geom_factor * geom) (APD.cu:918-927, 960-969, 1464-1471)
__device__ float weak_cost(const Args &a, const RefConst &rc, const ViewConst *sv, const float4 pl, const Anchors &an, const float *rcol,
                           const VW &vw, int px, int py, float inv36, float inv9, float inv_wn, float limit) {
	// `limit`: the caller adopts the plane only if the final weighted cost is < limit. Costs, the geometric term
	// (with geom_factor >= 0) and weights are non-negative and partial sums are rounded monotonically, so once a
	// partial sum reaches the limit the remaining views cannot change the outcome and are skipped (limit = +inf
	// disables this).
	const bool prune = a.geom_factor >= 0.0f;
	float acc = 0.0f;
	for (int v = 0; v < a.S; ++v) {
		const int w = vw_get(vw, v);
		if (w == 0) continue;
		float c = ncc_deform(a, rc, sv[v], v, pl, an, rcol, px, py, inv36, inv9);
		if (a.geom) c = fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, pl, (float)px, (float)py), c);
		acc = fmaf((float)w, c, acc);
		if (prune && acc * inv_wn >= limit) break;
	}
	return acc;
}

constexpr int kWeakTW = 32, kWeakTH = 8;      // 128 pixels of one colour per block

__global__ void __launch_bounds__(kWeakNT, 3) k_weak(const Args a, const int iter, const int color) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RefConst *sr = reinterpret_cast<RefConst *>(smem_raw);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	float *rcache = reinterpret_cast<float *>(sv + a.S);      // [kRefCacheRows][NT] reference taps + window sums
	int *s_slot = reinterpret_cast<int *>(rcache + kRefCacheRows * kWeakNT);
	const int tid = threadIdx.x;
	slab_acquire(a, s_slot, tid);
	{
		const int nv = a.S * (int)(sizeof(ViewConst) / 4);
		const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views); uint32_t *s = reinterpret_cast<uint32_t *>(sv);
		for (int i = tid; i < nv; i += kWeakNT) s[i] = g[i];
		const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.ref); uint32_t *srr = reinterpret_cast<uint32_t *>(sr);
		for (int i = tid; i < (int)(sizeof(RefConst) / 4); i += kWeakNT) srr[i] = gr[i];
	}
	__syncthreads();
	// [9*S][NT] cost matrix + probabilities of this block, in its slab of the L1/L2-resident pool
	float *cm = slab_ptr(a, s_slot);
	const int lx = tid & 31, row = tid >> 5;                   // 4 row pairs of 32 columns
	const int px = blockIdx.x * kWeakTW + lx;
	const int ybase = blockIdx.y * kWeakTH + 2 * row;
	const int py = ybase + ((px + ybase + color) & 1);
	if (px >= a.W || py >= a.H || py >= a.half_rows) { slab_exit(a, s_slot, kWeakNT); return; }
	const int W = a.W, S = a.S; const size_t n = (size_t)W * a.H;
	const int center = py * W + px;
	if (a.states[center] != APD_WEAK) { slab_exit(a, s_slot, kWeakNT); return; }
	const RefConst &rc = *sr;
	const float xf = (float)px, yf = (float)py;
	const float inv36 = a.inv_w[0], inv9 = a.inv_w[1];
	float *cmt = cm + tid;
#define CM(k, v) cmt[((k) * S + (v)) * kWeakNT]
#define PROB(v) cmt[(8 * S + (v)) * kWeakNT]
	Anchors an;
	for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) an.p[k] = a.anchors[(size_t)k * n + center];
	float *rcol = rcache + tid;
	cache_window<2, false>(a, an.p[0].x, an.p[0].y, rcol, rcol + kRefTaps * kWeakNT);
#pragma unroll 1
	for (int k = 1; k < APD_NEIGHBOUR_NUM; ++k) {
		const short2 q = an.p[k];
		if (q.x == -1 || q.y == -1) continue;
		cache_window<5, true>(a, q.x, q.y, rcol + 9 * (k - 1) * kWeakNT, rcol + (kRefTaps + 2 * k) * kWeakNT);
	}
	const bool own_window = an.p[0].x == px && an.p[0].y == py;    // always, K3 puts the pixel itself in slot 0

	// candidates = current planes of the anchors that are (still) STRONG (APD.cu:1352-1363)
	unsigned flags = 0u; int pos[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const short2 q = an.p[k + 1];
		pos[k] = 0;
		if (!(q.x == -1 || q.y == -1) && a.states[q.x + q.y * W] == APD_STRONG) { flags |= 1u << k; pos[k] = q.x + q.y * W; }
	}
	// view-major order: the eight candidate planes are neighbouring surfaces, so their footprints in ONE source view
	// overlap in L1; plane-major order walked through all views between two visits of the same texels
#pragma unroll 1
	for (int v = 0; v < S; ++v) {
#pragma unroll 1
		for (int k = 0; k < 8; ++k) {
			if ((flags >> k) & 1u) CM(k, v) = ncc_deform(a, rc, sv[v], v, a.planes[pos[k]], an, rcol, px, py, inv36, inv9);
			else CM(k, v) = (k == 0 && v == 0) ? 2.0f : 0.0f;                            // `= {2.0f}` quirk, APD.cu:1345
		}
	}
	// view selection (APD.cu:1365-1434): priors from every existing anchor, STRONG or not
	const float thr = 0.8 * __expf((float)(unsigned)(iter * iter) * -0.011111111380159854889f);
	const float thr_fallback = __expf((thr * thr) * -3.125f);
	float prob_sum = 0.0f;
	for (int v = 0; v < S; ++v) {
		float prior = 0.0f;
		for (int k = 1; k < APD_NEIGHBOUR_NUM; ++k) {
			const short2 q = an.p[k];
			if (q.x == -1 || q.y == -1) continue;
			prior += ((a.sel_views[q.x + q.y * W] >> v) & 1u) ? 0.9f : 0.1f;
		}
		float count = 0.0f, tmpw = 0.0f; int count_false = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const float c = CM(k, v);
			if (c < thr) { tmpw += __expf((c * c) * -5.5555553436279296875f); count += 1.0f; }
			if (c > 1.2f) count_false++;
		}
		float p = 0.0f;
		if (count > 2.0f && count_false < 3) p = tmpw * rcpf(count);
		else if (count_false < 3) p = thr_fallback;
		p = p * prior;
		PROB(v) = p; prob_sum += p;
	}
	Rng rng = rng_load(a.rng, center);
	VW vw; vw.lo = 0ull; vw.hi = 0ull;
	{
		const float inv = rcpf(prob_sum); float cum = 0.0f;
		for (int v = 0; v < S; ++v) { cum = fmaf(inv, PROB(v), cum); PROB(v) = cum; }
		for (int s = 0; s < 15; ++s) {
			const float r = rng_uniform(rng) - 1.1920928955078125e-07f;
			for (int v = 0; v < S; ++v) if (PROB(v) > r) { vw_add(vw, v); break; }
		}
	}
	uint32_t temp_sel = 0u; float weight_norm = 0.0f;
	for (int v = 0; v < S; ++v) { const int w = vw_get(vw, v); if (w > 0) { temp_sel |= 1u << v; weight_norm += (float)w; } }
	const float inv_wn = rcpf(weight_norm);

	float best_cost; int best_k;
	{
		float fc[8];
		const float miss = a.geom_factor * 3.0f;
#pragma unroll 1
		for (int k = 0; k < 8; ++k) {
			float acc = 0.0f;
			const bool fl = (flags >> k) & 1u;
			float4 pl = make_float4(0.f, 0.f, 0.f, 1.f);
			if (fl) pl = a.planes[pos[k]];
			for (int v = 0; v < S; ++v) {
				const int w = vw_get(vw, v);
				if (w == 0) continue;
				float c = CM(k, v);
				if (a.geom) c = fl ? fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, pl, xf, yf), c) : fmaf(a.geom_factor, 3.0f, c);
				acc = fmaf((float)w, c, acc);
			}
			(void)miss;
			fc[k] = acc * inv_wn;
		}
		best_cost = fc[0]; best_k = 0;
#pragma unroll
		for (int k = 1; k < 8; ++k) if (fc[k] <= best_cost) { best_cost = fc[k]; best_k = k; }
	}

	// One loop (and one inlined copy of the deformable NCC: the instruction cache is a measured limiter of this kernel)
	// over the pixel's plane evaluations: h = -1 the current plane, h = 0 the fit plane, h = 1..5 the five hypotheses of
	// PlaneHypothesisRefinementWeak (APD.cu:892-980), which returns before any draw when the fit plane is all-zero.
	float4 pl_now = a.planes[center];
	float cost_now = 0.0f, cost_stored = 0.0f, depth_now = 0.0f;
	const float4 fit = a.fit_planes[center];
	const bool have_fit = !(fit.x == 0.0f && fit.y == 0.0f && fit.z == 0.0f);
	float depth_rand = 0.0f, depth_pert = 0.0f, d0 = 0.0f;
	float4 n_rand = pl_now, n_pert = pl_now, n0 = pl_now;
#pragma unroll 1
	for (int h = -1; h < (have_fit ? 6 : 0); ++h) {
		float4 t = pl_now; float d = 0.0f; bool in_range = true;
		if (h == 0) { t = fit; d = plane_depth(rc, t, xf, yf); in_range = d >= a.depth_min && d <= a.depth_max; }
		else if (h > 0) {
			if (h == 1) {
				depth_rand = fmaf(rng_uniform(rng), a.depth_max - a.depth_min, a.depth_min);
				n_rand = random_normal(rc, xf, yf, rng, depth_now);
				const float lo = depth_now * (1.0f - 0.02f);
				const float span = fmaf(depth_now, 1.0f + 0.02f, -lo);
				depth_pert = fmaf(span, rng_uniform(rng), lo);
				n_pert = perturbed_normal(rc, xf, yf, pl_now, rng);
				n0 = pl_now; d0 = depth_now;
			}
			const int i = h - 1;
			const float di = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : d0);
			t = (i == 1 || i == 2) ? n_rand : (i == 3 ? n_pert : n0);
			t.w = plane_offset(rc, xf, yf, di, t.x, t.y, t.z);
			d = plane_depth(rc, t, xf, yf);
			in_range = d >= a.depth_min && d <= a.depth_max;
		}
		if (!in_range) continue;                 // an out-of-range plane is never adopted: not evaluated
		const float tc = weak_cost(a, rc, sv, t, an, rcol, vw, px, py, inv36, inv9, inv_wn, h < 0 ? __int_as_float(0x7f800000) : cost_now) * inv_wn;
		if (h >= 0) {
			if (tc < cost_now) { depth_now = d; pl_now = t; cost_now = tc; }
		} else {
			cost_now = tc; cost_stored = tc;
			depth_now = plane_depth(rc, pl_now, xf, yf);
			if ((flags >> best_k) & 1u) {
				int bp = pos[0];
#pragma unroll
				for (int k = 1; k < 8; ++k) if (best_k == k) bp = pos[k];
				const float4 cand = a.planes[bp];
				const float dc = plane_depth(rc, cand, xf, yf);
				if (dc >= a.depth_min && dc <= a.depth_max && best_cost < cost_now) {
					depth_now = dc; pl_now = cand; cost_now = best_cost; a.sel_views[center] = temp_sel;
				}
			}
		}
	}
	rng_store(a.rng, center, rng);
	vw_store(a.view_w, center, vw);
	float4 final_plane = a.planes[center];
	if (a.state == APD_REFINE_INIT) {
		if ((double)cost_now < (double)cost_stored - 0.1) { final_plane = pl_now; a.planes[center] = pl_now; }
	} else { final_plane = pl_now; a.planes[center] = pl_now; }
	{   // "update cost with old method" (APD.cu:1499-1507): plain NCC of the committed plane
		float acc = 0.0f;
		for (int v = 0; v < S; ++v) {
			const int w = vw_get(vw, v);
			if (w == 0) continue;
			const Homog Hm = make_homography(rc, sv[v], final_plane);
			float c = kCostMax;
			if (centre_inside(Hm, sv[v], xf, yf))
				c = own_window ? ncc_window_cached<2, false>(a, v + 1, Hm.h, px, py, inv36, rcol, rcol + kRefTaps * kWeakNT)
				               : ncc_window<2>(a, v + 1, Hm.h, px, py, inv36);
			acc = fmaf((float)w, c, acc);
		}
		a.costs[center] = acc * inv_wn;
	}
	slab_exit(a, s_slot, kWeakNT);
#undef CM
#undef PROB
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_nearest_strong(cudaStream_t st, const Args &a) {
	// row scratch: reuse the fit-plane buffer (16 B/px, rewritten by K8 before any use)
	int8_t *rowdx = reinterpret_cast<int8_t *>(a.fit_planes);
	dim3 b(32, 8), g((a.W + 31) / 32, (a.H + 7) / 8);
	k_row_nearest<<<g, b, 0, st>>>(a.states, a.W, a.H, rowdx);
	k_nearest_strong<<<g, b, 0, st>>>(a, rowdx);
	return cudaGetLastError();
}
size_t anchor_consts_bytes() { return sizeof(AnchorConsts); }
// `consts`: the handle's own AnchorConsts buffer (no process-global state: handles on one device may run concurrently)
cudaError_t launch_weak_lists(cudaStream_t st, const Args &a, bool split);       // apd_kernels_weakq.cu
cudaError_t launch_gen_anchors(cudaStream_t st, const Args &a, void *consts) {
	AnchorConsts *ac = static_cast<AnchorConsts *>(consts);
	launch_anchor_consts(st, a.rotate_time, ac);
	// the list of all WEAK pixels (K3 is a full launch in the reference, APD.cu:2415); the grid covers the worst case,
	// threads beyond the list's length leave at once
	cudaError_t e = launch_weak_lists(st, a, false);
	if (e != cudaSuccess) return e;
	const size_t n = (size_t)a.W * a.H;
	const unsigned g = (unsigned)((n + 127) / 128);
	// shift_range = max((int)(tan(angle / 2) * 20), 1) with angle = 45 deg / rotate_time (APD.cu:1790-1796): 8, 3, 2, 1
	// APD_K3_GENERIC=1: run-time jitter range and four tries per radius whatever rotate_time is (diagnostic)
	static const bool generic = getenv("APD_K3_GENERIC") != nullptr;
	if (generic) k_gen_anchors<0><<<g, 128, 0, st>>>(a, ac);
	else if (a.rotate_time == 4) k_gen_anchors<1><<<g, 128, 0, st>>>(a, ac);
	else if (a.rotate_time == 2) k_gen_anchors<3><<<g, 128, 0, st>>>(a, ac);
	else if (a.rotate_time == 1) k_gen_anchors<8><<<g, 128, 0, st>>>(a, ac);
	else k_gen_anchors<0><<<g, 128, 0, st>>>(a, ac);
	return cudaGetLastError();
}
cudaError_t launch_demote_unreliable(cudaStream_t st, const Args &a) {
	dim3 b(32, 8), g((a.W + 31) / 32, (a.H + 7) / 8);
	k_demote_unreliable<<<g, b, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_fit_plane(cudaStream_t st, const Args &a) {
	dim3 b(32, 4), g((a.W + 31) / 32, (a.H + 3) / 4);
	k_fit_plane<<<g, b, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_weak(cudaStream_t st, const Args &a, int iter, int color) {
	const size_t smem = sizeof(RefConst) + (size_t)a.S * sizeof(ViewConst) + (size_t)kRefCacheRows * kWeakNT * 4 + 16;
	if (smem > 227 * 1024) return cudaErrorInvalidValue;
	cudaFuncSetAttribute(k_weak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	dim3 g((a.W + kWeakTW - 1) / kWeakTW, (a.H + kWeakTH - 1) / kWeakTH);
	k_weak<<<g, kWeakNT, smem, st>>>(a, iter, color);
	return cudaGetLastError();
}

}  // namespace apd
