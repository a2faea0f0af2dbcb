// sm_100a kernels of the WEAK-pixel propagation, second design (round 2):
//   k_weak_lists   compacted lists of the WEAK pixels (one per checkerboard colour, or one for K3)
//   k_weak_q       K9/K10 Black/RedPixelUpdateWeak  APD.cu:1510-1545 -> CheckerboardPropagationWeak :1323-1508
//                  -> PlaneHypothesisRefinementWeak :892-980, deformable NCC ComputeBilateralNCCNew :400-528
//
// Why a second design. The first k_weak (apd_kernels_weak.cu, kept selectable for A/B timing) maps one thread to one
// pixel, like the reference. Its fetches are then the slow case of the texture unit measured in tools/tex_probe3.cu:
// the four lanes of a texture quad belong to four different pixels with four different anchor sets, so one TEX
// instruction touches four unrelated places (ncu, profiles/r01w_ncu_summary.json: 2.13 data-pipe wavefronts per quad
// request, texture data pipe 77 % busy at 36 % of the fetch peak, 26 % at full resolution).
//
// Here ONE PIXEL IS OWNED BY ONE TEXTURE QUAD and the four lanes evaluate FOUR PLANE HYPOTHESES of that pixel against the
// same source view in lock step: same reference taps, same anchor windows, only the plane differs. The hypotheses of
// one pixel are planes of anchors that RANSAC found to lie on one surface (or small perturbations of the current plane),
// so the four bilinear footprints of a TEX instruction fall within a texel or two of each other: the fast case.
//   * cost matrix (8 candidates x S views): lane l evaluates candidates l and l+4, interleaved window by window so the
//     second pass over a window hits the lines the first one brought into L1;
//   * current plane + fit plane, then the five refinement hypotheses: the same evaluation, lanes = hypotheses, each quad
//     walking its own list of sampled views; a quad is a small state machine (matrix -> current/fit -> refinement), all
//     quads of a warp run the ONE inlined copy of the deformable NCC convergently whatever their phase;
//   * the reference side of all nine windows (taps and their sums) is gathered once per pixel into shared memory and
//     broadcast to the quad; the 9xS cost matrix sits in shared memory too (8 pixels per warp instead of 32);
//   * warps are persistent and independent (no block barrier in the loop): they pull chunks of eight WEAK pixels from a
//     compacted list, so a warp is always full whatever the shape of the WEAK regions.
// Every arithmetic expression is the one of the first design (bit-identical to the reference); only who evaluates what,
// and when, changed. Exact skips kept: hypotheses out of the depth range are not evaluated, a hypothesis stops as soon as
// its partial weighted sum can no longer beat the current cost.
#include <cfloat>
#include <cstdlib>
#include "apd_device.cuh"
#include "apd_pair.cuh"

namespace apd {

constexpr int kWqNT = 128;                 // 4 independent warps
constexpr int kWqPix = 8;                  // pixels per warp = texture quads per warp
constexpr int kOwnTaps = 36, kAncTaps = 8 * 9;
constexpr int kRowSum = kOwnTaps + kAncTaps;                       // rows [kRowSum + 2k, +1] = (sum, sum of squares) of window k
constexpr int kRefRows = kRowSum + 2 * APD_NEIGHBOUR_NUM;          // 126 floats per pixel
__host__ __device__ constexpr int wq_warp_words(int S) { return (kRefRows + 9 * S + APD_NEIGHBOUR_NUM + 8 + 8) * kWqPix; }

// ------------------------------------------------------------------------------------------------
// Compacted WEAK lists. One warp owns an 8x8 pixel area, ordered as four 4x4 sub-areas so that eight consecutive list
// entries (= the eight pixels a k_weak_q warp processes together) are a compact cluster; a block owns a 32x8 tile.
// Blocks are numbered super-tile by super-tile (256x256 pixels = 8x32 tiles), not in raster order: the list comes out in
// (roughly) that order, so the ~1800 warps of k_weak_q that work on neighbouring list entries at the same time fetch
// their far-away anchor windows from a few hundred pixels around one place instead of from a band across the whole image.
// (Measured at 6221x4146, nine source images = 0.9 GB against 126 MB of L2: no effect on k_weak_q, 124 vs 126 ms - its
// full-resolution penalty of 1.3x per pixel against 1555x1036 is not an L2-locality problem; kept because it is free.)
// SPLIT: one list per colour, restricted to the rows the reference's half launch reaches; otherwise one list of all
// WEAK pixels (K3 is a full launch). The order of the entries does not affect any result.
constexpr int kSuperX = 8, kSuperY = 32;      // tiles per super-tile
template <bool SPLIT>
__global__ void __launch_bounds__(128) k_weak_lists(const Args a, int tiles_x, int tiles_y, int supers_x) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int st = blockIdx.x / (kSuperX * kSuperY), within = blockIdx.x % (kSuperX * kSuperY);
	const int tx = (st % supers_x) * kSuperX + within % kSuperX, ty = (st / supers_x) * kSuperY + within / kSuperX;
	if (tx >= tiles_x || ty >= tiles_y) return;
	if (!SPLIT) {
		// K3's list: a warp takes two rows of 32 pixels of the tile, so 32 consecutive entries are (mostly) neighbours in a
		// row - its threads walk rays through `states` / `nearest`, which are row-major byte / short2 maps
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			const int x = tx * 32 + lane, y = ty * 8 + warp * 2 + c;
			bool weak = x < a.W && y < a.H;
			if (weak) weak = a.states[(size_t)y * a.W + x] == APD_WEAK;
			const unsigned b = __ballot_sync(0xffffffffu, weak);
			if (b == 0u) continue;
			int base = 0;
			if (lane == 0) base = atomicAdd(&a.wctrl[2], __popc(b));
			base = __shfl_sync(0xffffffffu, base, 0);
			if (weak) a.wlist[base + __popc(b & ((1u << lane) - 1u))] = y * a.W + x;
		}
		return;
	}
	const int ax = tx * 32 + warp * 8, ay = ty * 8;
	const int sub = lane >> 3, j = lane & 7;
	const int y = ay + (sub >> 1) * 4 + (j >> 1);
	const int xb = ax + (sub & 1) * 4 + (j & 1) * 2;
	const int odd = (xb + y) & 1;
#pragma unroll
	for (int c = 0; c < 2; ++c) {
		const int x = xb + (c ^ odd);                     // colour c: (x + y + c) even
		bool weak = x < a.W && y < a.H && y < a.half_rows;
		if (weak) weak = a.states[(size_t)y * a.W + x] == APD_WEAK;
		const unsigned b = __ballot_sync(0xffffffffu, weak);
		if (b == 0u) continue;
		int base = 0;
		if (lane == 0) base = atomicAdd(&a.wctrl[c], __popc(b));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (weak) a.wlist[c * a.wlist_stride + base + __popc(b & ((1u << lane) - 1u))] = y * a.W + x;
	}
}

// per-pixel shared-memory columns of the owning warp (stride kWqPix between rows)
struct WqPixel {
	const float *refc;        // [kRefRows]
	const short2 *anc;        // [9] anchors, slot 0 = the pixel itself
	const uint32_t *asel;     // [8] selected-view bitmasks of anchors 1..8
};

// ComputeBilateralNCCNew (APD.cu:400-528) for the two slots of this lane: plane P0 against source view v0, P1 against v1.
__device__ __forceinline__ void wq_deform_pair(const Args &a, const RefConst &rc, const ViewConst *sv, int v0, int v1, const float4 P0, const float4 P1,
                                               bool on0, bool on1, const WqPixel &px, float inv36, float inv9, float &out0, float &out1) {
	const ViewConst &vc0 = sv[v0], &vc1 = sv[v1];
	const Homog2 H = make_homography2(rc, vc0, vc1, P0, P1);
	const short2 self = px.anc[0];
	bool live0, live1;
	{
		float x0, y0, x1, y1;
		project2(H, (float)self.x, (float)self.y, x0, y0, x1, y1);
		live0 = on0 && !(x0 >= vc0.wf || x0 < 0.0f || y0 >= vc0.hf || y0 < 0.0f);           // APD.cu:424-435
		live1 = on1 && !(x1 >= vc1.wf || x1 < 0.0f || y1 >= vc1.hf || y1 < 0.0f);
	}
	float cc0 = 0.f, sc0 = 0.f, cc1 = 0.f, sc1 = 0.f; int n0 = 0, n1 = 0;
	const float Wf = (float)a.W, Hf = (float)a.H;
#pragma unroll 1
	for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) {
		const short2 q = px.anc[k * kWqPix];
		const bool valid = !(q.x == -1 || q.y == -1);
		bool w0 = false, w1 = false;
		if (__any_sync(0xffffffffu, valid && (live0 || live1))) {
			float sx0, sy0, sx1, sy1;
			project2(H, (float)q.x, (float)q.y, sx0, sy0, sx1, sy1);
			const uint32_t sel = (k == 0) ? 0u : px.asel[(k - 1) * kWqPix];
			if (valid && live0) {
				if (sx0 < 0.0f || sy0 < 0.0f || sx0 >= Wf || sy0 >= Hf) {
					if (k == 0) live0 = false;                                                    // APD.cu:437-438
					else if ((sel >> v0) & 1u) { sc0 += kCostMax; ++n0; }                          // APD.cu:439-446
				} else w0 = true;
			}
			if (valid && live1) {
				if (sx1 < 0.0f || sy1 < 0.0f || sx1 >= Wf || sy1 >= Hf) {
					if (k == 0) live1 = false;
					else if ((sel >> v1) & 1u) { sc1 += kCostMax; ++n1; }
				} else w1 = true;
			}
		}
		if (!__any_sync(0xffffffffu, w0 || w1)) continue;          // warp-uniform
		float c0 = 0.f, c1 = 0.f;
		if (k == 0)
			wq_window2<2, kWqPix>(a.img_tex, v0 + 1, v1 + 1, H, w0, w1, q.x, q.y, inv36, px.refc, px.refc[kRowSum * kWqPix], px.refc[(kRowSum + 1) * kWqPix], c0, c1);
		else
			wq_window2<5, kWqPix>(a.img_tex, v0 + 1, v1 + 1, H, w0, w1, q.x, q.y, inv9, px.refc + (kOwnTaps + 9 * (k - 1)) * kWqPix,
			              px.refc[(kRowSum + 2 * k) * kWqPix], px.refc[(kRowSum + 2 * k + 1) * kWqPix], c0, c1);
		if (w0) { if (k == 0) cc0 = c0; else { sc0 += c0; ++n0; } }
		if (w1) { if (k == 0) cc1 = c1; else { sc1 += c1; ++n1; } }
	}
	// APD.cu:513-527
	float r0 = kCostMax, r1 = kCostMax;
	if (live0) {
		if (n0 == 0) r0 = cc0;
		else { float s = sc0 * rcpf((float)n0); s = (s > kCostMax) ? kCostMax : s; r0 = (float)fma((double)cc0, 0.25, (double)s * 0.75); }
	}
	if (live1) {
		if (n1 == 0) r1 = cc1;
		else { float s = sc1 * rcpf((float)n1); s = (s > kCostMax) ? kCostMax : s; r1 = (float)fma((double)cc1, 0.25, (double)s * 0.75); }
	}
	out0 = r0; out1 = r1;
}

// the five hypotheses of PlaneHypothesisRefinementWeak in the reference's order (APD.cu:947-979)
struct Refine5 { float depth_rand, depth_pert, d0; float4 n_rand, n_pert, n0; };
__device__ __forceinline__ void wq_hypothesis(const Args &a, const RefConst &rc, const Refine5 &r, int i, float xf, float yf, float4 &t, float &d, bool &in_range) {
	const float di = (i == 0 || i == 2) ? r.depth_rand : (i == 4 ? r.depth_pert : r.d0);
	t = (i == 1 || i == 2) ? r.n_rand : (i == 3 ? r.n_pert : r.n0);
	t.w = plane_offset(rc, xf, yf, di, t.x, t.y, t.z);
	d = plane_depth(rc, t, xf, yf);
	in_range = d >= a.depth_min && d <= a.depth_max;
}

// MINB = resident blocks per SM the register allocation is sized for (4: 128 registers, 3: 168)
template <int MINB>
__global__ void __launch_bounds__(kWqNT, MINB) k_weak_q(const Args a, const int iter, const int color, int *work) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RefConst *sr = reinterpret_cast<RefConst *>(smem_raw);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int S = a.S, W = a.W;
	float *wbase = reinterpret_cast<float *>(sv + S) + (size_t)warp * wq_warp_words(S);
	{
		const int nv = S * (int)(sizeof(ViewConst) / 4);
		const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views); uint32_t *s = reinterpret_cast<uint32_t *>(sv);
#pragma unroll 1
		for (int i = tid; i < nv; i += kWqNT) s[i] = g[i];
		const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.ref); uint32_t *srr = reinterpret_cast<uint32_t *>(sr);
		for (int i = tid; i < (int)(sizeof(RefConst) / 4); i += kWqNT) srr[i] = gr[i];
	}
	__syncthreads();
	const int ql = lane & 3, pq = lane >> 2;
	const unsigned qmask = 0xFu << (lane & ~3);
	// this pixel's columns in the warp's shared memory
	float *refc = wbase + pq;                                              // [kRefRows][8]
	float *cm = wbase + kRefRows * kWqPix + pq;                            // [9*S][8]: 8xS cost matrix + S probabilities
	short2 *anc = reinterpret_cast<short2 *>(wbase + (kRefRows + 9 * S) * kWqPix) + pq;                           // [9][8]
	uint32_t *asel = reinterpret_cast<uint32_t *>(wbase + (kRefRows + 9 * S + APD_NEIGHBOUR_NUM) * kWqPix) + pq;   // [8][8]
	uint32_t *astrong = asel + 8 * kWqPix;                                 // [1][8] bit k-1: anchor k is STRONG
	const RefConst &rc = *sr;
	const int count = a.wctrl[color];
	const int *list = a.wlist + (size_t)color * a.wlist_stride;
	const size_t n = (size_t)W * a.H;
	const float inv36 = a.inv_w[0], inv9 = a.inv_w[1];
	const bool prune = a.geom_factor >= 0.0f;
	const float kInf = __int_as_float(0x7f800000);
	WqPixel wp; wp.refc = refc; wp.anc = anc; wp.asel = asel;
#define CM(k, v) cm[((k) * S + (v)) * kWqPix]
#define PROB(v) cm[(8 * S + (v)) * kWqPix]

#pragma unroll 1
	for (;;) {
		int chunk = 0;
		if (lane == 0) chunk = atomicAdd(work, 1);
		chunk = __shfl_sync(0xffffffffu, chunk, 0);
		if (chunk * kWqPix >= count) break;
		// a quad without a pixel (tail of the list) repeats the last pixel and writes nothing
		const bool alive = chunk * kWqPix + pq < count;
		const int center = list[min(chunk * kWqPix + pq, count - 1)];
		const int py = center / W, px = center - py * W;
		const float xf = (float)px, yf = (float)py;
		__syncwarp();                                    // the previous chunk's shared-memory reads are over
		// ---- anchors, their selected views and states (APD.cu:1352-1363, 1370-1380)
		if (ql == 0) astrong[0] = 0u;
		__syncwarp();
		for (int k = ql; k < APD_NEIGHBOUR_NUM; k += 4) {
			const short2 q = a.anchors[(size_t)k * n + center];
			anc[k * kWqPix] = q;
			if (k >= 1) {
				uint32_t s = 0u;
				if (!(q.x == -1 || q.y == -1)) {
					const int qc = q.x + q.y * W;
					s = a.sel_views[qc];
					if (a.states[qc] == APD_STRONG) atomicOr(astrong, 1u << (k - 1));
				}
				asel[(k - 1) * kWqPix] = s;
			}
		}
		__syncwarp();
		// ---- reference side of the nine windows: lane 0 the own 6x6 window, lanes 1..3 the anchor windows
		if (ql == 0) wq_cache_window<2, kWqPix>(a, px, py, refc, refc + kRowSum * kWqPix);
		else {
#pragma unroll 1
			for (int k = ql; k < APD_NEIGHBOUR_NUM; k += 3) {
				const short2 q = anc[k * kWqPix];
				if (q.x == -1 || q.y == -1) continue;
				wq_cache_window<5, kWqPix>(a, q.x, q.y, refc + (kOwnTaps + 9 * (k - 1)) * kWqPix, refc + (kRowSum + 2 * k) * kWqPix);
			}
		}
		__syncwarp();
		// candidates = current planes of the anchors that are (still) STRONG
		const unsigned flags = astrong[0] & 0xFFu;
		auto cand_pos = [&](int k) { const short2 q = anc[(k + 1) * kWqPix]; return q.x + q.y * W; };

		// ---- quad state machine; every step each lane carries two evaluations ("slots") through wq_deform_pair.
		//  phase 0  cost matrix: lane l = candidates l (slot 0) and l+4 (slot 1) against view v, one view per step;
		//  phase 1  current plane (even lanes) and fit plane (odd lanes): lane pair l>>1 and the slot pick one of the next four
		//           sampled views, so a step covers four views of both planes; sums are formed in view order afterwards;
		//  phase 2  the five refinement hypotheses: every lane is a worker that walks the sampled views of ITS hypothesis,
		//           two per step (slots), and stops when the partial sum can no longer beat the current cost; hypothesis 2
		//           waits for the first lane that becomes free;   phase 3  done.
		int phase = 0;
		uint32_t m = (S >= 32) ? 0xffffffffu : ((1u << S) - 1u);      // phases 0/1: views still to do (quad-uniform)
		float4 T = make_float4(0.f, 0.f, 1.f, 1.f), Tb = T;
		const bool cand0 = (flags >> ql) & 1u, cand1 = (flags >> (ql + 4)) & 1u;
		if (cand0) T = a.planes[cand_pos(ql)];
		if (cand1) Tb = a.planes[cand_pos(ql + 4)];
		float limit = kInf;
		Rng rng; rng.v0 = rng.v1 = rng.v2 = rng.v3 = rng.v4 = rng.d = 0u;
		VW vw; vw.lo = 0ull; vw.hi = 0ull;
		uint32_t temp_sel = 0u; float inv_wn = 0.f, best_cost = 0.f; int best_k = 0;
		float4 pl_now = make_float4(0.f, 0.f, 1.f, 1.f);
		float cost_now = 0.f, cost_stored = 0.f, depth_now = 0.f, d_fit = 0.f;
		bool have_fit = false, fit_ok = false, sel_write = false;
		float acc_cur = 0.f, acc_fit = 0.f; bool pr_cur = false, pr_fit = false;                 // phase 1 (replicated in the quad)
		uint32_t my_m = 0u; float my_acc = 0.f, res_first = 0.f; bool my_busy = false, second = false, pending2 = false;   // phase 2 worker
		int owner2 = -1; unsigned q_busy = 0u;
		Refine5 rf; rf.depth_rand = rf.depth_pert = rf.d0 = 0.f; rf.n_rand = rf.n_pert = rf.n0 = pl_now;

#pragma unroll 1
		for (int step = 0; step < 4 * S + 16; ++step) {      // S matrix steps + at most S/4 + 2S + a few (never reached)
			// ---- phase transitions (quad-uniform; quads of a warp may be in different phases)
#pragma unroll 1
			while (phase < 3 && ((phase < 2) ? (m == 0u) : (q_busy == 0u && !pending2))) {
				if (phase == 0) {
					__syncwarp(qmask);                                     // the quad's matrix entries are visible
					// view selection (APD.cu:1365-1434): priors from every existing anchor, STRONG or not
					const float thr = 0.8 * __expf((float)(unsigned)(iter * iter) * -0.011111111380159854889f);
					const float thr_fallback = __expf((thr * thr) * -3.125f);
					float prob_sum = 0.0f;
#pragma unroll 1
					for (int v = 0; v < S; ++v) {
						float prior = 0.0f;
						for (int k = 1; k < APD_NEIGHBOUR_NUM; ++k) {
							const short2 q = anc[k * kWqPix];
							if (q.x == -1 || q.y == -1) continue;
							prior += ((asel[(k - 1) * kWqPix] >> v) & 1u) ? 0.9f : 0.1f;
						}
						float cnt = 0.0f, tmpw = 0.0f; int count_false = 0;
#pragma unroll
						for (int k = 0; k < 8; ++k) {
							const float c = CM(k, v);
							if (c < thr) { tmpw += __expf((c * c) * -5.5555553436279296875f); cnt += 1.0f; }
							if (c > 1.2f) count_false++;
						}
						float p = 0.0f;
						if (cnt > 2.0f && count_false < 3) p = tmpw * rcpf(cnt);
						else if (count_false < 3) p = thr_fallback;
						p = p * prior;
						if (ql == 0) PROB(v) = p;
						prob_sum += p;
					}
					__syncwarp(qmask);
					rng = rng_load(a.rng, center);
					{
						const float inv = rcpf(prob_sum); float cum = 0.0f;
						// the four lanes hold identical copies of the CDF; lane 0 keeps it in shared memory
#pragma unroll 1
						for (int v = 0; v < S; ++v) { cum = fmaf(inv, PROB(v), cum); __syncwarp(qmask); if (ql == 0) PROB(v) = cum; __syncwarp(qmask); }
#pragma unroll 1
						for (int s = 0; s < 15; ++s) {
							const float r = rng_uniform(rng) - 1.1920928955078125e-07f;
#pragma unroll 1
							for (int v = 0; v < S; ++v) if (PROB(v) > r) { vw_add(vw, v); break; }
						}
					}
					float weight_norm = 0.0f;
#pragma unroll 1
					for (int v = 0; v < S; ++v) { const int w = vw_get(vw, v); if (w > 0) { temp_sel |= 1u << v; weight_norm += (float)w; } }
					inv_wn = rcpf(weight_norm);
					// final costs of the eight candidates (APD.cu:1436-1452): lane l computes candidates l and l+4
					float fcr0 = 0.f, fcr1 = 0.f;
					{
						float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
						for (int v = 0; v < S; ++v) {
							const int w = vw_get(vw, v);
							if (w == 0) continue;
							float c0 = CM(ql, v), c1 = CM(ql + 4, v);
							if (a.geom) {
								c0 = cand0 ? fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, T, xf, yf), c0) : fmaf(a.geom_factor, 3.0f, c0);
								c1 = cand1 ? fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, Tb, xf, yf), c1) : fmaf(a.geom_factor, 3.0f, c1);
							}
							s0 = fmaf((float)w, c0, s0); s1 = fmaf((float)w, c1, s1);
						}
						fcr0 = s0 * inv_wn; fcr1 = s1 * inv_wn;
					}
					best_cost = __shfl_sync(qmask, fcr0, 0, 4); best_k = 0;             // FindMinCostIndex: `<=`, last minimum wins
#pragma unroll
					for (int k = 1; k < 8; ++k) {
						const float f = __shfl_sync(qmask, (k < 4) ? fcr0 : fcr1, k & 3, 4);
						if (f <= best_cost) { best_cost = f; best_k = k; }
					}
					// phase 1: the current plane and the fit plane (PlaneHypothesisRefinementWeak returns before any draw
					// when the fit plane is all-zero, APD.cu:912-914)
					pl_now = a.planes[center];
					const float4 fit = a.fit_planes[center];
					have_fit = !(fit.x == 0.0f && fit.y == 0.0f && fit.z == 0.0f);
					d_fit = plane_depth(rc, fit, xf, yf);
					fit_ok = have_fit && d_fit >= a.depth_min && d_fit <= a.depth_max;
					T = (ql & 1) ? fit : pl_now;
					acc_cur = acc_fit = 0.f; pr_cur = pr_fit = false; limit = kInf;
					m = temp_sel; phase = 1;
				} else if (phase == 1) {
					const float tc_cur = acc_cur * inv_wn, tc_fit = acc_fit * inv_wn;
					cost_now = tc_cur; cost_stored = tc_cur;
					depth_now = plane_depth(rc, pl_now, xf, yf);
					if ((flags >> best_k) & 1u) {                                       // APD.cu:1474-1486
						const float4 cand = a.planes[cand_pos(best_k)];
						const float dc = plane_depth(rc, cand, xf, yf);
						if (dc >= a.depth_min && dc <= a.depth_max && best_cost < cost_now) {
							depth_now = dc; pl_now = cand; cost_now = best_cost; sel_write = true;
						}
					}
					if (have_fit) {
						if (fit_ok && tc_fit < cost_now) { depth_now = d_fit; pl_now = a.fit_planes[center]; cost_now = tc_fit; }   // APD.cu:916-935
						rf.depth_rand = fmaf(rng_uniform(rng), a.depth_max - a.depth_min, a.depth_min);
						rf.n_rand = random_normal(rc, xf, yf, rng, depth_now);
						const float lo = depth_now * (1.0f - 0.02f);
						const float span = fmaf(depth_now, 1.0f + 0.02f, -lo);
						rf.depth_pert = fmaf(span, rng_uniform(rng), lo);
						rf.n_pert = perturbed_normal(rc, xf, yf, pl_now, rng);
						rf.n0 = pl_now; rf.d0 = depth_now;
						// lanes 0..3 start with hypotheses 3, 4 (the two near the current plane), 0, 1; hypothesis 2 is pending
						float d; bool ok, ok2; float4 t2;
						wq_hypothesis(a, rc, rf, (ql + 3) % 5, xf, yf, T, d, ok);
						wq_hypothesis(a, rc, rf, 2, xf, yf, t2, d, ok2);
						my_m = temp_sel; my_acc = 0.f; my_busy = ok && temp_sel != 0u;        // out of range: never evaluated
						pending2 = ok2 && temp_sel != 0u; owner2 = -1; second = false; res_first = 0.f;
						limit = cost_now;
						q_busy = __ballot_sync(qmask, my_busy) & qmask;
						phase = 2;
					} else phase = 3;
				} else {   // phase == 2: adopt in the reference's order, strict <
					const float mine = second ? res_first : my_acc;
#pragma unroll 1
					for (int i = 0; i < 5; ++i) {
						float val;
						if (i == 2) val = __shfl_sync(qmask, my_acc, owner2 < 0 ? 0 : owner2, 4);
						else val = __shfl_sync(qmask, mine, (i + 2) % 5, 4);               // lane that started with hypothesis i
						const float tci = val * inv_wn;
						float4 t; float d; bool ok;
						wq_hypothesis(a, rc, rf, i, xf, yf, t, d, ok);
						if (ok && tci < cost_now) { depth_now = d; pl_now = t; cost_now = tci; }
					}
					phase = 3;
				}
			}
			if (!__any_sync(0xffffffffu, phase < 3)) break;
			// ---- hypothesis 2 takes the first free lane of its quad
			if (phase == 2 && pending2) {
				const unsigned free_lanes = (~q_busy) & qmask;
				if (free_lanes != 0u) {
					owner2 = (__ffs(free_lanes) - 1) & 3;
					if (ql == owner2) {
						float d; bool ok;
						res_first = my_acc; second = true;
						wq_hypothesis(a, rc, rf, 2, xf, yf, T, d, ok);
						my_m = temp_sel; my_acc = 0.f; my_busy = true;
					}
					pending2 = false;
					q_busy |= 1u << ((lane & ~3) + owner2);
				}
			}
			// ---- this step's two slots
			int v0 = 0, v1 = 0; bool s0 = false, s1 = false;
			float4 P1 = T;
			if (phase == 0) {
				v0 = v1 = __ffs(m) - 1; m &= m - 1u;
				s0 = cand0; s1 = cand1; P1 = Tb;
			} else if (phase == 1) {
				uint32_t mm = m;
				if (ql >> 1) mm &= mm - 1u;
				const bool mine = ((ql & 1) == 0) ? !pr_cur : (fit_ok && !pr_fit);
				if (mm) { v0 = __ffs(mm) - 1; s0 = mine; }
				mm &= mm - 1u; mm &= mm - 1u;
				if (mm) { v1 = __ffs(mm) - 1; s1 = mine; }
			} else if (phase == 2 && my_busy) {
				v0 = __ffs(my_m) - 1; my_m &= my_m - 1u; s0 = true;
				if (my_m) { v1 = __ffs(my_m) - 1; my_m &= my_m - 1u; s1 = true; }
			}
			__syncwarp();
			float c0, c1;
			wq_deform_pair(a, rc, sv, v0, v1, T, P1, s0, s1, wp, inv36, inv9, c0, c1);
			if (phase == 0) {
				// a missing candidate keeps the reference's partially initialised row: `cost_array[8][32] = {2.0f}` sets
				// [0][0] only (APD.cu:1345)
				CM(ql, v0) = cand0 ? c0 : ((ql == 0 && v0 == 0) ? 2.0f : 0.0f);
				CM(ql + 4, v0) = cand1 ? c1 : 0.0f;
			} else if (phase == 1) {
				if (a.geom) {      // both slots hold the same plane: its world point is computed once (geom_point)
					const GeomPoint gp = geom_point(rc, T, xf, yf);
					if (s0) c0 = fmaf(a.geom_factor, geom_cost_at(a, rc, sv[v0], v0 + 1, gp, xf, yf), c0);
					if (s1) c1 = fmaf(a.geom_factor, geom_cost_at(a, rc, sv[v1], v1 + 1, gp, xf, yf), c1);
				}
				// the four views of this step in ascending order: position p sits in slot p>>1 of lane pair p&1
#pragma unroll
				for (int p = 0; p < 4; ++p) {
					const float cc = __shfl_sync(qmask, (p >> 1) ? c1 : c0, 2 * (p & 1), 4);
					const float cf = __shfl_sync(qmask, (p >> 1) ? c1 : c0, 2 * (p & 1) + 1, 4);
					if (m != 0u) {
						const int v = __ffs(m) - 1; m &= m - 1u;
						const float w = (float)vw_get(vw, v);
						if (!pr_cur) { acc_cur = fmaf(w, cc, acc_cur); if (prune && acc_cur * inv_wn >= limit) pr_cur = true; }
						if (fit_ok && !pr_fit) { acc_fit = fmaf(w, cf, acc_fit); if (prune && acc_fit * inv_wn >= limit) pr_fit = true; }
					}
				}
			} else if (phase == 2) {
				GeomPoint gp; gp.x = gp.y = gp.z = 0.f;
				if (a.geom && s0) gp = geom_point(rc, T, xf, yf);
				if (s0) {
					if (a.geom) c0 = fmaf(a.geom_factor, geom_cost_at(a, rc, sv[v0], v0 + 1, gp, xf, yf), c0);
					my_acc = fmaf((float)vw_get(vw, v0), c0, my_acc);
					if (prune && my_acc * inv_wn >= limit) my_busy = false;
				}
				if (s1 && my_busy) {
					if (a.geom) c1 = fmaf(a.geom_factor, geom_cost_at(a, rc, sv[v1], v1 + 1, gp, xf, yf), c1);
					my_acc = fmaf((float)vw_get(vw, v1), c1, my_acc);
					if (prune && my_acc * inv_wn >= limit) my_busy = false;
				}
				if (my_busy && my_m == 0u) my_busy = false;
				q_busy = __ballot_sync(qmask, my_busy) & qmask;
			}
		}

		// ---- commit (APD.cu:1488-1507)
		float4 final_plane = a.planes[center];
		const bool adopt = (a.state == APD_REFINE_INIT) ? ((double)cost_now < (double)cost_stored - 0.1) : true;
		if (adopt) final_plane = pl_now;
		__syncwarp();
		if (alive && ql == 0) {
			rng_store(a.rng, center, rng);
			vw_store(a.view_w, center, vw);
			if (sel_write) a.sel_views[center] = temp_sel;
			if (adopt) a.planes[center] = pl_now;
		}
		// "update cost with old method": plain NCC of the committed plane over the sampled views; the lanes and their two
		// slots take eight views at a time, the sum is formed in view order
		float facc = 0.0f;
		uint32_t fm = temp_sel;
#pragma unroll 1
		while (__any_sync(0xffffffffu, fm != 0u)) {
			uint32_t mm = fm;
			for (int t = 0; t < ql; ++t) mm &= mm - 1u;
			const int va = mm ? (__ffs(mm) - 1) : -1;
			mm &= mm - 1u; mm &= mm - 1u; mm &= mm - 1u; mm &= mm - 1u;
			const int vb = mm ? (__ffs(mm) - 1) : -1;
			const int ia = va < 0 ? 0 : va, ib = vb < 0 ? 0 : vb;
			const Homog2 H = make_homography2(rc, sv[ia], sv[ib], final_plane, final_plane);
			float x0, y0, x1, y1;
			project2(H, xf, yf, x0, y0, x1, y1);
			const bool wa = va >= 0 && !(x0 >= sv[ia].wf || x0 < 0.0f || y0 >= sv[ia].hf || y0 < 0.0f);
			const bool wb = vb >= 0 && !(x1 >= sv[ib].wf || x1 < 0.0f || y1 >= sv[ib].hf || y1 < 0.0f);
			float ca = kCostMax, cb = kCostMax;
			if (__any_sync(0xffffffffu, wa || wb)) {
				float ta = 0.f, tb = 0.f;
				wq_window2<2, kWqPix>(a.img_tex, ia + 1, ib + 1, H, wa, wb, px, py, inv36, refc, refc[kRowSum * kWqPix], refc[(kRowSum + 1) * kWqPix], ta, tb);
				if (wa) ca = ta;
				if (wb) cb = tb;
			}
#pragma unroll
			for (int t = 0; t < 8; ++t) {
				const float ct = __shfl_sync(0xffffffffu, (t < 4) ? ca : cb, t & 3, 4);
				const int vt = __shfl_sync(0xffffffffu, (t < 4) ? va : vb, t & 3, 4);
				if (vt >= 0) facc = fmaf((float)vw_get(vw, vt), ct, facc);
			}
#pragma unroll
			for (int t = 0; t < 8; ++t) fm &= fm - 1u;
		}
		if (alive && ql == 0) a.costs[center] = facc * inv_wn;
	}
#undef CM
#undef PROB
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_weak_lists(cudaStream_t st, const Args &a, bool split) {
	const int tiles_x = (a.W + 31) / 32, tiles_y = (a.H + 7) / 8;
	const int supers_x = (tiles_x + kSuperX - 1) / kSuperX, supers_y = (tiles_y + kSuperY - 1) / kSuperY;
	const unsigned g = (unsigned)(supers_x * supers_y * kSuperX * kSuperY);
	if (split) {
		cudaError_t e = cudaMemsetAsync(a.wctrl, 0, 2 * sizeof(int), st);
		if (e != cudaSuccess) return e;
		k_weak_lists<true><<<g, 128, 0, st>>>(a, tiles_x, tiles_y, supers_x);
	} else {
		cudaError_t e = cudaMemsetAsync(a.wctrl + 2, 0, sizeof(int), st);
		if (e != cudaSuccess) return e;
		k_weak_lists<false><<<g, 128, 0, st>>>(a, tiles_x, tiles_y, supers_x);
	}
	return cudaGetLastError();
}

template <int MINB>
static cudaError_t launch_weak_q_t(cudaStream_t st, const Args &a, int iter, int color, int work_slot, int num_sms, size_t smem) {
	cudaError_t e = cudaFuncSetAttribute(k_weak_q<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_weak_q<MINB>, kWqNT, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) per_sm = 1;
	k_weak_q<MINB><<<num_sms * per_sm, kWqNT, smem, st>>>(a, iter, color, a.wctrl + kWorkBase + work_slot);
	return cudaGetLastError();
}
cudaError_t launch_weak_q(cudaStream_t st, const Args &a, int iter, int color, int work_slot, int num_sms) {
	const size_t smem = sizeof(RefConst) + (size_t)a.S * sizeof(ViewConst) + (size_t)(kWqNT / 32) * wq_warp_words(a.S) * 4;
	if (smem > 227 * 1024) return cudaErrorInvalidValue;
	static const int minb = [] { const char *e = getenv("APD_WQ_BLOCKS"); return e ? atoi(e) : 3; }();   // measured: 168 registers without spills (3 blocks) beat 128 with spills (4): 5.8 vs 6.8 ms at 1555x1036
	return minb == 3 ? launch_weak_q_t<3>(st, a, iter, color, work_slot, num_sms, smem) : launch_weak_q_t<4>(st, a, iter, color, work_slot, num_sms, smem);
}

}  // namespace apd
