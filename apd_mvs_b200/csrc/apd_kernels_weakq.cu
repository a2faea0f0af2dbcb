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

namespace apd {

constexpr int kWqNT = 128;                 // 4 independent warps
constexpr int kWqPix = 8;                  // pixels per warp = texture quads per warp
constexpr int kOwnTaps = 36, kAncTaps = 8 * 9;
constexpr int kRowSum = kOwnTaps + kAncTaps;                       // rows [kRowSum + 2k, +1] = (sum, sum of squares) of window k
constexpr int kRefRows = kRowSum + 2 * APD_NEIGHBOUR_NUM;          // 126 floats per pixel
__host__ __device__ constexpr int wq_warp_words(int S) { return (kRefRows + 9 * S + APD_NEIGHBOUR_NUM + 8 + 8) * kWqPix; }

// ------------------------------------------------------------------------------------------------
// Compacted WEAK lists. One warp owns an 8x8 pixel area, ordered as four 4x4 sub-areas so that eight consecutive list
// entries (= the eight pixels a k_weak_q warp processes together) are a compact cluster. SPLIT: one list per colour,
// restricted to the rows the reference's half launch reaches; otherwise one list of all WEAK pixels (K3 is a full launch).
template <bool SPLIT>
__global__ void __launch_bounds__(128) k_weak_lists(const Args a) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int ax = blockIdx.x * 32 + warp * 8, ay = blockIdx.y * 8;
	const int sub = lane >> 3, j = lane & 7;
	const int y = ay + (sub >> 1) * 4 + (j >> 1);
	const int xb = ax + (sub & 1) * 4 + (j & 1) * 2;
	const int odd = (xb + y) & 1;
#pragma unroll
	for (int c = 0; c < 2; ++c) {
		const int x = xb + (c ^ odd);                     // colour c: (x + y + c) even
		bool weak = x < a.W && y < a.H && (!SPLIT || y < a.half_rows);
		if (weak) weak = a.states[(size_t)y * a.W + x] == APD_WEAK;
		const unsigned b = __ballot_sync(0xffffffffu, weak);
		if (b == 0u) continue;
		int base = 0;
		if (lane == 0) base = atomicAdd(&a.wctrl[SPLIT ? c : 2], __popc(b));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (weak) a.wlist[(SPLIT ? c * a.wlist_stride : 0) + base + __popc(b & ((1u << lane) - 1u))] = y * a.W + x;
	}
}

// ------------------------------------------------------------------------------------------------
// reference side of one window: taps in the evaluation order (x-offset outer, y-offset inner) and their sum / sum of
// squares accumulated exactly as the evaluation would (APD.cu:456-487)
template <int INC>
__device__ __forceinline__ void wq_cache_window(const Args &a, int cx, int cy, float *col, float *sums) {
	const float *base = a.ref_pad + (size_t)(cy + kRefPad) * a.ref_pitch + (cx + kRefPad);
	float R = 0.f, RR = 0.f;
	int t = 0;
#pragma unroll
	for (int i = -5; i <= 5; i += INC) {
		float r = 0.f, rr = 0.f;
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = __ldg(base + (ptrdiff_t)j * a.ref_pitch + i);
			col[t * kWqPix] = rp;
			++t;
			r += rp; rr = fmaf(rp, rp, rr);
		}
		R += r; RR += rr;
	}
	sums[0] = R; sums[kWqPix] = RR;
}

// NCC of one window for up to NR planes of this lane at once (same reference taps; APD.cu:456-505 / :556-610)
template <int INC, int NR>
__device__ __forceinline__ void wq_window(cudaTextureObject_t tex, int layer, const float *h0, const float *h1, bool w0, bool w1,
                                          int cx, int cy, float inv_w, const float *col, float sum_r, float sum_rr, float &o0, float &o1) {
	NccSums t0 = {sum_r, sum_rr, 0.f, 0.f, 0.f}, t1 = t0;
	int n = 0;
#pragma unroll(INC == 5 ? 3 : 1)
	for (int i = -5; i <= 5; i += INC) {
		const float xf = (float)(cx + i);
		const float ax0 = h0[0] * xf, ay0 = h0[3] * xf, az0 = h0[6] * xf;
		float ax1 = 0.f, ay1 = 0.f, az1 = 0.f;
		if (NR > 1) { ax1 = h1[0] * xf; ay1 = h1[3] * xf; az1 = h1[6] * xf; }
		float rs0 = 0.f, ss0 = 0.f, sm0 = 0.f, rs1 = 0.f, ss1 = 0.f, sm1 = 0.f;
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = col[n * kWqPix];
			++n;
			const float yf = (float)(cy + j);
			float sp0 = 0.f, sp1 = 0.f;
			if (w0) sp0 = src_tap(tex, layer, h0, ax0, ay0, az0, yf);
			if (NR > 1) { if (w1) sp1 = src_tap(tex, layer, h1, ax1, ay1, az1, yf); }
			rs0 = fmaf(rp, sp0, rs0); sm0 += sp0; ss0 = fmaf(sp0, sp0, ss0);
			if (NR > 1) { rs1 = fmaf(rp, sp1, rs1); sm1 += sp1; ss1 = fmaf(sp1, sp1, ss1); }
		}
		t0.s += sm0; t0.ss += ss0; t0.rs += rs0;
		if (NR > 1) { t1.s += sm1; t1.ss += ss1; t1.rs += rs1; }
	}
	o0 = ncc_cost(t0, inv_w);
	if (NR > 1) o1 = ncc_cost(t1, inv_w);
}

// per-pixel shared-memory columns of the owning warp (stride kWqPix between rows)
struct WqPixel {
	const float *refc;        // [kRefRows]
	const short2 *anc;        // [9] anchors, slot 0 = the pixel itself
	const uint32_t *asel;     // [8] selected-view bitmasks of anchors 1..8
};

// ComputeBilateralNCCNew (APD.cu:400-528) for the two planes of this lane against source view v.
__device__ __forceinline__ void wq_deform_pair(const Args &a, const RefConst &rc, const ViewConst &vc, int v, const float4 P0, const float4 P1,
                                               bool on0, bool on1, const WqPixel &px, float inv36, float inv9, float &out0, float &out1) {
	const Homog H0 = make_homography(rc, vc, P0), H1 = make_homography(rc, vc, P1);
	const short2 self = px.anc[0];
	bool live0 = on0 && centre_inside(H0, vc, (float)self.x, (float)self.y);
	bool live1 = on1 && centre_inside(H1, vc, (float)self.x, (float)self.y);
	float cc0 = 0.f, sc0 = 0.f, cc1 = 0.f, sc1 = 0.f; int n0 = 0, n1 = 0;
	const float Wf = (float)a.W, Hf = (float)a.H;
#pragma unroll 1
	for (int k = 0; k < APD_NEIGHBOUR_NUM; ++k) {
		const short2 q = px.anc[k * kWqPix];
		const bool valid = !(q.x == -1 || q.y == -1);
		const float xf = (float)q.x, yf = (float)q.y;
		bool w0 = false, w1 = false;
		if (valid && live0) {
			const float *h = H0.h;
			const float rz = rcpf(h[8] + fmaf(h[6], xf, h[7] * yf));
			const float sx = (h[2] + fmaf(h[0], xf, h[1] * yf)) * rz;
			const float sy = (h[5] + fmaf(h[3], xf, h[4] * yf)) * rz;
			if (sx < 0.0f || sy < 0.0f || sx >= Wf || sy >= Hf) {
				if (k == 0) live0 = false;                                                        // APD.cu:437-438
				else if ((px.asel[(k - 1) * kWqPix] >> v) & 1u) { sc0 += kCostMax; ++n0; }           // APD.cu:439-446
			} else w0 = true;
		}
		if (valid && live1) {
			const float *h = H1.h;
			const float rz = rcpf(h[8] + fmaf(h[6], xf, h[7] * yf));
			const float sx = (h[2] + fmaf(h[0], xf, h[1] * yf)) * rz;
			const float sy = (h[5] + fmaf(h[3], xf, h[4] * yf)) * rz;
			if (sx < 0.0f || sy < 0.0f || sx >= Wf || sy >= Hf) {
				if (k == 0) live1 = false;
				else if ((px.asel[(k - 1) * kWqPix] >> v) & 1u) { sc1 += kCostMax; ++n1; }
			} else w1 = true;
		}
		if (!__any_sync(0xffffffffu, w0 || w1)) continue;          // warp-uniform
		float c0 = 0.f, c1 = 0.f;
		if (k == 0)
			wq_window<2, 2>(a.img_tex, v + 1, H0.h, H1.h, w0, w1, q.x, q.y, inv36, px.refc, px.refc[kRowSum * kWqPix], px.refc[(kRowSum + 1) * kWqPix], c0, c1);
		else
			wq_window<5, 2>(a.img_tex, v + 1, H0.h, H1.h, w0, w1, q.x, q.y, inv9, px.refc + (kOwnTaps + 9 * (k - 1)) * kWqPix,
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

__global__ void __launch_bounds__(kWqNT, 4) k_weak_q(const Args a, const int iter, const int color, int *work) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RefConst *sr = reinterpret_cast<RefConst *>(smem_raw);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int S = a.S, W = a.W;
	float *wbase = reinterpret_cast<float *>(sv + S) + (size_t)warp * wq_warp_words(S);
	{
		const int nv = S * (int)(sizeof(ViewConst) / 4);
		const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views); uint32_t *s = reinterpret_cast<uint32_t *>(sv);
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
		if (ql == 0) wq_cache_window<2>(a, px, py, refc, refc + kRowSum * kWqPix);
		else {
#pragma unroll 1
			for (int k = ql; k < APD_NEIGHBOUR_NUM; k += 3) {
				const short2 q = anc[k * kWqPix];
				if (q.x == -1 || q.y == -1) continue;
				wq_cache_window<5>(a, q.x, q.y, refc + (kOwnTaps + 9 * (k - 1)) * kWqPix, refc + (kRowSum + 2 * k) * kWqPix);
			}
		}
		__syncwarp();
		// candidates = current planes of the anchors that are (still) STRONG
		const unsigned flags = astrong[0] & 0xFFu;
		auto cand_pos = [&](int k) { const short2 q = anc[(k + 1) * kWqPix]; return q.x + q.y * W; };

		// ---- quad state machine. phase 0: cost matrix (lane l: candidates l, l+4; all S views);
		//      phase 1: current plane (lane 0) + fit plane (lane 1) over the sampled views;
		//      phase 2: the five refinement hypotheses (lanes 0..3 + lane 0's second slot); phase 3: done
		int phase = 0;
		uint32_t m = (S >= 32) ? 0xffffffffu : ((1u << S) - 1u);
		float4 T0 = make_float4(0.f, 0.f, 1.f, 1.f), T1 = T0;
		bool on0 = (flags >> ql) & 1u, on1 = (flags >> (ql + 4)) & 1u;
		if (on0) T0 = a.planes[cand_pos(ql)];
		if (on1) T1 = a.planes[cand_pos(ql + 4)];
		float acc0 = 0.f, acc1 = 0.f; bool pruned0 = false, pruned1 = false;
		float limit = kInf;
		Rng rng; rng.v0 = rng.v1 = rng.v2 = rng.v3 = rng.v4 = rng.d = 0u;
		VW vw; vw.lo = 0ull; vw.hi = 0ull;
		uint32_t temp_sel = 0u; float inv_wn = 0.f, best_cost = 0.f; int best_k = 0;
		float4 pl_now = make_float4(0.f, 0.f, 1.f, 1.f), fit = pl_now;
		float cost_now = 0.f, cost_stored = 0.f, depth_now = 0.f, d_fit = 0.f;
		bool have_fit = false, fit_ok = false, sel_write = false;
		Refine5 rf; rf.depth_rand = rf.depth_pert = rf.d0 = 0.f; rf.n_rand = rf.n_pert = rf.n0 = pl_now;

#pragma unroll 1
		for (int step = 0; step < 3 * S + 8; ++step) {
			// ---- phase transitions (quad-uniform; quads of a warp may be in different phases)
#pragma unroll 1
			while (phase < 3 && m == 0u) {
				if (phase == 0) {
					__syncwarp(qmask);                                     // the quad's matrix entries are visible
					// view selection (APD.cu:1365-1434): priors from every existing anchor, STRONG or not
					const float thr = 0.8 * __expf((float)(unsigned)(iter * iter) * -0.011111111380159854889f);
					const float thr_fallback = __expf((thr * thr) * -3.125f);
					float prob_sum = 0.0f;
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
						for (int v = 0; v < S; ++v) { cum = fmaf(inv, PROB(v), cum); __syncwarp(qmask); if (ql == 0) PROB(v) = cum; __syncwarp(qmask); }
						for (int s = 0; s < 15; ++s) {
							const float r = rng_uniform(rng) - 1.1920928955078125e-07f;
							for (int v = 0; v < S; ++v) if (PROB(v) > r) { vw_add(vw, v); break; }
						}
					}
					float weight_norm = 0.0f;
					for (int v = 0; v < S; ++v) { const int w = vw_get(vw, v); if (w > 0) { temp_sel |= 1u << v; weight_norm += (float)w; } }
					inv_wn = rcpf(weight_norm);
					// final costs of the eight candidates (APD.cu:1436-1452): lane l computes candidates l and l+4
					float fcr0 = 0.f, fcr1 = 0.f;
					{
						float s0 = 0.f, s1 = 0.f;
						for (int v = 0; v < S; ++v) {
							const int w = vw_get(vw, v);
							if (w == 0) continue;
							float c0 = CM(ql, v), c1 = CM(ql + 4, v);
							if (a.geom) {
								c0 = on0 ? fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, T0, xf, yf), c0) : fmaf(a.geom_factor, 3.0f, c0);
								c1 = on1 ? fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, T1, xf, yf), c1) : fmaf(a.geom_factor, 3.0f, c1);
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
					fit = a.fit_planes[center];
					have_fit = !(fit.x == 0.0f && fit.y == 0.0f && fit.z == 0.0f);
					d_fit = plane_depth(rc, fit, xf, yf);
					fit_ok = have_fit && d_fit >= a.depth_min && d_fit <= a.depth_max;
					T0 = (ql == 1) ? fit : pl_now;
					on0 = (ql == 0) || (ql == 1 && fit_ok); on1 = false;
					acc0 = acc1 = 0.f; pruned0 = pruned1 = false; limit = kInf;
					m = temp_sel; phase = 1;
				} else if (phase == 1) {
					const float tc_cur = __shfl_sync(qmask, acc0, 0, 4) * inv_wn;
					const float tc_fit = __shfl_sync(qmask, acc0, 1, 4) * inv_wn;
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
						if (fit_ok && tc_fit < cost_now) { depth_now = d_fit; pl_now = fit; cost_now = tc_fit; }   // APD.cu:916-935
						rf.depth_rand = fmaf(rng_uniform(rng), a.depth_max - a.depth_min, a.depth_min);
						rf.n_rand = random_normal(rc, xf, yf, rng, depth_now);
						const float lo = depth_now * (1.0f - 0.02f);
						const float span = fmaf(depth_now, 1.0f + 0.02f, -lo);
						rf.depth_pert = fmaf(span, rng_uniform(rng), lo);
						rf.n_pert = perturbed_normal(rc, xf, yf, pl_now, rng);
						rf.n0 = pl_now; rf.d0 = depth_now;
						// lanes 0..3 take hypotheses 3, 4 (the two near the current plane), 0, 1; lane 0's second slot takes 2
						float d; bool ok;
						wq_hypothesis(a, rc, rf, (ql + 3) % 5, xf, yf, T0, d, ok);
						on0 = ok;
						on1 = false;
						if (ql == 0) { wq_hypothesis(a, rc, rf, 2, xf, yf, T1, d, ok); on1 = ok; }
						acc0 = acc1 = 0.f; pruned0 = pruned1 = false; limit = cost_now;
						m = temp_sel; phase = 2;
						// nothing in range anywhere in the quad: no evaluation at all
						if ((__ballot_sync(qmask, on0 || on1) & qmask) == 0u) m = 0u;
					} else phase = 3;
				} else {   // phase == 2: adopt in the reference's order, strict <
#pragma unroll 1
					for (int i = 0; i < 5; ++i) {
						const int src = (i == 2) ? 0 : ((i + 2) % 5);           // lane that evaluated hypothesis i
						const float tci = __shfl_sync(qmask, (i == 2) ? acc1 : acc0, src, 4) * inv_wn;
						float4 t; float d; bool ok;
						wq_hypothesis(a, rc, rf, i, xf, yf, t, d, ok);
						if (ok && tci < cost_now) { depth_now = d; pl_now = t; cost_now = tci; }
					}
					phase = 3;
				}
			}
			if (!__any_sync(0xffffffffu, phase < 3)) break;
			const bool want = phase < 3;
			const int v = want ? (__ffs(m) - 1) : 0;
			m &= m - 1u;
			const bool l0 = want && on0 && !pruned0, l1 = want && on1 && !pruned1;
			float c0, c1;
			wq_deform_pair(a, rc, sv[v], v, T0, T1, l0, l1, wp, inv36, inv9, c0, c1);
			if (phase == 0) {
				// a missing candidate keeps the reference's partially initialised row: `cost_array[8][32] = {2.0f}` sets
				// [0][0] only (APD.cu:1345)
				CM(ql, v) = on0 ? c0 : ((ql == 0 && v == 0) ? 2.0f : 0.0f);
				CM(ql + 4, v) = on1 ? c1 : 0.0f;
			} else if (want) {
				const float w = (float)vw_get(vw, v);
				if (l0) {
					if (a.geom) c0 = fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, T0, xf, yf), c0);
					acc0 = fmaf(w, c0, acc0);
					if (prune && acc0 * inv_wn >= limit) pruned0 = true;
				}
				if (l1) {
					if (a.geom) c1 = fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, T1, xf, yf), c1);
					acc1 = fmaf(w, c1, acc1);
					if (prune && acc1 * inv_wn >= limit) pruned1 = true;
				}
			}
			// a quad whose hypotheses are all out of the race stops walking its views
			const unsigned running = __ballot_sync(0xffffffffu, (on0 && !pruned0) || (on1 && !pruned1));
			if (phase != 0 && (running & qmask) == 0u) m = 0u;
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
		// "update cost with old method": plain NCC of the committed plane over the sampled views, lanes = views
		float facc = 0.0f;
		uint32_t fm = temp_sel;
#pragma unroll 1
		while (__any_sync(0xffffffffu, fm != 0u)) {
			uint32_t mm = fm;
			for (int t = 0; t < ql; ++t) mm &= mm - 1u;
			const int v = mm ? (__ffs(mm) - 1) : -1;
			const int vv = v < 0 ? 0 : v;
			const Homog Hm = make_homography(rc, sv[vv], final_plane);
			const bool w = v >= 0 && centre_inside(Hm, sv[vv], xf, yf);
			float c = kCostMax, dummy = 0.f;
			if (__any_sync(0xffffffffu, w)) {
				float cw = 0.f;
				wq_window<2, 1>(a.img_tex, vv + 1, Hm.h, Hm.h, w, false, px, py, inv36, refc, refc[kRowSum * kWqPix], refc[(kRowSum + 1) * kWqPix], cw, dummy);
				if (w) c = cw;
			}
#pragma unroll
			for (int t = 0; t < 4; ++t) {
				const float ct = __shfl_sync(0xffffffffu, c, t, 4);
				const int vt = __shfl_sync(0xffffffffu, v, t, 4);
				if (vt >= 0) facc = fmaf((float)vw_get(vw, vt), ct, facc);
			}
			fm &= fm - 1u; fm &= fm - 1u; fm &= fm - 1u; fm &= fm - 1u;
		}
		if (alive && ql == 0) a.costs[center] = facc * inv_wn;
	}
#undef CM
#undef PROB
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_weak_lists(cudaStream_t st, const Args &a, bool split) {
	dim3 g((a.W + 31) / 32, (a.H + 7) / 8);
	if (split) {
		cudaError_t e = cudaMemsetAsync(a.wctrl, 0, 2 * sizeof(int), st);
		if (e != cudaSuccess) return e;
		k_weak_lists<true><<<g, 128, 0, st>>>(a);
	} else {
		cudaError_t e = cudaMemsetAsync(a.wctrl + 2, 0, sizeof(int), st);
		if (e != cudaSuccess) return e;
		k_weak_lists<false><<<g, 128, 0, st>>>(a);
	}
	return cudaGetLastError();
}

cudaError_t launch_weak_q(cudaStream_t st, const Args &a, int iter, int color, int work_slot, int num_sms) {
	const size_t smem = sizeof(RefConst) + (size_t)a.S * sizeof(ViewConst) + (size_t)(kWqNT / 32) * wq_warp_words(a.S) * 4;
	if (smem > 227 * 1024) return cudaErrorInvalidValue;
	cudaError_t e = cudaFuncSetAttribute(k_weak_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_weak_q, kWqNT, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) per_sm = 1;
	k_weak_q<<<num_sms * per_sm, kWqNT, smem, st>>>(a, iter, color, a.wctrl + kWorkBase + work_slot);
	return cudaGetLastError();
}

}  // namespace apd
