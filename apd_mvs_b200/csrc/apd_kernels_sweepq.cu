// sm_100a kernel of the depth sweep, second design (round 2):
//   k_sweep_q      K14 DepthToWeak APD.cu:1990-2144 + K15 LocalRefine APD.cu:2146-2232 (fused)
//
// The first k_sweep (apd_kernels_strong.cu, kept selectable with APD_SWEEP_IMPL=old for A/B timing) maps one thread to
// one pixel and fetches quad-cooperatively through a shared-memory slab. ncu (profiles/kernel_counters.json) shows it
// ISSUE bound: 44 warp instructions per texture instruction (shuffles that broadcast the owner's homography, the slab
// stores and loads, scalar homographies and geometric terms), issue slots 70 % busy with the texture data pipe at 53 %.
//
// Here ONE PIXEL IS OWNED BY ONE TEXTURE QUAD, as in k_weak_q: the four lanes evaluate FOUR CONSECUTIVE DEPTH STEPS of
// the pixel's plane against the same source view (the four bilinear footprints of a TEX instruction lie a pixel apart
// on the epipolar line: the fast case of the texture unit, 1.02 wavefronts per request), and every lane carries two
// evaluations ("slots") through the NCC as packed fp32 pairs (apd_pair.cuh): homography, projective warp, accumulation
// and the geometric term are issued once for both. No slab, no shuffles in the evaluation; the reference window (36 taps
// and their sums) and the K14 cost profile sit in shared memory per pixel, 8 pixels per warp.
//   * A pixel's work is a list of 60 items: item 0 = LocalRefine's cost of the current depth (APD.cu:2173-2182), items
//     1..59 = the sweep steps in the centre-out order of k_sweep (centre window, then right, then left). Group g = items
//     4g..4g+3 = the four lanes. After a group the quad replays the classification rules on the new profile entries in
//     order, so the exact early decisions of k_sweep are kept (at a granularity of four steps).
//   * The evaluations of a quad form one stream: group g against its sampled views in ascending order, then group g+1,
//     ... A step takes the next TWO entries of the stream for the two slots. With an odd number of sampled views a step
//     straddles two groups (last view of g, first view of g+1: two planes, two views), so both slots always carry work
//     and the fetches need no predicates; the group opened ahead is discarded if the classification ends the pixel.
//     A lane's two group contexts (plane offset, world point, accumulators) live in shared memory.
//   * The window loop is software-pipelined (wq_window6_pipe): the fetches of the next tap column are in flight while
//     the current one is accumulated.
//   * Quads are independent state machines: a quad that has finished its pixel takes the next one of the warp's chunk
//     (8x4 pixels, pulled through an atomic counter by persistent warps), so lanes stay busy although pixels need between
//     3 and 15 groups.
// Every arithmetic expression is the one of k_sweep (bit-identical to the reference); only who evaluates what changed.
#include <cstdlib>
#include "apd_device.cuh"
#include "apd_pair.cuh"

namespace apd {

constexpr int kSqNT = 128;                 // 4 independent warps
constexpr int kSqPix = 8;                  // pixels in flight per warp = texture quads per warp
constexpr int kSqRef = 38;                 // 36 reference taps + their sum + sum of squares
constexpr int kSqProf = 61;                // K14 profile, entry 0 = cost of the current depth (the reference never reads profile[0])
constexpr int kSqNear = 11;                // LocalRefine's 11 candidate costs (disparity steps -5..5)
constexpr int kSqCtxRows = 7;              // a lane's group context: plane offset, world point (3), acc14, acc15, flags
constexpr int kSqWarpWords = (kSqRef + kSqProf + kSqNear) * kSqPix + 2 * kSqCtxRows * 32;
// context flags
constexpr int kFNeed = 1, kFInRange = 2, kFStore15 = 4, kFWant14 = 8;
constexpr int kSqTileW = 8, kSqTileH = 4;  // one chunk
constexpr int kSqBlocks = 4;               // resident blocks per SM the register allocation is sized for (128 registers, no spills;
                                           // measured: 5 blocks at 96 registers 291 ms, 3 at 168 341 ms, 4: 274 ms at 6221x4146)

template <bool DO14, bool DO15>
__global__ void __launch_bounds__(kSqNT, kSqBlocks) k_sweep_q(const Args a, int *work, const int tiles_x, const int nchunks) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RefConst *sr = reinterpret_cast<RefConst *>(smem_raw);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int S = a.S, W = a.W;
	float *wbase = reinterpret_cast<float *>(sv + S) + (size_t)warp * kSqWarpWords;
	{
		const int nv = S * (int)(sizeof(ViewConst) / 4);
		const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views); uint32_t *s = reinterpret_cast<uint32_t *>(sv);
#pragma unroll 1
		for (int i = tid; i < nv; i += kSqNT) s[i] = g[i];
		const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.ref); uint32_t *srr = reinterpret_cast<uint32_t *>(sr);
		for (int i = tid; i < (int)(sizeof(RefConst) / 4); i += kSqNT) srr[i] = gr[i];
	}
	const int ql = lane & 3, pq = lane >> 2;
	const unsigned qmask = 0xFu << (lane & ~3);
	float *refc = wbase + pq;                                  // [kSqRef][8]
	float *prof = wbase + kSqRef * kSqPix + pq;                // [kSqProf][8]
	float *near15 = wbase + (kSqRef + kSqProf) * kSqPix + pq;  // [kSqNear][8]
	float *lctx = wbase + (kSqRef + kSqProf + kSqNear) * kSqPix + lane;   // [2][kSqCtxRows][32]: this lane's two group contexts
	for (int i = 0; i < 2 * kSqCtxRows; ++i) lctx[i * 32] = (i % kSqCtxRows == 0) ? 1.0f : 0.0f;   // a quad without a pixel evaluates this (finite) plane
	__syncthreads();
	const RefConst &rc = *sr;
	const float inv36 = a.inv_w[0];
	const int rad = a.weak_peak_radius;
	const int R = min(29, max(rad + 1, 5));                    // centre window half-width (k_sweep)
	const int nA = 2 * R + 1, nSide = 29 - R;
	const int last15 = R + 6;                                  // item of the last step LocalRefine reads (k = +5)
	const float kNaN = __int_as_float(0x7fc00000);

	// the warp's queue of pixels: chunk id and cursor into its 32 pixels
	int chunk = 0, cursor = kSqTileW * kSqTileH; bool exhausted = false;
	// quad state (replicated in the four lanes)
	bool busy = false, on14 = false, on15 = false, decided = true, cur_open = false, nxt_open = false;
	int px = 0, py = 0, g = 0, ci = 0; size_t center = 0;
	float xf = 0.f, yf = 0.f, inv_wn = 0.f, c_in = 3.0f;
	SweepCtx c; c.pl = make_float4(0.f, 0.f, 1.f, 1.f); c.depth = 1.0f; c.weight_normal = 0.f; c.kb = 1.f; c.disp = 1.f; c.valid = 0;
	uint32_t act = 0u, mcur = 0u, mnxt = 0u; VW vw; vw.lo = 0ull; vw.hi = 0ull;
	uint8_t early = APD_UNKNOWN;

	// item index -> profile index of a sweep step (items 1..59)
	auto profile_index = [&](int step) {
		const bool phaseA = step < nA, right = !phaseA && step < nA + nSide;
		return phaseA ? (30 - R + step) : right ? (30 + R + 1 + (step - nA)) : (30 - R - 1 - (step - nA - nSide));
	};
	// open context `idx` for group `grp`: lane ql takes item 4*grp + ql. Returns the views the group has to visit.
	auto open_ctx = [&](int idx, int grp) -> uint32_t {
		const int item = 4 * grp + ql;
		const bool want14 = DO14 && !decided;
		int flags = want14 ? kFWant14 : 0;
		float tw = c.pl.w;
		bool need = false;
		if (item == 0) {
			if (DO15 && on15) {
				// APD.cu:2173-2182: the compiler hoisted normal.z * depth out of the view loop there as a rounded product
				float X0, X1; backproject(rc, xf, yf, c.depth, X0, X1);
				tw = -((c.depth * c.pl.z) + fmaf(X0, c.pl.x, X1 * c.pl.y));
				need = true; flags |= kFInRange | (1 << 8);
			}
		} else if (item <= 59) {
			const int i = profile_index(item - 1), kk = i - 30;
			const bool want15 = DO15 && on15 && (kk >= -5 && kk <= 5);
			const float d = c.kb * rcpf(c.disp + (float)kk);
			const bool in_range = !(d < a.depth_min || d > a.depth_max);
			need = in_range && (want14 || want15);
			tw = plane_offset(rc, xf, yf, d, c.pl.x, c.pl.y, c.pl.z);
			flags |= (in_range ? kFInRange : 0) | (want15 ? kFStore15 : 0) | ((i + 1) << 8) | ((kk + 32) << 16);
		}
		if (need) flags |= kFNeed;
		float *cx = lctx + idx * (kSqCtxRows * 32);
		cx[0] = tw;
		if (a.geom && need) {
			const GeomPoint gp = geom_point(rc, make_float4(c.pl.x, c.pl.y, c.pl.z, tw), xf, yf);
			cx[1 * 32] = gp.x; cx[2 * 32] = gp.y; cx[3 * 32] = gp.z;
		}
		cx[4 * 32] = 0.0f; cx[5 * 32] = 0.0f;
		cx[6 * 32] = __int_as_float(flags);
		return (__ballot_sync(qmask, need) & qmask) ? act : 0u;
	};

#pragma unroll 1
	for (;;) {
		// ---- group management. Every quad passes through the same straight-line sequence once per step (close, refill, open,
		// open ahead), so the quads that need a piece execute it TOGETHER; a second pass runs only when some quad still has
		// no evaluation to do (a single sampled view closes two groups per step; a group nobody needs).
#pragma unroll 1
		for (int pass = 0; pass < 2; ++pass) {
			if (pass == 1 && !__any_sync(0xffffffffu, busy && (!cur_open || mcur == 0u))) break;
			// (1) close group g, classify
			if (busy && cur_open && mcur == 0u) {
				const float *cx = lctx + ci * (kSqCtxRows * 32);
				const int flags = __float_as_int(cx[6 * 32]);
				const int pi = ((flags >> 8) & 0xff) - 1;
				const bool in_range = flags & kFInRange;
				const float acc14 = cx[4 * 32];
				if (pi > 0) { if (flags & kFWant14) prof[pi * kSqPix] = in_range ? ((2.0f > acc14 * inv_wn) ? acc14 * inv_wn : 2.0f) : 2.0f; }   // OpenCV MIN(2.0f, p_cost), APD.cu:2082
				else if (pi == 0) prof[0] = acc14;                                                                                         // cost of the current depth
				if (flags & kFStore15) near15[(((flags >> 16) & 0xff) - 32 + 5) * kSqPix] = in_range ? cx[5 * 32] * inv_wn : kNaN;
				__syncwarp(qmask);
				if (DO14 && !decided) {
					// the early rules of k_sweep (APD.cu:2092-2143 decided early, exactly). Every early decision is "WEAK", so the four
					// new entries can be tested by the four lanes at once: the outcome is the one of testing them in sweep order.
					const int step = 4 * g + ql - 1;
					if (4 * g - 1 <= nA - 1 && nA - 1 <= 4 * g + 2) {      // the centre window is complete: cheapest peak within the radius
						for (int j = max(2, 30 - rad); j <= min(58, 30 + rad); ++j) {
							const float cj = prof[j * kSqPix];
							if (prof[(j - 1) * kSqPix] > cj && prof[(j + 1) * kSqPix] > cj && cj < c_in) c_in = cj;
						}
						if (c_in > 0.5f) { decided = true; early = APD_WEAK; }     // none (3.0) or too dear
					}
					bool hit = false;
					if (step >= nA && step <= 58) {
						const float pc = prof[pi * kSqPix];
						if (step < nA + nSide) {                              // newly testable peak, outside the radius on the right
							const int j = pi - 1;
							const float cj = prof[j * kSqPix];
							hit = j <= 58 && prof[(j - 1) * kSqPix] > cj && pc > cj && cj < c_in;
						} else {                                              // outside on the left: wins ties (smaller index)
							const int j = pi + 1;
							const float cj = prof[j * kSqPix];
							hit = j >= 2 && pc > cj && prof[(j + 1) * kSqPix] > cj && cj <= c_in;
						}
					}
					if (__ballot_sync(qmask, hit) & qmask) { decided = true; early = APD_WEAK; }
				}
				++g;
				if ((g == 15) || ((!on14 || decided) && (!(DO15 && on15) || 4 * g > last15))) {
					// ---- the pixel is finished
					if (DO15 && on15) {
						float min_cost15 = 2.0f, best_depth = c.depth;
#pragma unroll 1
						for (int k = -5; k <= 5; ++k) {
							const float tc = near15[(k + 5) * kSqPix];
							if (tc < min_cost15) { min_cost15 = tc; best_depth = c.kb * rcpf(c.disp + (float)k); }
						}
						const float diff = fmaf(inv_wn, prof[0], -min_cost15);     // (cost_now / weight_normal) - min_cost, one FFMA
						if ((double)diff > 0.1 && ql == 0) a.planes[center].w = best_depth;
					}
					if (DO14) {
						uint8_t out = on14 ? early : (uint8_t)APD_UNKNOWN;
						if (on14 && !decided) {   // peak analysis, APD.cu:2092-2143: the lanes scan every fourth entry, then combine
							unsigned lo = 0u, hi = 0u; int min_peak = 0; float min_cost = 2.0f;
#pragma unroll 1
							for (int i = 2 + ql; i < 59; i += 4) {
								const float ci_ = prof[i * kSqPix];
								if (prof[(i - 1) * kSqPix] > ci_ && prof[(i + 1) * kSqPix] > ci_) {
									if (i < 32) lo |= 1u << i; else hi |= 1u << (i - 32);
									if (ci_ < min_cost) { min_peak = i; min_cost = ci_; }
								}
							}
#pragma unroll
							for (int d = 1; d <= 2; d <<= 1) {
								lo |= __shfl_xor_sync(qmask, lo, d); hi |= __shfl_xor_sync(qmask, hi, d);
								const float oc = __shfl_xor_sync(qmask, min_cost, d); const int op = __shfl_xor_sync(qmask, min_peak, d);
								// the reference keeps the FIRST index of the cheapest peak (strict <, ascending scan)
								if (oc < min_cost || (oc == min_cost && op < min_peak)) { min_cost = oc; min_peak = op; }
							}
							const unsigned long long peaks = ((unsigned long long)hi << 32) | lo;
							const int peak_count = __popcll(peaks);
							if (abs(min_peak - 30) > rad || prof[min_peak * kSqPix] > 0.5f) out = APD_WEAK;
							else if (peak_count == 1) out = (prof[min_peak * kSqPix] <= 0.15f) ? APD_STRONG : APD_WEAK;
							else {
								float var = 0.0f;
								unsigned long long rest = peaks & ~(1ull << min_peak);
#pragma unroll 1
								while (rest) {                               // ascending index = the reference's summation order
									const int i = __ffsll((long long)rest) - 1; rest &= rest - 1ull;
									const float dd = prof[i * kSqPix] - min_cost; var = fmaf(dd, dd, var);
								}
								var = sqrtaf(var) * rcpf((float)(peak_count - 1));
								out = (var > 0.2f) ? APD_STRONG : APD_WEAK;
							}
						}
						if (ql == 0) a.states[center] = out;
					}
					busy = false; cur_open = false; nxt_open = false;
				} else if (nxt_open) { ci ^= 1; mcur = mnxt; nxt_open = false; }      // the group opened ahead becomes the current one
				else cur_open = false;
			}
			// (2) quads without a pixel take the next pixels of the warp's chunk
			if (pass == 0) {
#pragma unroll 1
				for (;;) {
					const unsigned nb = __ballot_sync(0xffffffffu, !busy && ql == 0);
					if (nb == 0u || exhausted) break;
					if (cursor == kSqTileW * kSqTileH) {
						int cn = 0;
						if (lane == 0) cn = atomicAdd(work, 1);
						cn = __shfl_sync(0xffffffffu, cn, 0);
						if (cn >= nchunks) { exhausted = true; break; }
						chunk = cn; cursor = 0;
					}
					const int rank = __popc(nb & ((1u << (lane & ~3)) - 1u));
					const int avail = kSqTileW * kSqTileH - cursor;
					const bool take = !busy && rank < avail;
					const int j = cursor + rank;
					cursor += min(__popc(nb), avail);
					if (take) {
						// 8x4 chunk as four 4x2 clusters: eight consecutive pixels (= what the warp works on together) are neighbours
						const int cl = j >> 3;
						px = (chunk % tiles_x) * kSqTileW + 4 * (cl & 1) + (j & 3);
						py = (chunk / tiles_x) * kSqTileH + 2 * (cl >> 1) + ((j >> 2) & 1);
						if (px < W && py < a.H) {
							center = (size_t)py * W + px;
							xf = (float)px; yf = (float)py;
							const uint32_t bits = a.sel_views[center];
							vw = vw_load(a.view_w, center);
							c.weight_normal = 0.f; c.valid = 0;
							const bool has_depth = sweep_setup(a, rc, sv, center, bits, vw, c);
							const bool border = px < 6 || py < 6 || px >= W - 6 || py >= a.H - 6;
							on14 = DO14 && !border && has_depth && c.valid > 0;
							on15 = DO15 && has_depth && c.valid > 0 && c.weight_normal != 0.0f;
							if (on14 || on15) {
								inv_wn = rcpf(c.weight_normal);
								act = bits & vw_mask(vw, S);
								__syncwarp(qmask);                  // the quad's reads of the previous pixel's columns are over
								if (ql == 0) wq_cache_window<2, kSqPix>(a, px, py, refc, refc + 36 * kSqPix);
								__syncwarp(qmask);
								busy = true; decided = !on14; early = APD_UNKNOWN; c_in = 3.0f; g = 0; ci = 0; mcur = 0u; cur_open = false; nxt_open = false;
							} else if (DO14 && ql == 0) a.states[center] = APD_UNKNOWN;
						}
					}
				}
			}
			// (3) open the current group; (4) if its last evaluation is next, the second slot takes the first view of the group
			// after it, which is opened ahead (and discarded if the classification ends the pixel first)
#pragma unroll 1
			for (int o = 0; o < 2; ++o) {
				bool want;
				if (o == 0) want = busy && !cur_open;
				else {
					const bool last_group = (g + 1 == 15) || ((!on14 || decided) && (!(DO15 && on15) || 4 * (g + 1) > last15));
					want = busy && cur_open && mcur != 0u && (mcur & (mcur - 1u)) == 0u && !nxt_open && !last_group;
				}
				if (want) {
					const uint32_t mm = open_ctx(o ? (ci ^ 1) : ci, g + o);
					if (o) { mnxt = mm; nxt_open = true; } else { mcur = mm; cur_open = true; }
				}
			}
		}
		// a quad that finished in the second pass has not been given a new pixel yet: leave only when the queue is empty too
		if (!__any_sync(0xffffffffu, busy)) { if (exhausted) break; else continue; }

		// ---- one evaluation step: the next two entries of the quad's stream. Slot 0 always belongs to the current group;
		// slot 1 to the current group, or to the one opened ahead, or (nothing left) it repeats slot 0 and is discarded.
		int v0 = 0, v1 = 0, i1 = ci; bool h0 = false, h1 = false;
		if (busy && mcur != 0u) {
			v0 = __ffs(mcur) - 1; mcur &= mcur - 1u; h0 = true;
			if (mcur != 0u) { v1 = __ffs(mcur) - 1; mcur &= mcur - 1u; h1 = true; }
			else if (nxt_open && mnxt != 0u) { v1 = __ffs(mnxt) - 1; mnxt &= mnxt - 1u; h1 = true; i1 = ci ^ 1; }
			else v1 = v0;
		}
		float *cx0 = lctx + ci * (kSqCtxRows * 32), *cx1 = lctx + i1 * (kSqCtxRows * 32);
		const int f0 = __float_as_int(cx0[6 * 32]), f1 = __float_as_int(cx1[6 * 32]);
		const bool s0 = h0 && (f0 & kFNeed), s1 = h1 && (f1 & kFNeed);
		const ViewConst &vc0 = sv[v0], &vc1 = sv[v1];
		const Homog2 H = make_homography2(rc, vc0, vc1, make_float4(c.pl.x, c.pl.y, c.pl.z, cx0[0]), make_float4(c.pl.x, c.pl.y, c.pl.z, cx1[0]));
		bool in0, in1;
		{
			float x0, y0, x1, y1;
			project2(H, xf, yf, x0, y0, x1, y1);
			in0 = !(x0 >= vc0.wf || x0 < 0.0f || y0 >= vc0.hf || y0 < 0.0f);
			in1 = !(x1 >= vc1.wf || x1 < 0.0f || y1 >= vc1.hf || y1 < 0.0f);
		}
		float c0, c1;
		wq_window6<kSqPix>(a.img_tex, v0 + 1, v1 + 1, H, px, py, inv36, refc, refc[36 * kSqPix], refc[37 * kSqPix], c0, c1);
		if (!in0) c0 = kCostMax;
		if (!in1) c1 = kCostMax;
		float g0 = 0.f, g1 = 0.f;
		if (a.geom) {
			GeomPoint P0, P1;
			P0.x = cx0[1 * 32]; P0.y = cx0[2 * 32]; P0.z = cx0[3 * 32];
			P1.x = cx1[1 * 32]; P1.y = cx1[2 * 32]; P1.z = cx1[3 * 32];
			geom_cost_at2(a, rc, vc0, vc1, v0 + 1, v1 + 1, P0, P1, xf, yf, g0, g1);
		}
		if (s0) {
			const float w = (float)vw_get(vw, v0);
			float a14 = cx0[4 * 32];
			a14 = fmaf(w, a.geom ? fmaf(a.geom_factor, g0, c0) : c0, a14);                                      // APD.cu:2074-2078 / :2180
			cx0[4 * 32] = a14;
			if (DO15) { float a15 = fmaf(w, c0, cx0[5 * 32]); if (a.geom) a15 = fmaf(w, a.geom_factor * g0, a15); cx0[5 * 32] = a15; }   // :2217-2220
		}
		if (s1) {
			const float w = (float)vw_get(vw, v1);
			float a14 = cx1[4 * 32];
			a14 = fmaf(w, a.geom ? fmaf(a.geom_factor, g1, c1) : c1, a14);
			cx1[4 * 32] = a14;
			if (DO15) { float a15 = fmaf(w, c1, cx1[5 * 32]); if (a.geom) a15 = fmaf(w, a.geom_factor * g1, a15); cx1[5 * 32] = a15; }
		}
	}
}

// ------------------------------------------------------------------------------------------------
template <bool DO14, bool DO15>
static cudaError_t launch_sweep_q_t(cudaStream_t st, const Args &a, int *work, int num_sms, size_t smem) {
	cudaError_t e = cudaFuncSetAttribute(k_sweep_q<DO14, DO15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_q<DO14, DO15>, kSqNT, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) per_sm = 1;
	const int tiles_x = (a.W + kSqTileW - 1) / kSqTileW, tiles_y = (a.H + kSqTileH - 1) / kSqTileH;
	k_sweep_q<DO14, DO15><<<num_sms * per_sm, kSqNT, smem, st>>>(a, work, tiles_x, tiles_x * tiles_y);
	return cudaGetLastError();
}
// mode 0: K14 only, 1: K15 only, 2: K14+K15 fused (as launch_sweep)
cudaError_t launch_sweep_q(cudaStream_t st, const Args &a, int mode, int num_sms) {
	const size_t smem = sizeof(RefConst) + (size_t)a.S * sizeof(ViewConst) + (size_t)(kSqNT / 32) * kSqWarpWords * 4;
	int *work = a.wctrl + 3 + mode;                        // wctrl[3..5]: zeroed at the start of every run
	return mode == 0 ? launch_sweep_q_t<true, false>(st, a, work, num_sms, smem)
	     : mode == 1 ? launch_sweep_q_t<false, true>(st, a, work, num_sms, smem) : launch_sweep_q_t<true, true>(st, a, work, num_sms, smem);
}

}  // namespace apd
