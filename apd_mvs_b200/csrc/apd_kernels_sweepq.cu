// sm_100a kernel of the depth sweep, second design (round 2):
//   k_sweep_q      K14 DepthToWeak APD.cu:1990-2144 + K15 LocalRefine APD.cu:2146-2232 (fused)
//
// The first k_sweep (apd_kernels_strong.cu, kept selectable with APD_SWEEP_IMPL=old for A/B timing) maps one thread to
// one pixel and fetches quad-cooperatively through a shared-memory slab. ncu (profiles/r02_cfg3s_ncu_summary.json) shows it
// ISSUE bound: ~30 warp instructions per texture instruction (shuffles that broadcast the owner's homography, the slab
// stores and loads, scalar homographies and geometric terms), issue slots 70 % busy with the texture data pipe at 53 %.
//
// Here ONE PIXEL IS OWNED BY ONE TEXTURE QUAD, as in k_weak_q: the four lanes evaluate FOUR CONSECUTIVE DEPTH STEPS of
// the pixel's plane against the same source view (the four bilinear footprints of a TEX instruction lie a pixel apart
// on the epipolar line: the fast case of the texture unit), and every lane carries two source views ("slots") through
// the NCC as packed fp32 pairs (apd_pair.cuh): homography, projective warp, accumulation and the geometric term are
// issued once for both. No slab, no shuffles in the evaluation; the reference window (36 taps and their sums) and the
// K14 cost profile sit in shared memory per pixel, 8 pixels per warp.
//   * A pixel's work is a list of 60 items: item 0 = LocalRefine's cost of the current depth (APD.cu:2173-2182), items
//     1..59 = the sweep steps in the centre-out order of k_sweep (centre window, then right, then left). Group g = items
//     4g..4g+3 = the four lanes. After a group the quad replays the classification rules on the new profile entries in
//     order, so the exact early decisions of k_sweep are kept (at a granularity of four steps).
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
constexpr int kSqWarpWords = (kSqRef + kSqProf + kSqNear) * kSqPix;
constexpr int kSqTileW = 8, kSqTileH = 4;  // one chunk

// ComputeGeomConsistencyCost's view-dependent half (geom_cost_at, apd_device.cuh) for the two slots: one world point
// against source views va / vb. Negations sit on an operand ((-x)*y == -(x*y) exactly); subtractions are additions of
// the negated operand.
__device__ __forceinline__ void geom_cost_at2(const Args &a, const RefConst &rc, const ViewConst &va, const ViewConst &vb, int la, int lb,
                                              const GeomPoint P, float xf, float yf, float &g0, float &g1) {
	const apd_camera &s0 = va.cam, &s1 = vb.cam;
#define PK(f) pk2(s0.f, s1.f)
	const f32x2 Px = pk2(P.x, P.x), Py = pk2(P.y, P.y), Pz = pk2(P.z, P.z);
	const f32x2 tx = add2(PK(t[0]), fma2(PK(R[2]), Pz, fma2(PK(R[0]), Px, mul2(PK(R[1]), Py))));
	const f32x2 ty = add2(PK(t[1]), fma2(PK(R[5]), Pz, fma2(PK(R[3]), Px, mul2(PK(R[4]), Py))));
	const f32x2 tz = add2(PK(t[2]), fma2(PK(R[8]), Pz, fma2(PK(R[6]), Px, mul2(PK(R[7]), Py))));
	float d0, d1;
	unpk2(fma2(PK(K[8]), tz, fma2(PK(K[6]), tx, mul2(PK(K[7]), ty))), d0, d1);
	const f32x2 rd = pk2(rcpf(d0), rcpf(d1));
	const f32x2 SX = mul2(fma2(PK(K[2]), tz, fma2(PK(K[0]), tx, mul2(PK(K[1]), ty))), rd);
	const f32x2 SY = mul2(fma2(PK(K[5]), tz, fma2(PK(K[3]), tx, mul2(PK(K[4]), ty))), rd);
	float sx0, sx1, sy0, sy1;
	unpk2(SX, sx0, sx1); unpk2(SY, sy0, sy1);
	const float sd0 = tex2DLayered<float>(a.depth_tex, (float)(int)sx0 + 0.5f, (float)(int)sy0 + 0.5f, la);
	const float sd1 = tex2DLayered<float>(a.depth_tex, (float)(int)sx1 + 0.5f, (float)(int)sy1 + 0.5f, lb);
	const f32x2 SD = pk2(sd0, sd1);
	const f32x2 rsK0 = pk2(rcpf(s0.K[0]), rcpf(s1.K[0])), rsK4 = pk2(rcpf(s0.K[4]), rcpf(s1.K[4]));
	const f32x2 Y0 = mul2(mul2(SD, add2(SX, pk2(-s0.K[2], -s1.K[2]))), rsK0);
	const f32x2 Y1 = mul2(mul2(SD, add2(SY, pk2(-s0.K[5], -s1.K[5]))), rsK4);
	const f32x2 Qx = add2(PK(c[0]), fma2(PK(R[6]), SD, fma2(PK(R[0]), Y0, mul2(PK(R[3]), Y1))));
	const f32x2 Qy = add2(PK(c[1]), fma2(PK(R[7]), SD, fma2(PK(R[1]), Y0, mul2(PK(R[4]), Y1))));
	const f32x2 Qz = add2(PK(c[2]), fma2(PK(R[8]), SD, fma2(PK(R[2]), Y0, mul2(PK(R[5]), Y1))));
#undef PK
	const float *R = rc.cam.R; const float *K = rc.cam.K; const float *t = rc.cam.t;
#define BC(x) pk2(x, x)
	const f32x2 ux = add2(BC(t[0]), fma2(BC(R[2]), Qz, fma2(BC(R[0]), Qx, mul2(BC(R[1]), Qy))));
	const f32x2 uy = add2(BC(t[1]), fma2(BC(R[5]), Qz, fma2(BC(R[3]), Qx, mul2(BC(R[4]), Qy))));
	const f32x2 uz = add2(BC(t[2]), fma2(BC(R[8]), Qz, fma2(BC(R[6]), Qx, mul2(BC(R[7]), Qy))));
	float b0, b1;
	unpk2(fma2(BC(K[8]), uz, fma2(BC(K[6]), ux, mul2(BC(K[7]), uy))), b0, b1);
	const f32x2 nrb = pk2(-rcpf(b0), -rcpf(b1));
	const f32x2 DC = fma2(fma2(BC(K[2]), uz, fma2(BC(K[0]), ux, mul2(BC(K[1]), uy))), nrb, BC(xf));
	const f32x2 DR = fma2(fma2(BC(K[5]), uz, fma2(BC(K[3]), ux, mul2(BC(K[4]), uy))), nrb, BC(yf));
#undef BC
	float e0, e1;
	unpk2(fma2(DC, DC, mul2(DR, DR)), e0, e1);
	g0 = (sd0 == 0.0f) ? 3.0f : fminf(sqrtaf(e0), 3.0f);
	g1 = (sd1 == 0.0f) ? 3.0f : fminf(sqrtaf(e1), 3.0f);
}

// MINB = resident blocks per SM the register allocation is sized for
template <bool DO14, bool DO15, int MINB>
__global__ void __launch_bounds__(kSqNT, MINB) k_sweep_q(const Args a, int *work, const int tiles_x, const int nchunks) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RefConst *sr = reinterpret_cast<RefConst *>(smem_raw);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int S = a.S, W = a.W;
	float *wbase = reinterpret_cast<float *>(sv + S) + (size_t)warp * kSqWarpWords;
	{
		const int nv = S * (int)(sizeof(ViewConst) / 4);
		const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views); uint32_t *s = reinterpret_cast<uint32_t *>(sv);
		for (int i = tid; i < nv; i += kSqNT) s[i] = g[i];
		const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.ref); uint32_t *srr = reinterpret_cast<uint32_t *>(sr);
		for (int i = tid; i < (int)(sizeof(RefConst) / 4); i += kSqNT) srr[i] = gr[i];
	}
	__syncthreads();
	const int ql = lane & 3, pq = lane >> 2;
	const unsigned qmask = 0xFu << (lane & ~3);
	float *refc = wbase + pq;                                  // [kSqRef][8]
	float *prof = wbase + kSqRef * kSqPix + pq;                // [kSqProf][8]
	float *near15 = wbase + (kSqRef + kSqProf) * kSqPix + pq;  // [kSqNear][8]
	const RefConst &rc = *sr;
	const float inv36 = a.inv_w[0];
	const int rad = a.weak_peak_radius;
	const int R = min(29, max(rad + 1, 5));                    // centre window half-width (k_sweep)
	const int nA = 2 * R + 1, nSide = 29 - R;
	const int last15 = R + 6;                                  // item of the last step LocalRefine reads (k = +5)
	const float kNaN = __int_as_float(0x7fc00000);

	// the warp's queue of pixels: chunk id and cursor into its 32 pixels
	int chunk = 0, cursor = 32; bool exhausted = false;
	// quad state (replicated in the four lanes)
	bool busy = false, on14 = false, on15 = false, decided = true, group_open = false, want14 = false;
	int px = 0, py = 0, g = 0; size_t center = 0;
	float xf = 0.f, yf = 0.f, inv_wn = 0.f, c_in = 3.0f;
	SweepCtx c; c.pl = make_float4(0.f, 0.f, 1.f, 1.f); c.depth = 1.0f; c.weight_normal = 0.f; c.kb = 1.f; c.disp = 1.f; c.valid = 0;
	uint32_t act = 0u, m = 0u; VW vw; vw.lo = 0ull; vw.hi = 0ull;
	uint8_t early = APD_UNKNOWN;
	// this lane's item of the open group
	float4 t = c.pl; bool need = false, in_range = false, store15 = false; int pi = 0, kk = 0;
	float acc14 = 0.f, acc15 = 0.f;
	GeomPoint gp; gp.x = gp.y = gp.z = 0.f;

#pragma unroll 1
	for (;;) {
#pragma unroll 1
		for (int pass = 0; pass < 2; ++pass) {
			// ---- close the finished group, classify, open the next group (quad-uniform)
#pragma unroll 1
			while (busy && m == 0u) {
				if (group_open) {
					if (pi > 0) { if (want14) prof[pi * kSqPix] = in_range ? ((2.0f > acc14 * inv_wn) ? acc14 * inv_wn : 2.0f) : 2.0f; }   // OpenCV MIN(2.0f, p_cost), APD.cu:2082
					else if (pi == 0) prof[0] = acc14;                                                                                 // cost of the current depth
					if (store15) near15[(kk + 5) * kSqPix] = in_range ? acc15 * inv_wn : kNaN;
					__syncwarp(qmask);
					if (DO14 && !decided) {
						// the rules of k_sweep, entry by entry in sweep order (APD.cu:2092-2143 decided early, exactly)
#pragma unroll 1
						for (int l = 0; l < 4 && !decided; ++l) {
							const int step = 4 * g + l - 1;
							if (step < 0) continue;
							if (step > 58) break;
							const bool phaseA = step < nA, right = !phaseA && step < nA + nSide;
							const int i = phaseA ? (30 - R + step) : right ? (30 + R + 1 + (step - nA)) : (30 - R - 1 - (step - nA - nSide));
							const float pc = prof[i * kSqPix];
							if (phaseA) {
								if (step == nA - 1) {
									for (int j = max(2, 30 - rad); j <= min(58, 30 + rad); ++j) {
										const float cj = prof[j * kSqPix];
										if (prof[(j - 1) * kSqPix] > cj && prof[(j + 1) * kSqPix] > cj && cj < c_in) c_in = cj;
									}
									if (c_in > 0.5f) { decided = true; early = APD_WEAK; }
								}
							} else if (right) {
								const int j = i - 1;
								const float cj = prof[j * kSqPix];
								if (j <= 58 && prof[(j - 1) * kSqPix] > cj && pc > cj && cj < c_in) { decided = true; early = APD_WEAK; }
							} else {
								const int j = i + 1;
								const float cj = prof[j * kSqPix];
								if (j >= 2 && pc > cj && prof[(j + 1) * kSqPix] > cj && cj <= c_in) { decided = true; early = APD_WEAK; }
							}
						}
					}
					group_open = false; ++g;
					const bool finished = (g == 15) || ((!on14 || decided) && (!(DO15 && on15) || 4 * g > last15));
					if (finished) {
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
							if (on14 && !decided) {   // peak analysis, APD.cu:2092-2143
								int peak_count = 0, min_peak = 0; float min_cost = 2.0f;
								unsigned long long peaks = 0ull;
								for (int i = 2; i < 59; ++i) {
									const float ci = prof[i * kSqPix];
									if (prof[(i - 1) * kSqPix] > ci && prof[(i + 1) * kSqPix] > ci) {
										peaks |= 1ull << i; peak_count++;
										if (ci < min_cost) { min_peak = i; min_cost = ci; }
									}
								}
								if (abs(min_peak - 30) > rad || prof[min_peak * kSqPix] > 0.5f) out = APD_WEAK;
								else if (peak_count == 1) out = (prof[min_peak * kSqPix] <= 0.15f) ? APD_STRONG : APD_WEAK;
								else {
									float var = 0.0f;
									for (int i = 2; i < 59; ++i) if (((peaks >> i) & 1ull) && i != min_peak) { const float dd = prof[i * kSqPix] - min_cost; var = fmaf(dd, dd, var); }
									var = sqrtaf(var) * rcpf((float)(peak_count - 1));
									out = (var > 0.2f) ? APD_STRONG : APD_WEAK;
								}
							}
							if (ql == 0) a.states[center] = out;
						}
						busy = false;
						break;
					}
				}
				// open group g: lane ql takes item 4g + ql
				{
					const int item = 4 * g + ql;
					want14 = DO14 && !decided;
					need = false; in_range = false; store15 = false; pi = -1; kk = 0;
					acc14 = 0.f; acc15 = 0.f;
					t = c.pl;
					if (item == 0) {
						if (DO15 && on15) {
							// APD.cu:2173-2182: the compiler hoisted normal.z * depth out of the view loop there as a rounded product
							float X0, X1; backproject(rc, xf, yf, c.depth, X0, X1);
							t.w = -((c.depth * t.z) + fmaf(X0, t.x, X1 * t.y));
							need = true; in_range = true; pi = 0;
						}
					} else if (item <= 59) {
						const int step = item - 1;
						const bool phaseA = step < nA, right = !phaseA && step < nA + nSide;
						const int i = phaseA ? (30 - R + step) : right ? (30 + R + 1 + (step - nA)) : (30 - R - 1 - (step - nA - nSide));
						kk = i - 30;
						const bool want15 = DO15 && on15 && (kk >= -5 && kk <= 5);
						const float d = c.kb * rcpf(c.disp + (float)kk);
						in_range = !(d < a.depth_min || d > a.depth_max);
						need = in_range && (want14 || want15);
						t.w = plane_offset(rc, xf, yf, d, t.x, t.y, t.z);
						pi = i; store15 = want15;
					}
					if (a.geom && need) gp = geom_point(rc, t, xf, yf);
					group_open = true;
					m = (__ballot_sync(qmask, need) & qmask) ? act : 0u;
				}
			}
			if (pass == 1) break;
			// ---- quads without a pixel take the next pixels of the warp's chunk
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
							busy = true; decided = !on14; early = APD_UNKNOWN; c_in = 3.0f; g = 0; m = 0u; group_open = false;
						} else if (DO14 && ql == 0) a.states[center] = APD_UNKNOWN;
					}
				}
			}
		}
		if (!__any_sync(0xffffffffu, busy)) break;

		// ---- one evaluation step: this lane's plane against the next two sampled views of its quad
		int v0 = 0, v1 = 0; bool s0 = false, s1 = false;
		if (busy && m != 0u) {
			v0 = __ffs(m) - 1; m &= m - 1u; s0 = need;
			if (m != 0u) { v1 = __ffs(m) - 1; m &= m - 1u; s1 = need; }
		}
		const ViewConst &vc0 = sv[v0], &vc1 = sv[v1];
		const Homog2 H = make_homography2(rc, vc0, vc1, t, t);
		bool w0, w1;
		{
			float x0, y0, x1, y1;
			project2(H, xf, yf, x0, y0, x1, y1);
			w0 = s0 && !(x0 >= vc0.wf || x0 < 0.0f || y0 >= vc0.hf || y0 < 0.0f);
			w1 = s1 && !(x1 >= vc1.wf || x1 < 0.0f || y1 >= vc1.hf || y1 < 0.0f);
		}
		float c0 = kCostMax, c1 = kCostMax;
		if (__any_sync(0xffffffffu, w0 || w1)) {
			float ta = 0.f, tb = 0.f;
			wq_window2<2, kSqPix>(a.img_tex, v0 + 1, v1 + 1, H, w0, w1, px, py, inv36, refc, refc[36 * kSqPix], refc[37 * kSqPix], ta, tb);
			if (w0) c0 = ta;
			if (w1) c1 = tb;
		}
		if (a.geom) {
			if (__any_sync(0xffffffffu, s0)) {
				float g0 = 0.f, g1 = 0.f;
				geom_cost_at2(a, rc, vc0, vc1, v0 + 1, v1 + 1, gp, xf, yf, g0, g1);
				if (s0) {
					const float w = (float)vw_get(vw, v0);
					acc14 = fmaf(w, fmaf(a.geom_factor, g0, c0), acc14);                                   // APD.cu:2074-2078 / :2180
					if (DO15) { acc15 = fmaf(w, c0, acc15); acc15 = fmaf(w, a.geom_factor * g0, acc15); }   // :2217-2220
				}
				if (s1) {
					const float w = (float)vw_get(vw, v1);
					acc14 = fmaf(w, fmaf(a.geom_factor, g1, c1), acc14);
					if (DO15) { acc15 = fmaf(w, c1, acc15); acc15 = fmaf(w, a.geom_factor * g1, acc15); }
				}
			}
		} else {
			if (s0) { const float w = (float)vw_get(vw, v0); acc14 = fmaf(w, c0, acc14); if (DO15) acc15 = fmaf(w, c0, acc15); }
			if (s1) { const float w = (float)vw_get(vw, v1); acc14 = fmaf(w, c1, acc14); if (DO15) acc15 = fmaf(w, c1, acc15); }
		}
	}
}

// ------------------------------------------------------------------------------------------------
template <bool DO14, bool DO15, int MINB>
static cudaError_t launch_sweep_q_t(cudaStream_t st, const Args &a, int *work, int num_sms, size_t smem) {
	cudaError_t e = cudaFuncSetAttribute(k_sweep_q<DO14, DO15, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_q<DO14, DO15, MINB>, kSqNT, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) per_sm = 1;
	const int tiles_x = (a.W + kSqTileW - 1) / kSqTileW, tiles_y = (a.H + kSqTileH - 1) / kSqTileH;
	k_sweep_q<DO14, DO15, MINB><<<num_sms * per_sm, kSqNT, smem, st>>>(a, work, tiles_x, tiles_x * tiles_y);
	return cudaGetLastError();
}
// mode 0: K14 only, 1: K15 only, 2: K14+K15 fused (as launch_sweep)
cudaError_t launch_sweep_q(cudaStream_t st, const Args &a, int mode, int num_sms) {
	const size_t smem = sizeof(RefConst) + (size_t)a.S * sizeof(ViewConst) + (size_t)(kSqNT / 32) * kSqWarpWords * 4;
	int *work = a.wctrl + 3 + mode;                        // wctrl[3..5]: zeroed at the start of every run
	static const int minb = [] { const char *e = getenv("APD_SQ_BLOCKS"); return e ? atoi(e) : 4; }();
#define SQ(A, B) (minb == 3 ? launch_sweep_q_t<A, B, 3>(st, a, work, num_sms, smem) : launch_sweep_q_t<A, B, 4>(st, a, work, num_sms, smem))
	return mode == 0 ? SQ(true, false) : mode == 1 ? SQ(false, true) : SQ(true, true);
#undef SQ
}

}  // namespace apd
