// sm_100a kernels of the all-pixels / strong-pixel part of the PatchMatch path:
//   k_setup_views      per-view constants (hoists APD.cu:305-331 out of every NCC call)
//   k_pad_ref          reference image -> pitch-linear copy with replicated border (TMA/smem source)
//   k_rng_seed         K1  InitRandomStates           APD.cu:791-804
//   k_init_planes      K5  RandomInitialization       APD.cu:806-835 (+616-693)
//   k_strong           K6/K7 Black/RedPixelUpdateStrong APD.cu:1547-1585 -> :982-1321 -> :837-890
//   k_depth_normal     K11 GetDepthandNormal          APD.cu:1587-1602
//   k_median           K12/K13 Black/RedPixelFilterStrong APD.cu:1604-1748
//   k_sweep            K14 DepthToWeak + K15 LocalRefine (fused)  APD.cu:1990-2144, 2146-2232 -- FIRST design, selectable with
//                      APD_SWEEP_IMPL=old for A/B timing; the shipped depth sweep is k_sweep_q (apd_kernels_sweepq.cu)
// Design (not a translation):
//   * every NCC evaluation is fetched by the four lanes of a quad (each takes a 3x3 quadrant of the 6x6 window, so a
//     texture instruction covers 2x2 clusters of neighbouring taps), staged through a per-warp shared-memory slab, and
//     accumulated by its owner lane in the reference's order (ncc6_quad, apd_device.cuh);
//   * the reference window lives in shared memory (one tile + halo per block, loaded by TMA from a replicated-border
//     copy of the image) and its sum / sum of squares are computed once per pixel;
//   * source views are layers of ONE layered texture (uniform handle, per-lane layer index), so lanes walk their own
//     lists of sampled views;
//   * all camera algebra that does not depend on the hypothesis is precomputed per view and read from shared memory;
//     the 8xS cost matrices and the K14 profile sit in a slab pool indexed by resident block (L1/L2, not shared memory,
//     which would come out of the L1 the texture fetches live on); the RNG state is loaded/stored once per kernel;
//   * work that cannot change the result is skipped exactly (zero-weight views, hypotheses that can no longer win,
//     profile entries the classification never reads);
//   * no host synchronisation between launches.
#include <curand_kernel.h>
#include <cuda.h>
#include <cstdlib>
#include "apd_device.cuh"

namespace apd {

// ------------------------------------------------------------------------------------------------
__global__ void k_setup_views(const apd_camera *cams, int S, ViewConst *vout, RefConst *rout, float *inv_w_out) {
	const int v = threadIdx.x;
	const apd_camera rc = cams[0];
	if (v == 0) {
		RefConst r; r.cam = rc; r.rK0 = rcpf(rc.K[0]); r.rK4 = rcpf(rc.K[4]); r.kk = rc.K[0] * r.rK4;
		*rout = r;
		// 1/sum(weights) of the 6x6 and 3x3 windows: the reference accumulates 36 (9) times 1.0f and
		// takes MUFU.RCP of the sum at run time (APD.cu:592, :488)
		float w36 = 0.f, w9 = 0.f;
		for (int i = 0; i < 36; ++i) w36 += 1.0f;
		for (int i = 0; i < 9; ++i) w9 += 1.0f;
		inv_w_out[0] = rcpf(w36); inv_w_out[1] = rcpf(w9);
	}
	if (v >= S) return;
	const apd_camera sc = cams[v + 1];
	ViewConst o;
	// -C = R^T t, accumulated as FMUL(middle) / FFMA / FFMA (APD.cu:307-312)
	float a[3], b[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		a[k] = fmaf(rc.R[6 + k], rc.t[2], fmaf(rc.R[k], rc.t[0], rc.R[3 + k] * rc.t[1]));
		b[k] = fmaf(sc.R[6 + k], sc.t[2], fmaf(sc.R[k], sc.t[0], sc.R[3 + k] * sc.t[1]));
	}
	float C[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) C[k] = b[k] - a[k];   // ref_C - src_C
#pragma unroll
	for (int r = 0; r < 3; ++r) {
#pragma unroll
		for (int c = 0; c < 3; ++c)
			o.Rrel[3 * r + c] = fmaf(sc.R[3 * r + 2], rc.R[3 * c + 2], fmaf(sc.R[3 * r + 0], rc.R[3 * c + 0], sc.R[3 * r + 1] * rc.R[3 * c + 1]));
		o.trel[r] = fmaf(sc.R[3 * r + 2], C[2], fmaf(sc.R[3 * r + 0], C[0], sc.R[3 * r + 1] * C[1]));
	}
	o.K0 = sc.K[0]; o.K2 = sc.K[2]; o.K4 = sc.K[4]; o.K5 = sc.K[5]; o.K8 = sc.K[8];
	o.wf = (float)sc.width; o.hf = (float)sc.height;
	{   // APD.cu:2037-2042
		float d0 = rc.c[0] - sc.c[0], d1 = rc.c[1] - sc.c[1], d2 = rc.c[2] - sc.c[2];
		o.baseline = sqrtaf(fmaf(d2, d2, fmaf(d0, d0, d1 * d1)));
	}
	o.cam = sc;
	o.pad_ = 0.0f;
	vout[v] = o;
}

__global__ void k_pad_ref(const float *src, int W, int H, int src_pitch, float *dst, int dst_pitch, int rows) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= dst_pitch || y >= rows) return;
	const int sx = min(max(x - kRefPad, 0), W - 1), sy = min(max(y - kRefPad, 0), H - 1);
	dst[(size_t)y * dst_pitch + x] = src[(size_t)sy * src_pitch + sx];
}

// ------------------------------------------------------------------------------------------------
// K1. curand_init(seed, subsequence = y, offset = x) for every pixel. The subsequence skip-ahead is
// done once per row chunk; inside a chunk the state advances by one generator step per pixel,
// which is what `offset` means for XORWOW.
constexpr int kRngChunk = 32;
__global__ void k_rng_seed(uint2 *rng, int W, int H, unsigned long long seed) {
	const int chunks = (W + kRngChunk - 1) / kRngChunk;
	const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long)chunks * H) return;
	const int y = (int)(gid / chunks), x0 = (int)(gid % chunks) * kRngChunk;
	curandStateXORWOW_t st;
	curand_init(seed, (unsigned long long)y, (unsigned long long)x0, &st);
	Rng r; r.v0 = st.v[0]; r.v1 = st.v[1]; r.v2 = st.v[2]; r.v3 = st.v[3]; r.v4 = st.v[4]; r.d = st.d;
	const int x1 = min(x0 + kRngChunk, W);
	for (int x = x0; x < x1; ++x) {
		rng_store(rng, (size_t)y * W + x, r);
		(void)rng_next(r);
	}
}

// ------------------------------------------------------------------------------------------------
// Shared-memory staging of the reference tile (+5 px halo) and of the per-view constants.
template <int TW, int TH>
struct TileCfg {
	// Row pitch (floats), multiple of 16 B, chosen so that the reference-window reads of a warp are bank-conflict
	// free: the checkerboard kernels (TW = 32: a warp = 16x4 pixels of one colour, rows of alternating parity) need
	// pitch = 8 (mod 16); the full-grid kernels (TW = 16: a warp = two rows of 16 pixels) need pitch = 16 (mod 32).
	static constexpr int PW = (TW == 16) ? 48 : ((TW + 2 * kHalo + 7) / 16) * 16 + 8;
	static constexpr int PH = TH + 2 * kHalo;
	static constexpr int ELEMS = PW * PH;
};

template <int TW, int TH, int NT>
__device__ __forceinline__ void load_tile(const Args &a, float *tile, int x0, int y0, int tid) {
	using C = TileCfg<TW, TH>;
	const float *src = a.ref_pad + (size_t)(y0 - kHalo + kRefPad) * a.ref_pitch + (x0 - kHalo + kRefPad);
	const int max_x = a.W + 2 * kRefPad - (x0 - kHalo + kRefPad);   // columns available to the right
	const int max_y = a.H + 2 * kRefPad - (y0 - kHalo + kRefPad);
	for (int i = tid; i < C::ELEMS; i += NT) {
		const int ly = i / C::PW, lx = i - ly * C::PW;
		float v = 0.f;
		if (lx < max_x && ly < max_y) v = src[(size_t)ly * a.ref_pitch + lx];
		tile[i] = v;
	}
}

__device__ __forceinline__ void load_views(const Args &a, ViewConst *sv, RefConst *sr, int tid, int nt) {
	const int n = a.S * (int)(sizeof(ViewConst) / 4);
	const uint32_t *g = reinterpret_cast<const uint32_t *>(a.views);
	uint32_t *s = reinterpret_cast<uint32_t *>(sv);
#pragma unroll 1
	for (int i = tid; i < n; i += nt) s[i] = g[i];
	const uint32_t *gr = reinterpret_cast<const uint32_t *>(a.ref);
	uint32_t *srr = reinterpret_cast<uint32_t *>(sr);
	for (int i = tid; i < (int)(sizeof(RefConst) / 4); i += nt) srr[i] = gr[i];
}

// ---- TMA staging of the reference tile ------------------------------------------------------------
// The tile (+5 px halo) is one box of a 2-D tensor map over `ref_pad` (the reference image with a
// replicated 8-px border, so clamp-to-edge needs no per-element bounds test); the part of a box that
// overhangs the padded image is zero-filled by the hardware and never read. One elected thread arms an
// mbarrier with the byte count and issues cp.async.bulk.tensor; every thread then waits on the barrier.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_barrier_init(uint64_t *mbar, int tid) {
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
}
__device__ __forceinline__ void tma_load_tile(const CUtensorMap *tmap, float *tile, uint64_t *mbar, int cx, int cy, uint32_t bytes, int tid, uint32_t parity) {
	if (tid == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		             ::"r"(smem_u32(tile)), "l"(tmap), "r"(cx), "r"(cy), "r"(smem_u32(mbar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------
// K5. Full grid, 16x16 pixel tiles.
constexpr int kFullTW = 16, kFullTH = 16, kFullNT = 256;

template <bool FIRST>
__global__ void __launch_bounds__(kFullNT) k_init_planes(const Args a) {
	using C = TileCfg<kFullTW, kFullTH>;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float *tile = reinterpret_cast<float *>(smem_raw);
	float *patch = tile + C::ELEMS;
	RefConst *sr = reinterpret_cast<RefConst *>(patch + (kFullNT / 32) * kPatchFloats);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	float *cm = reinterpret_cast<float *>(sv + a.S);          // [S][NT] costs of this pixel
	const int tid = threadIdx.y * kFullTW + threadIdx.x;
	const int x0 = blockIdx.x * kFullTW, y0 = blockIdx.y * kFullTH;
	load_tile<kFullTW, kFullTH, kFullNT>(a, tile, x0, y0, tid);
	load_views(a, sv, sr, tid, kFullNT);
	__syncthreads();
	QuadCtx qc = make_quad_ctx(patch, tid);
	const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
	const int lx = threadIdx.x, ly = threadIdx.y;
	set_ref_sums(qc, tile, C::PW, lx, ly);
	const bool alive = px < a.W && py < a.H;                 // dead lanes stay: they fetch for their quad
	const size_t center = (size_t)py * a.W + px;
	const RefConst &rc = *sr;
	const float xf = (float)px, yf = (float)py;
	const int S = a.S;
	const float inv36 = a.inv_w[0];
	if (FIRST) {
		float4 pl = make_float4(0.f, 0.f, 1.f, 1.f);
		if (alive) {
			Rng rng = rng_load(a.rng, center);
			// GenerateRandomPlaneHypothesis, APD.cu:276-282
			const float depth = fmaf(rng_uniform(rng), a.depth_max - a.depth_min, a.depth_min);
			pl = random_normal(rc, xf, yf, rng, depth);
			pl.w = plane_offset(rc, xf, yf, depth, pl.x, pl.y, pl.z);
			rng_store(a.rng, center, rng);
			a.planes[center] = pl;
		}
		// ComputeMultiViewInitialCostandSelectedViews, APD.cu:616-662
		int valid = 0;
#pragma unroll 1
		for (int v = 0; v < S; ++v) {
			const float c = ncc6_quad<4, true>(qc, a.img_tex, v + 1, make_homography(rc, sv[v], pl), sv[v], alive, tile, C::PW, lx, ly, px, py, inv36);
			cm[v * kFullNT + tid] = c;
			if (c < kCostMax) valid++;
		}
		if (!alive) return;
		const int top_k = min(valid, a.top_k);
		uint32_t bits = 0u;
		float cost = kCostMax;
		if (top_k > 0) {
			// the top_k smallest costs in ascending order == prefix of the reference's insertion sort
			uint32_t taken = 0u;
			float sum = 0.0f, last = 0.0f;
			for (int r = 0; r < top_k; ++r) {
				int best = -1; float bv = 0.f;
				for (int v = 0; v < S; ++v) {
					if ((taken >> v) & 1u) continue;
					const float c = cm[v * kFullNT + tid];
					if (best < 0 || c < bv) { best = v; bv = c; }
				}
				taken |= 1u << best;
				sum += bv; last = bv;
			}
			for (int v = 0; v < S; ++v) if (cm[v * kFullNT + tid] <= last) bits |= 1u << v;
			cost = sum * rcpf((float)top_k);
		}
		a.sel_views[center] = bits;
		a.costs[center] = cost;
	} else {
		// prior (world normal, depth) -> plane in the reference camera frame, APD.cu:827-832
		float4 pl = make_float4(0.f, 0.f, 1.f, 1.f);
		uint32_t bits = 0u;
		if (alive) {
			const float4 in = a.planes[center];
			const float *R = rc.cam.R;
			pl.x = fmaf(in.z, R[2], fmaf(in.x, R[0], in.y * R[1]));
			pl.y = fmaf(in.z, R[5], fmaf(in.x, R[3], in.y * R[4]));
			pl.z = fmaf(in.z, R[8], fmaf(in.x, R[6], in.y * R[7]));
			pl.w = plane_offset(rc, xf, yf, in.w, pl.x, pl.y, pl.z);
			a.planes[center] = pl;
			bits = a.sel_views[center];
		}
		// ComputeMultiViewInitialCost, APD.cu:664-693 (incl. the unSetBit quirk :47-50); each lane walks its own
		// selected views in ascending order
		int count = 0; float sum = 0.0f;
		uint32_t m = bits & ((S >= 32) ? 0xffffffffu : ((1u << S) - 1u));
#pragma unroll 1
		while (__any_sync(0xffffffffu, m != 0u)) {
			const bool want = m != 0u;
			const int v = want ? (__ffs(m) - 1) : 0;
			m &= m - 1u;
			const float c = ncc6_quad<4, true>(qc, a.img_tex, v + 1, make_homography(rc, sv[v], pl), sv[v], want, tile, C::PW, lx, ly, px, py, inv36);
			if (want) {
				if (c < kCostMax) { count++; sum += c; }
				else bits &= (0xFFFFFFFEu << v);
			}
		}
		if (!alive) return;
		a.sel_views[center] = bits;
		a.costs[center] = (count == 0) ? kCostMax : sum * rcpf((float)count);
	}
}

// ------------------------------------------------------------------------------------------------
// K6/K7. One colour of the checkerboard per launch; a block owns a 32x16 pixel tile (256 pixels of
// that colour); a warp covers a 16x4 patch so that its texture footprint stays compact.
constexpr int kHalfTW = 32, kHalfTH = 8;      // 128 pixels of one colour per block (4 warps of 16x4)

__device__ __forceinline__ void half_pixel(int tid, int x0, int y0, int color, int &px, int &py, int &lx, int &ly) {
	// Lanes 4q..4q+3 (one texture quad) own the four same-colour pixels of a 4x2 block, a diamond
	// (x,y) (x+2,y) (x+1,y+1) (x+3,y+1): measured on B200 (tools/tex_probe3.cu) a quad whose four
	// bilinear footprints stay within ~4x2 texels streams at the full 4 fetches/clk/SM, whereas four
	// pixels of one row at stride 2 reach only 66 % of that. 8 quads = 16x4 pixels per warp.
	const int warp = tid >> 5, lane = tid & 31;
	const int q = lane >> 2, k = lane & 3;
	ly = (warp >> 1) * 4 + (q >> 2) * 2 + (k >> 1);
	py = y0 + ly;
	const int xb = (warp & 1) * 16 + (q & 3) * 4 + 2 * (k & 1);
	lx = xb + ((x0 + xb + py + color) & 1);
	px = x0 + lx;
}

// smallest stored cost along a propagation arm (strict <, first wins), APD.cu:1022-1199
struct ArmMin { float c; int pos; };
__device__ __forceinline__ void arm_try(ArmMin &m, const float *costs, int pos) {
	const float c = costs[pos];
	if (c < m.c) { m.c = c; m.pos = pos; }
}

template <int NT>
__global__ void __launch_bounds__(NT, 4) k_strong(const Args a, const int iter, const int color, const __grid_constant__ CUtensorMap tmap) {
	using C = TileCfg<kHalfTW, kHalfTH>;
	extern __shared__ __align__(128) unsigned char smem_raw[];      // no static smem in this kernel: the TMA destination must be 128-B aligned
	float *tile = reinterpret_cast<float *>(smem_raw);
	uint64_t &tile_bar = *reinterpret_cast<uint64_t *>(tile + C::ELEMS);
	float *xchg = tile + C::ELEMS + 4;                              // [NT/32][36][32] warped source patches
	RefConst *sr = reinterpret_cast<RefConst *>(xchg + (NT / 32) * kPatchFloats);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.x;
	const int x0 = blockIdx.x * kHalfTW, y0 = blockIdx.y * kHalfTH;
	int *s_slot = reinterpret_cast<int *>(&tile_bar + 1);         // 8 spare bytes behind the mbarrier
	tma_barrier_init(&tile_bar, tid);
	slab_acquire(a, s_slot, tid);
	load_views(a, sv, sr, tid, NT);
	__syncthreads();
	// [9*S][NT] per block: 8xS cost matrix + S probabilities, in this block's slab of the L1/L2-resident pool
	float *cm = slab_ptr(a, s_slot);
	tma_load_tile(&tmap, tile, &tile_bar, x0 - kHalo + kRefPad, y0 - kHalo + kRefPad, C::ELEMS * 4, tid, 0);
	QuadCtx qc = make_quad_ctx(xchg, tid);
	int px, py, lx, ly;
	half_pixel(tid, x0, y0, color, px, py, lx, ly);
	set_ref_sums(qc, tile, C::PW, lx, ly);
	const int W = a.W, H = a.H, S = a.S;
	const int center = py * W + px;
	// threads without a pixel to update stay: their lanes still fetch for the other lanes of their quad
	const bool alive = px < W && py < H && py < a.half_rows && a.states[center] != APD_WEAK;
	const RefConst &rc = *sr;
	const float xf = (float)px, yf = (float)py;
	const float inv36 = a.inv_w[0];
	const float *costs = a.costs;
	float *cmt = cm + tid;
#define CM(k, v) cmt[((k) * S + (v)) * NT]
#define PROB(v) cmt[(8 * S + (v)) * NT]
#define NCC(v, plane, want) ncc6_quad<4, true>(qc, a.img_tex, (v) + 1, make_homography(rc, sv[v], plane), sv[v], want, tile, C::PW, lx, ly, px, py, inv36)

	// ---- adaptive checkerboard sampling: 8 candidates (0 up_near 1 up_far 2 down_near 3 down_far
	//      4 left_near 5 left_far 6 right_near 7 right_far)
	int pos[8]; unsigned flags = 0u;
#pragma unroll
	for (int k = 0; k < 8; ++k) pos[k] = 0;
	if (alive) {
		ArmMin m;
		if (py > 2) { flags |= 2u; m.pos = center - 3 * W; m.c = costs[m.pos];
			for (int i = 1; i < 11; ++i) if (py > 2 + 2 * i) arm_try(m, costs, center - 3 * W - 2 * i * W);
			pos[1] = m.pos; }
		if (py < H - 3) { flags |= 8u; m.pos = center + 3 * W; m.c = costs[m.pos];
			for (int i = 1; i < 11; ++i) if (py < H - 3 - 2 * i) arm_try(m, costs, center + 3 * W + 2 * i * W);
			pos[3] = m.pos; }
		if (px > 2) { flags |= 32u; m.pos = center - 3; m.c = costs[m.pos];
			for (int i = 1; i < 11; ++i) if (px > 2 + 2 * i) arm_try(m, costs, center - 3 - 2 * i);
			pos[5] = m.pos; }
		if (px < W - 3) { flags |= 128u; m.pos = center + 3; m.c = costs[m.pos];
			for (int i = 1; i < 11; ++i) if (px < W - 3 - 2 * i) arm_try(m, costs, center + 3 + 2 * i);
			pos[7] = m.pos; }
		if (py > 0) { flags |= 1u; m.pos = center - W; m.c = costs[m.pos];
			for (int i = 0; i < 3; ++i) {
				if (py > 1 + i && px > i) arm_try(m, costs, center - W - (1 + i) * W - (1 + i));
				if (py > 1 + i && px < W - 1 - i) arm_try(m, costs, center - W - (1 + i) * W + (1 + i));
			}
			pos[0] = m.pos; }
		if (py < H - 1) { flags |= 4u; m.pos = center + W; m.c = costs[m.pos];
			for (int i = 0; i < 3; ++i) {
				if (py < H - 2 - i && px > i) arm_try(m, costs, center + W + (1 + i) * W - (1 + i));
				if (py < H - 2 - i && px < W - 1 - i) arm_try(m, costs, center + W + (1 + i) * W + (1 + i));
			}
			pos[2] = m.pos; }
		if (px > 0) { flags |= 16u; m.pos = center - 1; m.c = costs[m.pos];
			for (int i = 0; i < 3; ++i) {
				if (px > 1 + i && py > i) arm_try(m, costs, center - 1 - (1 + i) - (1 + i) * W);
				if (px > 1 + i && py < H - 1 - i) arm_try(m, costs, center - 1 - (1 + i) + (1 + i) * W);
			}
			pos[4] = m.pos; }
		if (px < W - 1) { flags |= 64u; m.pos = center + 1; m.c = costs[m.pos];
			for (int i = 0; i < 3; ++i) {
				if (px < W - 2 - i && py > i) arm_try(m, costs, center + 1 + (1 + i) - (1 + i) * W);
				if (px < W - 2 - i && py < H - 1 - i) arm_try(m, costs, center + 1 + (1 + i) + (1 + i) * W);
			}
			pos[6] = m.pos; }
	}

	// ---- 8 x S matching costs. A missing candidate keeps the reference's partially initialised
	//      row: `float cost_array[8][32] = {2.0f}` sets [0][0] only (APD.cu:1004)
#pragma unroll 1
	for (int k = 0; k < 8; ++k) {
		const bool fl = (flags >> k) & 1u;
		float4 pl = make_float4(0.f, 0.f, 1.f, 1.f);
		if (fl) pl = a.planes[pos[k]];
#pragma unroll 1
		for (int v = 0; v < S; ++v) {
			const float c = ncc6_quad<4, false>(qc, a.img_tex, v + 1, make_homography(rc, sv[v], pl), sv[v], fl, tile, C::PW, lx, ly, px, py, inv36);
			CM(k, v) = fl ? c : ((k == 0 && v == 0) ? 2.0f : 0.0f);
		}
	}

	// ---- multi-hypothesis joint view selection, APD.cu:1203-1259
	Rng rng; VW vw; vw.lo = 0ull; vw.hi = 0ull;
	uint32_t temp_sel = 0u; float inv_wn = 0.0f;
	float best_cost = 0.0f; int best_k = 0;
	if (alive) {
		const float thr = 0.8 * __expf((float)(unsigned)(iter * iter) * -0.011111111380159854889f);
		const float thr_fallback = __expf((thr * thr) * -3.125f);
		uint32_t nb_bits[4]; unsigned nb_ok = 0u;
		{
			const int nb_pos[4] = {center - W, center + W, center - 1, center + 1};
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				nb_bits[i] = 0u;
				if ((flags >> (2 * i)) & 1u) { nb_ok |= 1u << i; nb_bits[i] = a.sel_views[nb_pos[i]]; }
			}
		}
		float prob_sum = 0.0f;
		for (int v = 0; v < S; ++v) {
			float prior = 0.0f;
#pragma unroll
			for (int i = 0; i < 4; ++i)
				if ((nb_ok >> i) & 1u) prior += ((nb_bits[i] >> v) & 1u) ? 0.9f : 0.1f;
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
			PROB(v) = p;
			prob_sum += p;
		}
		rng = rng_load(a.rng, center);
		{
			const float inv = rcpf(prob_sum);
			float cum = 0.0f;
#pragma unroll 1
			for (int v = 0; v < S; ++v) { cum = fmaf(inv, PROB(v), cum); PROB(v) = cum; }   // TransformPDFToCDF
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
		// final costs of the eight candidates and their minimum. Rolled on purpose: this runs once per pixel, and unrolled (8 x S with the
		// nibble extraction of the weights) it was 2 900 of the kernel's 7 600 SASS instructions - instruction-cache footprint the hot loops pay for
		best_cost = 0.0f; best_k = 0;             // FindMinCostIndex: `<=`, last minimum wins (APD.cu:29-40)
#pragma unroll 1
		for (int k = 0; k < 8; ++k) {
			float acc = 0.0f;
#pragma unroll 1
			for (int v = 0; v < S; ++v) { const int w = vw_get(vw, v); if (w > 0) acc = fmaf((float)w, CM(k, v), acc); }
			const float fck = acc * inv_wn;
			if (k == 0 || fck <= best_cost) { best_cost = fck; best_k = k; }
		}
	}

	// ---- current hypothesis under the sampled views (views with weight 0 contribute exactly 0)
	float4 pl_now = make_float4(0.f, 0.f, 1.f, 1.f);
	if (alive) pl_now = a.planes[center];
	// Each lane walks ITS OWN list of sampled views (ascending, as the reference's sum does); lanes of a warp
	// are at different views at the same time, which the layered texture (per-lane layer) permits. The warp
	// iterates max-over-lanes(#sampled views) times instead of S times.
	const uint32_t wmask = alive ? temp_sel : 0u;
	float cost_now = 0.0f, cost_stored = 0.0f;            // costs[center] = cost_now (APD.cu:1295)
	float depth_now = 1.0f;
	uint32_t sel_out = 0u; bool sel_write = false;
	float depth_rand = 1.0f, depth_pert = 1.0f, d0 = 1.0f;
	float4 n_rand = pl_now, n_pert = pl_now, n0 = pl_now;
	// i = -1: the current plane; i = 0..4: PlaneHypothesisRefinementStrong's five hypotheses (APD.cu:837-890). One loop,
	// one inlined NCC body (the instruction cache is a measured limiter of this kernel).
#pragma unroll 1
	for (int i = -1; i < 5; ++i) {
		float4 t = pl_now; float d = depth_now; bool in_range = true;
		if (i >= 0) {
			const float di = (i == 0 || i == 2) ? depth_rand : (i == 4 ? depth_pert : d0);
			t = (i == 1 || i == 2) ? n_rand : (i == 3 ? n_pert : n0);
			t.w = plane_offset(rc, xf, yf, di, t.x, t.y, t.z);
			// A hypothesis is adopted only if its depth is in range and its weighted cost is below cost_now. Costs and
			// weights are non-negative and every partial sum is rounded monotonically, so (a) an out-of-range
			// hypothesis and (b) one whose partial sum already reaches cost_now can never be adopted: their remaining
			// views are not evaluated (the reference evaluates and then discards them).
			d = plane_depth(rc, t, xf, yf);
			in_range = d >= a.depth_min && d <= a.depth_max;
		}
		float acc = 0.0f;
		uint32_t m = in_range ? wmask : 0u;
#pragma unroll 1
		while (__any_sync(0xffffffffu, m != 0u)) {
			const bool want = m != 0u;
			const int v = want ? (__ffs(m) - 1) : 0;
			m &= m - 1u;
			const float c = NCC(v, t, want);
			if (want) { acc = fmaf((float)vw_get(vw, v), c, acc); if (i >= 0 && acc * inv_wn >= cost_now) m = 0u; }
		}
		const float tc = acc * inv_wn;
		if (i >= 0) {
			if (in_range && tc < cost_now) { depth_now = d; pl_now = t; cost_now = tc; }
		} else {
			cost_now = tc; cost_stored = tc;
			if (alive) {
				depth_now = plane_depth(rc, pl_now, xf, yf);
				if ((flags >> best_k) & 1u) {
					int bp = pos[0];
#pragma unroll
					for (int k = 1; k < 8; ++k) if (best_k == k) bp = pos[k];
					const float4 cand = a.planes[bp];
					const float dc = plane_depth(rc, cand, xf, yf);
					if (dc >= a.depth_min && dc <= a.depth_max && best_cost < cost_now) {
						depth_now = dc; pl_now = cand; cost_now = best_cost; sel_out = temp_sel; sel_write = true;
					}
				}
				depth_rand = fmaf(rng_uniform(rng), a.depth_max - a.depth_min, a.depth_min);
				n_rand = random_normal(rc, xf, yf, rng, depth_now);
				const float lo = depth_now * (1.0f - 0.02f);
				const float span = fmaf(depth_now, 1.0f + 0.02f, -lo);
				depth_pert = fmaf(span, rng_uniform(rng), lo);     // the do/while never repeats (:860-862)
				n_pert = perturbed_normal(rc, xf, yf, pl_now, rng);
			}
			n0 = pl_now; d0 = depth_now;
		}
	}
	slab_exit(a, s_slot, NT);
	if (!alive) return;
	rng_store(a.rng, center, rng);
	vw_store(a.view_w, center, vw);
	if (sel_write) a.sel_views[center] = sel_out;
	if (a.state == APD_REFINE_INIT) {
		if ((double)cost_now < (double)cost_stored - 0.1) { a.costs[center] = cost_now; a.planes[center] = pl_now; }
		else a.costs[center] = cost_stored;
	} else {
		a.costs[center] = cost_now; a.planes[center] = pl_now;
	}
#undef CM
#undef PROB
#undef NCC
}

// ------------------------------------------------------------------------------------------------
// K11
__global__ void k_depth_normal(const Args a) {
	const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
	if (px >= a.W || py >= a.H) return;
	const size_t center = (size_t)py * a.W + px;
	const RefConst rc = *a.ref;
	const float4 pl = a.planes[center];
	const float depth = plane_depth(rc, pl, (float)px, (float)py);
	const float *R = rc.cam.R;
	float4 o;   // TransformNormal, APD.cu:374-382 (R^T n)
	o.x = fmaf(pl.z, R[6], fmaf(pl.x, R[0], pl.y * R[3]));
	o.y = fmaf(pl.z, R[7], fmaf(pl.x, R[1], pl.y * R[4]));
	o.z = fmaf(pl.z, R[8], fmaf(pl.x, R[2], pl.y * R[5]));
	o.w = depth;
	a.planes[center] = o;
}

// K12/K13: median of own depth and up to 20 STRONG opposite-colour neighbours, APD.cu:1604-1714
__global__ void k_median(const Args a, const int color) {
	const int px = blockIdx.x * blockDim.x + threadIdx.x;
	const int yy = blockIdx.y * blockDim.y + threadIdx.y;
	const int py = 2 * yy + ((px + color) & 1);
	const int W = a.W, H = a.H;
	if (px >= W || py >= H || py >= a.half_rows) return;
	const int center = py * W + px;
	if (a.states[center] == APD_WEAK) return;
	if (a.costs[center] < 0.001f) return;
	float f[21]; int n = 0;
	f[n++] = a.planes[center].w;
	const int dx[20] = {0, 0, 0, 0, 0, 0, -1, -3, -5, 1, 3, 5, 2, 2, -2, -2, -1, 1, -1, 1};
	const int dy[20] = {-1, -3, -5, 1, 3, 5, 0, 0, 0, 0, 0, 0, -1, 1, -1, 1, -2, -2, 2, 2};
#pragma unroll
	for (int k = 0; k < 20; ++k) {
		const int x = px + dx[k], y = py + dy[k];
		// the reference's bounds (APD.cu:1642-1703): identical to "inside the image" except for the
		// two (+-1, -2) taps which demand `p.y > 2` (APD.cu:1691, :1695)
		bool ok = x >= 0 && x < W && y >= 0 && y < H;
		if (dy[k] == -2) ok = ok && (py > 2);
		if (ok && a.states[y * W + x] == APD_STRONG) f[n++] = a.planes[y * W + x].w;
	}
	for (int i = 1; i < n; ++i) {   // sort_small
		const float t = f[i]; int j = i;
		for (; j >= 1 && t < f[j - 1]; --j) f[j] = f[j - 1];
		f[j] = t;
	}
	const int m = n / 2;
	a.planes[center].w = (n % 2 == 0) ? (f[m - 1] + f[m]) * 0.5f : f[m];
}

// ------------------------------------------------------------------------------------------------
// K14 + K15 fused (used by full runs). LocalRefine's 11 disparity steps are the centre of DepthToWeak's
// 61-step sweep: same planes, same NCC values (and the same geometric term), only the weighted
// accumulation differs when geom_consistency is on (APD.cu:2074-2078 vs :2217-2220). Each NCC is
// therefore evaluated once and fed to both accumulators. The two kernels touch disjoint state (K14
// writes pixel states, K15 the depth of its own pixel; neither reads what the other writes), so the
// fusion is exact. All threads stay in the loops: fetches are quad-cooperative (ncc6_quad).
constexpr int kSweepTW = 16, kSweepTH = 8, kSweepNT = 128;

template <bool DO14, bool DO15>
__global__ void __launch_bounds__(kSweepNT, 4) k_sweep(const Args a, const __grid_constant__ CUtensorMap tmap) {
	using C = TileCfg<kSweepTW, kSweepTH>;
	extern __shared__ __align__(128) unsigned char smem_raw[];      // no static smem in this kernel: the TMA destination must be 128-B aligned
	float *tile = reinterpret_cast<float *>(smem_raw);
	uint64_t &tile_bar = *reinterpret_cast<uint64_t *>(tile + C::ELEMS);
	float *patch = tile + C::ELEMS + 4;
	RefConst *sr = reinterpret_cast<RefConst *>(patch + (kSweepNT / 32) * kPatchFloats);
	ViewConst *sv = reinterpret_cast<ViewConst *>(sr + 1);
	const int tid = threadIdx.y * kSweepTW + threadIdx.x;
	const int x0 = blockIdx.x * kSweepTW, y0 = blockIdx.y * kSweepTH;
	int *s_slot = reinterpret_cast<int *>(&tile_bar + 1);         // 8 spare bytes behind the mbarrier
	tma_barrier_init(&tile_bar, tid);
	slab_acquire(a, s_slot, tid);
	load_views(a, sv, sr, tid, kSweepNT);
	__syncthreads();
	// [61][NT] cost profile (K14) in this block's slab: 31 KB of shared memory per block would leave the texture
	// fetches almost no L1
	float *prof = slab_ptr(a, s_slot);
	tma_load_tile(&tmap, tile, &tile_bar, x0 - kHalo + kRefPad, y0 - kHalo + kRefPad, C::ELEMS * 4, tid, 0);
	QuadCtx qc = make_quad_ctx(patch, tid);
	const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
	const int lx = threadIdx.x, ly = threadIdx.y;
	set_ref_sums(qc, tile, C::PW, lx, ly);
	const bool in_img = px < a.W && py < a.H;
	const size_t center = (size_t)py * a.W + px;
	const RefConst &rc = *sr;
	const float xf = (float)px, yf = (float)py;
	const float inv36 = a.inv_w[0];
	const int S = a.S;
	uint32_t bits = 0u; VW vw; vw.lo = 0ull; vw.hi = 0ull;
	SweepCtx c; c.pl = make_float4(0.f, 0.f, 1.f, 1.f); c.depth = 1.0f; c.weight_normal = 0.0f; c.kb = 1.0f; c.disp = 1.0f; c.valid = 0;
	bool has_depth = false;
	if (in_img) {
		bits = a.sel_views[center];
		vw = vw_load(a.view_w, center);
		has_depth = sweep_setup(a, rc, sv, center, bits, vw, c);
	}
	const bool border = px < 6 || py < 6 || px >= a.W - 6 || py >= a.H - 6;
	const bool on14 = DO14 && in_img && !border && has_depth && c.valid > 0;
	const bool on15 = DO15 && in_img && has_depth && c.valid > 0 && c.weight_normal != 0.0f;
	const float inv_wn = rcpf(c.weight_normal);
	const uint32_t act = bits & vw_mask(vw, S);
	float *p = prof + tid;
	float min_cost15 = 2.0f, best_depth = c.depth;
	// Sweep order: the centre window first, then outwards. The classification rules (APD.cu:2092-2143) make
	// many pixels decidable before the whole 59-entry profile exists, and the decision is the same as the
	// reference's (which always evaluates everything):
	//  * min_peak is the FIRST index of the cheapest local minimum ("peak"); the pixel is WEAK whenever that
	//    peak lies farther than weak_peak_radius from the centre, costs more than 0.5, or does not exist;
	//  * so, with c_in = cheapest peak inside the radius (needs entries 30-r-1 .. 30+r+1 only): no inside
	//    peak, or c_in > 0.5  =>  WEAK at once; otherwise the first outside peak cheaper than c_in (left side:
	//    not dearer, because the first index wins ties) => WEAK.
	// Only pixels that survive all of that need the complete profile (peak count and spread).
	const int rad = a.weak_peak_radius;
	const int R = min(29, max(rad + 1, 5));              // centre window half-width; >= 5 covers LocalRefine's steps
	const int nA = 2 * R + 1, nSide = 29 - R;
	bool decided = !on14;                                // K14 outcome already known (UNKNOWN or early WEAK)
	uint8_t early = APD_UNKNOWN;
	float c_in = 3.0f;
#pragma unroll 1
	for (int step = 0; step < 59; ++step) {
		const bool phaseA = step < nA;
		const bool right = !phaseA && step < nA + nSide;
		const int i = phaseA ? (30 - R + step) : right ? (30 + R + 1 + (step - nA)) : (30 - R - 1 - (step - nA - nSide));
		const int k = i - 30;
		const bool near = (k >= -5 && k <= 5);
		const bool want14 = DO14 && !decided;
		const bool want15 = DO15 && on15 && near;
		if (!__any_sync(0xffffffffu, want14 || want15)) { if (phaseA) continue; else break; }
		const float d = c.kb * rcpf(c.disp + (float)k);
		const bool in_range = !(d < a.depth_min || d > a.depth_max);
		const bool need = in_range && (want14 || want15);
		float4 t = c.pl;
		t.w = plane_offset(rc, xf, yf, d, t.x, t.y, t.z);
		float acc14 = 0.0f, acc15 = 0.0f;
		// per-lane list of the views that count: selected AND sampled (a zero weight multiplies the cost by 0 in
		// the reference); ascending order = the reference's summation order
		uint32_t m = need ? act : 0u;
#pragma unroll 1
		while (__any_sync(0xffffffffu, m != 0u)) {
			const bool want = m != 0u;
			const int v = want ? (__ffs(m) - 1) : 0;
			m &= m - 1u;
			const int w = vw_get(vw, v);
			float ncc = kCostMax;
			ncc = ncc6_quad<4, false>(qc, a.img_tex, v + 1, make_homography(rc, sv[v], t), sv[v], want, tile, C::PW, lx, ly, px, py, inv36);
			if (want) {
				float g = 0.0f;
				if (a.geom) g = geom_cost(a, rc, sv[v], v + 1, t, xf, yf);       // hoisting its view-independent half (geom_point) costs registers: 372 vs 368 ms at cfg3
				if (DO14) acc14 = fmaf((float)w, a.geom ? fmaf(a.geom_factor, g, ncc) : ncc, acc14);          // APD.cu:2074-2078
				if (DO15) { acc15 = fmaf((float)w, ncc, acc15); if (a.geom) acc15 = fmaf((float)w, a.geom_factor * g, acc15); }   // :2217-2220
			}
		}
		if (want15 && in_range) {
			const float tc = acc15 * inv_wn;
			if (tc < min_cost15) { min_cost15 = tc; best_depth = d; }
		}
		if (want14) {
			float pc = 2.0f;
			if (in_range) { pc = acc14 * inv_wn; pc = (2.0f > pc) ? pc : 2.0f; }      // OpenCV MIN(2.0f, p_cost)
			p[i * kSweepNT] = pc;
			if (phaseA) {
				if (step == nA - 1) {                    // centre window complete: cheapest peak within the radius
					for (int j = max(2, 30 - rad); j <= min(58, 30 + rad); ++j) {
						const float cj = p[j * kSweepNT];
						if (p[(j - 1) * kSweepNT] > cj && p[(j + 1) * kSweepNT] > cj && cj < c_in) c_in = cj;
					}
					if (c_in > 0.5f) { decided = true; early = APD_WEAK; }     // none (3.0) or too dear
				}
			} else if (right) {
				const int j = i - 1;                     // newly testable peak, outside the radius on the right
				const float cj = p[j * kSweepNT];
				if (j <= 58 && p[(j - 1) * kSweepNT] > cj && pc > cj && cj < c_in) { decided = true; early = APD_WEAK; }
			} else {
				const int j = i + 1;                     // outside on the left: wins ties (smaller index)
				const float cj = p[j * kSweepNT];
				if (j >= 2 && pc > cj && p[(j + 1) * kSweepNT] > cj && cj <= c_in) { decided = true; early = APD_WEAK; }
			}
		}
	}
	const bool full_profile = on14 && !decided;
	if (DO15) {
		// cost of the current depth (APD.cu:2173-2182): the compiler hoisted normal.z * depth out of the view
		// loop there as a rounded product, so the plane offset is -((d*nz) + fma(X0,nx,X1*ny))
		float4 t = c.pl;
		{ float X0, X1; backproject(rc, xf, yf, c.depth, X0, X1); t.w = -((c.depth * t.z) + fmaf(X0, t.x, X1 * t.y)); }
		float cost_sum = 0.0f;
		uint32_t m = on15 ? act : 0u;
#pragma unroll 1
		while (__any_sync(0xffffffffu, m != 0u)) {
			const bool want = m != 0u;
			const int v = want ? (__ffs(m) - 1) : 0;
			m &= m - 1u;
			const int w = vw_get(vw, v);
			float tc = kCostMax;
			tc = ncc6_quad<4, false>(qc, a.img_tex, v + 1, make_homography(rc, sv[v], t), sv[v], want, tile, C::PW, lx, ly, px, py, inv36);
			if (want) {
				if (a.geom) tc = fmaf(a.geom_factor, geom_cost(a, rc, sv[v], v + 1, t, xf, yf), tc);
				cost_sum = fmaf((float)w, tc, cost_sum);
			}
		}
		if (on15) {
			const float diff = fmaf(inv_wn, cost_sum, -min_cost15);   // (cost_now / weight_normal) - min_cost, one FFMA
			if ((double)diff > 0.1) a.planes[center].w = best_depth;
		}
	}
	if (DO14 && in_img) {
		uint8_t out = on14 ? early : (uint8_t)APD_UNKNOWN;
		if (full_profile) {   // peak analysis, APD.cu:2092-2143
			int peak_count = 0, min_peak = 0; float min_cost = 2.0f;
			unsigned long long peaks = 0ull;
			for (int i = 2; i < 59; ++i) {
				const float ci = p[i * kSweepNT];
				if (p[(i - 1) * kSweepNT] > ci && p[(i + 1) * kSweepNT] > ci) {
					peaks |= 1ull << i; peak_count++;
					if (ci < min_cost) { min_peak = i; min_cost = ci; }
				}
			}
			if (abs(min_peak - 30) > a.weak_peak_radius || p[min_peak * kSweepNT] > 0.5f) out = APD_WEAK;
			else if (peak_count == 1) out = (p[min_peak * kSweepNT] <= 0.15f) ? APD_STRONG : APD_WEAK;
			else {
				float var = 0.0f;
				for (int i = 2; i < 59; ++i) if (((peaks >> i) & 1ull) && i != min_peak) { const float dd = p[i * kSweepNT] - min_cost; var = fmaf(dd, dd, var); }
				var = sqrtaf(var) * rcpf((float)(peak_count - 1));
				out = (var > 0.2f) ? APD_STRONG : APD_WEAK;
			}
		}
		a.states[center] = out;
	}
	slab_exit(a, s_slot, kSweepNT);
}

// ------------------------------------------------------------------------------------------------
// launchers
static inline size_t smem_common(int S) { return sizeof(RefConst) + (size_t)S * sizeof(ViewConst); }

void launch_setup_views(cudaStream_t st, const apd_camera *cams, int S, ViewConst *v, RefConst *r, float *inv_w) {
	k_setup_views<<<1, 32, 0, st>>>(cams, S, v, r, inv_w);
}
void launch_pad_ref(cudaStream_t st, const float *src, int W, int H, int src_pitch, float *dst, int dst_pitch, int rows) {
	dim3 b(32, 8), g((dst_pitch + 31) / 32, (rows + 7) / 8);
	k_pad_ref<<<g, b, 0, st>>>(src, W, H, src_pitch, dst, dst_pitch, rows);
}
void launch_rng_seed(cudaStream_t st, const Args &a, unsigned long long seed) {
	const long long n = (long long)((a.W + kRngChunk - 1) / kRngChunk) * a.H;
	k_rng_seed<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a.rng, a.W, a.H, seed);
}
cudaError_t launch_init_planes(cudaStream_t st, const Args &a) {
	using C = TileCfg<kFullTW, kFullTH>;
	const size_t smem = C::ELEMS * 4 + (kFullNT / 32) * kPatchFloats * 4 + smem_common(a.S) + (size_t)a.S * kFullNT * 4;
	dim3 b(kFullTW, kFullTH), g((a.W + kFullTW - 1) / kFullTW, (a.H + kFullTH - 1) / kFullTH);
	if (a.state == APD_FIRST_INIT) {
		cudaFuncSetAttribute(k_init_planes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_init_planes<true><<<g, b, smem, st>>>(a);
	} else {
		cudaFuncSetAttribute(k_init_planes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_init_planes<false><<<g, b, smem, st>>>(a);
	}
	return cudaGetLastError();
}
cudaError_t launch_strong(cudaStream_t st, const Args &a, int iter, int color, const CUtensorMap *tmap) {
	using C = TileCfg<kHalfTW, kHalfTH>;
	constexpr int NT = 128;
	const size_t smem = C::ELEMS * 4 + 16 + (NT / 32) * kPatchFloats * 4 + smem_common(a.S);
	if (smem > 227 * 1024) return cudaErrorInvalidValue;
	cudaFuncSetAttribute(k_strong<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	// four resident blocks need 4 x 25.5 KB: ask for the 100 KB shared-memory configuration (the driver would take 132 KB), the
	// difference goes to the L1 the texture fetches live on
	if (4 * (smem + 1024) <= 100 * 1024) cudaFuncSetAttribute(k_strong<NT>, cudaFuncAttributePreferredSharedMemoryCarveout, 43);
	dim3 g((a.W + kHalfTW - 1) / kHalfTW, (a.H + kHalfTH - 1) / kHalfTH);
	k_strong<NT><<<g, NT, smem, st>>>(a, iter, color, *tmap);
	return cudaGetLastError();
}
void launch_depth_normal(cudaStream_t st, const Args &a) {
	dim3 b(32, 8), g((a.W + 31) / 32, (a.H + 7) / 8);
	k_depth_normal<<<g, b, 0, st>>>(a);
}
void launch_median(cudaStream_t st, const Args &a, int color) {
	dim3 b(32, 8), g((a.W + 31) / 32, ((a.H + 1) / 2 + 7) / 8);
	k_median<<<g, b, 0, st>>>(a, color);
}
// mode 0: K14 only, 1: K15 only, 2: K14+K15 fused
cudaError_t launch_sweep(cudaStream_t st, const Args &a, int mode, const CUtensorMap *tmap) {
	using C = TileCfg<kSweepTW, kSweepTH>;
	const size_t smem = C::ELEMS * 4 + 16 + (kSweepNT / 32) * kPatchFloats * 4 + smem_common(a.S);
	dim3 b(kSweepTW, kSweepTH), g((a.W + kSweepTW - 1) / kSweepTW, (a.H + kSweepTH - 1) / kSweepTH);
#define SWEEP_LAUNCH(A, B) do { cudaFuncSetAttribute(k_sweep<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
		if (4 * (smem + 1024) <= 100 * 1024) cudaFuncSetAttribute(k_sweep<A, B>, cudaFuncAttributePreferredSharedMemoryCarveout, 43); \
		k_sweep<A, B><<<g, b, smem, st>>>(a, *tmap); } while (0)
	if (mode == 0) SWEEP_LAUNCH(true, false); else if (mode == 1) SWEEP_LAUNCH(false, true); else SWEEP_LAUNCH(true, true);
#undef SWEEP_LAUNCH
	return cudaGetLastError();
}


// Tensor maps over the padded reference image, one per tile shape (box = tile + halo, row pitch of the box = TileCfg::PW).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_map(PFN_encodeTiled enc, CUtensorMap *m, const float *base, int pitch_elems, int rows, int box_w, int box_h) {
	cuuint64_t dims[2] = {(cuuint64_t)pitch_elems, (cuuint64_t)rows};
	cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 4};
	cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
	cuuint32_t estr[2] = {1, 1};
	return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}
int make_tensor_maps(const float *ref_pad, int pitch_elems, int rows, CUtensorMap *strong, CUtensorMap *sweep) {
	void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
	PFN_encodeTiled enc = (PFN_encodeTiled)fn;
	using CS = TileCfg<kHalfTW, kHalfTH>; using CW = TileCfg<kSweepTW, kSweepTH>;
	if (encode_map(enc, strong, ref_pad, pitch_elems, rows, CS::PW, CS::PH)) return -1;
	if (encode_map(enc, sweep, ref_pad, pitch_elems, rows, CW::PW, CW::PH)) return -1;
	return 0;
}

}  // namespace apd
