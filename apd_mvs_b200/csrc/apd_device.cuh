// Device-side building blocks of the B200 PatchMatch engine.
//
// Numerical contract. The reference is built with `--use_fast_math` (CMakeLists.txt:20) and its
// outputs are the result of discrete decisions on fp32 costs, so parity needs the SAME rounded
// operations, not merely the same formulas. This TU is compiled with
//   -ftz=true -prec-div=false -prec-sqrt=false -fmad=false
// so the compiler never contracts on its own: every fused multiply-add below is an explicit
// fmaf() placed where the reference's sm_100 SASS has an FFMA, every plain `*`/`+` is a rounded
// FMUL/FADD, divisions are written as multiplications by rcpf() (MUFU.RCP) exactly as
// div.approx is lowered, and sqrtf()/rsqrtf()/__expf()/__sinf()/__cosf() map to the same MUFU ops.
// Comments of the form "APD.cu:NNN" name the reference lines whose arithmetic is being matched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/apd_b200.h"

namespace apd {

constexpr int kRefPad = 9;        // replicated border of the pitch-linear reference-image copy; 9 - kHalo = 4 keeps every
                                  // TMA box origin (x0 - 5 + 9, x0 a multiple of 16) 16-byte aligned, which the hardware demands
constexpr int kHalo = 5;          // strong_radius (main.h:83)
constexpr float kCostMax = 2.0f;

// ---- per-view constants (built once per run by k_setup_views) ---------------------------------
struct ViewConst {
	float Rrel[9];      // R_src * R_ref^T            (APD.cu:317-325)
	float trel[3];      // R_src * (C_ref - C_src)    (APD.cu:326-331)
	float K0, K2, K4, K5, K8;   // source intrinsics used by the homography (APD.cu:354-362)
	float wf, hf;       // float(width), float(height) of the source image (APD.cu:546)
	float baseline;     // |c_ref - c_src| as DepthToWeak/LocalRefine compute it (APD.cu:2037-2042)
	apd_camera cam;     // full source camera for the geometric-consistency term (APD.cu:752-789)
	float pad_;         // 49-float stride: lanes reading the same field of different views hit different banks
};
static_assert(sizeof(ViewConst) == 49 * 4, "ViewConst stride");

struct RefConst {
	apd_camera cam;
	float rK0, rK4;     // MUFU.RCP(K[0]), MUFU.RCP(K[4])
	float kk;           // K[0] * rcp(K[4])            (APD.cu:208)
};

// ---- XORWOW (curand default generator), 24 B of live state ------------------------------------
// curandState is 48 B (APD.cpp:640); only v[5] and d are ever live on this path because the
// reference uses curand()/curand_uniform() only (no Box-Muller). Same recurrence as
// curand_kernel.h: identical streams given identical (v, d).
struct Rng { uint32_t v0, v1, v2, v3, v4, d; };

__device__ __forceinline__ uint32_t rng_next(Rng &s) {
	uint32_t t = s.v0 ^ (s.v0 >> 2);
	s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
	s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
	s.d += 362437u;
	return s.v4 + s.d;
}
__device__ __forceinline__ float rng_uniform(Rng &s) {
	// curand_uniform: x * 2^-32 + 2^-33, one FFMA in the reference SASS
	return fmaf((float)rng_next(s), 2.3283064365386962890625e-10f, 1.16415321826934814453125e-10f);
}
__device__ __forceinline__ Rng rng_load(const uint2 *g, size_t idx) {
	const uint2 *p = g + idx * 3;
	uint2 a = p[0], b = p[1], c = p[2];
	Rng s; s.v0 = a.x; s.v1 = a.y; s.v2 = b.x; s.v3 = b.y; s.v4 = c.x; s.d = c.y;
	return s;
}
__device__ __forceinline__ void rng_store(uint2 *g, size_t idx, const Rng &s) {
	uint2 *p = g + idx * 3;
	p[0] = make_uint2(s.v0, s.v1); p[1] = make_uint2(s.v2, s.v3); p[2] = make_uint2(s.v4, s.d);
}

// ---- approximate math, named after the SASS they become --------------------------------------
__device__ __forceinline__ float rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrtaf(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrtaf(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- packed fp32 (sm_100 FFMA2/FADD2): two independent IEEE operations per instruction, same rounding as
// the scalar forms, half the issue slots (measured: tools/ffma2_probe.cu)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- geometry helpers --------------------------------------------------------------------------
// ComputeDepthfromPlaneHypothesis, APD.cu:206-209
__device__ __forceinline__ float plane_depth(const RefConst &rc, const float4 pl, const float xf, const float yf) {
	const float *K = rc.cam.K;
	float b = pl.y * (rc.kk * (yf - K[5]));
	float t = fmaf(pl.x, xf - K[2], b);
	float den = fmaf(K[0], pl.z, t);
	return (K[0] * -pl.w) * rcpf(den);
}
// Get3DPoint, APD.cu:159-164 (X[2] = depth)
__device__ __forceinline__ void backproject(const RefConst &rc, float xf, float yf, float depth, float &X0, float &X1) {
	X0 = (depth * (xf - rc.cam.K[2])) * rc.rK0;
	X1 = (depth * (yf - rc.cam.K[5])) * rc.rK4;
}
// GetDistance2Origin, APD.cu:187-192
__device__ __forceinline__ float plane_offset(const RefConst &rc, float xf, float yf, float depth, float nx, float ny, float nz) {
	float X0, X1; backproject(rc, xf, yf, depth, X0, X1);
	float a = X1 * ny;
	a = fmaf(X0, nx, a);
	a = fmaf(depth, nz, a);
	return -a;
}
__device__ __forceinline__ void normalize3(float &x, float &y, float &z) {   // NormalizeVec3, APD.cu:126-133
	float n2 = y * y; n2 = fmaf(x, x, n2); n2 = fmaf(z, z, n2);
	float r = rsqrtaf(n2);
	x *= r; y *= r; z *= r;
}
// GenerateRandomNormal, APD.cu:211-237
__device__ __forceinline__ float4 random_normal(const RefConst &rc, float xf, float yf, Rng &rng, float depth) {
	float q1, q2, s;
	do {
		q1 = fmaf(rng_uniform(rng), 2.0f, -1.0f);
		q2 = fmaf(rng_uniform(rng), 2.0f, -1.0f);
		s = fmaf(q1, q1, q2 * q2);
	} while (s >= 1.0f);
	float sq = sqrtaf(1.0f - s);
	float nx = sq * (q1 + q1);
	float ny = sq * (q2 + q2);
	float nz = 1.0f - (s + s);
	float X0, X1; backproject(rc, xf, yf, depth, X0, X1);
	float n2 = X1 * X1; n2 = fmaf(X0, X0, n2); n2 = fmaf(depth, depth, n2);
	float rn = rcpf(sqrtaf(n2));
	float vx = X0 * rn, vy = X1 * rn, vz = depth * rn;
	float dot = ny * vy; dot = fmaf(nx, vx, dot); dot = fmaf(nz, vz, dot);
	if (dot > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
	normalize3(nx, ny, nz);
	return make_float4(nx, ny, nz, 0.0f);
}
// GeneratePerturbedNormal, APD.cu:239-274 (perturbation = 0.02f * M_PI folded to a float constant)
__device__ __forceinline__ float4 perturbed_normal(const RefConst &rc, float xf, float yf, const float4 n, Rng &rng) {
	const float pert = (float)(0.02f * 3.14159265358979323846);
	float a1 = (rng_uniform(rng) - 0.5f) * pert;
	float a2 = (rng_uniform(rng) - 0.5f) * pert;
	float a3 = (rng_uniform(rng) - 0.5f) * pert;
	float s1 = __sinf(a1), c1 = __cosf(a1);
	float s2 = __sinf(a2), c2 = __cosf(a2);
	float s3 = __sinf(a3), c3 = __cosf(a3);
	float c1c3 = c1 * c3, s1c3 = s1 * c3, s3c1 = s3 * c1;
	float R0 = c2 * c3;
	float R1 = fmaf(s2, s1c3, -s3c1);
	float R2 = fmaf(s1, s3, s2 * c1c3);
	float R3 = s3 * c2;
	float R4 = fmaf(s3, s1 * s2, c1c3);
	float R5 = fmaf(s3, s2 * c1, -s1c3);
	float R7 = s1 * c2;
	float R8 = c1 * c2;
	float px = fmaf(n.z, R2, fmaf(n.x, R0, n.y * R1));
	float py = fmaf(n.z, R5, fmaf(n.x, R3, n.y * R4));
	float pz = fmaf(n.z, R8, fmaf(n.y, R7, -(n.x * s2)));
	// view direction at depth 1 (APD.cu:241)
	float X0 = (xf - rc.cam.K[2]) * rc.rK0;
	float X1 = (yf - rc.cam.K[5]) * rc.rK4;
	float n2 = fmaf(X0, X0, X1 * X1) + 1.0f;
	float rn = rcpf(sqrtaf(n2));
	float vx = X0 * rn, vy = X1 * rn;
	float dot = vx * px; dot = fmaf(vy, py, dot); dot = fmaf(pz, rn, dot);
	if (dot >= 0.0f) { px = n.x; py = n.y; pz = n.z; }
	normalize3(px, py, pz);
	return make_float4(px, py, pz, 0.0f);
}

// ---- homography (APD.cu:303-363) with the per-view part hoisted ---------------------------------
struct Homog { float h[9]; };
__device__ __forceinline__ Homog make_homography(const RefConst &rc, const ViewConst &vc, const float4 pl) {
	const float rw = rcpf(pl.w);
	float H[9];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		const float t = vc.trel[r];
		H[3 * r + 0] = fmaf(-(pl.x * t), rw, vc.Rrel[3 * r + 0]);
		H[3 * r + 1] = fmaf(-(pl.y * t), rw, vc.Rrel[3 * r + 1]);
		H[3 * r + 2] = fmaf(-(pl.z * t), rw, vc.Rrel[3 * r + 2]);
	}
	const float K2 = rc.cam.K[2], K5 = rc.cam.K[5];
	float T[9];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		T[3 * r + 0] = H[3 * r + 0] * rc.rK0;
		T[3 * r + 1] = H[3 * r + 1] * rc.rK4;
		T[3 * r + 2] = H[3 * r + 2] + fmaf(K2 * -H[3 * r + 0], rc.rK0, -((K5 * H[3 * r + 1]) * rc.rK4));
	}
	Homog o;
	o.h[0] = fmaf(vc.K0, T[0], vc.K2 * T[6]);
	o.h[1] = fmaf(vc.K0, T[1], vc.K2 * T[7]);
	o.h[2] = fmaf(vc.K0, T[2], vc.K2 * T[8]);
	o.h[3] = fmaf(vc.K4, T[3], vc.K5 * T[6]);
	o.h[4] = fmaf(vc.K4, T[4], vc.K5 * T[7]);
	o.h[5] = fmaf(vc.K4, T[5], vc.K5 * T[8]);
	o.h[6] = vc.K8 * T[6];
	o.h[7] = vc.K8 * T[7];
	o.h[8] = vc.K8 * T[8];
	return o;
}
// ComputeCorrespondingPoint for the patch centre (APD.cu:365-372 as inlined at APD.cu:545)
__device__ __forceinline__ bool centre_inside(const Homog &Hm, const ViewConst &vc, float xf, float yf) {
	const float *h = Hm.h;
	float z = h[8] + fmaf(xf, h[6], yf * h[7]);
	float rz = rcpf(z);
	float x = (h[2] + fmaf(xf, h[0], yf * h[1])) * rz;
	float y = (h[5] + fmaf(xf, h[3], yf * h[4])) * rz;
	return !(x >= vc.wf || x < 0.0f || y >= vc.hf || y < 0.0f);
}

// ---- NCC epilogue (APD.cu:592-610) -----------------------------------------------------------------
struct NccSums { float r, rr, s, ss, rs; };
__device__ __forceinline__ float ncc_cost(const NccSums &a, float inv_w) {
	float mr = inv_w * a.r;
	float ms = inv_w * a.s;
	float var_r = fmaf(inv_w, a.rr, -(mr * mr));
	float var_s = fmaf(inv_w, a.ss, -(ms * ms));
	const float kMinVar = 1e-5f;
	if (var_r < kMinVar || var_s < kMinVar) return kCostMax;
	float covar = fmaf(-mr, ms, inv_w * a.rs);
	float den = sqrtaf(var_r * var_s);
	float c = fmaf(-covar, rcpf(den), 1.0f);
	return fmaxf(0.0f, fminf(kCostMax, c));
}

// One source tap (APD.cu:570-573): projective warp with the x-part hoisted out of the row loop,
// exactly as the reference compiles it (FMUL / FFMA / FADD / MUFU.RCP / FFMA +0.5).
__device__ __forceinline__ float src_tap(cudaTextureObject_t tex, int layer, const float *h, float ax, float ay, float az, float yf) {
	float xs = h[2] + fmaf(h[1], yf, ax);
	float ys = h[5] + fmaf(h[4], yf, ay);
	float zs = h[8] + fmaf(h[7], yf, az);
	float rz = rcpf(zs);
	return tex2DLayered<float>(tex, fmaf(xs, rz, 0.5f), fmaf(ys, rz, 0.5f), layer);
}

// ---- quad-cooperative NCC ------------------------------------------------------------------------
// Measured on B200 (tools/tex_probe3.cu, profiles/): the texture data pipe delivers the full
// 4 bilinear fetches/clk/SM only when the four lanes of a quad touch a compact (<= ~4x4 texel)
// footprint; four lanes sampling four unrelated places (= four different plane hypotheses, the
// natural thread-per-pixel mapping) run at ~35 %. So the four lanes of a quad fetch TOGETHER:
// evaluation e (owned by lane e of the quad: its homography, pixel and view) is fetched by all four
// lanes, lane s taking the 3x3 taps of quadrant s ^ e of the 6x6 window (taps (2c+a, 2d+b)), so every
// TEX instruction covers a 2x2 cluster of neighbouring taps. The warped source patch is staged in
// shared memory (one [36 taps][32 lanes] slab per warp) and each lane then accumulates its OWN
// evaluation's 36 taps from there in the reference's order (x-offset outer, y-offset inner, row
// sums) -- bit-identical sums. Slab addressing: tap T of evaluation e of quad Q sits at
// T*32 + Q*4 + (e ^ quadrant(T)); writers therefore store at T*32 + lane and readers load from
// T*32 + (lane ^ quadrant(T)): both conflict-free.
// All four lanes of a quad must call this convergently; `want` = false lanes only help.
constexpr int kPatchFloats = 36 * 32;            // per warp
struct QuadCtx {
	float *slab; unsigned qmask; int lane, ql;
	float ref_r, ref_rr;        // the pixel's own reference-patch sums (independent of view and plane): set_ref_sums
	bool qx, qy;                // quadrant bits of this lane
};
__device__ __forceinline__ QuadCtx make_quad_ctx(float *patch_base, int tid) {
	QuadCtx q; q.slab = patch_base + (tid >> 5) * kPatchFloats; q.lane = tid & 31; q.ql = tid & 3; q.qmask = 0xFu << (q.lane & ~3);
	q.ref_r = 0.f; q.ref_rr = 0.f;
	q.qx = (q.ql & 1) != 0; q.qy = (q.ql & 2) != 0;
	return q;
}
// Sum and sum of squares of the 6x6 reference patch in exactly the order ncc6_quad / the reference accumulate them
// (APD.cu:556-590): they do not depend on the source view or the plane, so a pixel computes them once.
__device__ __forceinline__ void set_ref_sums(QuadCtx &q, const float *tile, int pitch, int lx, int ly) {
	const float *base = tile + (ly + kHalo) * pitch + (lx + kHalo) - 5 * pitch - 5;
	float r = 0.f, rr = 0.f;
#pragma unroll 1
	for (int i2 = 0; i2 < 3; ++i2) {
		const float *rb0 = base + 2 * (2 * i2), *rb1 = rb0 + 2;
		f32x2 R = 0ull, RR = 0ull;
#pragma unroll
		for (int j = 0; j < 6; ++j) {
			const f32x2 RP = pk2(rb0[2 * j * pitch], rb1[2 * j * pitch]);
			R = add2(RP, R); RR = fma2(RP, RP, RR);
		}
		float a0, a1;
		unpk2(R, a0, a1); r += a0; r += a1;
		unpk2(RR, a0, a1); rr += a0; rr += a1;
	}
	q.ref_r = r; q.ref_rr = rr;
}

// The evaluation in three pieces, so that a caller with a queue of evaluations can keep the fetches of the NEXT one in
// flight while it accumulates the current one (ncc6_quad below is the plain sequence issue, store, accumulate).
struct NccFetch { float v[4][9]; unsigned my_quad; bool active; };

// Piece 1: decide who is live, broadcast each owner's homography to its quad, issue all 36 fetches of this lane.
__device__ __forceinline__ void ncc6_issue(const QuadCtx &q, cudaTextureObject_t tex, int layer, const Homog &Hm, const ViewConst &vc, bool want,
                                           int px, int py, NccFetch &f) {
	// must be called by all 32 lanes of the warp convergently
	const float pxf = (float)px, pyf = (float)py;
	f.active = want && centre_inside(Hm, vc, pxf, pyf);
	const unsigned ballot = __ballot_sync(0xffffffffu, f.active);
	f.my_quad = (ballot >> (q.lane & ~3)) & 0xFu;                 // which of my quad's four evaluation slots are live
#pragma unroll
	for (int e = 0; e < 4; ++e) {
		if ((ballot & (0x11111111u << e)) == 0u) continue;            // warp-uniform: nobody owns a live slot e
		float h[9];
#pragma unroll
		for (int i = 0; i < 9; ++i) h[i] = __shfl_sync(0xffffffffu, Hm.h[i], e, 4);
		// tap coordinates are small integers: float sums of them are exact, no conversions per tap
		const float pxe = __shfl_sync(0xffffffffu, pxf, e, 4), pye = __shfl_sync(0xffffffffu, pyf, e, 4);
		const int lay = __shfl_sync(0xffffffffu, layer, e, 4);
		if ((f.my_quad >> e) & 1u) {
			const float x0 = pxe + ((q.qx != ((e & 1) != 0)) ? -3.f : -5.f), y0 = pye + ((q.qy != ((e & 2) != 0)) ? -3.f : -5.f);
			// x and y of the projective warp advance as one packed pair (same operations as src_tap)
			const f32x2 H03 = pk2(h[0], h[3]), H14 = pk2(h[1], h[4]), H25 = pk2(h[2], h[5]);
			float yf[3]; f32x2 YF[3];
#pragma unroll
			for (int d = 0; d < 3; ++d) { yf[d] = y0 + (float)(4 * d); YF[d] = pk2(yf[d], yf[d]); }
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float xf = x0 + (float)(4 * c);
				const f32x2 AXY = mul2(H03, pk2(xf, xf));
				const float az = h[6] * xf;
#pragma unroll
				for (int d = 0; d < 3; ++d) {
					const f32x2 XY = add2(H25, fma2(H14, YF[d], AXY));
					const float rz = rcpf(h[8] + fmaf(h[7], yf[d], az));
					// scalar here: the texture instruction wants (layer, x, y) in consecutive registers, which a
					// packed result (even/odd pair) can only reach through extra moves
					float xs, ys; unpk2(XY, xs, ys);
					f.v[e][c * 3 + d] = tex2DLayered<float>(tex, fmaf(xs, rz, 0.5f), fmaf(ys, rz, 0.5f), lay);
				}
			}
		}
	}
}

// Piece 2: stage the warped patches: tap (2c+qa, 2d+qb) -> T = 12c + 2d + 6qa + qb, slot T*32 + lane. The caller
// synchronises the warp afterwards.
__device__ __forceinline__ void ncc6_store(const QuadCtx &q, const NccFetch &f) {
#pragma unroll
	for (int e = 0; e < 4; ++e) {
		if ((f.my_quad >> e) & 1u) {
			const int qd = q.ql ^ e;
			float *dst = q.slab + (6 * (qd & 1) + (qd >> 1)) * 32 + q.lane;
#pragma unroll
			for (int c = 0; c < 3; ++c)
#pragma unroll
				for (int d = 0; d < 3; ++d) dst[(12 * c + 2 * d) * 32] = f.v[e][c * 3 + d];
		}
	}
}

// Piece 3: the owner accumulates its own evaluation from the slab in the reference's order. The caller synchronises
// the warp before the slab is written again.
template <bool ROLL>
__device__ __forceinline__ float ncc6_accum(const QuadCtx &q, bool active, const float *tile, int pitch, int lx, int ly, float inv_w) {
	float cost = kCostMax;
	if (active) {
		NccSums t = {q.ref_r, q.ref_rr, 0.f, 0.f, 0.f};
		const float *base = tile + (ly + kHalo) * pitch + (lx + kHalo) - 5 * pitch - 5;
		const float *s0 = q.slab + q.lane, *s1 = q.slab + (q.lane ^ 1), *s2 = q.slab + (q.lane ^ 2), *s3 = q.slab + (q.lane ^ 3);
		// ROLL keeps the accumulation loop small for the instruction cache (two x-offsets per trip)
#pragma unroll(ROLL ? 1 : 3)
		for (int i2 = 0; i2 < 3; ++i2) {
			// the two x-offsets of a trip are independent row sums: their running sums advance in lock
			// step as packed pairs (lo = column 2*i2, hi = column 2*i2+1)
			const float *rb0 = base + 2 * (2 * i2), *rb1 = rb0 + 2;
			const float *sa0 = s0 + (2 * i2) * 6 * 32, *sb0 = s2 + (2 * i2) * 6 * 32;             // column 2*i2:   quadrants 0 (j even), 2 (j odd)
			const float *sa1 = s1 + (2 * i2 + 1) * 6 * 32, *sb1 = s3 + (2 * i2 + 1) * 6 * 32;     // column 2*i2+1: quadrants 1, 3
			f32x2 RS = 0ull, S = 0ull, SS = 0ull;
#pragma unroll
			for (int j = 0; j < 6; ++j) {
				const f32x2 RP = pk2(rb0[2 * j * pitch], rb1[2 * j * pitch]);
				const f32x2 SP = pk2(((j & 1) ? sb0 : sa0)[j * 32], ((j & 1) ? sb1 : sa1)[j * 32]);
				RS = fma2(RP, SP, RS);
				S = add2(SP, S); SS = fma2(SP, SP, SS);
			}
			float a0, a1;
			unpk2(S, a0, a1); t.s += a0; t.s += a1;
			unpk2(SS, a0, a1); t.ss += a0; t.ss += a1;
			unpk2(RS, a0, a1); t.rs += a0; t.rs += a1;
		}
		cost = ncc_cost(t, inv_w);
	}
	return cost;
}

template <int SPT = 4, bool ROLL = true>
__device__ __forceinline__ float ncc6_quad(const QuadCtx &q, cudaTextureObject_t tex, int layer, const Homog &Hm, const ViewConst &vc, bool want,
                                           const float *tile, int pitch, int lx, int ly, int px, int py, float inv_w) {
	// must be called by all 32 lanes of the warp convergently; all 36 fetches are in flight before any result is consumed
	NccFetch f;
	ncc6_issue(q, tex, layer, Hm, vc, want, px, py, f);
	ncc6_store(q, f);
	__syncwarp();
	const float cost = ncc6_accum<ROLL>(q, f.active, tile, pitch, lx, ly, inv_w);
	__syncwarp();
	return cost;
}

// ---- kernel arguments -------------------------------------------------------------------------------
struct Args {
	int W, H, S;                 // S = number of source views = num_images - 1
	int half_rows;               // rows reachable by the reference's half launch (APD.cu:2400-2403)
	int ref_pitch;               // elements per row of ref_pad
	float depth_min, depth_max;
	int top_k, state, geom, weak_peak_radius, rotate_time;
	float geom_factor, ransac_threshold;
	const float *inv_w;          // [0] = MUFU.RCP(36.0f) strong window, [1] = MUFU.RCP(9.0f) anchor window
	cudaTextureObject_t img_tex;     // layered: layer v = image v (0 = reference)
	cudaTextureObject_t depth_tex;   // layered depth maps (geom consistency)
	const float *ref_pad;        // reference image with kRefPad replicated border
	const ViewConst *views;      // [S]
	const RefConst *ref;
	float4 *planes; float4 *fit_planes; float *costs;
	uint32_t *sel_views; uint8_t *states; uint2 *rng; uint4 *view_w;
	short2 *anchors;             // [9][W*H]: slot k of pixel p at anchors[k*W*H + p] (reference: compact [weak_idx*9+k])
	short2 *nearest; uint8_t *reliable;
	float *scratch;              // slab pool for per-block work arrays (cost matrices, K14 profiles): one slab per RESIDENT block
	int *slab_slots;             // [kSlabSMs * kSlabPerSM] 0 = free, 1 = taken
	int slab_stride;             // floats per slab
	// compacted WEAK-pixel lists (k_weak_lists): pixel indices y*W+x; wlist[c*wlist_stride + i] is entry i of colour c
	// (c = 0: (x+y) even = the reference's "black" launch, APD.cu:1512-1519); wctrl[0..1] = entries per colour,
	// wctrl[2] = entries of the single pre-demotion list K3 walks, wctrl[kWorkBase + j] = work counter of launch j
	int *wlist; int wlist_stride; int *wctrl;
};
constexpr int kWorkBase = 8, kWorkSlots = 248;

// ---- slab pool: work arrays too big for shared memory (which would come out of the L1 the texture fetches live on)
// stay in a small pool indexed by (SM, resident block), so they are rewritten in place in L1/L2 and never stream to DRAM.
constexpr int kSlabSMs = 256, kSlabPerSM = 8;
// Thread 0 claims a slab of its SM; the caller's next __syncthreads() publishes s_slot[0] (slab index), s_slot[1] (exit count).
__device__ __forceinline__ void slab_acquire(const Args &a, int *s_slot, int tid) {
	if (tid == 0) {
		unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		const int base = (int)(smid % kSlabSMs) * kSlabPerSM;
		int k = 0;
		while (atomicCAS(&a.slab_slots[base + k], 0, 1) != 0) k = (k + 1) % kSlabPerSM;   // fewer resident blocks per SM than slots
		s_slot[0] = base + k; s_slot[1] = 0;
	}
}
__device__ __forceinline__ float *slab_ptr(const Args &a, const int *s_slot) { return a.scratch + (size_t)s_slot[0] * a.slab_stride; }
// Every thread of the block calls this exactly once, after its last access to the slab; the last one frees it.
__device__ __forceinline__ void slab_exit(const Args &a, int *s_slot, int nthreads) {
	__threadfence_block();
	if (atomicAdd(&s_slot[1], 1) == nthreads - 1) atomicExch(&a.slab_slots[s_slot[0]], 0);
}

// ComputeGeomConsistencyCost, APD.cu:752-789, in two pieces: the world point of the pixel at the hypothesis' depth does not
// depend on the source view (APD.cu:758-764), so callers that visit several views with one plane compute it once.
struct GeomPoint { float x, y, z; };
__device__ __forceinline__ GeomPoint geom_point(const RefConst &rc, const float4 pl, float xf, float yf) {
	const float depth = plane_depth(rc, pl, xf, yf);
	float X0, X1; backproject(rc, xf, yf, depth, X0, X1);
	const float *R = rc.cam.R;
	GeomPoint P;
	P.x = rc.cam.c[0] + fmaf(R[6], depth, fmaf(R[0], X0, R[3] * X1));
	P.y = rc.cam.c[1] + fmaf(R[7], depth, fmaf(R[1], X0, R[4] * X1));
	P.z = rc.cam.c[2] + fmaf(R[8], depth, fmaf(R[2], X0, R[5] * X1));
	return P;
}
__device__ __forceinline__ float geom_cost_at(const Args &a, const RefConst &rc, const ViewConst &vc, int layer, const GeomPoint P, float xf, float yf) {
	const float *R = rc.cam.R;
	const float Px = P.x, Py = P.y, Pz = P.z;
	const apd_camera &s = vc.cam;
	float tx = s.t[0] + fmaf(s.R[2], Pz, fmaf(s.R[0], Px, s.R[1] * Py));
	float ty = s.t[1] + fmaf(s.R[5], Pz, fmaf(s.R[3], Px, s.R[4] * Py));
	float tz = s.t[2] + fmaf(s.R[8], Pz, fmaf(s.R[6], Px, s.R[7] * Py));
	float rd = rcpf(fmaf(s.K[8], tz, fmaf(s.K[6], tx, s.K[7] * ty)));
	float sx = fmaf(s.K[2], tz, fmaf(s.K[0], tx, s.K[1] * ty)) * rd;
	float sy = fmaf(s.K[5], tz, fmaf(s.K[3], tx, s.K[4] * ty)) * rd;
	const float sd = tex2DLayered<float>(a.depth_tex, (float)(int)sx + 0.5f, (float)(int)sy + 0.5f, layer);
	if (sd == 0.0f) return 3.0f;
	const float rsK0 = rcpf(s.K[0]), rsK4 = rcpf(s.K[4]);
	float Y0 = (sd * (sx - s.K[2])) * rsK0;
	float Y1 = (sd * (sy - s.K[5])) * rsK4;
	float Qx = s.c[0] + fmaf(s.R[6], sd, fmaf(s.R[0], Y0, s.R[3] * Y1));
	float Qy = s.c[1] + fmaf(s.R[7], sd, fmaf(s.R[1], Y0, s.R[4] * Y1));
	float Qz = s.c[2] + fmaf(s.R[8], sd, fmaf(s.R[2], Y0, s.R[5] * Y1));
	const float *K = rc.cam.K; const float *t = rc.cam.t;
	float ux = t[0] + fmaf(R[2], Qz, fmaf(R[0], Qx, R[1] * Qy));
	float uy = t[1] + fmaf(R[5], Qz, fmaf(R[3], Qx, R[4] * Qy));
	float uz = t[2] + fmaf(R[8], Qz, fmaf(R[6], Qx, R[7] * Qy));
	float rb = rcpf(fmaf(K[8], uz, fmaf(K[6], ux, K[7] * uy)));
	float dc = fmaf(-fmaf(K[2], uz, fmaf(K[0], ux, K[1] * uy)), rb, xf);
	float dr = fmaf(-fmaf(K[5], uz, fmaf(K[3], ux, K[4] * uy)), rb, yf);
	return fminf(sqrtaf(fmaf(dc, dc, dr * dr)), 3.0f);
}
__device__ __forceinline__ float geom_cost(const Args &a, const RefConst &rc, const ViewConst &vc, int layer, const float4 pl, float xf, float yf) {
	return geom_cost_at(a, rc, vc, layer, geom_point(rc, pl, xf, yf), xf, yf);
}

// view weights: 32 nibbles (sum <= 15) in one uint4 per pixel (reference: 32 bytes, APD.cpp:645)
struct VW { unsigned long long lo, hi; };
__device__ __forceinline__ void vw_add(VW &w, int v) { if (v < 16) w.lo += 1ull << (4 * v); else w.hi += 1ull << (4 * (v - 16)); }
__device__ __forceinline__ int vw_get(const VW &w, int v) { return (int)(((v < 16) ? (w.lo >> (4 * v)) : (w.hi >> (4 * (v - 16)))) & 15ull); }
__device__ __forceinline__ VW vw_load(const uint4 *g, size_t i) { uint4 u = g[i]; VW w; w.lo = ((unsigned long long)u.y << 32) | u.x; w.hi = ((unsigned long long)u.w << 32) | u.z; return w; }
__device__ __forceinline__ void vw_store(uint4 *g, size_t i, const VW &w) { g[i] = make_uint4((uint32_t)w.lo, (uint32_t)(w.lo >> 32), (uint32_t)w.hi, (uint32_t)(w.hi >> 32)); }

// bitmask of the views with a non-zero sampling weight
__device__ __forceinline__ uint32_t vw_mask(const VW &w, int S) {
	uint32_t m = 0u;
#pragma unroll 1
	for (int v = 0; v < S; ++v) if (vw_get(w, v) > 0) m |= 1u << v;
	return m;
}

// Shared front end of K14 and K15: plane back into the reference camera frame, mean baseline and
// summed weights over the selected views (APD.cu:2012-2052, :2160-2199).
struct SweepCtx { float4 pl; float depth, weight_normal, kb, disp; int valid; };
__device__ __forceinline__ bool sweep_setup(const Args &a, const RefConst &rc, const ViewConst *sv, size_t center, uint32_t bits, const VW &vw, SweepCtx &c) {
	const float4 in = a.planes[center];
	const float *R = rc.cam.R;
	c.pl.x = fmaf(in.z, R[2], fmaf(in.x, R[0], in.y * R[1]));
	c.pl.y = fmaf(in.z, R[5], fmaf(in.x, R[3], in.y * R[4]));
	c.pl.z = fmaf(in.z, R[8], fmaf(in.x, R[6], in.y * R[7]));
	c.pl.w = in.w; c.depth = in.w;
	if (c.depth == 0.0f) return false;
	float base = 0.0f; c.weight_normal = 0.0f; c.valid = 0;
#pragma unroll 1
	for (int v = 0; v < a.S; ++v) if ((bits >> v) & 1u) { base += sv[v].baseline; c.weight_normal += (float)vw_get(vw, v); c.valid++; }
	if (c.valid == 0) return true;
	base = rcpf((float)c.valid) * base;
	c.kb = base * rc.cam.K[0];
	c.disp = c.kb * rcpf(c.depth);
	return true;
}

}  // namespace apd
