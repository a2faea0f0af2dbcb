// Two evaluations per lane, computed as packed fp32 pairs: shared by the quad-per-pixel kernels (k_weak_q, k_sweep_q).
#pragma once
#include "apd_device.cuh"

namespace apd {

// ---- two evaluations per lane ("slots" 0 and 1), computed as packed fp32 pairs ---------------------------------------
// A lane always carries two evaluations through the deformable NCC: two candidate planes against one view (cost matrix),
// one plane against two views (current / fit / refinement hypotheses). Both slots execute the same operations on
// different data, so every FMUL/FADD/FFMA of the evaluation is issued ONCE as an f32x2 instruction (lo = slot 0,
// hi = slot 1; same IEEE roundings as the scalar forms, apd_device.cuh), which halves the issue slots of a kernel that ncu
// showed to be issue bound once its fetches were compact (profiles/r02w_*).
struct Homog2 { f32x2 h[9]; };

// make_homography (apd_device.cuh) for both slots. Negations are moved onto an operand: -(a*b) == (-a)*b exactly.
__device__ __forceinline__ Homog2 make_homography2(const RefConst &rc, const ViewConst &v0, const ViewConst &v1, const float4 p0, const float4 p1) {
	const f32x2 RW = pk2(rcpf(p0.w), rcpf(p1.w));
	const f32x2 NX = pk2(-p0.x, -p1.x), NY = pk2(-p0.y, -p1.y), NZ = pk2(-p0.z, -p1.z);
	f32x2 H[9];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		const f32x2 t = pk2(v0.trel[r], v1.trel[r]);
		H[3 * r + 0] = fma2(mul2(NX, t), RW, pk2(v0.Rrel[3 * r + 0], v1.Rrel[3 * r + 0]));
		H[3 * r + 1] = fma2(mul2(NY, t), RW, pk2(v0.Rrel[3 * r + 1], v1.Rrel[3 * r + 1]));
		H[3 * r + 2] = fma2(mul2(NZ, t), RW, pk2(v0.Rrel[3 * r + 2], v1.Rrel[3 * r + 2]));
	}
	const f32x2 rK0 = pk2(rc.rK0, rc.rK0), rK4 = pk2(rc.rK4, rc.rK4);
	const f32x2 nK2 = pk2(-rc.cam.K[2], -rc.cam.K[2]), nK5 = pk2(-rc.cam.K[5], -rc.cam.K[5]);
	f32x2 T[9];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		T[3 * r + 0] = mul2(H[3 * r + 0], rK0);
		T[3 * r + 1] = mul2(H[3 * r + 1], rK4);
		T[3 * r + 2] = add2(H[3 * r + 2], fma2(mul2(nK2, H[3 * r + 0]), rK0, mul2(mul2(nK5, H[3 * r + 1]), rK4)));
	}
	const f32x2 K0 = pk2(v0.K0, v1.K0), K2 = pk2(v0.K2, v1.K2), K4 = pk2(v0.K4, v1.K4), K5 = pk2(v0.K5, v1.K5), K8 = pk2(v0.K8, v1.K8);
	Homog2 o;
	o.h[0] = fma2(K0, T[0], mul2(K2, T[6]));
	o.h[1] = fma2(K0, T[1], mul2(K2, T[7]));
	o.h[2] = fma2(K0, T[2], mul2(K2, T[8]));
	o.h[3] = fma2(K4, T[3], mul2(K5, T[6]));
	o.h[4] = fma2(K4, T[4], mul2(K5, T[7]));
	o.h[5] = fma2(K4, T[5], mul2(K5, T[8]));
	o.h[6] = mul2(K8, T[6]);
	o.h[7] = mul2(K8, T[7]);
	o.h[8] = mul2(K8, T[8]);
	return o;
}
// (H (x, y, 1))_xy / z for both slots: ComputeCorrespondingPoint as inlined at APD.cu:426-432 and :545
__device__ __forceinline__ void project2(const Homog2 &H, float xf, float yf, float &x0, float &y0, float &x1, float &y1) {
	const f32x2 XF = pk2(xf, xf), YF = pk2(yf, yf);
	const f32x2 Z = add2(H.h[8], fma2(H.h[6], XF, mul2(H.h[7], YF)));
	float z0, z1; unpk2(Z, z0, z1);
	const f32x2 RZ = pk2(rcpf(z0), rcpf(z1));
	const f32x2 X = mul2(add2(H.h[2], fma2(H.h[0], XF, mul2(H.h[1], YF))), RZ);
	const f32x2 Y = mul2(add2(H.h[5], fma2(H.h[3], XF, mul2(H.h[4], YF))), RZ);
	unpk2(X, x0, x1); unpk2(Y, y0, y1);
}

// NCC of one window for the two slots of this lane (same reference taps; APD.cu:456-505 / :556-610)
// `col`: the window's reference taps in evaluation order, STRIDE floats apart (a per-pixel shared-memory column)
template <int INC, int STRIDE>
__device__ __forceinline__ void wq_window2(cudaTextureObject_t tex, int lay0, int lay1, const Homog2 &H, bool w0, bool w1,
                                           int cx, int cy, float inv_w, const float *col, float sum_r, float sum_rr, float &o0, float &o1) {
	f32x2 TS = 0ull, TSS = 0ull, TRS = 0ull;
	const float cxf = (float)cx, cyf = (float)cy;         // tap coordinates are small integers: cxf + i == (float)(cx + i) exactly
	int n = 0;
#pragma unroll(INC == 5 ? 3 : 1)
	for (int i = -5; i <= 5; i += INC) {
		const float xf = cxf + (float)i;
		const f32x2 XF = pk2(xf, xf);
		const f32x2 AX = mul2(H.h[0], XF), AY = mul2(H.h[3], XF), AZ = mul2(H.h[6], XF);
		f32x2 RS = 0ull, S = 0ull, SS = 0ull;
#pragma unroll
		for (int j = -5; j <= 5; j += INC) {
			const float rp = col[n * STRIDE];
			++n;
			const float yf = cyf + (float)j;
			const f32x2 YF = pk2(yf, yf);
			const f32x2 XS = add2(H.h[2], fma2(H.h[1], YF, AX));
			const f32x2 YS = add2(H.h[5], fma2(H.h[4], YF, AY));
			const f32x2 ZS = add2(H.h[8], fma2(H.h[7], YF, AZ));
			float xs0, xs1, ys0, ys1, zs0, zs1;
			unpk2(XS, xs0, xs1); unpk2(YS, ys0, ys1); unpk2(ZS, zs0, zs1);
			float sp0 = 0.f, sp1 = 0.f;
			if (w0) { const float rz = rcpf(zs0); sp0 = tex2DLayered<float>(tex, fmaf(xs0, rz, 0.5f), fmaf(ys0, rz, 0.5f), lay0); }
			if (w1) { const float rz = rcpf(zs1); sp1 = tex2DLayered<float>(tex, fmaf(xs1, rz, 0.5f), fmaf(ys1, rz, 0.5f), lay1); }
			const f32x2 SP = pk2(sp0, sp1), RP = pk2(rp, rp);
			RS = fma2(RP, SP, RS); S = add2(S, SP); SS = fma2(SP, SP, SS);
		}
		TS = add2(TS, S); TSS = add2(TSS, SS); TRS = add2(TRS, RS);
	}
	NccSums t0 = {sum_r, sum_rr, 0.f, 0.f, 0.f}, t1 = t0;
	unpk2(TS, t0.s, t1.s); unpk2(TSS, t0.ss, t1.ss); unpk2(TRS, t0.rs, t1.rs);
	o0 = ncc_cost(t0, inv_w); o1 = ncc_cost(t1, inv_w);
}

// The plain 6x6 window (INC = 2) for the two slots, WITHOUT per-slot predicates: unconditional fetches keep the code the
// compiler emits per TEX down to the coordinate arithmetic (a predicated TEX makes ptxas rebuild the layer clamp, the LOD
// register and the texture handle for every fetch: profiles/r02d_*); callers make sure both slots carry useful work and
// discard the result of a slot that does not. The loop over the six tap columns stays ROLLED (12 fetches per trip): fully
// unrolled and software-pipelined the kernel no longer fits the instruction cache (383 vs 295 ms at 6221x4146), and a rolled
// two-column pipeline was no faster either (309 ms: the other warps already cover the fetch latency).
template <int STRIDE>
__device__ __forceinline__ void wq6_issue(cudaTextureObject_t tex, int lay0, int lay1, const Homog2 &H, float xf, const float (&yf)[6], f32x2 (&b)[6]) {
	const f32x2 XF = pk2(xf, xf);
	const f32x2 AX = mul2(H.h[0], XF), AY = mul2(H.h[3], XF), AZ = mul2(H.h[6], XF);
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		const f32x2 YF = pk2(yf[j], yf[j]);
		const f32x2 XS = add2(H.h[2], fma2(H.h[1], YF, AX));
		const f32x2 YS = add2(H.h[5], fma2(H.h[4], YF, AY));
		const f32x2 ZS = add2(H.h[8], fma2(H.h[7], YF, AZ));
		float xs0, xs1, ys0, ys1, zs0, zs1;
		unpk2(XS, xs0, xs1); unpk2(YS, ys0, ys1); unpk2(ZS, zs0, zs1);
		const float rz0 = rcpf(zs0), rz1 = rcpf(zs1);
		const float sp0 = tex2DLayered<float>(tex, fmaf(xs0, rz0, 0.5f), fmaf(ys0, rz0, 0.5f), lay0);
		const float sp1 = tex2DLayered<float>(tex, fmaf(xs1, rz1, 0.5f), fmaf(ys1, rz1, 0.5f), lay1);
		b[j] = pk2(sp0, sp1);
	}
}
template <int STRIDE>
__device__ __forceinline__ void wq6_accum(const float *col, const f32x2 (&b)[6], f32x2 &TS, f32x2 &TSS, f32x2 &TRS) {
	f32x2 RS = 0ull, S = 0ull, SS = 0ull;
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		const float rp = col[j * STRIDE];
		const f32x2 SP = b[j], RP = pk2(rp, rp);
		RS = fma2(RP, SP, RS); S = add2(S, SP); SS = fma2(SP, SP, SS);
	}
	TS = add2(TS, S); TSS = add2(TSS, SS); TRS = add2(TRS, RS);
}
template <int STRIDE>
__device__ __forceinline__ void wq_window6(cudaTextureObject_t tex, int lay0, int lay1, const Homog2 &H, int cx, int cy, float inv_w,
                                           const float *col, float sum_r, float sum_rr, float &o0, float &o1) {
	const float cxf = (float)cx, cyf = (float)cy;
	float yf[6];
#pragma unroll
	for (int j = 0; j < 6; ++j) yf[j] = cyf + (float)(2 * j - 5);
	f32x2 TS = 0ull, TSS = 0ull, TRS = 0ull;
	float xf = cxf - 5.0f;
#pragma unroll 1
	for (int c = 0; c < 6; ++c) {
		f32x2 b[6];
		wq6_issue<STRIDE>(tex, lay0, lay1, H, xf, yf, b);
		wq6_accum<STRIDE>(col + 6 * c * STRIDE, b, TS, TSS, TRS);
		xf += 2.0f;
	}
	NccSums t0 = {sum_r, sum_rr, 0.f, 0.f, 0.f}, t1 = t0;
	unpk2(TS, t0.s, t1.s); unpk2(TSS, t0.ss, t1.ss); unpk2(TRS, t0.rs, t1.rs);
	o0 = ncc_cost(t0, inv_w); o1 = ncc_cost(t1, inv_w);
}

// reference side of one window: taps in the evaluation order (x-offset outer, y-offset inner) and their sum / sum of
// squares accumulated exactly as the evaluation would (APD.cu:456-487)
template <int INC, int STRIDE>
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
			col[t * STRIDE] = rp;
			++t;
			r += rp; rr = fmaf(rp, rp, rr);
		}
		R += r; RR += rr;
	}
	sums[0] = R; sums[STRIDE] = RR;
}


// ComputeGeomConsistencyCost's view-dependent half (geom_cost_at, apd_device.cuh) for the two slots: world point P0
// against source view va, P1 against vb. Negations sit on an operand ((-x)*y == -(x*y) exactly); subtractions are additions of
// the negated operand.
__device__ __forceinline__ void geom_cost_at2(const Args &a, const RefConst &rc, const ViewConst &va, const ViewConst &vb, int la, int lb,
                                              const GeomPoint P0, const GeomPoint P1, float xf, float yf, float &g0, float &g1) {
	const apd_camera &s0 = va.cam, &s1 = vb.cam;
#define PK(f) pk2(s0.f, s1.f)
	const f32x2 Px = pk2(P0.x, P1.x), Py = pk2(P0.y, P1.y), Pz = pk2(P0.z, P1.z);
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


}  // namespace apd
