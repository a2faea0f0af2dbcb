/*
 * oracle/apd_cpu.c — plain-C CPU restatement of the reference PatchMatch path (whoiszzj/APD-MVS,
 * APD.cu:791-2495): all 15 kernels of APD::RunPatchMatch, i.e. the strong-pixel schedule (K1, K5-K7,
 * K11-K15) and the adaptive-patch-deformation path of WEAK pixels (K2-K4, K8-K10).
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under apd_mvs_b200/ links, loads or calls this file; it is
 * used by tests/ (as the checker), by __graft_entry__.smoke() and by bench.py's cpu_baseline leg.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself: tests/golden/ holds per-stage
 * dumps produced by oracle/_ref/libapd_ref.so (the reference's own APD.cu compiled for sm_100,
 * see oracle/ref_wrapper.cu) on a B200, together with the script that made them. The RNG stage is
 * compared bit-for-bit; floating-point stages are compared distributionally because the GPU build
 * uses --use_fast_math (MUFU.RCP/SQRT/EX2/SIN approximations, hardware bilinear filtering with
 * 8-bit weights) which a CPU can only approximate. Tolerances are written in tests/test_oracle_cpu.py.
 *
 * Every function cites the reference lines it follows. Texture reads follow the CUDA programming
 * guide's linear-filtering definition (clamp addressing, 1.8 fixed-point weights).
 */
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846   /* APD.h:7 */
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define MAX_IMAGES 32
enum { FIRST_INIT = 0, REFINE_INIT = 1, REFINE_ITER = 2 };
enum { WEAK = 0, STRONG = 1, UNKNOWN = 2 };

typedef struct { float K[9], R[9], t[3], c[3]; int height, width; float depth_min, depth_max; } Camera; /* main.h:47-56 */
typedef struct {                                                                                      /* main.h:75-94 */
	int max_iterations, num_images; float sigma_spatial, sigma_color; int top_k; float depth_min, depth_max;
	unsigned char geom_consistency, pad0[3]; int strong_radius, strong_increment, weak_radius, weak_increment;
	unsigned char use_APD, pad1[3]; int weak_peak_radius, rotate_time; float ransac_threshold, geom_factor; int state;
} Params;
typedef struct { float x, y, z, w; } f4;
typedef struct { uint32_t v[5], d; } Rng;

typedef struct {
	int W, H, N;
	const float *images;      /* [N][H][W] */
	const float *depths;      /* [N][H][W] or NULL */
	const Camera *cams;       /* [N] */
	Params p;
	f4 *planes; float *costs; uint32_t *views; uint8_t *states; uint8_t *vw; /* [H*W*32] */ Rng *rng;
	/* deformation path (WEAK pixels) */
	int16_t *anchors;         /* [H*W][9][2]: slot 0 = the pixel itself, (-1,-1) = absent (reference: compact weak index) */
	int16_t *nearest;         /* [H*W][2] nearest STRONG pixel */
	uint8_t *reliable;        /* [H*W] */
	f4 *fit;                  /* [H*W] RANSAC plane through the anchors */
} Ctx;

/* ---- curand XORWOW (curand_kernel.h) --------------------------------------------------------- */
static uint32_t rng_next(Rng *s) {
	uint32_t t = s->v[0] ^ (s->v[0] >> 2);
	s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
	s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
	s->d += 362437u;
	return s->v[4] + s->d;
}
static float rng_uniform(Rng *s) { return fmaf((float)rng_next(s), 2.3283064e-10f, 2.3283064e-10f / 2.0f); }

/* The v-part of XORWOW is linear over GF(2): one 160x160 bit matrix per step. Subsequence y starts
 * 2^67 * y steps into the stream (curand documentation; _skipahead_sequence_scratch). */
typedef struct { uint32_t m[160][5]; } Mat;   /* row r = image of basis vector r */
static void vec_step(uint32_t v[5]) {
	uint32_t t = v[0] ^ (v[0] >> 2);
	v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
	v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}
static void mat_apply(const Mat *M, const uint32_t in[5], uint32_t out[5]) {
	uint32_t o[5] = {0, 0, 0, 0, 0};
	for (int r = 0; r < 160; ++r)
		if ((in[r >> 5] >> (r & 31)) & 1u) for (int k = 0; k < 5; ++k) o[k] ^= M->m[r][k];
	memcpy(out, o, sizeof(o));
}
static void mat_square(const Mat *A, Mat *out) { for (int r = 0; r < 160; ++r) mat_apply(A, A->m[r], out->m[r]); }

/* InitRandomStates, APD.cu:791-804: curand_init(seed, subsequence = y, offset = x) */
static void init_rng(Ctx *c, unsigned long long seed) {
	uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
	uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
	Rng base; base.d = 6615241u + t1 + t0;
	base.v[0] = 123456789u + t0; base.v[1] = 362436069u ^ t0; base.v[2] = 521288629u + t1;
	base.v[3] = 88675123u ^ t1; base.v[4] = 5783321u + t0;
	Mat *A = (Mat *)malloc(sizeof(Mat)), *B = (Mat *)malloc(sizeof(Mat));
	for (int r = 0; r < 160; ++r) { memset(A->m[r], 0, 20); A->m[r][r >> 5] = 1u << (r & 31); vec_step(A->m[r]); }
	for (int k = 0; k < 67; ++k) { mat_square(A, B); Mat *t = A; A = B; B = t; }   /* A = M^(2^67) */
	Rng row = base;
	for (int y = 0; y < c->H; ++y) {
		Rng s = row;
		for (int x = 0; x < c->W; ++x) { c->rng[(size_t)y * c->W + x] = s; (void)rng_next(&s); }
		mat_apply(A, row.v, row.v);   /* next subsequence; d is unchanged (2^67 * 362437 = 0 mod 2^32) */
	}
	free(A); free(B);
}

/* ---- texture unit model -------------------------------------------------------------------------
 * Clamp addressing, linear filter. Measured on B200 (tools/tex_probe*.cu, 2.3e5 random samples): the
 * coordinate minus 0.5 is rounded to nearest in 1/256 units (integer part i, 8-bit fraction a8, b8), the
 * four weights are the integers w11 = (a8*b8 + 128) >> 8, w10 = a8 - w11, w01 = b8 - w11,
 * w00 = 256 - a8 - b8 + w11 (sum exactly 256), and the result is the weighted texel sum / 256 rounded
 * once (99.65 % bit-identical to the hardware, the rest 1 ulp off). */
static float tex2d(const float *img, int W, int H, float x, float y) {
	double xs = floor(((double)x - 0.5) * 256.0 + 0.5), ys = floor(((double)y - 0.5) * 256.0 + 0.5);
	if (!(xs > -1e12)) xs = -1e12; if (xs > 1e12) xs = 1e12;      /* NaN / huge coordinates clamp to an edge */
	if (!(ys > -1e12)) ys = -1e12; if (ys > 1e12) ys = 1e12;
	long long xq = (long long)xs, yq = (long long)ys;
	long long i = xq >> 8, j = yq >> 8; int a8 = (int)(xq & 255), b8 = (int)(yq & 255);
	long long i0 = i < 0 ? 0 : (i > W - 1 ? W - 1 : i), i1 = i + 1 < 0 ? 0 : (i + 1 > W - 1 ? W - 1 : i + 1);
	long long j0 = j < 0 ? 0 : (j > H - 1 ? H - 1 : j), j1 = j + 1 < 0 ? 0 : (j + 1 > H - 1 ? H - 1 : j + 1);
	double t00 = img[j0 * W + i0], t10 = img[j0 * W + i1], t01 = img[j1 * W + i0], t11 = img[j1 * W + i1];
	int w11 = (a8 * b8 + 128) >> 8, w10 = a8 - w11, w01 = b8 - w11, w00 = 256 - a8 - b8 + w11;
	return (float)((w00 * t00 + w10 * t10 + w01 * t01 + w11 * t11) / 256.0);
}

/* ---- small helpers --------------------------------------------------------------------------- */
static void get3d(const Camera *cam, float px, float py, float depth, float X[3]) {      /* APD.cu:159-164 */
	X[0] = depth * (px - cam->K[2]) / cam->K[0]; X[1] = depth * (py - cam->K[5]) / cam->K[4]; X[2] = depth;
}
static float dist2origin(const Camera *cam, int px, int py, float depth, f4 n) {          /* APD.cu:187-192 */
	float X[3]; get3d(cam, (float)px, (float)py, depth, X);
	return -(n.x * X[0] + n.y * X[1] + n.z * X[2]);
}
static float depth_from_plane(const Camera *cam, f4 pl, int px, int py) {                /* APD.cu:206-209 */
	return -pl.w * cam->K[0] / ((px - cam->K[2]) * pl.x + (cam->K[0] / cam->K[4]) * (py - cam->K[5]) * pl.y + cam->K[0] * pl.z);
}
static void normalize3(f4 *v) { float r = 1.0f / sqrtf(v->x * v->x + v->y * v->y + v->z * v->z); v->x *= r; v->y *= r; v->z *= r; }
static f4 view_dir(const Camera *cam, int px, int py, float depth) {                      /* APD.cu:173-185 */
	float X[3]; get3d(cam, (float)px, (float)py, depth, X);
	float n = sqrtf(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
	f4 v = {X[0] / n, X[1] / n, X[2] / n, 0}; return v;
}
static f4 random_normal(const Camera *cam, int px, int py, Rng *r, float depth) {         /* APD.cu:211-237 */
	float q1 = 1, q2 = 1, s = 2;
	while (s >= 1.0f) { q1 = 2.0f * rng_uniform(r) - 1.0f; q2 = 2.0f * rng_uniform(r) - 1.0f; s = q1 * q1 + q2 * q2; }
	float sq = sqrtf(1.0f - s);
	f4 n = {2.0f * q1 * sq, 2.0f * q2 * sq, 1.0f - 2.0f * s, 0};
	f4 v = view_dir(cam, px, py, depth);
	if (n.x * v.x + n.y * v.y + n.z * v.z > 0.0f) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
	normalize3(&n); return n;
}
static f4 perturbed_normal(const Camera *cam, int px, int py, f4 n, Rng *r, float pert) { /* APD.cu:239-274 */
	f4 v = view_dir(cam, px, py, 1.0f);
	float a1 = (rng_uniform(r) - 0.5f) * pert, a2 = (rng_uniform(r) - 0.5f) * pert, a3 = (rng_uniform(r) - 0.5f) * pert;
	float s1 = sinf(a1), s2 = sinf(a2), s3 = sinf(a3), c1 = cosf(a1), c2 = cosf(a2), c3 = cosf(a3);
	float R[9] = {c2 * c3, c3 * s1 * s2 - c1 * s3, s1 * s3 + c1 * c3 * s2, c2 * s3, c1 * c3 + s1 * s2 * s3,
	              c1 * s2 * s3 - c3 * s1, -s2, c2 * s1, c1 * c2};
	f4 o = {R[0] * n.x + R[1] * n.y + R[2] * n.z, R[3] * n.x + R[4] * n.y + R[5] * n.z, R[6] * n.x + R[7] * n.y + R[8] * n.z, 0};
	if (o.x * v.x + o.y * v.y + o.z * v.z >= 0.0f) o = n;
	normalize3(&o); o.w = 0; return o;
}

/* ComputeHomography, APD.cu:303-363 */
static void homography(const Camera *rc, const Camera *sc, f4 pl, float H[9]) {
	float rC[3], sC[3], Rr[9], Cr[3], tr[3], T[9];
	for (int k = 0; k < 3; ++k) {
		rC[k] = -(rc->R[k] * rc->t[0] + rc->R[3 + k] * rc->t[1] + rc->R[6 + k] * rc->t[2]);
		sC[k] = -(sc->R[k] * sc->t[0] + sc->R[3 + k] * sc->t[1] + sc->R[6 + k] * sc->t[2]);
	}
	for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q)
		Rr[3 * r + q] = sc->R[3 * r] * rc->R[3 * q] + sc->R[3 * r + 1] * rc->R[3 * q + 1] + sc->R[3 * r + 2] * rc->R[3 * q + 2];
	for (int k = 0; k < 3; ++k) Cr[k] = rC[k] - sC[k];
	for (int r = 0; r < 3; ++r) tr[r] = sc->R[3 * r] * Cr[0] + sc->R[3 * r + 1] * Cr[1] + sc->R[3 * r + 2] * Cr[2];
	for (int r = 0; r < 3; ++r) {
		H[3 * r + 0] = Rr[3 * r + 0] - tr[r] * pl.x / pl.w;
		H[3 * r + 1] = Rr[3 * r + 1] - tr[r] * pl.y / pl.w;
		H[3 * r + 2] = Rr[3 * r + 2] - tr[r] * pl.z / pl.w;
	}
	for (int r = 0; r < 3; ++r) {
		T[3 * r + 0] = H[3 * r + 0] / rc->K[0];
		T[3 * r + 1] = H[3 * r + 1] / rc->K[4];
		T[3 * r + 2] = -H[3 * r + 0] * rc->K[2] / rc->K[0] - H[3 * r + 1] * rc->K[5] / rc->K[4] + H[3 * r + 2];
	}
	for (int q = 0; q < 3; ++q) {
		H[q] = sc->K[0] * T[q] + sc->K[2] * T[6 + q];
		H[3 + q] = sc->K[4] * T[3 + q] + sc->K[5] * T[6 + q];
		H[6 + q] = sc->K[8] * T[6 + q];
	}
}
static void warp(const float H[9], float x, float y, float *ox, float *oy) {               /* APD.cu:365-372 */
	float z = H[6] * x + H[7] * y + H[8];
	*ox = (H[0] * x + H[1] * y + H[2]) / z; *oy = (H[3] * x + H[4] * y + H[5]) / z;
}

/* ComputeBilateralNCCOld, APD.cu:530-614 (weight is the constant 1.0f, :575) */
static float ncc_old(const Ctx *c, int px, int py, int src, f4 pl) {
	const Camera *rc = &c->cams[0], *sc = &c->cams[src];
	const float *ref = c->images, *img = c->images + (size_t)src * c->W * c->H;
	float H[9]; homography(rc, sc, pl, H);
	float cx, cy; warp(H, (float)px, (float)py, &cx, &cy);
	if (cx >= sc->width || cx < 0.0f || cy >= sc->height || cy < 0.0f) return 2.0f;
	const int radius = c->p.strong_radius, inc = c->p.strong_increment;
	float sr = 0, srr = 0, ss = 0, sss = 0, srs = 0, sw = 0;
	for (int i = -radius; i <= radius; i += inc) {
		float a = 0, aa = 0, b = 0, bb = 0, ab = 0, w = 0;
		for (int j = -radius; j <= radius; j += inc) {
			float rp = tex2d(ref, c->W, c->H, px + i + 0.5f, py + j + 0.5f);
			float sx, sy; warp(H, (float)(px + i), (float)(py + j), &sx, &sy);
			float sp = tex2d(img, c->W, c->H, sx + 0.5f, sy + 0.5f);
			a += rp; aa += rp * rp; b += sp; bb += sp * sp; ab += rp * sp; w += 1.0f;
		}
		sr += a; srr += aa; ss += b; sss += bb; srs += ab; sw += w;
	}
	float inv = 1.0f / sw; sr *= inv; srr *= inv; ss *= inv; sss *= inv; srs *= inv;
	float vr = srr - sr * sr, vs = sss - ss * ss;
	if (vr < 1e-5f || vs < 1e-5f) return 2.0f;
	float cost = 1.0f - (srs - sr * ss) / sqrtf(vr * vs);
	return fmaxf(0.0f, fminf(2.0f, cost));
}

/* ComputeGeomConsistencyCost, APD.cu:752-789 (+ Get3DPointonWorld_cu :718, ProjectonCamera_cu :740) */
static void to_world(const Camera *cam, float x, float y, float d, float P[3]) {
	float X[3]; get3d(cam, x, y, d, X);
	for (int k = 0; k < 3; ++k) P[k] = cam->R[k] * X[0] + cam->R[3 + k] * X[1] + cam->R[6 + k] * X[2] + cam->c[k];
}
static void project(const Camera *cam, const float P[3], float *x, float *y) {
	float t[3];
	for (int k = 0; k < 3; ++k) t[k] = cam->R[3 * k] * P[0] + cam->R[3 * k + 1] * P[1] + cam->R[3 * k + 2] * P[2] + cam->t[k];
	float d = cam->K[6] * t[0] + cam->K[7] * t[1] + cam->K[8] * t[2];
	*x = (cam->K[0] * t[0] + cam->K[1] * t[1] + cam->K[2] * t[2]) / d;
	*y = (cam->K[3] * t[0] + cam->K[4] * t[1] + cam->K[5] * t[2]) / d;
}
static float geom_cost(const Ctx *c, int px, int py, int src, f4 pl) {
	const Camera *rc = &c->cams[0], *sc = &c->cams[src];
	float depth = depth_from_plane(rc, pl, px, py), P[3], sx, sy;
	to_world(rc, (float)px, (float)py, depth, P);
	project(sc, P, &sx, &sy);
	float sd = tex2d(c->depths + (size_t)src * c->W * c->H, c->W, c->H, (int)sx + 0.5f, (int)sy + 0.5f);
	if (sd == 0.0f) return 3.0f;
	float Q[3], bx, by; to_world(sc, sx, sy, sd, Q); project(rc, Q, &bx, &by);
	float dc = px - bx, dr = py - by;
	return fminf(3.0f, sqrtf(dc * dc + dr * dr));
}

static int isset(uint32_t v, int n) { return (v >> n) & 1u; }

/* RandomInitialization, APD.cu:806-835 with :616-693 */
static void init_pixel(Ctx *c, int px, int py) {
	const int S = c->p.num_images - 1; const size_t ctr = (size_t)py * c->W + px; const Camera *rc = &c->cams[0];
	if (c->p.state == FIRST_INIT) {
		Rng *r = &c->rng[ctr];
		float depth = rng_uniform(r) * (c->p.depth_max - c->p.depth_min) + c->p.depth_min;     /* :278 */
		f4 pl = random_normal(rc, px, py, r, depth); pl.w = dist2origin(rc, px, py, depth, pl);
		c->planes[ctr] = pl;
		float cv[MAX_IMAGES], cs[MAX_IMAGES]; int valid = 0;
		for (int v = 0; v < S; ++v) { cv[v] = cs[v] = ncc_old(c, px, py, v + 1, pl); if (cv[v] < 2.0f) valid++; }
		for (int i = 1; i < S; ++i) { float t = cs[i]; int j = i; for (; j >= 1 && t < cs[j - 1]; --j) cs[j] = cs[j - 1]; cs[j] = t; }
		c->views[ctr] = 0; int k = valid < c->p.top_k ? valid : c->p.top_k;
		if (k > 0) { float sum = 0; for (int i = 0; i < k; ++i) sum += cs[i];
			for (int v = 0; v < S; ++v) if (cv[v] <= cs[k - 1]) c->views[ctr] |= 1u << v;
			c->costs[ctr] = sum / k; } else c->costs[ctr] = 2.0f;
	} else {
		f4 in = c->planes[ctr], pl;                                                             /* :827-832 */
		pl.x = rc->R[0] * in.x + rc->R[1] * in.y + rc->R[2] * in.z; pl.y = rc->R[3] * in.x + rc->R[4] * in.y + rc->R[5] * in.z;
		pl.z = rc->R[6] * in.x + rc->R[7] * in.y + rc->R[8] * in.z; pl.w = dist2origin(rc, px, py, in.w, pl);
		c->planes[ctr] = pl;
		int cnt = 0; float sum = 0;
		for (int v = 0; v < S; ++v) if (isset(c->views[ctr], v)) {
			float cost = ncc_old(c, px, py, v + 1, pl);
			if (cost < 2.0f) { cnt++; sum += cost; } else c->views[ctr] &= (0xFFFFFFFEu << v);   /* unSetBit quirk :47-50 */
		}
		c->costs[ctr] = cnt == 0 ? 2.0f : sum / cnt;
	}
}

static void cost_vector(const Ctx *c, int px, int py, f4 pl, float *out) {                 /* APD.cu:707-716 */
	for (int v = 0; v < c->p.num_images - 1; ++v) out[v] = ncc_old(c, px, py, v + 1, pl);
}

/* CheckerboardPropagationStrong + PlaneHypothesisRefinementStrong, APD.cu:982-1321, :837-890 */
static void strong_pixel(Ctx *c, int px, int py, int iter) {
	const int W = c->W, H = c->H, S = c->p.num_images - 1, ctr = py * W + px; const Camera *rc = &c->cams[0];
	float ca[8][32]; memset(ca, 0, sizeof(ca)); ca[0][0] = 2.0f;          /* `= {2.0f}` sets [0][0] only (:1004) */
	int flag[8] = {0}, pos[8] = {0}; const float *costs = c->costs;
#define TRY(P) do { int q_ = (P); if (costs[q_] < cmin) { cmin = costs[q_]; best = q_; } } while (0)
	float cmin; int best;
	if (py > 2) { flag[1] = 1; best = ctr - 3 * W; cmin = costs[best]; for (int i = 1; i < 11; ++i) if (py > 2 + 2 * i) TRY(ctr - 3 * W - 2 * i * W); pos[1] = best; }
	if (py < H - 3) { flag[3] = 1; best = ctr + 3 * W; cmin = costs[best]; for (int i = 1; i < 11; ++i) if (py < H - 3 - 2 * i) TRY(ctr + 3 * W + 2 * i * W); pos[3] = best; }
	if (px > 2) { flag[5] = 1; best = ctr - 3; cmin = costs[best]; for (int i = 1; i < 11; ++i) if (px > 2 + 2 * i) TRY(ctr - 3 - 2 * i); pos[5] = best; }
	if (px < W - 3) { flag[7] = 1; best = ctr + 3; cmin = costs[best]; for (int i = 1; i < 11; ++i) if (px < W - 3 - 2 * i) TRY(ctr + 3 + 2 * i); pos[7] = best; }
	if (py > 0) { flag[0] = 1; best = ctr - W; cmin = costs[best];
		for (int i = 0; i < 3; ++i) { if (py > 1 + i && px > i) TRY(ctr - W - (1 + i) * W - (1 + i)); if (py > 1 + i && px < W - 1 - i) TRY(ctr - W - (1 + i) * W + (1 + i)); } pos[0] = best; }
	if (py < H - 1) { flag[2] = 1; best = ctr + W; cmin = costs[best];
		for (int i = 0; i < 3; ++i) { if (py < H - 2 - i && px > i) TRY(ctr + W + (1 + i) * W - (1 + i)); if (py < H - 2 - i && px < W - 1 - i) TRY(ctr + W + (1 + i) * W + (1 + i)); } pos[2] = best; }
	if (px > 0) { flag[4] = 1; best = ctr - 1; cmin = costs[best];
		for (int i = 0; i < 3; ++i) { if (px > 1 + i && py > i) TRY(ctr - 1 - (1 + i) - (1 + i) * W); if (px > 1 + i && py < H - 1 - i) TRY(ctr - 1 - (1 + i) + (1 + i) * W); } pos[4] = best; }
	if (px < W - 1) { flag[6] = 1; best = ctr + 1; cmin = costs[best];
		for (int i = 0; i < 3; ++i) { if (px < W - 2 - i && py > i) TRY(ctr + 1 + (1 + i) - (1 + i) * W); if (px < W - 2 - i && py < H - 1 - i) TRY(ctr + 1 + (1 + i) + (1 + i) * W); } pos[6] = best; }
#undef TRY
	for (int k = 0; k < 8; ++k) if (flag[k]) cost_vector(c, px, py, c->planes[pos[k]], ca[k]);

	uint8_t *vw = &c->vw[(size_t)ctr * MAX_IMAGES]; memset(vw, 0, MAX_IMAGES);
	float prior[32] = {0}, prob[32] = {0};
	const int nb[4] = {ctr - W, ctr + W, ctr - 1, ctr + 1};
	for (int i = 0; i < 4; ++i) if (flag[2 * i]) for (int v = 0; v < S; ++v) prior[v] += isset(c->views[nb[i]], v) ? 0.9f : 0.1f;
	float thr = (float)(0.8 * expf((iter * iter) / (-90.0f)));                               /* :1225 */
	for (int v = 0; v < S; ++v) {
		float count = 0, tmpw = 0; int bad = 0;
		for (int k = 0; k < 8; ++k) { if (ca[k][v] < thr) { tmpw += expf(ca[k][v] * ca[k][v] / (-0.18f)); count++; } if (ca[k][v] > 1.2f) bad++; }
		if (count > 2 && bad < 3) prob[v] = tmpw / count; else if (bad < 3) prob[v] = expf(thr * thr / (-0.32f));
		prob[v] *= prior[v];
	}
	{ float sum = 0; for (int v = 0; v < S; ++v) sum += prob[v]; float inv = 1.0f / sum, cum = 0;      /* TransformPDFToCDF :143 */
	  for (int v = 0; v < S; ++v) { cum += prob[v] * inv; prob[v] = cum; } }
	Rng *r = &c->rng[ctr];
	for (int s = 0; s < 15; ++s) { float u = rng_uniform(r) - FLT_EPSILON; for (int v = 0; v < S; ++v) if (prob[v] > u) { vw[v]++; break; } }
	uint32_t sel = 0; float wn = 0; for (int v = 0; v < S; ++v) if (vw[v] > 0) { sel |= 1u << v; wn += vw[v]; }
	float fc[8]; for (int k = 0; k < 8; ++k) { fc[k] = 0; for (int v = 0; v < S; ++v) if (vw[v] > 0) fc[k] += vw[v] * ca[k][v]; fc[k] /= wn; }
	int mi = 0; { float m = fc[0]; for (int k = 1; k < 8; ++k) if (fc[k] <= m) { m = fc[k]; mi = k; } }    /* :29-40 */
	float cvn[32]; cost_vector(c, px, py, c->planes[ctr], cvn);
	float cost_now = 0; for (int v = 0; v < S; ++v) cost_now += vw[v] * cvn[v]; cost_now /= wn;
	c->costs[ctr] = cost_now;
	float depth_now = depth_from_plane(rc, c->planes[ctr], px, py); f4 pl = c->planes[ctr];
	if (flag[mi]) { float d = depth_from_plane(rc, c->planes[pos[mi]], px, py);
		if (d >= c->p.depth_min && d <= c->p.depth_max && fc[mi] < cost_now) { depth_now = d; pl = c->planes[pos[mi]]; cost_now = fc[mi]; c->views[ctr] = sel; } }
	{   /* refinement */
		float dmin = c->p.depth_min, dmax = c->p.depth_max;
		float drand = rng_uniform(r) * (dmax - dmin) + dmin; f4 nrand = random_normal(rc, px, py, r, depth_now);
		float lo = (1 - 0.02f) * depth_now, hi = (1 + 0.02f) * depth_now, dpert = rng_uniform(r) * (hi - lo) + lo;   /* loop never repeats (:860-862) */
		f4 npert = perturbed_normal(rc, px, py, pl, r, (float)(0.02f * 3.14159265358979323846));
		float ds[5] = {drand, depth_now, drand, depth_now, dpert}; f4 ns[5] = {pl, nrand, nrand, npert, pl};
		for (int i = 0; i < 5; ++i) { f4 t = ns[i]; t.w = dist2origin(rc, px, py, ds[i], t);
			float cv[32]; cost_vector(c, px, py, t, cv); float tc = 0; for (int v = 0; v < S; ++v) if (vw[v] > 0) tc += vw[v] * cv[v]; tc /= wn;
			float d = depth_from_plane(rc, t, px, py);
			if (d >= dmin && d <= dmax && tc < cost_now) { depth_now = d; pl = t; cost_now = tc; } }
	}
	if (c->p.state == REFINE_INIT) { if (cost_now < c->costs[ctr] - 0.1) { c->costs[ctr] = cost_now; c->planes[ctr] = pl; } }
	else { c->costs[ctr] = cost_now; c->planes[ctr] = pl; }
}

/* CheckerboardFilterStrong, APD.cu:1604-1714 */
static void filter_pixel(Ctx *c, int px, int py) {
	const int W = c->W, H = c->H, ctr = py * W + px; float f[21]; int n = 0;
	f[n++] = c->planes[ctr].w;
	if (c->costs[ctr] < 0.001f) return;
#define ADD(cond, off) do { if ((cond) && c->states[ctr + (off)] == STRONG) f[n++] = c->planes[ctr + (off)].w; } while (0)
	ADD(py > 0, -W); ADD(py > 2, -3 * W); ADD(py > 4, -5 * W); ADD(py < H - 1, W); ADD(py < H - 3, 3 * W); ADD(py < H - 5, 5 * W);
	ADD(px > 0, -1); ADD(px > 2, -3); ADD(px > 4, -5); ADD(px < W - 1, 1); ADD(px < W - 3, 3); ADD(px < W - 5, 5);
	ADD(py > 0 && px < W - 2, -W + 2); ADD(py < H - 1 && px < W - 2, W + 2); ADD(py > 0 && px > 1, -W - 2); ADD(py < H - 1 && px > 1, W - 2);
	ADD(px > 0 && py > 2, -1 - 2 * W); ADD(px < W - 1 && py > 2, 1 - 2 * W); ADD(px > 0 && py < H - 2, -1 + 2 * W); ADD(px < W - 1 && py < H - 2, 1 + 2 * W);
#undef ADD
	for (int i = 1; i < n; ++i) { float t = f[i]; int j = i; for (; j >= 1 && t < f[j - 1]; --j) f[j] = f[j - 1]; f[j] = t; }
	c->planes[ctr].w = (n % 2 == 0) ? (f[n / 2 - 1] + f[n / 2]) / 2 : f[n / 2];
}

/* shared front end of DepthToWeak / LocalRefine, APD.cu:2012-2052 / :2160-2199 */
static int sweep_front(const Ctx *c, int px, int py, f4 *pl, float *depth, float *wn, float *base, float *cost_now) {
	const size_t ctr = (size_t)py * c->W + px; const Camera *rc = &c->cams[0]; const int S = c->p.num_images - 1;
	f4 in = c->planes[ctr];
	pl->x = rc->R[0] * in.x + rc->R[1] * in.y + rc->R[2] * in.z; pl->y = rc->R[3] * in.x + rc->R[4] * in.y + rc->R[5] * in.z;
	pl->z = rc->R[6] * in.x + rc->R[7] * in.y + rc->R[8] * in.z; pl->w = in.w; *depth = in.w;
	if (*depth == 0) return -1;
	int valid = 0; *wn = 0; *base = 0; *cost_now = 0;
	for (int v = 0; v < S; ++v) if (isset(c->views[ctr], v)) {
		f4 t = *pl; t.w = dist2origin(rc, px, py, *depth, t);
		float tc = ncc_old(c, px, py, v + 1, t);
		if (c->p.geom_consistency) tc += c->p.geom_factor * geom_cost(c, px, py, v + 1, t);
		uint8_t w = c->vw[ctr * MAX_IMAGES + v];
		*cost_now += tc * w; *wn += w;
		float d0 = rc->c[0] - c->cams[v + 1].c[0], d1 = rc->c[1] - c->cams[v + 1].c[1], d2 = rc->c[2] - c->cams[v + 1].c[2];
		*base += sqrtf(d0 * d0 + d1 * d1 + d2 * d2); valid++;
	}
	return valid;
}
static float sweep_cost(const Ctx *c, int px, int py, f4 pl, float d, float wn) {
	const size_t ctr = (size_t)py * c->W + px; const int S = c->p.num_images - 1;
	f4 t = pl; t.w = dist2origin(&c->cams[0], px, py, d, t); float pc = 0;
	for (int v = 0; v < S; ++v) if (isset(c->views[ctr], v)) {
		float tc = ncc_old(c, px, py, v + 1, t);
		if (c->p.geom_consistency) tc += c->p.geom_factor * geom_cost(c, px, py, v + 1, t);
		pc += tc * c->vw[ctr * MAX_IMAGES + v];
	}
	return pc / wn;
}
/* DepthToWeak, APD.cu:1990-2144 */
static void classify_pixel(Ctx *c, int px, int py) {
	const size_t ctr = (size_t)py * c->W + px;
	if (px < 6 || py < 6 || px >= c->W - 6 || py >= c->H - 6) { c->states[ctr] = UNKNOWN; return; }
	f4 pl; float depth, wn, base, cn; int valid = sweep_front(c, px, py, &pl, &depth, &wn, &base, &cn);
	if (valid <= 0) { c->states[ctr] = UNKNOWN; return; }
	base /= valid; float disp = c->cams[0].K[0] * base / depth; float pc[61];
	for (int k = -30; k <= 30; ++k) {
		float d = c->cams[0].K[0] * base / (disp + k);
		if (d < c->p.depth_min || d > c->p.depth_max) { pc[k + 30] = 2.0f; continue; }
		float v = sweep_cost(c, px, py, pl, d, wn); pc[k + 30] = (2.0f > v) ? v : 2.0f;      /* OpenCV MIN */
	}
	int peak[61] = {0}, peaks = 0, minp = 0; float minc = 2.0f;
	for (int i = 2; i < 59; ++i) if (pc[i - 1] > pc[i] && pc[i + 1] > pc[i]) { peak[i] = 1; peaks++; if (pc[i] < minc) { minp = i; minc = pc[i]; } }
	if (abs(minp - 30) > c->p.weak_peak_radius || pc[minp] > 0.5f) { c->states[ctr] = WEAK; return; }
	if (peaks == 1) { c->states[ctr] = pc[minp] <= 0.15f ? STRONG : WEAK; return; }
	float var = 0; for (int i = 2; i < 59; ++i) if (peak[i] && i != minp) { float d = pc[i] - minc; var += d * d; }
	var = sqrtf(var) / (peaks - 1);
	c->states[ctr] = var > 0.2f ? STRONG : WEAK;
}
/* LocalRefine, APD.cu:2146-2232 */
static void refine_pixel(Ctx *c, int px, int py) {
	const size_t ctr = (size_t)py * c->W + px;
	f4 pl; float depth, wn, base, cn; int valid = sweep_front(c, px, py, &pl, &depth, &wn, &base, &cn);
	if (valid <= 0 || wn == 0) return;
	cn /= wn; base /= valid; float disp = c->cams[0].K[0] * base / depth, minc = 2.0f, bestd = depth;
	for (int k = -5; k <= 5; ++k) {
		float d = c->cams[0].K[0] * base / (disp + k);
		if (d < c->p.depth_min || d > c->p.depth_max) continue;
		float v = sweep_cost(c, px, py, pl, d, wn); if (v < minc) { minc = v; bestd = d; }
	}
	if (cn - minc > 0.1) c->planes[ctr].w = bestd;
}

/* =================================================================================================
 * Adaptive patch deformation: the kernels that touch WEAK pixels.
 * ================================================================================================= */
#define NEIGHBOUR_NUM 9
#define MAX_SEARCH_RADIUS 4096

/* FindNearestStrongPoint, APD.cu:2234-2270 (scan order x outer, y inner; strict <) */
static void nearest_strong_pixel(Ctx *c, int px, int py) {
	const int W = c->W, H = c->H; const size_t ctr = (size_t)py * W + px;
	c->nearest[2 * ctr] = -1; c->nearest[2 * ctr + 1] = -1;
	if (c->states[ctr] != WEAK) return;
	float min_dist = 255.0f;
	for (int x = -100; x <= 100; ++x) for (int y = -100; y <= 100; ++y) {
		const int nx = px + x, ny = py + y;
		if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
		if (c->states[(size_t)ny * W + nx] == STRONG) {
			float dist = sqrtf((float)(x * x + y * y));
			if (dist < min_dist) { min_dist = dist; c->nearest[2 * ctr] = (int16_t)nx; c->nearest[2 * ctr + 1] = (int16_t)ny; }
		}
	}
}

/* NormalizeVec2, APD.cu:135-141. The FMAs below are where the reference's compiler contracts (read off its SASS);
 * its rsqrt/sqrt are MUFU approximations, which a CPU cannot reproduce: agreement with the GPU is statistical. */
static void normalize2(float *x, float *y) { float r = 1.0f / sqrtf(fmaf(*x, *x, *y * *y)); *x *= r; *y *= r; }
typedef struct { float x, y, z; } f3;
static f3 point3(const Camera *cam, int x, int y, float depth) { float X[3]; get3d(cam, (float)x, (float)y, depth, X); f3 p = {X[0], X[1], X[2]}; return p; }
/* PointinTriangle, APD.cu:91-112 */
static int point_in_triangle(const int16_t *A, const int16_t *B, const int16_t *C, int px, int py) {
	float abx = (float)(B[0] - A[0]), aby = (float)(B[1] - A[1]), bcx = (float)(C[0] - B[0]), bcy = (float)(C[1] - B[1]);
	float cax = (float)(A[0] - C[0]), cay = (float)(A[1] - C[1]);
	float ab = sqrtf(abx * abx + aby * aby), bc = sqrtf(bcx * bcx + bcy * bcy), ca = sqrtf(cax * cax + cay * cay);
	if (ab <= 2.0f || bc <= 2.0f || ca <= 2.0f) return 0;
	if (!(ab + bc > ca && bc + ca > ab && ab + ca > bc)) return 0;
	float pax = (float)(A[0] - px), pay = (float)(A[1] - py), pbx = (float)(B[0] - px), pby = (float)(B[1] - py);
	float pcx = (float)(C[0] - px), pcy = (float)(C[1] - py);
	float t1 = pax * pby - pay * pbx, t2 = pbx * pcy - pby * pcx, t3 = pcx * pay - pcy * pax;
	return t1 * t2 >= 0.0f && t1 * t3 >= 0.0f;
}
/* unit-normal plane through three points (APD.cu:1897-1907, 2338-2349); 0 if degenerate */
static int plane_from_points(f3 A, f3 B, f3 C, f4 *pl) {
	float acx = A.x - C.x, acy = A.y - C.y, acz = A.z - C.z, bcx = B.x - C.x, bcy = B.y - C.y, bcz = B.z - C.z;
	f4 n = {acy * bcz - bcy * acz, -(acx * bcz - bcx * acz), acx * bcy - bcx * acy, 0};
	if ((n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) || isnan(n.x) || isnan(n.y) || isnan(n.z)) return 0;
	normalize3(&n);
	n.w = -(n.x * A.x + n.y * A.y + n.z * A.z);
	*pl = n; return 1;
}
static float plane_dist(f4 pl, f3 p) { return fabsf(pl.x * p.x + pl.y * p.y + pl.z * p.z + pl.w); }

/* GenNeighbours, APD.cu:1750-1969: deformable anchors along 8 directions x rotate_time sub-rotations, then a
 * 50-draw RANSAC plane through the anchors' 3-D points; the anchors are sorted by their distance to that plane. */
static void gen_anchors_pixel(Ctx *c, int px, int py) {
	const int W = c->W, H = c->H; const size_t ctr = (size_t)py * W + px; const Camera *rc = &c->cams[0];
	if (c->states[ctr] != WEAK) return;
	int16_t *an = &c->anchors[ctr * NEIGHBOUR_NUM * 2];
	for (int k = 0; k < NEIGHBOUR_NUM * 2; ++k) an[k] = -1;
	an[0] = (int16_t)px; an[1] = (int16_t)py;
	Rng *r = &c->rng[ctr];
	const int rotate_time = c->p.rotate_time;
	const float angle = 45.0f / rotate_time;                                                  /* :1790-1795 */
	const float cos_a = (float)cos(angle * M_PI / 180.f), sin_a = (float)sin(angle * M_PI / 180.f);
	const float thresh = (float)cos((angle / 2.0f) * M_PI / 180.0f);
	int shift_range = (int)(tan((angle / 2.0f) * M_PI / 180.0f) * 20); if (shift_range < 1) shift_range = 1;
	int16_t sp[32][2]; unsigned valid = 0u; int found = 0;
	for (int i = 0; i < 32; ++i) sp[i][0] = sp[i][1] = -1;
	int base = -1;
	for (int ox = -1; ox <= 1; ++ox) for (int oy = -1; oy <= 1; ++oy) {
		if (ox == 0 && oy == 0) continue;
		float dx = (float)ox, dy = (float)oy; normalize2(&dx, &dy);
		++base;
		for (int rot = 0; rot < rotate_time; ++rot) {
			const int di = base * 4 + rot;
			for (int radius = 2; radius <= MAX_SEARCH_RADIUS; radius = (radius * 2 < radius + 25) ? radius * 2 : radius + 25) {
				const float tx = fmaf((float)radius, dx, (float)px), ty = fmaf((float)radius, dy, (float)py);
				if (tx < 0.0f || ty < 0.0f || tx >= (float)W || ty >= (float)H) break;
				for (int t = 0; t < 4; ++t) {
					/* (curand() % 2 == 0 ? 1 : -1) * curand() % shift_range, evaluated in unsigned arithmetic (:1813-1814) */
					const uint32_t d1 = rng_next(r), d2 = rng_next(r), d3 = rng_next(r), d4 = rng_next(r);
					const uint32_t xs = (((d1 & 1u) == 0u) ? d2 : (0u - d2)) % (uint32_t)shift_range;
					const uint32_t ys = (((d3 & 1u) == 0u) ? d4 : (0u - d4)) % (uint32_t)shift_range;
					float ddx = fmaf(dx, 20.0f, (float)xs), ddy = fmaf(dy, 20.0f, (float)ys); normalize2(&ddx, &ddy);
					int nx = (int16_t)(int)fmaf((float)radius, ddx, (float)px), ny = (int16_t)(int)fmaf((float)radius, ddy, (float)py);
					if (nx < 6 || ny < 6 || nx >= W - 6 || ny >= H - 6) continue;
					size_t nc = (size_t)ny * W + nx;
					if (c->states[nc] != STRONG) {
						const int16_t sx = c->nearest[2 * nc], sy = c->nearest[2 * nc + 1];
						if (sx == -1 || sy == -1) continue;
						nx = sx; ny = sy;
					}
					float tdx = (float)(nx - px), tdy = (float)(ny - py); normalize2(&tdx, &tdy);
					if (fmaf(tdx, dx, tdy * dy) > thresh) { sp[di][0] = (int16_t)nx; sp[di][1] = (int16_t)ny; valid |= 1u << di; ++found; break; }
				}
				if ((valid >> di) & 1u) break;
			}
			const float rx = fmaf(dx, cos_a, -(dy * sin_a)), ry = fmaf(dx, sin_a, dy * cos_a);
			dx = rx; dy = ry; normalize2(&dx, &dy);
		}
	}
	if (found <= 3) { c->reliable[ctr] = 0; return; }
	int16_t pts[32][2]; f3 p3[32]; int vc = 0;
	const f3 c3 = point3(rc, px, py, c->planes[ctr].w);                 /* planes[].w still holds the prior depth here */
	for (int i = 0; i < 32; ++i) {
		pts[i][0] = pts[i][1] = -1;
		if ((valid >> i) & 1u) { pts[vc][0] = sp[i][0]; pts[vc][1] = sp[i][1];
			p3[vc] = point3(rc, sp[i][0], sp[i][1], c->planes[(size_t)sp[i][1] * W + sp[i][0]].w); ++vc; }
	}
	const float dd = c->p.depth_max - c->p.depth_min, thr = c->p.ransac_threshold;
	f4 best = {0, 0, 0, 0}; int ua = -1, ub = -1, uc = -1, max_count = 3, has = 0; float min_cost = FLT_MAX;
	for (int it = 0; it < 50; ++it) {
		const int ia = (int)(rng_next(r) % (uint32_t)vc), ib = (int)(rng_next(r) % (uint32_t)vc), ic = (int)(rng_next(r) % (uint32_t)vc);
		if (ia == ib || ib == ic || ia == ic) continue;
		if (!point_in_triangle(pts[ia], pts[ib], pts[ic], px, py)) continue;
		f4 pl; if (!plane_from_points(p3[ia], p3[ib], p3[ic], &pl)) continue;
		int cnt = 0; for (int s = 0; s < vc; ++s) if (plane_dist(pl, p3[s]) / dd < thr) ++cnt;
		if (cnt < 6) continue;
		if (cnt > max_count) { max_count = cnt; min_cost = plane_dist(pl, c3); best = pl; has = 1; ua = ia; ub = ib; uc = ic; }
		else if (cnt == max_count) { const float cd = plane_dist(pl, c3); if (cd < min_cost) { min_cost = cd; best = pl; ua = ia; ub = ib; uc = ic; } }
	}
	if (!has) { c->reliable[ctr] = 0; return; }
	float wgt[32];
	for (int i = 0; i < vc; ++i) {
		float d = plane_dist(best, p3[i]);
		if (d / dd >= thr) { pts[i][0] = pts[i][1] = -1; wgt[i] = FLT_MAX; continue; }
		if (i == ua || i == ub || i == uc) d -= 1.0f;
		wgt[i] = d;
	}
	for (int i = 1; i < vc; ++i) {                                          /* sort_small_weighted, APD.cu:14-27 */
		const int16_t t0 = pts[i][0], t1 = pts[i][1]; const float tw = wgt[i]; int j = i;
		for (; j >= 1 && tw < wgt[j - 1]; --j) { pts[j][0] = pts[j - 1][0]; pts[j][1] = pts[j - 1][1]; wgt[j] = wgt[j - 1]; }
		pts[j][0] = t0; pts[j][1] = t1; wgt[j] = tw;
	}
	for (int k = 1; k < NEIGHBOUR_NUM; ++k) { an[2 * k] = pts[k - 1][0]; an[2 * k + 1] = pts[k - 1][1]; }
	c->reliable[ctr] = 1;
}

/* RANSACToGetFitPlane, APD.cu:2272-2384 */
static void fit_plane_pixel(Ctx *c, int px, int py) {
	const int W = c->W; const size_t ctr = (size_t)py * W + px; const Camera *rc = &c->cams[0];
	if (c->states[ctr] != WEAK) { c->fit[ctr] = c->planes[ctr]; return; }
	const int16_t *an = &c->anchors[ctr * NEIGHBOUR_NUM * 2];
	int16_t pts[8][2]; f3 p3[8]; int cnt = 0;
	for (int k = 1; k < NEIGHBOUR_NUM; ++k) {
		const int sx = an[2 * k], sy = an[2 * k + 1];
		if (sx == -1 || sy == -1) continue;
		const float depth = depth_from_plane(rc, c->planes[(size_t)sy * W + sx], sx, sy);
		pts[cnt][0] = (int16_t)sx; pts[cnt][1] = (int16_t)sy; p3[cnt] = point3(rc, sx, sy, depth); ++cnt;
	}
	if (cnt < 3) { c->fit[ctr] = c->planes[ctr]; return; }
	Rng *r = &c->rng[ctr];
	float min_cost = FLT_MAX; f4 best = {0, 0, 0, 0}; int has = 0;
	for (int it = 0; it < 50; ++it) {
		const int ia = (int)(rng_next(r) % (uint32_t)cnt), ib = (int)(rng_next(r) % (uint32_t)cnt), ic = (int)(rng_next(r) % (uint32_t)cnt);
		if (ia == ib || ib == ic || ia == ic) continue;
		if (!point_in_triangle(pts[ia], pts[ib], pts[ic], px, py)) continue;
		f4 pl; if (!plane_from_points(p3[ia], p3[ib], p3[ic], &pl)) continue;
		float cost = 0.0f;
		for (int s = 0; s < cnt; ++s) { if (s == ia || s == ib || s == ic) continue; cost += plane_dist(pl, p3[s]); }
		if (cost < min_cost) { min_cost = cost; best = pl; has = 1; }
		if (min_cost == 0.0f) break;
	}
	if (has) {
		const float depth = depth_from_plane(rc, c->planes[ctr], px, py);
		f4 v = view_dir(rc, px, py, depth);
		if (v.x * best.x + v.y * best.y + v.z * best.z > 0.0f) { best.x = -best.x; best.y = -best.y; best.z = -best.z; best.w = -best.w; }
		c->fit[ctr] = best;
	} else { f4 z = {0, 0, 0, 0}; c->fit[ctr] = z; }
}

/* one window of ComputeBilateralNCCNew: radius 5, step `inc` around (cx, cy), all warped by H (APD.cu:424-520) */
static float ncc_window(const Ctx *c, int src, const float H[9], int cx, int cy, int inc) {
	const float *ref = c->images, *img = c->images + (size_t)src * c->W * c->H;
	float sr = 0, srr = 0, ss = 0, sss = 0, srs = 0, sw = 0;
	for (int i = -5; i <= 5; i += inc) {
		float a = 0, aa = 0, b = 0, bb = 0, ab = 0, w = 0;
		for (int j = -5; j <= 5; j += inc) {
			float rp = tex2d(ref, c->W, c->H, cx + i + 0.5f, cy + j + 0.5f);
			float sx, sy; warp(H, (float)(cx + i), (float)(cy + j), &sx, &sy);
			float sp = tex2d(img, c->W, c->H, sx + 0.5f, sy + 0.5f);
			a += rp; aa += rp * rp; b += sp; bb += sp * sp; ab += rp * sp; w += 1.0f;
		}
		sr += a; srr += aa; ss += b; sss += bb; srs += ab; sw += w;
	}
	float inv = 1.0f / sw; sr *= inv; srr *= inv; ss *= inv; sss *= inv; srs *= inv;
	float vr = srr - sr * sr, vs = sss - ss * ss;
	if (vr < 1e-5f || vs < 1e-5f) return 2.0f;
	return fmaxf(0.0f, fminf(2.0f, 1.0f - (srs - sr * ss) / sqrtf(vr * vs)));
}

/* ComputeBilateralNCCNew, APD.cu:400-528: the pixel's own 6x6 window (weight 0.25) + 3x3 windows on the anchors
 * (mean, weight 0.75); an anchor that projects outside the source view counts as cost 2 if the anchor selected that view */
static float ncc_deform(const Ctx *c, int px, int py, int src, f4 pl) {
	const Camera *rc = &c->cams[0], *sc = &c->cams[src];
	const int16_t *an = &c->anchors[((size_t)py * c->W + px) * NEIGHBOUR_NUM * 2];
	float H[9]; homography(rc, sc, pl, H);
	float cx, cy; warp(H, (float)px, (float)py, &cx, &cy);
	if (cx >= sc->width || cx < 0.0f || cy >= sc->height || cy < 0.0f) return 2.0f;
	float center_cost = 0.0f, strong_cost = 0.0f; int cnt = 0;
	for (int k = 0; k < NEIGHBOUR_NUM; ++k) {
		const int qx = an[2 * k], qy = an[2 * k + 1];
		if (qx == -1 || qy == -1) continue;
		float sx, sy; warp(H, (float)qx, (float)qy, &sx, &sy);
		if (sx < 0.0f || sy < 0.0f || sx >= (float)c->W || sy >= (float)c->H) {
			if (k == 0) return 2.0f;
			if (isset(c->views[(size_t)qy * c->W + qx], src - 1)) { strong_cost += 2.0f; ++cnt; }
			continue;
		}
		if (k == 0) center_cost = ncc_window(c, src, H, qx, qy, 2);
		else { strong_cost += ncc_window(c, src, H, qx, qy, 5); ++cnt; }
	}
	if (cnt == 0) return center_cost;
	strong_cost /= (float)cnt;
	if (strong_cost > 2.0f) strong_cost = 2.0f;
	return (float)(center_cost * 0.25 + strong_cost * 0.75);
}

static float weak_cost(const Ctx *c, int px, int py, f4 pl, const uint8_t *vw, float wn) {      /* APD.cu:918-927, 1464-1471 */
	const int S = c->p.num_images - 1; float acc = 0.0f;
	for (int v = 0; v < S; ++v) {
		if (vw[v] == 0) continue;
		float cost = ncc_deform(c, px, py, v + 1, pl);
		if (c->p.geom_consistency) cost += c->p.geom_factor * geom_cost(c, px, py, v + 1, pl);
		acc += vw[v] * cost;
	}
	return acc / wn;
}

/* CheckerboardPropagationWeak + PlaneHypothesisRefinementWeak, APD.cu:1323-1508, :892-980 */
static void weak_pixel(Ctx *c, int px, int py, int iter) {
	const int W = c->W, S = c->p.num_images - 1; const size_t ctr = (size_t)py * W + px; const Camera *rc = &c->cams[0];
	const int16_t *an = &c->anchors[ctr * NEIGHBOUR_NUM * 2];
	float ca[8][32]; memset(ca, 0, sizeof(ca)); ca[0][0] = 2.0f;          /* `= {2.0f}` sets [0][0] only (:1345) */
	int flag[8] = {0}; size_t pos[8] = {0};
	for (int k = 0; k < 8; ++k) {                                          /* candidates: planes of the STRONG anchors (:1352-1363) */
		const int qx = an[2 * (k + 1)], qy = an[2 * (k + 1) + 1];
		if (qx == -1 || qy == -1) continue;
		if (c->states[(size_t)qy * W + qx] != STRONG) continue;
		flag[k] = 1; pos[k] = (size_t)qy * W + qx;
		for (int v = 0; v < S; ++v) ca[k][v] = ncc_deform(c, px, py, v + 1, c->planes[pos[k]]);
	}
	uint8_t *vw = &c->vw[ctr * MAX_IMAGES]; memset(vw, 0, MAX_IMAGES);
	float prob[32] = {0};
	float thr = (float)(0.8 * expf((iter * iter) / (-90.0f)));
	for (int v = 0; v < S; ++v) {                                          /* view selection (:1365-1434) */
		float prior = 0.0f;
		for (int k = 1; k < NEIGHBOUR_NUM; ++k) { const int qx = an[2 * k], qy = an[2 * k + 1]; if (qx == -1 || qy == -1) continue;
			prior += isset(c->views[(size_t)qy * W + qx], v) ? 0.9f : 0.1f; }
		float count = 0, tmpw = 0; int bad = 0;
		for (int k = 0; k < 8; ++k) { if (ca[k][v] < thr) { tmpw += expf(ca[k][v] * ca[k][v] / (-0.18f)); count++; } if (ca[k][v] > 1.2f) bad++; }
		if (count > 2 && bad < 3) prob[v] = tmpw / count; else if (bad < 3) prob[v] = expf(thr * thr / (-0.32f));
		prob[v] *= prior;
	}
	{ float sum = 0; for (int v = 0; v < S; ++v) sum += prob[v]; float inv = 1.0f / sum, cum = 0;
	  for (int v = 0; v < S; ++v) { cum += prob[v] * inv; prob[v] = cum; } }
	Rng *r = &c->rng[ctr];
	for (int s = 0; s < 15; ++s) { float u = rng_uniform(r) - FLT_EPSILON; for (int v = 0; v < S; ++v) if (prob[v] > u) { vw[v]++; break; } }
	uint32_t sel = 0; float wn = 0; for (int v = 0; v < S; ++v) if (vw[v] > 0) { sel |= 1u << v; wn += vw[v]; }
	float fc[8];
	for (int k = 0; k < 8; ++k) {
		float acc = 0.0f;
		for (int v = 0; v < S; ++v) { if (vw[v] == 0) continue; float cost = ca[k][v];
			if (c->p.geom_consistency) cost += c->p.geom_factor * (flag[k] ? geom_cost(c, px, py, v + 1, c->planes[pos[k]]) : 3.0f);
			acc += vw[v] * cost; }
		fc[k] = acc / wn;
	}
	int mi = 0; { float m = fc[0]; for (int k = 1; k < 8; ++k) if (fc[k] <= m) { m = fc[k]; mi = k; } }
	f4 pl = c->planes[ctr];
	float cost_now = weak_cost(c, px, py, pl, vw, wn); const float cost_stored = cost_now;
	float depth_now = depth_from_plane(rc, pl, px, py);
	if (flag[mi]) { f4 cand = c->planes[pos[mi]]; float d = depth_from_plane(rc, cand, px, py);
		if (d >= c->p.depth_min && d <= c->p.depth_max && fc[mi] < cost_now) { depth_now = d; pl = cand; cost_now = fc[mi]; c->views[ctr] = sel; } }
	{   /* PlaneHypothesisRefinementWeak: only with a fit plane (:892-980) */
		const f4 fit = c->fit[ctr];
		if (!(fit.x == 0.0f && fit.y == 0.0f && fit.z == 0.0f)) {
			const float dmin = c->p.depth_min, dmax = c->p.depth_max;
			{ float tc = weak_cost(c, px, py, fit, vw, wn); float d = depth_from_plane(rc, fit, px, py);
			  if (d >= dmin && d <= dmax && tc < cost_now) { depth_now = d; pl = fit; cost_now = tc; } }
			float drand = rng_uniform(r) * (dmax - dmin) + dmin; f4 nrand = random_normal(rc, px, py, r, depth_now);
			float lo = (1 - 0.02f) * depth_now, hi = (1 + 0.02f) * depth_now, dpert = rng_uniform(r) * (hi - lo) + lo;
			f4 npert = perturbed_normal(rc, px, py, pl, r, (float)(0.02f * 3.14159265358979323846));
			float ds[5] = {drand, depth_now, drand, depth_now, dpert}; f4 ns[5] = {pl, nrand, nrand, npert, pl};
			for (int i = 0; i < 5; ++i) { f4 t = ns[i]; t.w = dist2origin(rc, px, py, ds[i], t);
				float tc = weak_cost(c, px, py, t, vw, wn); float d = depth_from_plane(rc, t, px, py);
				if (d >= dmin && d <= dmax && tc < cost_now) { depth_now = d; pl = t; cost_now = tc; } }
		}
	}
	f4 final_plane = c->planes[ctr];
	if (c->p.state == REFINE_INIT) { if ((double)cost_now < (double)cost_stored - 0.1) { final_plane = pl; c->planes[ctr] = pl; } }
	else { final_plane = pl; c->planes[ctr] = pl; }
	{   /* "update cost with old method" (:1499-1507) */
		float acc = 0.0f;
		for (int v = 0; v < S; ++v) if (vw[v] > 0) acc += vw[v] * ncc_old(c, px, py, v + 1, final_plane);
		c->costs[ctr] = acc / wn;
	}
}

/* APD::RunPatchMatch, APD.cu:2386-2495. Returns the number of stages executed; stage numbering as in
 * include/apd_b200.h. anchors_out [H*W*9*2] int16, nearest_out [H*W*2] int16, reliable_out [H*W], fit_out [H*W*4]
 * may be NULL. */
static int run_impl(Ctx *cp, const float *prior_planes, const uint32_t *prior_views, const uint8_t *prior_states, unsigned long long seed, int stage_end);
int apd_cpu_run_apd(int W, int H, int N, const float *images, const float *depths, const Camera *cams, const Params *params,
                    const float *prior_planes, const uint32_t *prior_views, const uint8_t *prior_states, unsigned long long seed,
                    int stage_end, float *planes, float *costs, uint32_t *views, uint8_t *states, uint8_t *view_weights, uint32_t *rng6,
                    int16_t *anchors_out, int16_t *nearest_out, uint8_t *reliable_out, float *fit_out) {
	Ctx c; memset(&c, 0, sizeof(c));
	c.W = W; c.H = H; c.N = N; c.images = images; c.depths = depths; c.cams = cams; c.p = *params; c.p.num_images = N;
	const size_t n = (size_t)W * H;
	c.planes = (f4 *)planes; c.costs = costs; c.views = views; c.states = states; c.vw = view_weights; c.rng = (Rng *)rng6;
	c.anchors = anchors_out ? anchors_out : (int16_t *)malloc(n * NEIGHBOUR_NUM * 2 * sizeof(int16_t));
	c.nearest = nearest_out ? nearest_out : (int16_t *)malloc(n * 2 * sizeof(int16_t));
	c.reliable = reliable_out ? reliable_out : (uint8_t *)malloc(n);
	c.fit = fit_out ? (f4 *)fit_out : (f4 *)malloc(n * sizeof(f4));
	memset(c.anchors, 0xff, n * NEIGHBOUR_NUM * 2 * sizeof(int16_t)); memset(c.nearest, 0xff, n * 2 * sizeof(int16_t));
	memset(c.reliable, 0, n); memset(c.fit, 0, n * sizeof(f4));
	const int rc = run_impl(&c, prior_planes, prior_views, prior_states, seed, stage_end);
	if (!anchors_out) free(c.anchors);
	if (!nearest_out) free(c.nearest);
	if (!reliable_out) free(c.reliable);
	if (!fit_out) free(c.fit);
	return rc;
}
int apd_cpu_run(int W, int H, int N, const float *images, const float *depths, const Camera *cams, const Params *params,
                const float *prior_planes, const uint32_t *prior_views, const uint8_t *prior_states, unsigned long long seed,
                int stage_end, float *planes, float *costs, uint32_t *views, uint8_t *states, uint8_t *view_weights, uint32_t *rng6) {
	return apd_cpu_run_apd(W, H, N, images, depths, cams, params, prior_planes, prior_views, prior_states, seed, stage_end,
	                       planes, costs, views, states, view_weights, rng6, NULL, NULL, NULL, NULL);
}
static int run_impl(Ctx *cp, const float *prior_planes, const uint32_t *prior_views, const uint8_t *prior_states, unsigned long long seed, int stage_end) {
	Ctx c = *cp;
	const int W = c.W, H = c.H; const Camera *cams = c.cams;
	float *planes = (float *)c.planes; float *costs = c.costs; uint32_t *views = c.views; uint8_t *states = c.states; uint8_t *view_weights = c.vw;
#if 0
	c.W = W; c.H = H; c.N = N; c.images = images; c.depths = depths; c.cams = cams; c.p = *params; c.p.num_images = N;
	const size_t n = (size_t)W * H;
	c.planes = (f4 *)planes; c.costs = costs; c.views = views; c.states = states; c.vw = view_weights; c.rng = (Rng *)rng6;
#endif
	const size_t n = (size_t)W * H;
	const int apd_on = c.p.use_APD != 0;
	memset(costs, 0, n * 4); memset(view_weights, 0, n * MAX_IMAGES);
	for (size_t i = 0; i < n; ++i) states[i] = (c.p.use_APD && prior_states) ? prior_states[i] : STRONG;    /* APD.cpp:513-548 */
	if (c.p.state != FIRST_INIT) { memcpy(planes, prior_planes, n * 16); memcpy(views, prior_views, n * 4); }
	else { memset(planes, 0, n * 16); memset(views, 0, n * 4); }
	const int nstages = 10 + 5 * c.p.max_iterations;
	if (stage_end < 0 || stage_end >= nstages) stage_end = nstages - 1;
	const int half_rows = 32 * ((H / 2 + 15) / 16);          /* rows reached by the half launches, APD.cu:2400-2403 */
	int stage = 0;
#define DONE() do { if (stage == stage_end) return stage + 1; ++stage; } while (0)
	init_rng(&c, seed); DONE();
	if (apd_on) {
#pragma omp parallel for schedule(dynamic, 4)
		for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) nearest_strong_pixel(&c, x, y);                    /* K2 */
	}
	DONE();
	if (apd_on) {
#pragma omp parallel for schedule(dynamic, 4)
		for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) gen_anchors_pixel(&c, x, y);                       /* K3 */
	}
	DONE();
	if (apd_on)                                                                                                   /* K4 NeigbourUpdate, APD.cu:1971-1987 */
		for (size_t i = 0; i < n; ++i) if (states[i] == WEAK && c.reliable[i] != 1) states[i] = UNKNOWN;
	DONE();
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) init_pixel(&c, x, y);
	DONE();
	for (int it = 0; it < c.p.max_iterations; ++it) {
		for (int color = 0; color < 2; ++color) {
#pragma omp parallel for schedule(dynamic, 4)
			for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
				if (((x + y) & 1) == color && y < half_rows && states[(size_t)y * W + x] != WEAK) strong_pixel(&c, x, y, it);
			DONE();
		}
		if (apd_on) {
#pragma omp parallel for schedule(dynamic, 4)
			for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) fit_plane_pixel(&c, x, y);                     /* K8 */
		}
		DONE();
		for (int color = 0; color < 2; ++color) {                                                                 /* K9 / K10 */
			if (apd_on) {
#pragma omp parallel for schedule(dynamic, 4)
				for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
					if (((x + y) & 1) == color && y < half_rows && states[(size_t)y * W + x] == WEAK) weak_pixel(&c, x, y, it);
			}
			DONE();
		}
	}
	{   /* GetDepthandNormal, APD.cu:1587-1602 */
		const Camera *rc = &cams[0];
		for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) { f4 *p = &c.planes[(size_t)y * W + x]; f4 o;
			float d = depth_from_plane(rc, *p, x, y);
			o.x = rc->R[0] * p->x + rc->R[3] * p->y + rc->R[6] * p->z; o.y = rc->R[1] * p->x + rc->R[4] * p->y + rc->R[7] * p->z;
			o.z = rc->R[2] * p->x + rc->R[5] * p->y + rc->R[8] * p->z; o.w = d; *p = o; }
	}
	DONE();
	for (int color = 0; color < 2; ++color) {
		for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
			if (((x + y) & 1) == color && y < half_rows && states[(size_t)y * W + x] != WEAK) filter_pixel(&c, x, y);
		DONE();
	}
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) classify_pixel(&c, x, y);
	DONE();
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) refine_pixel(&c, x, y);
	DONE();
#undef DONE
	return stage;
}

/* Time one strong-propagation colour pass over a pixel sub-rectangle (bench.py cpu_baseline). State must
 * have been prepared by apd_cpu_run(..., stage_end = 4). Returns the number of pixels processed. */
long apd_cpu_strong_pass(int W, int H, int N, const float *images, const Camera *cams, const Params *params,
                         float *planes, float *costs, uint32_t *views, uint8_t *states, uint8_t *view_weights, uint32_t *rng6,
                         int iter, int color, int x0, int y0, int x1, int y1) {
	Ctx c; memset(&c, 0, sizeof(c));
	c.W = W; c.H = H; c.N = N; c.images = images; c.cams = cams; c.p = *params; c.p.num_images = N;
	c.planes = (f4 *)planes; c.costs = costs; c.views = views; c.states = states; c.vw = view_weights; c.rng = (Rng *)rng6;
	long cnt = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : cnt)
	for (int y = y0; y < y1; ++y) for (int x = x0; x < x1; ++x)
		if (((x + y) & 1) == color && states[(size_t)y * W + x] != WEAK) { strong_pixel(&c, x, y, iter); cnt++; }
	return cnt;
}
