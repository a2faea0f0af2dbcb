/*
 * apd_b200.h — C-ABI of the B200-native PatchMatch engine (libapd_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of whoiszzj/APD-MVS: everything that
 * happens between `APD::CudaSpaceInitialization()` and the end of `APD::RunPatchMatch()`
 * (reference: APD.h:67-145, APD.cpp:585-727, APD.cu:2386-2495). Plain pointers and sizes
 * only; no C++ / torch / OpenCV types. All functions return 0 on success and a negative
 * apd_status on failure (the reference prints and exit()s instead, APD.cpp:315-323); the
 * message is available from apd_last_error().
 *
 * Calling sequence that replaces the reference's per-(ref view, pass) object lifetime
 * (main.cpp:95-124):
 *
 *   reference                                   this library
 *   ---------------------------------------    -------------------------------------------
 *   APD apd(problem)              APD.cpp:356   apd_create(&h, device, W, H, N, &params, seed)
 *   InuputInitialization()        APD.cpp:399   (host side keeps reading the files; arrays
 *                                                are then handed over with the setters)
 *   CudaSpaceInitialization()     APD.cpp:585   apd_set_cameras / apd_set_images /
 *   SetDataPassHelperInCuda()     APD.cpp:673   apd_set_depths / apd_set_priors
 *   RunPatchMatch()               APD.cu:2386   apd_run(h)
 *   GetPlaneHypothesis(r,c) ...   APD.cpp:701   apd_get_planes / apd_get_states / apd_get_views
 *   ~APD()                        APD.cpp:361   apd_destroy(h)
 *
 * A handle may be re-run: apd_run() always starts from the priors last set (or from the
 * random initialisation when state == APD_FIRST_INIT), so repeated runs are identical.
 */
#ifndef APD_B200_H
#define APD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APD_MAX_IMAGES 32       /* main.h:37  (selected-view bitmask is 32 bit)            */
#define APD_NEIGHBOUR_NUM 9     /* main.h:38  (pixel itself + 8 deformable anchors)        */
#define APD_MAX_SEARCH_RADIUS 4096 /* main.h:39                                            */

typedef enum { APD_OK = 0, APD_E_ARG = -1, APD_E_CUDA = -2, APD_E_STATE = -3, APD_E_LIMIT = -4 } apd_status;

/* RunState, main.h:63-67 */
enum { APD_FIRST_INIT = 0, APD_REFINE_INIT = 1, APD_REFINE_ITER = 2 };
/* PixelState, main.h:69-73 */
enum { APD_WEAK = 0, APD_STRONG = 1, APD_UNKNOWN = 2 };

/* Binary-compatible with `struct Camera`, main.h:47-56 (112 bytes). R, K row-major. */
typedef struct apd_camera {
	float K[9];
	float R[9];
	float t[3];
	float c[3];
	int height;
	int width;
	float depth_min;
	float depth_max;
} apd_camera;

/* Binary-compatible with `struct PatchMatchParams`, main.h:75-94 (72 bytes). */
typedef struct apd_params {
	int max_iterations;            /* 3 */
	int num_images;                /* overwritten by apd_create */
	float sigma_spatial;           /* dead in the reference */
	float sigma_color;             /* dead in the reference */
	int top_k;                     /* 4 */
	float depth_min;               /* caller passes cam0.depth_min*0.6f (APD.cpp:454) */
	float depth_max;               /* caller passes cam0.depth_max*1.2f (APD.cpp:455) */
	unsigned char geom_consistency; unsigned char pad0_[3];
	int strong_radius;             /* 5 */
	int strong_increment;          /* 2 */
	int weak_radius;               /* 5 */
	int weak_increment;            /* 5 */
	unsigned char use_APD; unsigned char pad1_[3];
	int weak_peak_radius;          /* 2 */
	int rotate_time;               /* 4 */
	float ransac_threshold;        /* 0.005 */
	float geom_factor;             /* 0.2 */
	int state;                     /* APD_FIRST_INIT / APD_REFINE_INIT / APD_REFINE_ITER */
} apd_params;

/* Fills *p with the defaults of main.h:75-94. */
void apd_default_params(apd_params *p);

typedef struct apd_engine *apd_handle;

/* Allocates every device buffer once (replaces the ~25 cudaMalloc + N cudaMallocArray of
 * APD.cpp:585-671). `seed` replaces the reference's clock64() curand seed (APD.cu:803). */
int apd_create(apd_handle *out, int device, int width, int height, int num_images,
               const apd_params *params, uint64_t seed);
void apd_destroy(apd_handle h);
const char *apd_last_error(apd_handle h);

/* Change run-time parameters between runs of the same handle (main.cpp:171-211 re-tunes
 * state / geom_consistency / weak_peak_radius / ransac_threshold / rotate_time per pass). */
int apd_set_params(apd_handle h, const apd_params *params);
int apd_set_seed(apd_handle h, uint64_t seed);
/* Use fewer images than the handle was created for (problems of one scene have different numbers of source
 * views, main.cpp:36-46); images, cameras and depths must be set again afterwards. */
int apd_set_num_images(apd_handle h, int num_images);
/* Handle reuse across problems (the reference constructs and destroys one APD object per (view, pass), main.cpp:91-138;
 * a caller that keeps the handle instead calls this between problems): forgets every input that was set, so that the
 * preconditions of apd_run (cameras, images, depth maps, priors: APD.cpp:492-581) are checked against the NEW problem.
 * Device buffers, streams, texture objects and tensor maps are kept. */
int apd_reset_inputs(apd_handle h);
/* Image count the handle was created for (upper bound of apd_set_num_images). */
int apd_get_capacity(apd_handle h);

/* Upload mode of apd_set_cameras / apd_set_images / apd_set_depths / apd_set_priors with HOST pointers.
 *   0 (default): every call returns when its copy has completed; the host buffer may be reused at once.
 *   1 (asynchronous): the calls only enqueue their copies on the handle's copy stream and return; apd_run makes each
 *     launch wait for exactly the inputs it reads (cameras + priors before the first launch, images + depth maps before
 *     RandomInitialization), so the uploads of the image and depth stacks overlap InitRandomStates / FindNearestStrongPoint /
 *     GenNeighbours. The caller keeps the host buffers valid and unchanged until apd_run / apd_run_until returns (pinned
 *     host memory for a real overlap). The *_device variants are not affected. The reference uploads synchronously
 *     inside CudaSpaceInitialization (APD.cpp:585-671). */
int apd_set_upload_mode(apd_handle h, int asynchronous);

/* cams[num_images]; index 0 is the reference view (APD.cpp:633-634). */
int apd_set_cameras(apd_handle h, const apd_camera *cams);
/* images[num_images]: pointers to float32 grey images (0..255), all width x height, rows `pitch_bytes` apart
 * (APD.cpp:588-606). The pointers of apd_set_images / apd_set_depths / apd_set_priors may be host OR device
 * memory of the handle's device (unified addressing decides). */
int apd_set_images(apd_handle h, const float *const *images, size_t pitch_bytes);
/* Same, from ONE device buffer holding the num_images images back to back
 * (image i at dev_stack + i*image_stride_bytes). Used when the image stack already lives
 * in HBM (e.g. after the NCCL broadcast at multi-GPU setup). */
int apd_set_images_device(apd_handle h, const float *dev_stack, size_t pitch_bytes, size_t image_stride_bytes);
/* depths[num_images]: per-view depth maps for the geometric-consistency term (APD.cpp:608-630). */
int apd_set_depths(apd_handle h, const float *const *depths, size_t pitch_bytes);
int apd_set_depths_device(apd_handle h, const float *dev_stack, size_t pitch_bytes, size_t image_stride_bytes);
/* Priors of a refinement pass (APD.cpp:513-581): planes = float4 (world normal xyz, depth w),
 * views = selected-view bitmasks, states = PixelState per pixel. planes/views may be NULL
 * when state == APD_FIRST_INIT; states may be NULL when use_APD == 0 (all STRONG). */
int apd_set_priors(apd_handle h, const float *planes_xyzw, const uint32_t *views, const uint8_t *states);

/* The hot path: APD::RunPatchMatch(), APD.cu:2386-2495. Asynchronous launches on the
 * handle's stream, one synchronisation at the end. */
int apd_run(apd_handle h);
/* Test hook: run only the launches [0, stage_end] of the schedule (stage numbering = the
 * order of the reference's 25 launches: 0 InitRandomStates, 1 FindNearestStrongPoint,
 * 2 GenNeighbours, 3 NeigbourUpdate, 4 RandomInitialization, then per iteration
 * {strong black, strong red, fit plane, weak black, weak red}, then GetDepthandNormal,
 * filter black, filter red, DepthToWeak, LocalRefine). stage_end < 0 means all. */
int apd_run_until(apd_handle h, int stage_end);
int apd_num_stages(apd_handle h);

/* Outputs (APD.cu:2490-2492). planes: W*H*4 floats (world normal, depth) after a full run. */
int apd_get_planes(apd_handle h, float *planes_xyzw);
int apd_get_states(apd_handle h, uint8_t *states);
int apd_get_views(apd_handle h, uint32_t *views);
/* Extras the reference keeps on the device: */
int apd_get_costs(apd_handle h, float *costs);
int apd_get_view_weights(apd_handle h, uint8_t *weights32_per_pixel);   /* W*H*32, as view_weight_cuda */
int apd_get_rng(apd_handle h, uint32_t *v5_d_per_pixel);                /* W*H*6: XORWOW v[0..4], d   */
/* Deformable-anchor state: anchors = W*H*9 short2 (x,y), (-1,-1) when absent; valid for WEAK pixels. */
int apd_get_anchors(apd_handle h, int16_t *anchors_xy, int16_t *nearest_strong_xy, uint8_t *reliable, float *fit_planes);

/* Per-stage GPU time of the last run in ms (cudaEvents on the handle's stream). Returns the
 * number of stages; fills at most `capacity` entries. */
int apd_get_stage_ms(apd_handle h, float *ms, int capacity);
/* Number of kernels launched by the last apd_run / apd_run_until. */
int apd_get_launch_count(apd_handle h);
/* Raw stream (cudaStream_t) so a host framework can order its own work / record events. */
void *apd_get_stream(apd_handle h);

#ifdef __cplusplus
}
#endif
#endif /* APD_B200_H */
