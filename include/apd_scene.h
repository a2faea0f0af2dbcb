/*
 * apd_scene.h — C-ABI of the pass scheduler with a persistent per-scene device cache.
 *
 * The reference drives the hot path from `main()` (main.cpp:140-217): for every round (image pyramid level)
 * one pass A and three passes B over all problems, and for each (problem, pass) a fresh `APD` object that
 * re-decodes N JPEGs, re-allocates ~25 device buffers + 2N cudaArrays and exchanges every result with the next
 * pass through files (ProcessProblem, main.cpp:91-138; APD::InuputInitialization, APD.cpp:399-583).
 *
 * This layer keeps the whole scene on the GPU instead: full-resolution images and cameras of all views, the
 * round's down-scaled images, and each view's latest depth / normal / pixel-state / selected-view maps stay in
 * HBM across the 4*round_num passes; priors are re-sampled on the device; the PatchMatch engine handle
 * (apd_b200.h) is allocated once per round. Problems are processed in pair-list order on one stream, so a problem
 * sees the depth maps that earlier problems wrote in the same pass exactly as in the reference (main.cpp:169-213).
 *
 *   reference                                             this library
 *   --------------------------------------------------   ------------------------------------------------
 *   GenerateSampleList (pair.txt)        main.cpp:6-49    apd_scene_add_problem(ref, srcs)
 *   cv::imread + ReadCamera              APD.cpp:408-450  apd_scene_set_view(view, image, pitch, camera)
 *   ComputeRoundNum                      main.cpp:72-88   apd_scene_num_rounds
 *   for i, pass: for problem: ProcessProblem  :168-217    apd_scene_run  /  apd_scene_run_pass(round, pass)
 *   cv::resize(INTER_LINEAR), camera K scaling APD.cpp:464-489   device kernel, once per view and round
 *   RescaleMatToTargetSize<T>            APD.cpp:752-774  device kernels (nearest, with the reference's
 *                                                         row/column scale swap)
 *   depth range test + UNKNOWN marking   main.cpp:105-115 device kernel after every run
 *   depths.dmb / normals.dmb / weak.bin / selected_views.bin   apd_scene_get_depth / _normal / _states / _views
 *
 * All functions return 0 or a negative apd_status (apd_b200.h); apd_scene_last_error() gives the message.
 */
#ifndef APD_SCENE_H
#define APD_SCENE_H
#include "apd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct apd_scene *apd_scene_handle;

/* One scene = n_views images of identical size (CheckImages, main.cpp:51-70). `seed` replaces the reference's
 * clock64() seeds: run (round, pass, problem k) uses seed + (round*4 + pass)*65536 + k. */
int apd_scene_create(apd_scene_handle *out, int device, int n_views, int width, int height, uint64_t seed);
void apd_scene_destroy(apd_scene_handle s);
const char *apd_scene_last_error(apd_scene_handle s);

/* Full-resolution float32 grey image (0..255, what cv::imread(GRAYSCALE) + convertTo(CV_32F) yields,
 * APD.cpp:410-414) and camera (ReadCamera, APD.cpp:94-140) of one view. `image` may be host or device memory. */
int apd_scene_set_view(apd_scene_handle s, int view, const float *image, size_t pitch_bytes, const apd_camera *cam);

/* One entry of pair.txt: reference view and its source views (those with score > 0), main.cpp:19-47.
 * Problems run in the order they are added. */
int apd_scene_add_problem(apd_scene_handle s, int ref_view, const int *src_views, int n_src);

/* ComputeRoundNum, main.cpp:72-88: the image is halved until its larger side is <= 1000. That constant is a literal in the
 * reference (main.cpp:81); apd_scene_set_round_limit replaces it for this scene (before the first pass), e.g. 1920 to run
 * the single-round schedule BASELINE configs[3] describes at 1920x1080. Default 1000. */
int apd_scene_set_round_limit(apd_scene_handle s, int max_size);
int apd_scene_num_rounds(apd_scene_handle s);
/* Image size of a round: scale_size = 2^(rounds-1-round), size = round(full * 1/scale_size) (APD.cpp:466-468). */
int apd_scene_round_size(apd_scene_handle s, int round, int *width, int *height);
/* The parameters main.cpp:171-211 gives pass `pass` (0 = A, 1..3 = B) of round `round`. */
int apd_scene_pass_params(apd_scene_handle s, int round, int pass, apd_params *out);

/* Whole schedule, main.cpp:168-217. */
int apd_scene_run(apd_scene_handle s);
/* One pass over all problems (passes must be run in schedule order; exposed for tests and external schedulers). */
int apd_scene_run_pass(apd_scene_handle s, int round, int pass);
/* One (problem, pass): ProcessProblem, main.cpp:91-138. */
int apd_scene_run_problem(apd_scene_handle s, int round, int pass, int problem);

/* Latest results of a view = contents of its depths.dmb / normals.dmb / weak.bin / selected_views.bin
 * (main.cpp:117-124), at the size of the last pass that processed it (apd_scene_result_size). */
int apd_scene_result_size(apd_scene_handle s, int view, int *width, int *height);
int apd_scene_get_depth(apd_scene_handle s, int view, float *depth);
int apd_scene_get_normal(apd_scene_handle s, int view, float *normal_xyz);
int apd_scene_get_states(apd_scene_handle s, int view, uint8_t *states);
int apd_scene_get_views(apd_scene_handle s, int view, uint32_t *selected_views);
/* Checkpoint / resume: load a view's four maps as they were read back (or read from its .dmb / .bin files), so that a new
 * scene continues the schedule where a previous process stopped - the role the result files play between the reference's
 * passes (APD.cpp:492-581). Pointers may be host or device memory. */
int apd_scene_set_result(apd_scene_handle s, int view, int width, int height, const float *depth, const float *normal_xyz,
                         const uint8_t *states, const uint32_t *selected_views);
/* Multi-GPU hand-over: device pointers of a view's result buffers (each sized for the full
 * resolution; planes = float4 per pixel: normal xyz + depth w) so that a collective can write a peer's result in
 * place, and the declaration that they now hold a width x height result. Ranks shard the problems; between passes
 * only the depth maps travel (the geometric term reads the source views' depth, APD.cpp:492-510). */
int apd_scene_result_device(apd_scene_handle s, int view, void **planes, void **depth, void **states, void **selected_views);
int apd_scene_mark_result(apd_scene_handle s, int view, int width, int height);
/* The round's down-scaled image of a view (what cv::resize produced in the reference), for tests. */
int apd_scene_get_scaled_image(apd_scene_handle s, int round, int view, float *image);

/* GPU time of the PatchMatch launches of the last apd_scene_run* call (sum of the engine's stage events) and the
 * wall time of the call, both in ms; number of kernels launched. */
int apd_scene_get_timing(apd_scene_handle s, double *patchmatch_ms, double *wall_ms, long long *launches);

#ifdef __cplusplus
}
#endif
#endif /* APD_SCENE_H */
