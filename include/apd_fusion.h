/*
 * apd_fusion.h — C-ABI of the GPU depth-map fusion: the reference's RunFusion (ETH variant,
 * APD.cpp:826-977) with Get3DPointonWorld / ProjectCamera / GetAngle (APD.cpp:776-824), one thread per reference pixel.
 *
 * The reference walks the views in problem order and, inside a view, the pixels in raster order; a pixel that is
 * accepted marks the source pixels it used (masks[src]) so that later pixels - of this view and of later views - skip
 * them. That raster-order dependency is kept EXACTLY: candidates are computed in parallel, then pixels are decided in
 * rounds, a pixel being decidable once no earlier undecided pixel of the view contends for any of its source pixels
 * (atomicMin claims), so every pixel sees exactly the marks it sees in the sequential loop. Accepted points are
 * compacted in raster order: the point list has the reference's order.
 *
 *   reference                                        this library
 *   ---------------------------------------------   ------------------------------------------------
 *   per-view vectors images/cameras/depths/...       apd_fusion_create + apd_fusion_set_view (host or device pointers,
 *     APD.cpp:836-893                                  e.g. straight from apd_scene_result_device)
 *   problems[i].src_image_ids                        apd_fusion_add_problem
 *   the fusion loops APD.cpp:897-974                 apd_fusion_run
 *   std::vector<PointList> PointCloud                apd_fusion_num_points / apd_fusion_get_points
 *   ExportPointCloud APD.cpp:214-254                 apd_fusion_write_ply
 *
 * Images must already have the size of the depth maps (true after the full-resolution round; the reference's
 * RescaleImageAndCamera, APD.cpp:729-750, is a no-op then). All functions return 0 or a negative apd_status.
 */
#ifndef APD_FUSION_H
#define APD_FUSION_H
#include "apd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct apd_fusion *apd_fusion_handle;

int apd_fusion_create(apd_fusion_handle *out, int device, int n_views, int width, int height);
void apd_fusion_destroy(apd_fusion_handle f);
const char *apd_fusion_last_error(apd_fusion_handle f);

/* One view: BGR image (3 bytes per pixel, cv::IMREAD_COLOR layout), camera as read by ReadCamera (only K, R, t are
 * used), depth map, normal map (3 floats per pixel), pixel states (weak.bin), optional block mask (NULL = none;
 * pixels with block < 128 are skipped, APD.cpp:905-907). Pointers may be host or device memory. */
int apd_fusion_set_view(apd_fusion_handle f, int view, const uint8_t *bgr, const apd_camera *cam, const float *depth,
                        const float *normal_xyz, const uint8_t *states, const uint8_t *block);
/* Same with the normal taken from float4 planes (xyz = normal), the layout apd_scene_result_device hands out. */
int apd_fusion_set_view_planes(apd_fusion_handle f, int view, const uint8_t *bgr, const apd_camera *cam, const float *depth,
                               const float *planes_xyzw, const uint8_t *states, const uint8_t *block);
int apd_fusion_add_problem(apd_fusion_handle f, int ref_view, const int *src_views, int n_src);

/* Fuses all problems in order (RunFusion, the ETH variant: the only one main.cpp calls, main.cpp:219). */
int apd_fusion_run(apd_fusion_handle f);
/* The two Tanks-and-Temples variants the reference defines next to it: RunFusion_TAT_Intermediate (APD.cpp:979-1147:
 * at least k of the sources within k x (0.25 px, 1/3500 relative depth, 3 deg x k + 4 deg), k = 2.., colour averaged over
 * the agreeing sources) and RunFusion_TAT_advanced (APD.cpp:1149-1296: k x (0.25 px, 1/3000), reference colour). Both keep
 * the reference's carry-over of a source's last measurement to later pixels (`diff` is not reset per pixel) and mark only the
 * reference view's own pixels. Pixel states are not used by these variants. */
#define APD_FUSION_TAT_INTERMEDIATE 1
#define APD_FUSION_TAT_ADVANCED 2
int apd_fusion_run_tat(apd_fusion_handle f, int variant);
long long apd_fusion_num_points(apd_fusion_handle f);
/* xyz: 3 floats per point; color: 3 floats per point (blue, green, red averages, PointList.color). Either may be NULL. */
int apd_fusion_get_points(apd_fusion_handle f, float *xyz, float *color_bgr);
/* ExportPointCloud, APD.cpp:214-254 (binary little-endian PLY, colours truncated to uchar). */
int apd_fusion_write_ply(apd_fusion_handle f, const char *path);
/* GPU time of the last run in ms, and the number of decision rounds it took (max over views). */
int apd_fusion_get_timing(apd_fusion_handle f, double *gpu_ms, int *max_rounds);

#ifdef __cplusplus
}
#endif
#endif /* APD_FUSION_H */
