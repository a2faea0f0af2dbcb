/*
 * apd_io.h — the reference's on-disk formats without OpenCV / Boost: std-only C++ behind a C-ABI,
 * part of libapd_b200.so. Host code only; nothing here touches the GPU.
 *
 *   reference                                   this library
 *   ---------------------------------------    ----------------------------------------------
 *   ReadBinMat / WriteBinMat   APD.cpp:3-50     apd_io_read_mat_header / apd_io_read_mat / apd_io_write_mat
 *       (.dmb / .bin: int32 version = 1, rows, cols, OpenCV type code, then rows*cols elements, row-major)
 *   ReadCamera                 APD.cpp:52-92    apd_io_read_camera   (TAT & ETH flavour: depth_min interval depth_num depth_max)
 *   GenerateSampleList         main.cpp:6-49    apd_io_read_pairs    (pair.txt; sources with score <= 0 are dropped)
 *   ToFormatIndex              APD.cpp:350-354  apd_io_format_index  ("%08d")
 *
 * All functions return 0 or a negative apd_status (apd_b200.h).
 */
#ifndef APD_IO_H
#define APD_IO_H
#include "apd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* OpenCV type codes of the matrices the reference writes (CV_MAKETYPE(depth, channels)). */
enum { APD_IO_8UC1 = 0, APD_IO_32SC1 = 4, APD_IO_32FC1 = 5, APD_IO_32FC3 = 21 };

/* Bytes per element of a type code, 0 if the code is not one of the four above. */
size_t apd_io_elem_size(int type);

int apd_io_read_mat_header(const char *path, int *rows, int *cols, int *type);
/* Reads the payload into `data` (capacity_bytes must cover rows*cols*elem_size). */
int apd_io_read_mat(const char *path, void *data, size_t capacity_bytes, int *rows, int *cols, int *type);
int apd_io_write_mat(const char *path, const void *data, int rows, int cols, int type);

/* Fills R, t, K, c (= -R^T t, accumulated in double as the reference does), depth_min, depth_max; width/height = 0. */
int apd_io_read_camera(const char *path, apd_camera *cam);

/* pair.txt. ref_ids[max_problems], n_src[max_problems], src_ids[max_problems * max_src] (row p holds the sources of
 * problem p). *n_problems receives the number of problems in the file; returns APD_E_LIMIT if a capacity is too small
 * (call with max_problems = 0 to query the count only). */
int apd_io_read_pairs(const char *path, int *n_problems, int *ref_ids, int *n_src, int *src_ids, int max_problems, int max_src);

/* "%08d" into out[9]. */
void apd_io_format_index(int index, char *out);

#ifdef __cplusplus
}
#endif
#endif /* APD_IO_H */
