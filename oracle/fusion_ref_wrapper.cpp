// TEST INFRASTRUCTURE: C-ABI around the UNMODIFIED reference APD.cpp (RunFusion = the ETH fusion, APD.cpp:826-977),
// compiled from /root/reference where it lies against oracle/shim_host (no OpenCV / Boost in this image). Used by the
// tests and tools as the checker and CPU baseline of the GPU fusion (include/apd_fusion.h); never shipped or linked by
// the product. Build: oracle/Makefile target `fusion`.
#include APD_REF_CPP
#include <cstring>

extern "C" {
// Runs RunFusion(dense_folder, problems): reads images/%08d.jpg (raw container, see the shim), cams/%08d_cam.txt and
// APD/%08d/{depths.dmb, normals.dmb, weak.bin}; writes APD/APD.ply. Returns 0.
static int run_variant(const char *dense_folder, int n_problems, const int *ref_ids, const int *n_src, const int *src_ids, int max_src, int variant);
int apdfusion_ref_run(const char *dense_folder, int n_problems, const int *ref_ids, const int *n_src, const int *src_ids, int max_src) {
	return run_variant(dense_folder, n_problems, ref_ids, n_src, src_ids, max_src, 0);
}
// variant 0 = RunFusion (ETH), 1 = RunFusion_TAT_Intermediate, 2 = RunFusion_TAT_advanced (APD.cpp:979-1296)
int apdfusion_ref_run_variant(const char *dense_folder, int n_problems, const int *ref_ids, const int *n_src, const int *src_ids, int max_src, int variant) {
	return run_variant(dense_folder, n_problems, ref_ids, n_src, src_ids, max_src, variant);
}
static int run_variant(const char *dense_folder, int n_problems, const int *ref_ids, const int *n_src, const int *src_ids, int max_src, int variant) {
	std::vector<Problem> problems;
	const path dense(dense_folder);
	for (int i = 0; i < n_problems; ++i) {
		Problem p;
		p.index = i; p.ref_image_id = ref_ids[i];
		for (int j = 0; j < n_src[i]; ++j) p.src_image_ids.push_back(src_ids[(size_t)i * max_src + j]);
		p.dense_folder = dense;
		p.result_folder = dense / path("APD") / path(ToFormatIndex(p.ref_image_id));      // main.cpp:29
		problems.push_back(p);
	}
	std::streambuf *old = std::cout.rdbuf(nullptr);      // the reference prints a line per image
	if (variant == 1) RunFusion_TAT_Intermediate(dense, problems);
	else if (variant == 2) RunFusion_TAT_advanced(dense, problems);
	else RunFusion(dense, problems);
	std::cout.rdbuf(old);
	return 0;
}

// The reference's own matrix file I/O (ReadBinMat / WriteBinMat, APD.cpp:3-49) and camera parser (ReadCamera, APD.cpp:51-92)
// behind plain pointers, so that tests can pin include/apd_io.h against files the reference code itself wrote and read.
int apdref_write_bin_mat(const char *file, const void *data, int rows, int cols, int cv_type) {
	cv::Mat m(rows, cols, cv_type);
	memcpy(m.data, data, (size_t)m.step * rows);
	return WriteBinMat(path(file), m) ? 0 : -1;
}
int apdref_read_bin_mat(const char *file, void *out, size_t capacity, int *rows, int *cols, int *cv_type) {
	cv::Mat m;
	if (!ReadBinMat(path(file), m)) return -1;
	*rows = m.rows; *cols = m.cols; *cv_type = m.type();
	const size_t bytes = (size_t)m.step * m.rows;
	if (out) { if (bytes > capacity) return -2; memcpy(out, m.data, bytes); }
	return 0;
}
int apdref_read_camera(const char *file, void *camera_112_bytes) {
	Camera c;
	ReadCamera(path(file), c);
	memcpy(camera_112_bytes, &c, sizeof(Camera));
	return 0;
}
}
