// TEST INFRASTRUCTURE: C-ABI around the UNMODIFIED reference APD.cpp (RunFusion = the ETH fusion, APD.cpp:826-977),
// compiled from /root/reference where it lies against oracle/shim_host (no OpenCV / Boost in this image). Used by the
// tests and tools as the checker and CPU baseline of the GPU fusion (include/apd_fusion.h); never shipped or linked by
// the product. Build: oracle/Makefile target `fusion`.
#include APD_REF_CPP
#include <cstring>

extern "C" {
// Runs RunFusion(dense_folder, problems): reads images/%08d.jpg (raw container, see the shim), cams/%08d_cam.txt and
// APD/%08d/{depths.dmb, normals.dmb, weak.bin}; writes APD/APD.ply. Returns 0.
int apdfusion_ref_run(const char *dense_folder, int n_problems, const int *ref_ids, const int *n_src, const int *src_ids, int max_src) {
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
	RunFusion(dense, problems);
	std::cout.rdbuf(old);
	return 0;
}
}
