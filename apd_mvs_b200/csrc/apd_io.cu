// The reference's file formats in std-only C++ (include/apd_io.h): .dmb/.bin matrices (APD.cpp:3-50), _cam.txt
// (APD.cpp:52-92) and pair.txt (main.cpp:6-49). Host code only; lives in a .cu file so the one Makefile rule builds it.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/apd_io.h"

extern "C" size_t apd_io_elem_size(int type) {
	switch (type) {
		case APD_IO_8UC1: return 1;
		case APD_IO_32SC1: return 4;
		case APD_IO_32FC1: return 4;
		case APD_IO_32FC3: return 12;
		default: return 0;
	}
}

static int read_header(std::ifstream &in, int *rows, int *cols, int *type) {
	int v[4];
	in.read((char *)v, sizeof(v));
	if (!in || v[0] != 1) return APD_E_ARG;                   // "Version error", APD.cpp:17-21
	if (v[1] < 0 || v[2] < 0 || apd_io_elem_size(v[3]) == 0) return APD_E_ARG;
	*rows = v[1]; *cols = v[2]; *type = v[3];
	return APD_OK;
}

extern "C" int apd_io_read_mat_header(const char *path, int *rows, int *cols, int *type) {
	if (!path || !rows || !cols || !type) return APD_E_ARG;
	std::ifstream in(path, std::ios_base::binary);
	if (!in) return APD_E_STATE;
	return read_header(in, rows, cols, type);
}

extern "C" int apd_io_read_mat(const char *path, void *data, size_t capacity, int *rows, int *cols, int *type) {
	if (!path || !data || !rows || !cols || !type) return APD_E_ARG;
	std::ifstream in(path, std::ios_base::binary);
	if (!in) return APD_E_STATE;
	int rc = read_header(in, rows, cols, type);
	if (rc != APD_OK) return rc;
	const size_t bytes = (size_t)*rows * *cols * apd_io_elem_size(*type);
	if (bytes > capacity) return APD_E_LIMIT;
	in.read((char *)data, (std::streamsize)bytes);
	return in ? APD_OK : APD_E_ARG;
}

extern "C" int apd_io_write_mat(const char *path, const void *data, int rows, int cols, int type) {
	if (!path || !data || rows < 0 || cols < 0 || apd_io_elem_size(type) == 0) return APD_E_ARG;
	std::ofstream out(path, std::ios_base::binary);
	if (!out) return APD_E_STATE;
	const int v[4] = {1, rows, cols, type};
	out.write((const char *)v, sizeof(v));
	out.write((const char *)data, (std::streamsize)((size_t)rows * cols * apd_io_elem_size(type)));
	return out ? APD_OK : APD_E_STATE;
}

extern "C" int apd_io_read_camera(const char *path, apd_camera *cam) {
	if (!path || !cam) return APD_E_ARG;
	std::ifstream in(path);
	if (!in) return APD_E_STATE;
	memset(cam, 0, sizeof(*cam));
	std::string word;
	in >> word;                                                                     // "extrinsic"
	for (int i = 0; i < 3; ++i) in >> cam->R[3 * i + 0] >> cam->R[3 * i + 1] >> cam->R[3 * i + 2] >> cam->t[i];
	float last_row[4];
	in >> last_row[0] >> last_row[1] >> last_row[2] >> last_row[3];
	in >> word;                                                                     // "intrinsic"
	for (int i = 0; i < 3; ++i) in >> cam->K[3 * i + 0] >> cam->K[3 * i + 1] >> cam->K[3 * i + 2];
	for (int j = 0; j < 3; ++j)                                                     // camera centre, APD.cpp:73-78
		cam->c[j] = -float(double(cam->R[0 + j]) * double(cam->t[0]) + double(cam->R[3 + j]) * double(cam->t[1]) + double(cam->R[6 + j]) * double(cam->t[2]));
	float interval, depth_num;
	in >> cam->depth_min >> interval >> depth_num >> cam->depth_max;                // TAT & ETH flavour, APD.cpp:80-84
	return in.fail() ? APD_E_ARG : APD_OK;
}

extern "C" int apd_io_read_pairs(const char *path, int *n_problems, int *ref_ids, int *n_src, int *src_ids, int max_problems, int max_src) {
	if (!path || !n_problems) return APD_E_ARG;
	std::ifstream file(path);
	if (!file) return APD_E_STATE;
	std::string line;
	std::istringstream iss;
	int num_images = 0;
	std::getline(file, line); iss.str(line); iss >> num_images;
	if (iss.fail() || num_images < 0) return APD_E_ARG;
	*n_problems = num_images;
	if (max_problems == 0) return APD_OK;
	if (num_images > max_problems || !ref_ids || !n_src || !src_ids) return APD_E_LIMIT;
	for (int i = 0; i < num_images; ++i) {
		int ref = 0, count = 0;
		iss.clear(); std::getline(file, line); iss.str(line); iss >> ref;
		iss.clear(); std::getline(file, line); iss.str(line); iss >> count;
		if (iss.fail()) return APD_E_ARG;
		ref_ids[i] = ref;
		int kept = 0;
		for (int j = 0; j < count; ++j) {
			int id; float score;
			iss >> id >> score;
			if (iss.fail()) return APD_E_ARG;
			if (score <= 0.0f) continue;                                            // main.cpp:41-43
			if (kept >= max_src) return APD_E_LIMIT;
			src_ids[(size_t)i * max_src + kept++] = id;
		}
		n_src[i] = kept;
	}
	return APD_OK;
}

extern "C" void apd_io_format_index(int index, char *out) { if (out) snprintf(out, 9, "%08d", index); }
