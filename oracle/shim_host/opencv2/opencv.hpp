// Host-side stand-in for the part of OpenCV that the UNMODIFIED reference APD.cpp uses, so that its CPU functions
// (RunFusion, ReadBinMat, ReadCamera, ExportPointCloud, RescaleMatToTargetSize ...) can be compiled and run as the
// oracle in an image without the OpenCV C++ SDK. Test infrastructure only (oracle/_ref). Semantics kept: cv::Mat is a
// reference-counted row-major matrix (copy = shallow, clone = deep), at<T>(row, col), OpenCV's type codes.
// imread reads the raw container written by the tests (see apdraw_write in tests/fusion_tools.py), not JPEG;
// resize / imwrite exist only so that the file compiles (RunFusion never reaches them when image and depth sizes agree).
#ifndef APD_ORACLE_SHIM_HOST_OPENCV_HPP
#define APD_ORACLE_SHIM_HOST_OPENCV_HPP
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { INTER_LINEAR = 1 };
template <typename T> struct Size_ {
	T width, height;
	Size_() : width(0), height(0) {}
	Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size2i;
typedef Size2i Size;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
template <typename T, int N> struct Vec {
	T val[N];
	Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
	Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
	T &operator[](int i) { return val[i]; }
	const T &operator[](int i) const { return val[i]; }
};
template <typename T, int N> inline Vec<T, N> operator/(const Vec<T, N> &a, float d) { Vec<T, N> o; for (int i = 0; i < N; ++i) o.val[i] = (T)(a.val[i] / d); return o; }
typedef Vec<float, 3> Vec3f;
typedef Vec<uchar, 3> Vec3b;

inline size_t elem_size_of(int type) { const int depth = type & 7, cn = (type >> 3) + 1; const size_t d = depth == 0 ? 1 : 4; return d * cn; }

struct MatStep {                                              // cv::MatStep: converts to size_t, step[0] = row pitch
	size_t v;
	MatStep(size_t s = 0) : v(s) {}
	operator size_t() const { return v; }
	size_t operator[](int) const { return v; }
};

class Mat {
public:
	int rows, cols;
	uchar *data;
	MatStep step;
	int type_;
	std::shared_ptr<std::vector<uchar>> buf_;
	Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
	Mat(int r, int c, int type) { create(r, c, type); }
	Mat(Size s, int type) { create(s.height, s.width, type); }
	Mat(Size s, int type, const Scalar &v) { create(s.height, s.width, type); fill(v); }
	void create(int r, int c, int type) {
		rows = r; cols = c; type_ = type; step = (size_t)c * elem_size_of(type);
		buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step);
		data = buf_->data();
	}
	void fill(const Scalar &v) {
		const int depth = type_ & 7, cn = (type_ >> 3) + 1;
		for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) for (int k = 0; k < cn; ++k) {
			uchar *p = data + (size_t)r * step + ((size_t)c * cn + k) * (depth == 0 ? 1 : 4);
			if (depth == 0) *p = (uchar)v.val[k]; else if (depth == 4) *(int *)p = (int)v.val[k]; else *(float *)p = (float)v.val[k];
		}
	}
	static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }      // vector storage is value-initialised
	static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
	template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(data + (size_t)r * step)[c]; }
	template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(data + (size_t)r * step)[c]; }
	template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
	template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }
	Mat clone() const { Mat m; if (data) { m.create(rows, cols, type_); memcpy(m.data, data, (size_t)rows * step); } return m; }
	Size size() const { return Size(cols, rows); }
	int type() const { return type_; }
	bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
	void convertTo(Mat &dst, int rtype, double alpha = 1.0, double beta = 0.0) const {   // 8U -> 32F (APD.cpp:413), 32F -> 8U (visualisation)
		Mat out(rows, cols, rtype);
		const int cn = (type_ >> 3) + 1, sd = type_ & 7, dd = rtype & 7;
		for (int r = 0; r < rows; ++r) for (int c = 0; c < cols * cn; ++c) {
			const double v = (sd == 0 ? (double)ptr<uchar>(r)[c] : (double)ptr<float>(r)[c]) * alpha + beta;
			if (dd == 0) out.ptr<uchar>(r)[c] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : std::lrint(v))); else out.ptr<float>(r)[c] = (float)v;
		}
		dst = out;
	}
};
template <typename T> class Mat_ : public Mat {
public:
	Mat_() {}
	Mat_(const Mat &m) : Mat(m) {}
	Mat_ &operator=(const Mat &m) { Mat::operator=(m); return *this; }
};

// Raw container: "APDRAW\0\0", int32 rows, cols, channels, then rows*cols*channels bytes (BGR order for 3 channels).
inline Mat imread(const std::string &path, int flags = IMREAD_COLOR) {
	Mat m;
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return m;
	char magic[8]; int hdr[3];
	if (fread(magic, 1, 8, f) == 8 && memcmp(magic, "APDRAW\0\0", 8) == 0 && fread(hdr, 4, 3, f) == 3) {
		const int want = (flags == IMREAD_GRAYSCALE) ? 1 : 3;
		if (hdr[2] == want) { m.create(hdr[0], hdr[1], want == 1 ? CV_8UC1 : CV_8UC3); if (fread(m.data, 1, (size_t)hdr[0] * m.step, f) != (size_t)hdr[0] * m.step) m = Mat(); }
		else if (hdr[2] == 3 && want == 1) {      // BGR -> grey with OpenCV's fixed-point weights (B 1868, G 9617, R 4899, >> 14)
			std::vector<uchar> bgr((size_t)hdr[0] * hdr[1] * 3);
			if (fread(bgr.data(), 1, bgr.size(), f) == bgr.size()) {
				m.create(hdr[0], hdr[1], CV_8UC1);
				for (size_t i = 0; i < (size_t)hdr[0] * hdr[1]; ++i) m.data[i] = (uchar)((bgr[3 * i] * 1868 + bgr[3 * i + 1] * 9617 + bgr[3 * i + 2] * 4899 + 8192) >> 14);
			}
		}
	}
	fclose(f);
	return m;
}
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline void resize(const Mat &, Mat &, Size, double, double, int) { fprintf(stderr, "oracle shim: cv::resize is not available\n"); abort(); }
}  // namespace cv
#endif
