// Minimal stand-in for <opencv2/opencv.hpp>: just enough declarations for the
// UNMODIFIED reference headers (main.h / APD.h) and APD.cu to compile in an
// image that has no OpenCV C++ SDK. Test infrastructure only (oracle/_ref).
#ifndef APD_ORACLE_SHIM_OPENCV_HPP
#define APD_ORACLE_SHIM_OPENCV_HPP
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cstdio>

typedef unsigned char uchar;

// OpenCV's own definitions (opencv2/core/cvdef.h); device code in APD.cu relies
// on their exact NaN behaviour, so they are reproduced literally.
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

#define CV_8UC1 0
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {
template <typename T> struct Size_ {
	T width, height;
	Size_() : width(0), height(0) {}
	Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size2i;
typedef Size2i Size;

// POD view of a row-major matrix; owns nothing (the harness owns the storage).
struct Mat {
	int rows, cols;
	unsigned char *data;
	size_t step;
	int type_;
	Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
	template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
	template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }
	int type() const { return type_; }
};
template <typename T> struct Mat_ : public Mat {};
template <typename T, int N> struct Vec { T val[N]; T &operator[](int i) { return val[i]; } };
typedef Vec<float, 3> Vec3f;
}  // namespace cv
#endif
