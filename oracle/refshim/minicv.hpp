// minicv.hpp -- TEST INFRASTRUCTURE ONLY (oracle side, never linked into the product).
//
// A header-only stand-in for the slice of OpenCV's `cv::` namespace that the
// reference's six hot-path translation units touch, so that those files can be
// compiled UNMODIFIED, straight from /root/reference, into oracle/_ref/libdmc_ref.so
// (see oracle/Makefile).  No OpenCV C++ headers/libs exist in this image; the
// arithmetic of every stand-in below is pinned against Python cv2 4.13 by
// tests/test_refshim_vs_cv2.py (median, Gaussian, dilate/erode, copyMakeBorder,
// convertTo), which is the "oracle of record" for the third-party calls
// (SURVEY.md section 8c).
//
// Semantics that matter for bit parity:
//   * output arrays follow OpenCV's "create() keeps a matching buffer" rule;
//   * float -> integer conversion is cvRound (= cvtss2si: RNE, "integer
//     indefinite" 0x80000000 on NaN/overflow) followed by saturation;
//   * GaussianBlur(32FC1) reproduces OpenCV's scalar sepFilter2D op order;
//   * morphology ignores out-of-image taps; medianBlur replicates the border.
#ifndef DMC_MINICV_HPP
#define DMC_MINICV_HPP

#include <smmintrin.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cassert>
#include <vector>
#include <string>
#include <memory>
#include <algorithm>
#include <stdexcept>
#include <iostream>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef CV_SSE4_1
#define CV_SSE4_1 1
#endif
#define CV_SSE2 1

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef signed char schar;
typedef long long int64;
typedef unsigned long long uint64;

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAKE_TYPE CV_MAKETYPE
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_CPU_SSE4_1 6
#define CV_DECL_ALIGNED(x) __attribute__((aligned(x)))
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(#expr, __FILE__, __LINE__); } while (0)

namespace cv {

class Exception : public std::runtime_error {
public:
    Exception(const char* e, const char* f, int l)
        : std::runtime_error(std::string("CV_Assert failed: ") + e + " at " + f + ":" + std::to_string(l)) {}
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };

template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<int> Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<double> Point3d; typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {}
              int area() const { return width * height; }
              bool operator==(const Size& o) const { return width == o.width && height == o.height; }
              bool operator!=(const Size& o) const { return !(*this == o); } };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {}
              Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };
struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };

static inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
static inline int cvRound(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }
static inline int cvRound(int v) { return v; }
static inline bool checkHardwareSupport(int) { return true; }
static inline int getNumThreads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static inline int64 getTickCount() { return (int64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline double getTickFrequency() { return 1e9; }

static inline size_t depthSize(int depth) { static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return s[depth & 7]; }

template <class T> static inline T saturate_cast(int v);
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255u ? v : v > 0 ? 255 : 0); }
template <> inline schar saturate_cast<schar>(int v) { return (schar)((unsigned)(v + 128) <= 255u ? v : v > 0 ? 127 : -128); }
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= 65535u ? v : v > 0 ? 65535 : 0); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535u ? v : v > 0 ? 32767 : -32768); }
template <> inline int saturate_cast<int>(int v) { return v; }

class Mat {
public:
    uchar* data; int rows, cols; size_t step; int flags;
    std::shared_ptr<uchar> buf;

    Mat() : data(0), rows(0), cols(0), step(0), flags(0) {}
    Mat(int r, int c, int type) : data(0), rows(0), cols(0), step(0), flags(0) { create(r, c, type); }
    Mat(Size s, int type) : data(0), rows(0), cols(0), step(0), flags(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : data((uchar*)ext), rows(r), cols(c), flags(type) {
        step = st ? st : (size_t)c * depthSize(CV_MAT_DEPTH(type)) * CV_MAT_CN(type); }
    Mat(Size s, int type, void* ext, size_t st = 0) : data((uchar*)ext), rows(s.height), cols(s.width), flags(type) {
        step = st ? st : (size_t)s.width * depthSize(CV_MAT_DEPTH(type)) * CV_MAT_CN(type); }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return depthSize(depth()) * channels(); }
    size_t elemSize1() const { return depthSize(depth()); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type) return;   // OpenCV output-array rule
        rows = r; cols = c; flags = type; step = (size_t)c * elemSize();
        size_t bytes = step * (size_t)r;
        void* p = 0;
        if (posix_memalign(&p, 64, bytes + 64) != 0) throw std::bad_alloc();
        buf = std::shared_ptr<uchar>((uchar*)p, free);
        data = (uchar*)p;
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf.reset(); data = 0; rows = cols = 0; step = 0; }

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (!m.empty()) memset(m.data, 0, m.step * r); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(1)); return m; }
    static Mat ones(Size s, int type) { return ones(s.height, s.width, type); }
    static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.setElem(i, i, 1.0); return m; }

    uchar* ptr(int y = 0) { return data + step * y; }
    const uchar* ptr(int y = 0) const { return data + step * y; }
    template <class T> T* ptr(int y = 0) { return (T*)(data + step * y); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + step * y); }
    template <class T> T& at(int y, int x) { return ((T*)(data + step * y))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + step * y))[x]; }
    template <class T> T& at(int i) { return ((T*)data)[i]; }
    template <class T> const T& at(int i) const { return ((const T*)data)[i]; }

    Mat operator()(const Rect& r) const {
        Mat m; m.data = data + step * r.y + (size_t)r.x * elemSize(); m.rows = r.height; m.cols = r.width;
        m.step = step; m.flags = flags; m.buf = buf; return m; }

    void copyTo(Mat& dst) const {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.flags == flags && dst.step == step) return;
        dst.create(rows, cols, flags);
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; y++) memmove(dst.data + dst.step * y, data + step * y, rb);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }

    void setElem(int y, int x, double v) {
        switch (depth()) {
        case CV_8U: at<uchar>(y, x) = saturate_cast<uchar>(cvRound(v)); break;
        case CV_8S: at<schar>(y, x) = saturate_cast<schar>(cvRound(v)); break;
        case CV_16U: at<ushort>(y, x) = saturate_cast<ushort>(cvRound(v)); break;
        case CV_16S: at<short>(y, x) = saturate_cast<short>(cvRound(v)); break;
        case CV_32S: at<int>(y, x) = cvRound(v); break;
        case CV_32F: at<float>(y, x) = (float)v; break;
        default: at<double>(y, x) = v; break;
        }
    }
    Mat& setTo(const Scalar& s) {
        int cn = channels();
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) for (int c = 0; c < cn; c++) setElem(y, x * cn + c, s.val[c & 3]);
        return *this;
    }

    // convertTo: dst = saturate(src*alpha+beta); integer targets round with cvRound (RNE) then saturate.
    template <class S> void cvtRow(const S* s, uchar* d, int n, int ddepth, double alpha, double beta) const {
        bool scale = !(alpha == 1.0 && beta == 0.0);
        for (int i = 0; i < n; i++) {
            switch (ddepth) {
            case CV_32F: ((float*)d)[i] = scale ? (float)(s[i] * alpha + beta) : (float)s[i]; break;
            case CV_64F: ((double*)d)[i] = scale ? (double)(s[i] * alpha + beta) : (double)s[i]; break;
            default: {
                int iv;
                if (scale) iv = cvRound((double)s[i] * alpha + beta);
                else iv = cvRound(s[i]);
                if (ddepth == CV_8U) d[i] = saturate_cast<uchar>(iv);
                else if (ddepth == CV_8S) ((schar*)d)[i] = saturate_cast<schar>(iv);
                else if (ddepth == CV_16U) ((ushort*)d)[i] = saturate_cast<ushort>(iv);
                else if (ddepth == CV_16S) ((short*)d)[i] = saturate_cast<short>(iv);
                else ((int*)d)[i] = iv;
            } }
        }
    }
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const {
        int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
        int cn = channels();
        if (ddepth == depth() && alpha == 1.0 && beta == 0.0) { copyTo(dst); return; }
        Mat src = *this;                       // keeps the source alive if dst aliases it
        Mat out;
        if (dst.data == data) out.create(rows, cols, CV_MAKETYPE(ddepth, cn));
        else { dst.create(rows, cols, CV_MAKETYPE(ddepth, cn)); out = dst; }
        const bool plain = alpha == 1.0 && beta == 0.0;
        for (int y = 0; y < rows; y++) {
            const uchar* s = src.ptr(y); uchar* d = out.ptr(y); int n = cols * cn;
            if (plain && src.depth() == CV_8U && ddepth == CV_32F) {          // vector forms of the two conversions the chain uses
                int i = 0;
                for (; i + 4 <= n; i += 4) { int w; memcpy(&w, s + i, 4); _mm_storeu_ps((float*)d + i, _mm_cvtepi32_ps(_mm_cvtepu8_epi32(_mm_cvtsi32_si128(w)))); }
                for (; i < n; i++) ((float*)d)[i] = (float)s[i];
                continue;
            }
            if (plain && src.depth() == CV_32F && ddepth == CV_8U) {          // cvtps2dq rounds like cvRound; the packs saturate
                int i = 0;
                for (; i + 4 <= n; i += 4) {
                    __m128i v = _mm_cvtps_epi32(_mm_loadu_ps((const float*)s + i));
                    v = _mm_packs_epi32(v, v); v = _mm_packus_epi16(v, v);
                    int w = _mm_cvtsi128_si32(v); memcpy(d + i, &w, 4);
                }
                for (; i < n; i++) d[i] = saturate_cast<uchar>(cvRound(((const float*)s)[i]));
                continue;
            }
            switch (src.depth()) {
            case CV_8U: cvtRow((const uchar*)s, d, n, ddepth, alpha, beta); break;
            case CV_8S: cvtRow((const schar*)s, d, n, ddepth, alpha, beta); break;
            case CV_16U: cvtRow((const ushort*)s, d, n, ddepth, alpha, beta); break;
            case CV_16S: cvtRow((const short*)s, d, n, ddepth, alpha, beta); break;
            case CV_32S: cvtRow((const int*)s, d, n, ddepth, alpha, beta); break;
            case CV_32F: cvtRow((const float*)s, d, n, ddepth, alpha, beta); break;
            default: cvtRow((const double*)s, d, n, ddepth, alpha, beta); break;
            }
        }
        if (dst.data == data) dst = out;
    }

    Mat t() const {
        Mat m(cols, rows, flags); size_t es = elemSize();
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) memcpy(m.data + m.step * x + es * y, data + step * y + es * x, es);
        return m;
    }
};

static inline Mat operator*(const Mat& a, const Mat& b) {           // 64F matrix product (renderer only)
    CV_Assert(a.type() == CV_64F && b.type() == CV_64F && a.cols == b.rows);
    Mat c = Mat::zeros(a.rows, b.cols, CV_64F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < b.cols; j++) { double s = 0; for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j); c.at<double>(i, j) = s; }
    return c;
}
static inline Mat operator*(double s, const Mat& a) {
    Mat c; a.copyTo(c); int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < n; x++) {
        if (a.depth() == CV_32F) c.at<float>(y, x) = (float)(a.at<float>(y, x) * s);
        else if (a.depth() == CV_64F) c.at<double>(y, x) = a.at<double>(y, x) * s;
        else CV_Assert(!"scalar*Mat: float types only");
    }
    return c;
}
static inline Mat operator*(const Mat& a, double s) { return s * a; }

class ParallelLoopBody { public: virtual ~ParallelLoopBody() {} virtual void operator()(const Range& r) const = 0; };

// Row-stripe parallel loop (OpenCV's backend is TBB/pthreads; here: OpenMP over contiguous stripes).
static inline void parallel_for_(const Range& range, const ParallelLoopBody& body, double nstripes = -1.) {
    (void)nstripes;
    int n = range.end - range.start; if (n <= 0) return;
    int nt = getNumThreads(); int stripes = std::max(1, std::min(n, nt * 4));
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < stripes; s++) {
        int a = range.start + (int)((int64)n * s / stripes), b = range.start + (int)((int64)n * (s + 1) / stripes);
        if (b > a) body(Range(a, b));
    }
}

// OpenCV's borderInterpolate (modules/core/src/copy.cpp)
static inline int borderInterpolate(int p, int len, int borderType) {
    if ((unsigned)p < (unsigned)len) return p;
    if (borderType == BORDER_REPLICATE) return p < 0 ? 0 : len - 1;
    if (borderType == BORDER_REFLECT || borderType == BORDER_REFLECT_101) {
        int delta = borderType == BORDER_REFLECT_101;
        if (len == 1) return 0;
        do { if (p < 0) p = -p - 1 + delta; else p = len - 1 - (p - len) - delta; } while ((unsigned)p >= (unsigned)len);
        return p;
    }
    if (borderType == BORDER_WRAP) { if (p < 0) p -= ((p - len + 1) / len) * len; if (p >= len) p %= len; return p; }
    return -1;
}

static inline void copyMakeBorder(const Mat& src_, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar()) {
    Mat src = src_;
    borderType &= ~BORDER_ISOLATED;
    CV_Assert(borderType != BORDER_CONSTANT);
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    size_t es = src.elemSize();
    std::vector<int> xmap(out.cols);
    for (int x = 0; x < out.cols; x++) xmap[x] = borderInterpolate(x - left, src.cols, borderType);
    for (int y = 0; y < out.rows; y++) {
        const uchar* s = src.ptr(borderInterpolate(y - top, src.rows, borderType)); uchar* d = out.ptr(y);
        for (int x = 0; x < out.cols; x++) memcpy(d + es * x, s + es * xmap[x], es);
    }
    dst = out;
}


// OpenCV output-array rule for filters that computed into a scratch `out`: write into dst's existing
// buffer when it already has the right size/type (this covers in-place calls), else hand over `out`.
static inline void finishOutput(const Mat& out, Mat& dst) {
    if (!dst.empty() && dst.size() == out.size() && dst.type() == out.type()) out.copyTo(dst); else dst = out;
}

template <class T> static void medianBlur_(const Mat& src, Mat& dst, int k) {
    int r = k / 2, cn = src.channels(); std::vector<T> w((size_t)k * k);
    for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) for (int c = 0; c < cn; c++) {
        int n = 0;
        for (int dy = -r; dy <= r; dy++) { const T* s = src.ptr<T>(std::min(std::max(y + dy, 0), src.rows - 1));
            for (int dx = -r; dx <= r; dx++) w[n++] = s[std::min(std::max(x + dx, 0), src.cols - 1) * cn + c]; }
        std::nth_element(w.begin(), w.begin() + n / 2, w.end());
        dst.ptr<T>(y)[x * cn + c] = w[n / 2];
    }
}
// 8U, k = 3 / 5: min/max sorting networks on 16 pixels at a time over a replicate-padded copy (the same kind of
// code OpenCV runs for small kernels), so that the CPU baseline is not dominated by a slow stand-in.
#define DMC_MCE(a, b) { __m128i _t = _mm_min_epu8(a, b); b = _mm_max_epu8(a, b); a = _t; }
static inline void medianBlurSmall8u(const Mat& src, Mat& out, int k) {
    const int r = k / 2, W = src.cols, H = src.rows, PW = W + 2 * r + 16;
    Mat pad(H + 2 * r, PW, CV_8U);
    for (int y = 0; y < H + 2 * r; y++) {
        const uchar* s = src.ptr<uchar>(std::min(std::max(y - r, 0), H - 1)); uchar* d = pad.ptr<uchar>(y);
        for (int x = 0; x < PW; x++) d[x] = s[std::min(std::max(x - r, 0), W - 1)];
    }
    std::vector<uchar> rowbuf(W + 16);
    for (int y = 0; y < H; y++) {
        uchar* d = rowbuf.data();
        for (int x = 0; x < W; x += 16) {
            __m128i p[25];
            for (int dy = 0; dy < k; dy++) for (int dx = 0; dx < k; dx++) p[dy * k + dx] = _mm_loadu_si128((const __m128i*)(pad.ptr<uchar>(y + dy) + x + dx));
            if (k == 3) {
                DMC_MCE(p[1], p[2]) DMC_MCE(p[4], p[5]) DMC_MCE(p[7], p[8]) DMC_MCE(p[0], p[1]) DMC_MCE(p[3], p[4]) DMC_MCE(p[6], p[7])
                DMC_MCE(p[1], p[2]) DMC_MCE(p[4], p[5]) DMC_MCE(p[7], p[8]) DMC_MCE(p[0], p[3]) DMC_MCE(p[5], p[8]) DMC_MCE(p[4], p[7])
                DMC_MCE(p[3], p[6]) DMC_MCE(p[1], p[4]) DMC_MCE(p[2], p[5]) DMC_MCE(p[4], p[7]) DMC_MCE(p[4], p[2]) DMC_MCE(p[6], p[4])
                DMC_MCE(p[4], p[2])
                _mm_storeu_si128((__m128i*)(d + x), p[4]);
            } else {
                DMC_MCE(p[0], p[1]) DMC_MCE(p[3], p[4]) DMC_MCE(p[2], p[4]) DMC_MCE(p[2], p[3]) DMC_MCE(p[6], p[7]) DMC_MCE(p[5], p[7])
                DMC_MCE(p[5], p[6]) DMC_MCE(p[9], p[10]) DMC_MCE(p[8], p[10]) DMC_MCE(p[8], p[9]) DMC_MCE(p[12], p[13]) DMC_MCE(p[11], p[13])
                DMC_MCE(p[11], p[12]) DMC_MCE(p[15], p[16]) DMC_MCE(p[14], p[16]) DMC_MCE(p[14], p[15]) DMC_MCE(p[18], p[19]) DMC_MCE(p[17], p[19])
                DMC_MCE(p[17], p[18]) DMC_MCE(p[21], p[22]) DMC_MCE(p[20], p[22]) DMC_MCE(p[20], p[21]) DMC_MCE(p[23], p[24]) DMC_MCE(p[2], p[5])
                DMC_MCE(p[3], p[6]) DMC_MCE(p[0], p[6]) DMC_MCE(p[0], p[3]) DMC_MCE(p[4], p[7]) DMC_MCE(p[1], p[7]) DMC_MCE(p[1], p[4])
                DMC_MCE(p[11], p[14]) DMC_MCE(p[8], p[14]) DMC_MCE(p[8], p[11]) DMC_MCE(p[12], p[15]) DMC_MCE(p[9], p[15]) DMC_MCE(p[9], p[12])
                DMC_MCE(p[13], p[16]) DMC_MCE(p[10], p[16]) DMC_MCE(p[10], p[13]) DMC_MCE(p[20], p[23]) DMC_MCE(p[17], p[23]) DMC_MCE(p[17], p[20])
                DMC_MCE(p[21], p[24]) DMC_MCE(p[18], p[24]) DMC_MCE(p[18], p[21]) DMC_MCE(p[19], p[22]) DMC_MCE(p[8], p[17]) DMC_MCE(p[9], p[18])
                DMC_MCE(p[0], p[18]) DMC_MCE(p[0], p[9]) DMC_MCE(p[10], p[19]) DMC_MCE(p[1], p[19]) DMC_MCE(p[1], p[10]) DMC_MCE(p[11], p[20])
                DMC_MCE(p[2], p[20]) DMC_MCE(p[2], p[11]) DMC_MCE(p[12], p[21]) DMC_MCE(p[3], p[21]) DMC_MCE(p[3], p[12]) DMC_MCE(p[13], p[22])
                DMC_MCE(p[4], p[22]) DMC_MCE(p[4], p[13]) DMC_MCE(p[14], p[23]) DMC_MCE(p[5], p[23]) DMC_MCE(p[5], p[14]) DMC_MCE(p[15], p[24])
                DMC_MCE(p[6], p[24]) DMC_MCE(p[6], p[15]) DMC_MCE(p[7], p[16]) DMC_MCE(p[7], p[19]) DMC_MCE(p[13], p[21]) DMC_MCE(p[15], p[23])
                DMC_MCE(p[7], p[13]) DMC_MCE(p[7], p[15]) DMC_MCE(p[1], p[9]) DMC_MCE(p[3], p[11]) DMC_MCE(p[5], p[17]) DMC_MCE(p[11], p[17])
                DMC_MCE(p[9], p[17]) DMC_MCE(p[4], p[10]) DMC_MCE(p[6], p[12]) DMC_MCE(p[7], p[14]) DMC_MCE(p[4], p[6]) DMC_MCE(p[4], p[7])
                DMC_MCE(p[12], p[14]) DMC_MCE(p[10], p[14]) DMC_MCE(p[6], p[7]) DMC_MCE(p[10], p[12]) DMC_MCE(p[6], p[10]) DMC_MCE(p[6], p[17])
                DMC_MCE(p[12], p[17]) DMC_MCE(p[7], p[17]) DMC_MCE(p[7], p[10]) DMC_MCE(p[12], p[18]) DMC_MCE(p[7], p[12]) DMC_MCE(p[10], p[18])
                DMC_MCE(p[12], p[20]) DMC_MCE(p[10], p[20]) DMC_MCE(p[10], p[12])
                _mm_storeu_si128((__m128i*)(d + x), p[12]);
            }
        }
        memcpy(out.ptr<uchar>(y), d, W);
    }
}

// cv::medianBlur: exact median of the k x k window, BORDER_REPLICATE.
static inline void medianBlur(const Mat& src_, Mat& dst, int ksize) {
    CV_Assert(ksize % 2 == 1);
    Mat src = src_;
    if (ksize <= 1) { src.copyTo(dst); return; }
    Mat out(src.rows, src.cols, src.type());
    if (src.type() == CV_8UC1 && (ksize == 3 || ksize == 5)) medianBlurSmall8u(src, out, ksize);
    else if (src.depth() == CV_8U) medianBlur_<uchar>(src, out, ksize);
    else if (src.depth() == CV_16U) medianBlur_<ushort>(src, out, ksize);
    else if (src.depth() == CV_32F) medianBlur_<float>(src, out, ksize);
    else CV_Assert(!"medianBlur: unsupported depth");
    finishOutput(out, dst);
}

// cv::getGaussianKernel(n, sigma, CV_32F) as in OpenCV 4.x: taps in double, normalised in double, cast to float.
static inline std::vector<float> getGaussianKernel32f(int n, double sigma) {
    std::vector<double> t(n); std::vector<float> k(n);
    static const float small_tab[5][9] = {{1.f}, {0.25f, 0.5f, 0.25f}, {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
        {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f},
        {0.015625f, 0.05078125f, 0.1171875f, 0.19921875f, 0.234375f, 0.19921875f, 0.1171875f, 0.05078125f, 0.015625f}};
    if (sigma <= 0 && (n & 1) && n <= 9) { for (int i = 0; i < n; i++) k[i] = small_tab[n >> 1][i]; return k; }   // fixed taps of OpenCV 4.13
    double sigmaX = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2X = -0.5 / (sigmaX * sigmaX), sum = 0;
    for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; t[i] = std::exp(scale2X * x * x); sum += t[i]; }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) k[i] = (float)(t[i] * sum);
    return k;
}

// cv::GaussianBlur on 32FC1, OpenCV scalar sepFilter2D op order (rows first, then columns):
//   d <= 5 : x0*k0 + sum_i (x[-i]+x[+i])*k_i            (SymmRowSmallFilter / SymmColumnSmallFilter shape)
//   d >= 7 : rows  = sequential left-to-right  k0*x[-r] + k1*x[-r+1] + ...   (RowFilter)
//            cols  = x0*k0 + sum_i (x[-i]+x[+i])*k_i     (SymmColumnFilter)
static inline void GaussianBlur(const Mat& src_, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    Mat src = src_;
    CV_Assert(src.type() == CV_32FC1);
    if (sigmaY <= 0) sigmaY = sigmaX;
    int kw = ksize.width, kh = ksize.height;
    if (src.rows == 1) kh = 1;
    if (src.cols == 1) kw = 1;
    if (kw == 1 && kh == 1) { src.copyTo(dst); return; }
    std::vector<float> kx = getGaussianKernel32f(kw, sigmaX), ky = getGaussianKernel32f(kh, sigmaY);
    int rx = kw / 2, ry = kh / 2, W = src.cols, H = src.rows;
    Mat tmp(H, W, CV_32F), out(H, W, CV_32F);
    std::vector<int> xm(W + 2 * rx);
    for (int x = -rx; x < W + rx; x++) xm[x + rx] = borderInterpolate(x, W, borderType);
    for (int y = 0; y < H; y++) {
        const float* s = src.ptr<float>(y); float* d = tmp.ptr<float>(y);
        for (int x = 0; x < W; x++) {
            if (x == rx && W - rx - x >= 4) {       // interior: same operations in the same order, four columns per step
                for (; x + 4 <= W - rx; x += 4) {
                    __m128 acc;
                    if (kw <= 5) { acc = _mm_mul_ps(_mm_loadu_ps(s + x), _mm_set1_ps(kx[rx])); for (int i = 1; i <= rx; i++) acc = _mm_add_ps(acc, _mm_mul_ps(_mm_add_ps(_mm_loadu_ps(s + x - i), _mm_loadu_ps(s + x + i)), _mm_set1_ps(kx[rx + i]))); }
                    else { acc = _mm_mul_ps(_mm_set1_ps(kx[0]), _mm_loadu_ps(s + x - rx)); for (int i = 1; i < kw; i++) acc = _mm_add_ps(acc, _mm_mul_ps(_mm_set1_ps(kx[i]), _mm_loadu_ps(s + x + i - rx))); }
                    _mm_storeu_ps(d + x, acc);
                }
                if (x >= W) break;
            }
            const int* m = &xm[x + rx];
            float acc;
            if (kw <= 5) { acc = s[m[0]] * kx[rx]; for (int i = 1; i <= rx; i++) acc = acc + (s[m[-i]] + s[m[i]]) * kx[rx + i]; }
            else { acc = kx[0] * s[m[-rx]]; for (int i = 1; i < kw; i++) acc = acc + kx[i] * s[m[i - rx]]; }
            d[x] = acc;
        }
    }
    for (int y = 0; y < H; y++) {
        float* d = out.ptr<float>(y); const float* s0 = tmp.ptr<float>(y);
        std::vector<const float*> ra(ry + 1), rb(ry + 1);
        for (int i = 1; i <= ry; i++) { ra[i] = tmp.ptr<float>(borderInterpolate(y - i, H, borderType)); rb[i] = tmp.ptr<float>(borderInterpolate(y + i, H, borderType)); }
        int x = 0;
        for (; x + 4 <= W; x += 4) {
            __m128 acc = _mm_mul_ps(_mm_loadu_ps(s0 + x), _mm_set1_ps(ky[ry]));
            for (int i = 1; i <= ry; i++) acc = _mm_add_ps(acc, _mm_mul_ps(_mm_add_ps(_mm_loadu_ps(ra[i] + x), _mm_loadu_ps(rb[i] + x)), _mm_set1_ps(ky[ry + i])));
            _mm_storeu_ps(d + x, acc);
        }
        for (; x < W; x++) {
            float acc = s0[x] * ky[ry];
            for (int i = 1; i <= ry; i++) acc = acc + (ra[i][x] + rb[i][x]) * ky[ry + i];
            d[x] = acc;
        }
    }
    finishOutput(out, dst);
}

template <class T, bool IsMax> static void morph_(const Mat& src, Mat& out, int kw, int kh) {
    int rx = kw / 2, ry = kh / 2; Mat tmp(src.rows, src.cols, src.type());
    for (int y = 0; y < src.rows; y++) { const T* s = src.ptr<T>(y); T* d = tmp.ptr<T>(y);
        for (int x = 0; x < src.cols; x++) { T m = s[x]; for (int i = std::max(0, x - rx); i <= std::min(src.cols - 1, x + rx); i++) m = IsMax ? std::max(m, s[i]) : std::min(m, s[i]); d[x] = m; } }
    for (int y = 0; y < src.rows; y++) { T* d = out.ptr<T>(y);
        for (int x = 0; x < src.cols; x++) { T m = tmp.ptr<T>(y)[x]; for (int i = std::max(0, y - ry); i <= std::min(src.rows - 1, y + ry); i++) { T v = tmp.ptr<T>(i)[x]; m = IsMax ? std::max(m, v) : std::min(m, v); } d[x] = m; } }
}
template <bool IsMax> static void morph8u_fast(const Mat& src, Mat& out, int kw, int kh) {
    const int rx = kw / 2, ry = kh / 2, W = src.cols, H = src.rows, PW = W + 2 * rx + 16;
    Mat tmp(H, W + 16, CV_8U); std::vector<uchar> pad(PW);
    for (int y = 0; y < H; y++) {                                   // horizontal pass on a replicate-padded row
        const uchar* s = src.ptr<uchar>(y); uchar* d = tmp.ptr<uchar>(y);
        memset(pad.data(), s[0], rx); memcpy(pad.data() + rx, s, W); memset(pad.data() + rx + W, s[W - 1], PW - rx - W);
        for (int x = 0; x < W; x += 16) {
            __m128i m = _mm_loadu_si128((const __m128i*)(pad.data() + x));
            for (int i = 1; i <= 2 * rx; i++) { __m128i v = _mm_loadu_si128((const __m128i*)(pad.data() + x + i)); m = IsMax ? _mm_max_epu8(m, v) : _mm_min_epu8(m, v); }
            _mm_storeu_si128((__m128i*)(d + x), m);
        }
    }
    for (int y = 0; y < H; y++) {                                   // vertical pass, out-of-image rows ignored
        uchar* d = out.ptr<uchar>(y); int y0 = std::max(0, y - ry), y1 = std::min(H - 1, y + ry);
        for (int x = 0; x < W; x += 16) {
            __m128i m = _mm_loadu_si128((const __m128i*)(tmp.ptr<uchar>(y0) + x));
            for (int i = y0 + 1; i <= y1; i++) { __m128i v = _mm_loadu_si128((const __m128i*)(tmp.ptr<uchar>(i) + x)); m = IsMax ? _mm_max_epu8(m, v) : _mm_min_epu8(m, v); }
            if (x + 16 <= W) _mm_storeu_si128((__m128i*)(d + x), m);
            else { uchar t[16]; _mm_storeu_si128((__m128i*)t, m); memcpy(d + x, t, W - x); }
        }
    }
}
// cv::dilate / cv::erode with an all-ones rectangular element, default border (out-of-image taps ignored).
template <bool IsMax> static void morph(const Mat& src_, Mat& dst, const Mat& kernel) {
    Mat src = src_; CV_Assert(src.channels() == 1);
    Mat out(src.rows, src.cols, src.type()); int kw = kernel.cols, kh = kernel.rows;
    switch (src.depth()) {
    case CV_8U: morph8u_fast<IsMax>(src, out, kw, kh); break;
    case CV_16U: morph_<ushort, IsMax>(src, out, kw, kh); break;
    case CV_16S: morph_<short, IsMax>(src, out, kw, kh); break;
    case CV_32F: morph_<float, IsMax>(src, out, kw, kh); break;
    case CV_64F: morph_<double, IsMax>(src, out, kw, kh); break;
    default: CV_Assert(!"morph: unsupported depth");
    }
    finishOutput(out, dst);
}
static inline void dilate(const Mat& s, Mat& d, const Mat& k) { morph<true>(s, d, k); }
static inline void erode(const Mat& s, Mat& d, const Mat& k) { morph<false>(s, d, k); }

template <class T> static void absdiff_(const Mat& a, const Mat& b, Mat& c) { int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); const T* q = b.ptr<T>(y); T* d = c.ptr<T>(y); for (int x = 0; x < n; x++) d[x] = p[x] > q[x] ? (T)(p[x] - q[x]) : (T)(q[x] - p[x]); } }
template <> void absdiff_<short>(const Mat& a, const Mat& b, Mat& c) { int n = a.cols * a.channels();   // saturating, as OpenCV
    for (int y = 0; y < a.rows; y++) { const short* p = a.ptr<short>(y); const short* q = b.ptr<short>(y); short* d = c.ptr<short>(y); for (int x = 0; x < n; x++) d[x] = saturate_cast<short>(std::abs((int)p[x] - (int)q[x])); } }
template <> void absdiff_<float>(const Mat& a, const Mat& b, Mat& c) { int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) { const float* p = a.ptr<float>(y); const float* q = b.ptr<float>(y); float* d = c.ptr<float>(y); for (int x = 0; x < n; x++) d[x] = std::fabs(p[x] - q[x]); } }
template <> void absdiff_<double>(const Mat& a, const Mat& b, Mat& c) { int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) { const double* p = a.ptr<double>(y); const double* q = b.ptr<double>(y); double* d = c.ptr<double>(y); for (int x = 0; x < n; x++) d[x] = std::fabs(p[x] - q[x]); } }
static inline void absdiff(const Mat& a, const Mat& b, Mat& c) {
    CV_Assert(a.size() == b.size() && a.type() == b.type()); c.create(a.rows, a.cols, a.type());
    switch (a.depth()) { case CV_8U: absdiff_<uchar>(a, b, c); break; case CV_16U: absdiff_<ushort>(a, b, c); break; case CV_16S: absdiff_<short>(a, b, c); break;
        case CV_32F: absdiff_<float>(a, b, c); break; case CV_64F: absdiff_<double>(a, b, c); break; default: CV_Assert(!"absdiff depth"); }
}
// cv::min on floats is (a < b ? a : b)-like via minps(b-operand order); OpenCV: std::min(a,b) = (b < a) ? b : a.
template <class T> static void min_(const Mat& a, const Mat& b, Mat& c) { int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; y++) { const T* p = a.ptr<T>(y); const T* q = b.ptr<T>(y); T* d = c.ptr<T>(y); for (int x = 0; x < n; x++) d[x] = std::min(p[x], q[x]); } }
static inline void min(const Mat& a, const Mat& b, Mat& c) {
    CV_Assert(a.size() == b.size() && a.type() == b.type()); c.create(a.rows, a.cols, a.type());
    switch (a.depth()) { case CV_8U: min_<uchar>(a, b, c); break; case CV_16U: min_<ushort>(a, b, c); break; case CV_16S: min_<short>(a, b, c); break;
        case CV_32F: min_<float>(a, b, c); break; case CV_64F: min_<double>(a, b, c); break; default: CV_Assert(!"min depth"); }
}
static inline void split(const Mat& src, std::vector<Mat>& v) {
    int cn = src.channels(); size_t e1 = src.elemSize1(); v.resize(cn);
    for (int c = 0; c < cn; c++) { v[c] = Mat(src.rows, src.cols, src.depth());
        for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) memcpy(v[c].ptr(y) + e1 * x, src.ptr(y) + e1 * (x * cn + c), e1); }
}
static inline void merge(const std::vector<Mat>& v, Mat& dst) {
    int cn = (int)v.size(); CV_Assert(cn > 0); size_t e1 = v[0].elemSize1();
    dst.create(v[0].rows, v[0].cols, CV_MAKETYPE(v[0].depth(), cn));
    for (int c = 0; c < cn; c++) for (int y = 0; y < dst.rows; y++) for (int x = 0; x < dst.cols; x++) memcpy(dst.ptr(y) + e1 * (x * cn + c), v[c].ptr(y) + e1 * x, e1);
}

}  // namespace cv

#endif
