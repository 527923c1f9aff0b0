// filter.h -- drop-in for PostFilterSetForDepthCoding/filter.h of Wavelet303/DepthMapCompression.
//
// The declarations below are the reference's (filter.h:12-45), unchanged: same names, argument order, default
// arguments, the enum, the `using namespace` lines and the PostFilterSet class.  A translation unit that was written
// against the reference header (main.cpp:303, :485, :495, :526) compiles against this one unchanged; instead of the
// reference's SSE4.1 .cpp files it links libdmc_b200.so, and every operator below is ONE call through the C ABI of
// include/dmc_c.h to hand-written sm_100a CUDA kernels.  There is no CPU implementation behind this header (not even for the layout helper
// splitBGRLineInterleave).
//
// cv::Mat can come from a real OpenCV (>= 2.4.5, as the reference requires) or from any header that provides the
// same Mat surface (tests build against oracle/refshim/minicv.hpp because this image has no OpenCV C++ headers).
//
// Behaviour kept from the reference:
//   * output Mats are allocated exactly where the reference allocates them (create-if-empty, copyTo, zeros);
//   * in-place calls (src.data == dest.data) are legal everywhere;
//   * the (type, method) pairs the reference's dispatchers silently ignore leave `dst` untouched here too;
//   * a type / size mismatch that trips CV_Assert in the reference throws cv::Exception here.
// Deliberate differences (DESIGN.md "boundary"): FULL_KERNEL_PAIR computes the FULL_KERNEL result (the reference's
// PAIR code is racy and leaves columns unwritten); smallGaussianBlur and the PostFilterSet entry points accept the
// types the reference's call sites use (CV_8UC1 input); a PostFilterSet keeps its scratch in the per-thread context.
#ifndef _FILTER_H_
#define _FILTER_H_

#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <iostream>
#include <string>

#include "dmc_c.h"

using namespace cv;
using namespace std;

//rgb interleave function for bilateral filter
void splitBGRLineInterleave( const Mat& src, Mat& dest);

void smallGaussianBlur(const Mat& src, Mat& dest, const int d, const double sigma);

//max, min filter and blur remove filter by using min-max filter
void maxFilter(const Mat& src, Mat& dest, Size ksize, int borderType=cv::BORDER_REPLICATE);
void minFilter(const Mat& src, Mat& dest, Size ksize, int borderType=cv::BORDER_REPLICATE);
void blurRemoveMinMax(Mat& src, Mat& dest, const int r);
void blurRemoveMinMaxBase(Mat& src, Mat& dest, const int r);

//range filter functions
enum
{
	FULL_KERNEL = 0,
	FULL_KERNEL_PAIR,
	SEPARABLE_KERNEL
};
void binalyWeightedRangeFilter(const Mat& src, Mat& dst, Size kernelSize, float threshold, int method, int borderType=cv::BORDER_REPLICATE);

//post filter set class
class PostFilterSet
{
	Mat buff,bufff;
public:
	PostFilterSet();
	~PostFilterSet();
	void filterDisp8U2Depth32F(Mat& src, Mat& dest, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method=FULL_KERNEL);
	void filterDisp8U2Depth16U(Mat& src, Mat& dest, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method=FULL_KERNEL);
	void filterDisp8U2Disp32F(Mat& src, Mat& dest, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method=FULL_KERNEL);
	void operator()(Mat& src, Mat& dest, int median_r, int gaussian_r, int minmax_r, int brange_r, int brange_th, int brange_method=FULL_KERNEL);
};

//boundary reconstruction filter for lossy encoded depth maps
void boundaryReconstructionFilter(Mat& src, Mat& dest, Size ksize, const float frec, const float color, const float space);

// ---------------------------------------------------------------------------------------------------------------------
// Implementation: thin inline forwarding to the C ABI.  (The reference implements these in postFilterSet.cpp,
// binalyWeightedRangeFilter.cpp, minmaxFilter.cpp, boundaryReconstructionFilter.cpp and split.cpp.)
// ---------------------------------------------------------------------------------------------------------------------
namespace dmc_dropin {

struct ContextHolder {
    dmc_ctx* ctx;
    ContextHolder() : ctx(0) {
        int dev = 0;
        if (const char* e = getenv("DMC_DEVICE")) dev = atoi(e);
        int rc = dmc_create(dev, &ctx);
        if (rc != DMC_OK) { std::string m = std::string("libdmc_b200: ") + dmc_last_error(0); ctx = 0; throw std::runtime_error(m); }
    }
    ~ContextHolder() { if (ctx) dmc_destroy(ctx); }
};
// one context per host thread: contexts are not thread-safe, exactly like a PostFilterSet instance
inline dmc_ctx* context() { static thread_local ContextHolder h; return h.ctx; }

inline dmc_image wrap(const Mat& m) {
    dmc_image im; im.data = (void*)m.data; im.rows = m.rows; im.cols = m.cols; im.cvtype = m.type();
    im.step = m.rows > 1 ? (size_t)m.step : 0; im.mem = DMC_MEM_HOST; return im;
}
inline void raise(const char* what) {
    std::string msg = std::string(what) + ": " + dmc_last_error(context());
#ifdef DMC_MINICV_HPP
    throw cv::Exception(msg.c_str(), __FILE__, __LINE__);
#else
    CV_Error(cv::Error::StsAssert, msg);
#endif
}
inline void check(int rc, const char* what) { if (rc < 0) raise(what); }

}  // namespace dmc_dropin

inline void splitBGRLineInterleave(const Mat& src, Mat& dest)
{
	// split.cpp:167-177.  (The CUDA range filter reads interleaved pixels directly; this stays for callers of the header.)
	if (src.type() != CV_MAKE_TYPE(CV_8U,3) && src.type() != CV_MAKE_TYPE(CV_32F,3)) return;
	Mat s = src;
	dest.create(Size(s.cols, s.rows * 3), s.depth());                            // split.cpp:13 / :106
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_split_bgr_line_interleave(dmc_dropin::context(), &a, &b), "splitBGRLineInterleave");
}

inline void smallGaussianBlur(const Mat& src, Mat& dest, const int d, const double sigma)
{
	if (d == 0) { src.copyTo(dest); return; }                                   // postFilterSet.cpp:6-10
	Mat s = src;                                                                   // keeps the buffer alive if dest aliases src
	dest.create(s.size(), s.type());                                              // convertTo(dest, src.type()) :15
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_small_gaussian(dmc_dropin::context(), &a, &b, d, sigma), "smallGaussianBlur");
}

inline void maxFilter(const Mat& src, Mat& dest, Size ksize, int borderType)
{
	if (src.channels() != 1) return;                                               // minmaxFilter.cpp:259
	if (src.type() != CV_8U && src.type() != CV_16S && src.type() != CV_16U && src.type() != CV_32F) return;   // :316-333
	Mat s = src;
	if (dest.empty() || dest.size() != s.size() || dest.type() != s.type()) dest = Mat::zeros(s.size(), s.type());
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_max_filter(dmc_dropin::context(), &a, &b, ksize.width, ksize.height, borderType), "maxFilter");
}

inline void minFilter(const Mat& src, Mat& dest, Size ksize, int borderType)
{
	if (src.channels() != 1) return;
	if (src.type() != CV_8U && src.type() != CV_16S && src.type() != CV_16U && src.type() != CV_32F) return;
	Mat s = src;
	if (dest.empty() || dest.size() != s.size() || dest.type() != s.type()) dest = Mat::zeros(s.size(), s.type());
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_min_filter(dmc_dropin::context(), &a, &b, ksize.width, ksize.height, borderType), "minFilter");
}

inline void blurRemoveMinMax(Mat& src, Mat& dest, const int r)
{
	if (src.data != dest.data) src.copyTo(dest);                                   // minmaxFilter.cpp:52 (allocates dest)
	dmc_image a = dmc_dropin::wrap(src), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_blur_remove_minmax(dmc_dropin::context(), &a, &b, r), "blurRemoveMinMax");
}

inline void blurRemoveMinMaxBase(Mat& src, Mat& dest, const int r) { blurRemoveMinMax(src, dest, r); }   // scalar twin, same result

inline void binalyWeightedRangeFilter(const Mat& src, Mat& dst, Size kernelSize, float threshold, int method, int borderType)
{
	if (dst.empty()) dst.create(src.size(), src.type());                           // binalyWeightedRangeFilter.cpp:1108
	dmc_image a = dmc_dropin::wrap(src), b = dmc_dropin::wrap(dst);
	dmc_dropin::check(dmc_bwrf(dmc_dropin::context(), &a, &b, kernelSize.width, kernelSize.height, threshold, method, borderType), "binalyWeightedRangeFilter");
}

// extra overload of the reference (binalyWeightedRangeFilter.cpp:1101), not in its header
inline void binalyWeightedRangeFilter(const Mat& src, Mat& dst, int kernelSize, float threshold, int method, int borderType)
{
	binalyWeightedRangeFilter(src, dst, Size(kernelSize, kernelSize), threshold, method, borderType);
}

inline PostFilterSet::PostFilterSet(){;}
inline PostFilterSet::~PostFilterSet(){;}

inline void PostFilterSet::filterDisp8U2Depth16U(Mat& src, Mat& dest, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method)
{
	Mat s = src;
	dest.create(s.size(), CV_16U);                                                 // bufff.convertTo(dest,CV_16U) postFilterSet.cpp:31
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_filter_disp8u_depth16u(dmc_dropin::context(), &a, &b, focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method), "PostFilterSet::filterDisp8U2Depth16U");
}

inline void PostFilterSet::filterDisp8U2Depth32F(Mat& src, Mat& dest, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method)
{
	Mat s = src;
	if (dest.empty()) dest.create(s.size(), CV_32F);                               // dst.create(bufff.size(), bufff.type()) via :42 -> :1108
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_filter_disp8u_depth32f(dmc_dropin::context(), &a, &b, focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method), "PostFilterSet::filterDisp8U2Depth32F");
}

inline void PostFilterSet::filterDisp8U2Disp32F(Mat& src, Mat& dest, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int brange_method)
{
	Mat s = src;
	dest.create(s.size(), CV_16U);                                                 // bufff.convertTo(dest,CV_16U) postFilterSet.cpp:54
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_filter_disp8u_disp32f(dmc_dropin::context(), &a, &b, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method), "PostFilterSet::filterDisp8U2Disp32F");
}

inline void PostFilterSet::operator()(Mat& src, Mat& dest, int median_r, int gaussian_r, int minmax_r, int brange_r, int brange_th, int brange_method)
{
	Mat s = src;
	if (dest.empty()) dest.create(s.size(), s.type());                             // postFilterSet.cpp:62 -> :1108
	dmc_image a = dmc_dropin::wrap(s), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_post_filter_set(dmc_dropin::context(), &a, &b, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method), "PostFilterSet::operator()");
}

inline void boundaryReconstructionFilter(Mat& src, Mat& dest, Size ksize, const float frec, const float color, const float space)
{
	const int t = src.type();
	if (t != CV_8U && t != CV_16S && t != CV_16U && t != CV_32F && t != CV_64F) return;   // boundaryReconstructionFilter.cpp:133-155
	if (dest.empty()) dest.create(src.size(), src.type());                         // :15
	dmc_image a = dmc_dropin::wrap(src), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_boundary_reconstruction(dmc_dropin::context(), &a, &b, ksize.width, ksize.height, frec, color, space), "boundaryReconstructionFilter");
}

#endif
