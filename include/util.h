// util.h -- drop-in for the part of PostFilterSetForDepthCoding/util.h that lies on the post-filter path: the
// disparity <-> depth converters (util.h:25-28), fillOcclusion (util.h:24) and reprojectXYZ(depth, xyz, f) (util.h:11).
// Declarations are the reference's; the implementations forward to libdmc_b200.so (include/dmc_c.h).  The rest of
// the reference's util.h (GUI, point-cloud renderer, codec wrappers, timers) is out of scope.
#ifndef _UTIL_H_
#define _UTIL_H_

#include "filter.h"

//point cloud rendering
void reprojectXYZ(const Mat& depth, Mat& xyz, double f);

//oocclusion filling
enum
{
	FILL_DISPARITY =0,
	FILL_DEPTH =1
};
void fillOcclusion(Mat& src, int invalidvalue, int disp_or_depth=FILL_DEPTH);

//disparity depth converter
void depth32F2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void disp16S2depth16U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void depth16U2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void disp8U2depth32F(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);

// ---------------------------------------------------------------------------------------------------------------------
inline void reprojectXYZ(const Mat& depth, Mat& xyz, double f)
{
	const int t = depth.type();
	if (t != CV_8U && t != CV_16S && t != CV_16U && t != CV_32F) return;                    // depthmapUtil.cpp:483-501
	if (xyz.empty()) xyz = Mat::zeros(depth.size().area(), 1, CV_32FC3);                     // :453
	dmc_image a = dmc_dropin::wrap(depth), b = dmc_dropin::wrap(xyz);
	dmc_dropin::check(dmc_reproject_xyz(dmc_dropin::context(), &a, &b, f), "reprojectXYZ");
}

inline void fillOcclusion(Mat& src, int invalidvalue, int disp_or_depth)
{
	dmc_image a = dmc_dropin::wrap(src);
	dmc_dropin::check(dmc_fill_occlusion(dmc_dropin::context(), &a, invalidvalue, disp_or_depth), "fillOcclusion");
}

namespace dmc_dropin {
typedef int (*convert_fn)(dmc_ctx*, const dmc_image*, dmc_image*, float, float, float);
inline void convert(convert_fn fn, Mat& src, Mat& dest, int dtype, float fb, float a, float b, const char* what) {
	if (dest.empty()) dest = Mat::zeros(src.size(), dtype);                                  // depthmapUtil.cpp:925-926 etc.
	if (dest.type() != dtype) dest = Mat::zeros(src.size(), dtype);
	dmc_image s = wrap(src), d = wrap(dest);
	check(fn(context(), &s, &d, fb, a, b), what);
}
}  // namespace dmc_dropin

inline void depth32F2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a, float b) { dmc_dropin::convert(dmc_depth32f2disp8u, src, dest, CV_8U, focal_baseline, a, b, "depth32F2disp8U"); }
inline void disp16S2depth16U(Mat& src, Mat& dest, const float focal_baseline, float a, float b) { dmc_dropin::convert(dmc_disp16s2depth16u, src, dest, CV_16U, focal_baseline, a, b, "disp16S2depth16U"); }
inline void depth16U2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a, float b) { dmc_dropin::convert(dmc_depth16u2disp8u, src, dest, CV_8U, focal_baseline, a, b, "depth16U2disp8U"); }
inline void disp8U2depth32F(Mat& src, Mat& dest, const float focal_baseline, float a, float b) { dmc_dropin::convert(dmc_disp8u2depth32f, src, dest, CV_32F, focal_baseline, a, b, "disp8U2depth32F"); }

#endif
