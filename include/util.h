// util.h -- drop-in for the part of PostFilterSetForDepthCoding/util.h that lies on the post-filter path: the
// disparity <-> depth converters (util.h:25-28), fillOcclusion (util.h:24), reprojectXYZ(depth, xyz, f) (util.h:11) and the
// point-cloud renderer that consumes its output (projectImagefromXYZ util.h:12-13, fillSmallHole util.h:25,
// projectPointsSimple util.h:33).  Declarations are the reference's; the implementations forward to libdmc_b200.so
// (include/dmc_c.h).  The rest of the reference's util.h (GUI, camera helpers, codec wrappers, timers) is out of scope.
#ifndef _UTIL_H_
#define _UTIL_H_

#include "filter.h"

//point cloud rendering
void reprojectXYZ(const Mat& depth, Mat& xyz, double f);
void projectImagefromXYZ(const Mat& image, Mat& destimage, const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, const Mat& dist, Mat& mask, const bool isSub);
void projectImagefromXYZ(const Mat& image, Mat& destimage, const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, const Mat& dist, Mat& mask, const bool isSub, vector<Point2f>& pt, Mat& depth);

//oocclusion filling
enum
{
	FILL_DISPARITY =0,
	FILL_DEPTH =1
};
void fillOcclusion(Mat& src, int invalidvalue, int disp_or_depth=FILL_DEPTH);
void fillSmallHole(const Mat& src, Mat& dest);

//disparity depth converter
void depth32F2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void disp16S2depth16U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void depth16U2disp8U(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);
void disp8U2depth32F(Mat& src, Mat& dest, const float focal_baseline, float a=1.f, float b=0.f);

void projectPointsSimple(const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, vector<Point2f>& dest);//multi points projection

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

// ---- point-cloud render -----------------------------------------------------------------------------------------------
namespace dmc_dropin {
inline void camera(const Mat& R, const Mat& t, const Mat& K, double r[9], double tt[3], double k[9], const char* what) {
	if (K.type() != CV_64F || R.type() != CV_64F || t.type() != CV_64F) {                   // CV_Assert depthmapUtil.cpp:290-294
#ifdef DMC_MINICV_HPP
		throw cv::Exception((std::string(what) + ": only support 64F matrix type").c_str(), __FILE__, __LINE__);
#else
		CV_Error(cv::Error::StsAssert, std::string(what) + ": only support 64F matrix type");
#endif
	}
	for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) { r[3 * i + j] = R.at<double>(i, j); k[3 * i + j] = K.at<double>(i, j); } tt[i] = t.at<double>(i, 0); }
}
inline dmc_image wrap_points(vector<Point2f>& pt) {
	dmc_image im; im.data = (void*)&pt[0]; im.rows = (int)pt.size(); im.cols = 1; im.cvtype = DMC_MAKETYPE(DMC_32F, 2); im.step = 0; im.mem = DMC_MEM_HOST; return im;
}
}  // namespace dmc_dropin

inline void projectPointsSimple(const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, vector<Point2f>& dest)
{
	double r[9], tt[3], k[9]; dmc_dropin::camera(R, t, K, r, tt, k, "projectPointsSimple");
	const size_t n = (size_t)xyz.size().area();
	if (dest.size() < n) dest.resize(n);                                                      // (the reference writes through &dest[0] and expects the caller to have sized it)
	dmc_image a = dmc_dropin::wrap(xyz), b = dmc_dropin::wrap_points(dest); b.rows = (int)n;
	dmc_dropin::check(dmc_project_points(dmc_dropin::context(), &a, r, tt, k, &b, 0), "projectPointsSimple");
}

inline void projectImagefromXYZ(const Mat& image, Mat& destimage, const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, const Mat& dist, Mat& mask, const bool isSub, vector<Point2f>& pt, Mat& depth)
{
	if (destimage.empty()) destimage = Mat::zeros(Size(image.size()), image.type());        // depthmapUtil.cpp:287-288
	double r[9], tt[3], k[9]; dmc_dropin::camera(R, t, K, r, tt, k, "projectImagefromXYZ");
	const size_t n = (size_t)image.size().area();
	if (pt.size() < n) pt.resize(n);
	if (depth.empty() || depth.type() != CV_32F || depth.size() != image.size()) depth = Mat::zeros(image.size(), CV_32F);
	dmc_image a = dmc_dropin::wrap(image), d = dmc_dropin::wrap(destimage), x = dmc_dropin::wrap(xyz), z = dmc_dropin::wrap(depth), p = dmc_dropin::wrap_points(pt);
	p.rows = (int)n;
	dmc_dropin::check(dmc_project_image_from_xyz(dmc_dropin::context(), &a, &d, &x, r, tt, k, isSub ? 1 : 0, &z, &p, 0), "projectImagefromXYZ");
}

inline void projectImagefromXYZ(const Mat& image, Mat& destimage, const Mat& xyz, const Mat& R, const Mat& t, const Mat& K, const Mat& dist, Mat& mask, const bool isSub)
{
	if (destimage.empty()) destimage = Mat::zeros(Size(image.size()), image.type());
	double r[9], tt[3], k[9]; dmc_dropin::camera(R, t, K, r, tt, k, "projectImagefromXYZ");
	dmc_image a = dmc_dropin::wrap(image), d = dmc_dropin::wrap(destimage), x = dmc_dropin::wrap(xyz);
	dmc_dropin::check(dmc_project_image_from_xyz(dmc_dropin::context(), &a, &d, &x, r, tt, k, isSub ? 1 : 0, 0, 0, 0), "projectImagefromXYZ");
}

inline void fillSmallHole(const Mat& src, Mat& dest)
{
	dmc_image a = dmc_dropin::wrap(src), b = dmc_dropin::wrap(dest);                         // in place (main.cpp:355) or into a pre-filled dest
	dmc_dropin::check(dmc_fill_small_hole(dmc_dropin::context(), &a, &b), "fillSmallHole");
}

#endif
