// filter_ext.h -- operators of libdmc_b200 that have NO counterpart in the reference's filter.h.  They live in their
// own header so that include/filter.h stays a truthful drop-in of the reference API.
#pragma once
#include "filter.h"

// Joint ("colour-guided") binary-weighted range filter: `src` (CV_8UC1 depth / disparity) is averaged over the window
// of binalyWeightedRangeFilter with the binary weights computed on `guide` (CV_8UC3: the saturated L1 colour distance of
// binalyWeightedRangeFilter.cpp:297-301; CV_8UC1: absolute difference) instead of on src itself.  Window, border
// (REPLICATE), (uchar)threshold, FP32 division and rounding are those of the reference's 8UC1 filter; with guide == src
// the result equals binalyWeightedRangeFilter(src, ...).  Only FULL_KERNEL is defined.
inline void jointBinalyWeightedRangeFilter(const Mat& src, const Mat& guide, Mat& dst, Size kernelSize, float threshold, int method = FULL_KERNEL)
{
	if (dst.empty()) dst.create(src.size(), src.type());
	dmc_image a = dmc_dropin::wrap(src), g = dmc_dropin::wrap(guide), b = dmc_dropin::wrap(dst);
	dmc_dropin::check(dmc_joint_bwrf(dmc_dropin::context(), &a, &g, &b, kernelSize.width, kernelSize.height, threshold, method), "jointBinalyWeightedRangeFilter");
}

// Fused min-max -> boundary reconstruction: blurRemoveMinMax(src, tmp, r) followed by boundaryReconstructionFilter(tmp,
// dest, ksize, frec, color, space) in one kernel (the intermediate image lives in shared memory only; src is read once).
// Bit-identical to the two reference calls (minmaxFilter.cpp:48-174, boundaryReconstructionFilter.cpp:12-131).
// Single channel CV_8U / CV_16U / CV_16S.
inline void minmaxBoundaryReconstructionFilter(const Mat& src, Mat& dest, const int r, Size ksize, const float frec, const float color, const float space)
{
	if (dest.empty()) dest.create(src.size(), src.type());
	dmc_image a = dmc_dropin::wrap(src), b = dmc_dropin::wrap(dest);
	dmc_dropin::check(dmc_minmax_boundary_reconstruction(dmc_dropin::context(), &a, &b, r, ksize.width, ksize.height, frec, color, space), "minmaxBoundaryReconstructionFilter");
}
