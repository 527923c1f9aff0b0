// dmc_kernels.cuh -- host-side launchers of the sm_100a kernels (implemented in dmc_kernels_*.cu).
// All images are dense device buffers; `n` frames of H x W (x cn) lie back to back and are processed by one
// launch (frame = blockIdx.z).  Launchers return the number of kernels they launched (0 = nothing to do).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmc {

constexpr int kMaxRadius = 10;
constexpr int kMaxTapsRow = 2 * kMaxRadius + 1;

struct GaussTaps { float kx[kMaxTapsRow]; float ky[kMaxTapsRow]; int rx, ry; };   // getGaussianKernel taps (host computed)
struct RowSpan { signed char hw[kMaxTapsRow]; int rV, rH, ntaps; };               // circular window: half-width per row

RowSpan make_rowspan(int kw, int kh);                    // binalyWeightedRangeFilter.cpp:1066-1076
bool make_gauss_taps(int d, double sigma, int rows, int cols, GaussTaps* t);   // cv::getGaussianKernel (OpenCV 4.x)

// cv::medianBlur, 8UC1, BORDER_REPLICATE (postFilterSet.cpp:23,36,47,59)
int launch_median8u(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s);
// smallGaussianBlur on 8UC1: 8U -> 32F -> GaussianBlur(REFLECT_101) -> RNE/saturate 8U (postFilterSet.cpp:4-16)
int launch_gauss8u(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s);
// blurRemoveMinMax_<T> (minmaxFilter.cpp:48-174); depth in {8U,16U,16S,32F,64F}, cn channels interleaved
int launch_minmax(const void* src, void* dst, int n, int H, int W, int depth, int cn, int r, cudaStream_t s);
// windowed max / min with out-of-image taps ignored (cv::dilate/erode; maxFilter/minFilter for integer types)
int launch_morph(const void* src, void* dst, int n, int H, int W, int depth, int kw, int kh, int is_max, cudaStream_t s);
// maxFilter/minFilter 32F with the reference's FLT_MIN / FLT_MAX seeds (minmaxFilter.cpp:256-414), row-serial
int launch_minmax_filter_f32_seeded(const float* src, float* dst, float* tmp, int H, int W, int kw, int kh, int is_max, cudaStream_t s);

// binalyWeightedRangeFilter_8u (binalyWeightedRangeFilter.cpp:1031-1082), C1 and C3
int launch_bwrf8u(const uint8_t* src, uint8_t* dst, int n, int H, int W, int cn, const RowSpan& rs, int th, cudaStream_t s);

// binalyWeightedRangeFilter_32f (:978-1029) with fused input / output conversions
enum LoadOp { LOAD_F32 = 0, LOAD_U16 = 1, LOAD_S16 = 2, LOAD_U8 = 3, LOAD_U8_DISP2DEPTH = 4 };
enum StoreOp { STORE_F32 = 0, STORE_U16 = 1, STORE_S16 = 2 };
int launch_bwrf32f(const void* src, void* dst, int n, int H, int W, int cn, const RowSpan& rs, float th,
                   int load_op, float maf, int store_op, cudaStream_t s);

// boundaryReconstructionFilter_<T> (boundaryReconstructionFilter.cpp:12-131)
int launch_brf(const void* src, void* dst, int H, int W, int depth, int kw, int kh, float frec, float color, float space, cudaStream_t s);
// extension (filter_ext.h): blurRemoveMinMax(minmax_r) fused into the tile staging of the boundary reconstruction filter; 8U / 16U / 16S
int launch_minmax_brf(const void* src, void* dst, int H, int W, int depth, int minmax_r, int kw, int kh, float frec, float color, float space, cudaStream_t s);

// converters (depthmapUtil.cpp:685-1014); kind: 0 disp8U2depth32F, 1 depth32F2disp8U, 2 depth16U2disp8U, 3 disp16S2depth16U
int launch_convert(int kind, const void* src, void* dst, long n, float fb, float a, float b, cudaStream_t s);
int launch_f32_to_u16(const float* src, uint16_t* dst, long n, cudaStream_t s);   // Mat::convertTo(CV_16U)
// fillOcclusion_<T> / Inv_ (depthmapUtil.cpp:548-636): `pristine` is an untouched copy of `img`
int launch_fill_occlusion(void* img, const void* pristine, int H, int W, int depth, double invalid, int inv, cudaStream_t s);
// reprojectXYZ_<T> (depthmapUtil.cpp:450-481): xtab[i] = running FP32 sum along the row, computed by the host
int launch_transpose(const void* src, void* dst, int H, int W, int elem_size, cudaStream_t s);   // cv::transpose, single channel
int launch_reproject(const void* depth, float* xyz, const float* xtab, int H, int W, int dtype, float fyinv, float ch, cudaStream_t s);

}  // namespace dmc

namespace dmc {
// Fast path of the 8UC1 range filter for square kernels (radius 1..6): packed half2 arithmetic, two pixels per
// instruction, register-tiled over rows.  Exact (all intermediate values are integers below 2048), taken when
// ntaps * th <= 2048; returns 0 if the configuration is not covered (the caller then uses the generic kernel).
int launch_bwrf8u_h2(const uint8_t* src, uint8_t* dst, int n, int H, int W, int radius, int th, int ntaps, cudaStream_t s);
}

namespace dmc {
// Packed-SIMD fast paths of the three 8-bit front stages (dmc_front8u.cu); each returns 0 when the configuration is
// not covered and the caller falls back to the generic kernel of dmc_kernels_8u.cu.
int launch_median8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s);          // r = 1, 2
int launch_gauss8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s);   // d = 3, 5
int launch_minmax8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s);          // r = 1..5
}

namespace dmc {
// Register-tiled fast path of the single-channel 32-bit range filter, square window radius 1..7 (0 = not covered).
int launch_bwrf32f_tiled(const void* src, void* dst, int n, int H, int W, int radius, float th, int load_op, float maf, int store_op, cudaStream_t s);
int launch_bwrf32f_c3_tiled(const void* src, void* dst, int n, int H, int W, int radius, float th, int load_op, int store_op, cudaStream_t s);
}

#include "dmc_jpeg_parse.h"
namespace dmc {
// Baseline grayscale JPEG decoding (dmc_jpeg.cu).  `scratch`: sum of jpeg_scratch_bytes over the frames (FrameDesc::ds_offset
// points into it); `coefs` (n * blocks * 64 int16) is only needed when n_restart > 0 frames carry restart intervals.
size_t jpeg_scratch_bytes(uint64_t scan_bytes);
int launch_jpeg_decode(const uint8_t* blob, const void* desc, const void* hts, const void* qts, uint8_t* scratch, int16_t* coefs, uint8_t* dst,
                       int n, int n_restart, int H, int W, cudaStream_t s);
}

namespace dmc {
// point-cloud render (dmc_render.cu)
int launch_project_points(const float* xyz, float* pt, long n, const float kr[9], const float t[3], int exact_divide, cudaStream_t s);
size_t render_scratch_bytes(int rows, int cols);
int launch_render(const uint8_t* image, const float* xyz, const float* pt, int rows, int cols, int is_sub, uint8_t* dest, float* depth,
                  void* scratch, int* changed_dev, int* changed_host, cudaStream_t s);
int launch_fill_small_hole(const uint8_t* src, uint8_t* dst, int rows, int cols, cudaStream_t s);
int launch_split_line_interleave(const void* src, void* dst, int rows, int cols, int elem, cudaStream_t s);
}

namespace dmc {
// Packed-half fast path of the 8UC3 range filter, square window radius 1..5, exact while ntaps*th <= 2048 (0 = not covered).
int launch_bwrf8u_c3_h2(const uint8_t* src, uint8_t* dst, int n, int H, int W, int radius, int th, int ntaps, cudaStream_t s);
// joint (guided) range filter: 8UC1 image averaged with weights from a guide of gcn = 1 or 3 channels (dmc_joint_bwrf.cu)
int launch_joint_bwrf8u(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int n, int H, int W, int gcn, const RowSpan& rs, int th, cudaStream_t s);
}
