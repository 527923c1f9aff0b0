/* dmc_oracle.h -- TEST INFRASTRUCTURE ONLY: CPU restatement (plain C) of the reference's post-filter
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (depthmapcompression_b200, libdmc_b200.so) never does.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (tests/test_oracle.py) against
 *   (1) oracle/_ref/libdmc_ref.so = the unmodified reference sources compiled through oracle/refshim, and
 *   (2) the committed golden vectors in tests/golden/ generated from that build (tests/golden/make_golden.py),
 * and the third-party OpenCV stages (median, Gaussian, morphology, convertTo) against cv2 4.13.
 *
 * Images are continuous row-major; `cvtype` uses OpenCV's encoding depth + ((cn-1)<<3) with
 * depth 0=8U 2=16U 3=16S 5=32F 6=64F.  Return 0 = done, 1 = silent no-op (as the reference), <0 = error.
 */
#ifndef DMC_ORACLE_H
#define DMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FULL_KERNEL = 0, ORC_FULL_KERNEL_PAIR = 1, ORC_SEPARABLE_KERNEL = 2 };
enum { ORC_8U = 0, ORC_16U = 2, ORC_16S = 3, ORC_32F = 5, ORC_64F = 6 };

void orc_set_num_threads(int n);

/* third-party OpenCV stages on the path (call sites postFilterSet.cpp:23,36,47,59 and :4-16) */
int orc_median_blur_8u(const uint8_t* src, uint8_t* dst, int rows, int cols, int ksize);
int orc_gaussian_kernel32f(int n, double sigma, float* taps);
int orc_gaussian_blur_32f(const float* src, float* dst, int rows, int cols, int d, double sigma);
int orc_small_gaussian_8u(const uint8_t* src, uint8_t* dst, int rows, int cols, int d, double sigma);
int orc_morph(const void* src, void* dst, int rows, int cols, int depth, int kw, int kh, int is_max);

/* minmaxFilter.cpp */
int orc_blur_remove_minmax(const void* src, void* dst, int rows, int cols, int cvtype, int r);
int orc_max_filter(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh);
int orc_min_filter(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh);

/* binalyWeightedRangeFilter.cpp */
int orc_bwrf(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh, float threshold, int method);
/* extension without a reference counterpart: weights from `guide` (gcn channels), average of the 8UC1 `src` */
int orc_joint_bwrf(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int rows, int cols, int gcn, int kw, int kh, float threshold);

/* boundaryReconstructionFilter.cpp */
int orc_brf(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh, float frec, float color, float space);

/* depthmapUtil.cpp helpers on the path */
int orc_disp8u2depth32f(const uint8_t* src, float* dst, int rows, int cols, float focal_baseline, float a, float b);
int orc_depth32f2disp8u(const float* src, uint8_t* dst, int rows, int cols, float focal_baseline, float a, float b);
int orc_depth16u2disp8u(const uint16_t* src, uint8_t* dst, int rows, int cols, float focal_baseline, float a, float b);
int orc_disp16s2depth16u(const int16_t* src, uint16_t* dst, int rows, int cols, float focal_baseline, float a, float b);
int orc_fill_occlusion(void* data, int rows, int cols, int cvtype, int invalid, int disp_or_depth);
int orc_reproject_xyz(const void* depth, float* xyz, int rows, int cols, int cvtype, double f);
int orc_convert_32f_to_16u(const float* src, uint16_t* dst, long n);

/* point-cloud render (depthmapUtil.cpp:10-448); see the parity rule above these functions in dmc_oracle.c */
int orc_project_points(const float* xyz, long n, const double* R, const double* t, const double* K, float* pt, int rcp);
int orc_project_image_serial(const uint8_t* image, const float* xyz, int rows, int cols, const double* R, const double* t, const double* K, int is_sub, int rcp, uint8_t* dest, float* depth);
int orc_fill_small_hole(const uint8_t* src, uint8_t* dst, int rows, int cols);

/* postFilterSet.cpp */
int orc_post_filter_set(const uint8_t* src, uint8_t* dst, int rows, int cols, int median_r, int gaussian_r, int minmax_r, int brange_r, int brange_th, int method);
int orc_filter_disp8u_depth32f(const uint8_t* src, float* dst, int rows, int cols, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int method);
int orc_filter_disp8u_depth16u(const uint8_t* src, uint16_t* dst, int rows, int cols, double focus, double baseline, double amp, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int method);
int orc_filter_disp8u_disp32f(const uint8_t* src, uint16_t* dst, int rows, int cols, int median_r, int gaussian_r, int minmax_r, int brange_r, float brange_th, int method);

#ifdef __cplusplus
}
#endif
#endif
