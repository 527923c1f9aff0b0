/* dmc_oracle.c -- TEST INFRASTRUCTURE ONLY (see dmc_oracle.h for the rules and the parity status).
 *
 * Plain-C restatement of the reference's post filter set.  Paths in the comments are relative to
 * /root/reference/PostFilterSetForDepthCoding/.  Third-party OpenCV calls on the path (medianBlur,
 * GaussianBlur, dilate/erode, copyMakeBorder, convertTo; no version pinned by the reference, oracle of
 * record = cv2 4.13.0) are restated from their published semantics.
 *
 * Build: gcc -O2 -msse2 -mfpmath=sse -ffp-contract=off -fopenmp  (every float op is one IEEE RN op, in
 * source order; never -ffast-math / -mfma).
 */
#include "dmc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <xmmintrin.h>
#include <float.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CN(t) ((((t) >> 3) & 511) + 1)
#define DEPTH(t) ((t) & 7)
static size_t depth_size(int d) { static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return s[d & 7]; }

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
static inline int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do { if (p < 0) p = -p - 1 + 1; else p = len - 1 - (p - len) - 1; } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* cvRound == cvtss2si / cvtsd2si: round-half-even; NaN or out of int32 range -> 0x80000000 */
static inline int cvround_f(float v) { if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN; return (int)lrintf(v); }
static inline int cvround_d(double v) { if (!(v >= -2147483648.0 && v < 2147483648.0)) return INT_MIN; return (int)lrint(v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)((unsigned)v <= 255u ? v : v > 0 ? 255 : 0); }
static inline uint16_t sat_u16(int v) { return (uint16_t)((unsigned)v <= 65535u ? v : v > 0 ? 65535 : 0); }
static inline int16_t sat_s16(int v) { return (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v); }

int orc_convert_32f_to_16u(const float* src, uint16_t* dst, long n) {   /* Mat::convertTo(CV_16U) */
    for (long i = 0; i < n; i++) dst[i] = sat_u16(cvround_f(src[i]));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * cv::medianBlur(src, dst, ksize) for CV_8UC1 -- exact median of the ksize x ksize window,
 * BORDER_REPLICATE (call sites postFilterSet.cpp:23,36,47,59).  Huang sliding histogram.
 * ---------------------------------------------------------------------------------------------- */
int orc_median_blur_8u(const uint8_t* src, uint8_t* dst, int rows, int cols, int ksize) {
    if (ksize % 2 != 1) return -1;
    size_t n = (size_t)rows * cols;
    if (ksize <= 1) { if (dst != src) memmove(dst, src, n); return 0; }
    uint8_t* out = (uint8_t*)malloc(n);
    int r = ksize / 2, half = (ksize * ksize) / 2;
#pragma omp parallel for schedule(dynamic, 8)
    for (int y = 0; y < rows; y++) {
        int hist[256]; memset(hist, 0, sizeof hist);
        for (int dy = -r; dy <= r; dy++) { const uint8_t* s = src + (size_t)clampi(y + dy, 0, rows - 1) * cols;
            for (int dx = -r; dx <= r; dx++) hist[s[clampi(dx, 0, cols - 1)]]++; }
        for (int x = 0; x < cols; x++) {
            if (x > 0) for (int dy = -r; dy <= r; dy++) { const uint8_t* s = src + (size_t)clampi(y + dy, 0, rows - 1) * cols;
                hist[s[clampi(x - r - 1, 0, cols - 1)]]--; hist[s[clampi(x + r, 0, cols - 1)]]++; }
            int acc = 0, m = 0;
            for (; m < 256; m++) { acc += hist[m]; if (acc > half) break; }
            out[(size_t)y * cols + x] = (uint8_t)m;
        }
    }
    memcpy(dst, out, n); free(out);
    return 0;
}

static const float kSmallGaussianTab[5][9] = {      /* cv::getGaussianKernel: fixed taps for odd n <= 9 when sigma <= 0 (OpenCV 4.13 small_gaussian_tab) */
    {1.f}, {0.25f, 0.5f, 0.25f}, {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
    {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f},
    {0.015625f, 0.05078125f, 0.1171875f, 0.19921875f, 0.234375f, 0.19921875f, 0.1171875f, 0.05078125f, 0.015625f}};

/* cv::getGaussianKernel(n, sigma, CV_32F), OpenCV 4.x: taps and normalisation in double, then cast. */
int orc_gaussian_kernel32f(int n, double sigma, float* taps) {
    double t[64], sum = 0; if (n > 64 || n < 1) return -1;
    if (sigma <= 0 && (n & 1) && n <= 9) { for (int i = 0; i < n; i++) taps[i] = kSmallGaussianTab[n >> 1][i]; return 0; }
    double sx = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8, scale2x = -0.5 / (sx * sx);
    for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; t[i] = exp(scale2x * x * x); sum += t[i]; }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) taps[i] = (float)(t[i] * sum);
    return 0;
}

/* cv::GaussianBlur(src, dst, Size(d,d), sigma) on CV_32FC1, BORDER_REFLECT_101 (postFilterSet.cpp:14).
 * Scalar sepFilter2D order: rows then columns; d<=5: x0*k0 + sum (x[-i]+x[+i])*k_i on both passes;
 * d>=7: the row pass is a left-to-right running sum k0*x[-r] + k1*x[-r+1] + ..., the column pass
 * stays symmetric (SURVEY.md 8a "parity hazards"; equal to cv2 4.13 with setUseOptimized(False)). */
int orc_gaussian_blur_32f(const float* src, float* dst, int rows, int cols, int d, double sigma) {
    int kw = d, kh = d; size_t n = (size_t)rows * cols;
    if (rows == 1) kh = 1;
    if (cols == 1) kw = 1;
    if (kw <= 1 && kh <= 1) { if (dst != src) memmove(dst, src, n * sizeof(float)); return 0; }
    float kx[64], ky[64];
    if (orc_gaussian_kernel32f(kw, sigma, kx) || orc_gaussian_kernel32f(kh, sigma, ky)) return -1;
    int rx = kw / 2, ry = kh / 2;
    float* tmp = (float*)malloc(n * sizeof(float)); float* out = (float*)malloc(n * sizeof(float));
#pragma omp parallel for
    for (int y = 0; y < rows; y++) {
        const float* s = src + (size_t)y * cols; float* t = tmp + (size_t)y * cols;
        for (int x = 0; x < cols; x++) {
            float acc;
            if (kw <= 5) { acc = s[x] * kx[rx];
                for (int i = 1; i <= rx; i++) acc = acc + (s[reflect101(x - i, cols)] + s[reflect101(x + i, cols)]) * kx[rx + i]; }
            else { acc = kx[0] * s[reflect101(x - rx, cols)];
                for (int i = 1; i < kw; i++) acc = acc + kx[i] * s[reflect101(x - rx + i, cols)]; }
            t[x] = acc;
        }
    }
#pragma omp parallel for
    for (int y = 0; y < rows; y++) {
        float* o = out + (size_t)y * cols;
        for (int x = 0; x < cols; x++) {
            float acc = tmp[(size_t)y * cols + x] * ky[ry];
            for (int i = 1; i <= ry; i++)
                acc = acc + (tmp[(size_t)reflect101(y - i, rows) * cols + x] + tmp[(size_t)reflect101(y + i, rows) * cols + x]) * ky[ry + i];
            o[x] = acc;
        }
    }
    memcpy(dst, out, n * sizeof(float)); free(tmp); free(out);
    return 0;
}

/* smallGaussianBlur (postFilterSet.cpp:4-16) on CV_8UC1: 8U -> 32F -> GaussianBlur -> RNE+saturate 8U */
int orc_small_gaussian_8u(const uint8_t* src, uint8_t* dst, int rows, int cols, int d, double sigma) {
    size_t n = (size_t)rows * cols;
    if (d == 0) { if (dst != src) memmove(dst, src, n); return 0; }          /* :6-10 */
    float* f = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) f[i] = (float)src[i];                       /* :13 */
    int rc = orc_gaussian_blur_32f(f, f, rows, cols, d, sigma);              /* :14 */
    if (!rc) for (size_t i = 0; i < n; i++) dst[i] = sat_u8(cvround_f(f[i])); /* :15 */
    free(f); return rc;
}

/* cv::dilate / cv::erode with Mat::ones(kh,kw) and the default border: out-of-image taps are ignored
 * (minmaxFilter.cpp:56-58). */
#define MORPH_IMPL(NAME, T)                                                                                    \
    static void NAME(const T* src, T* dst, int rows, int cols, int kw, int kh, int is_max) {                   \
        int rx = kw / 2, ry = kh / 2; size_t n = (size_t)rows * cols; T* tmp = (T*)malloc(n * sizeof(T));      \
        T* out = (T*)malloc(n * sizeof(T));                                                                    \
        _Pragma("omp parallel for") for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {            \
            T m = src[(size_t)y * cols + x];                                                                   \
            for (int i = x - rx < 0 ? 0 : x - rx; i <= (x + rx > cols - 1 ? cols - 1 : x + rx); i++) {         \
                T v = src[(size_t)y * cols + i]; m = is_max ? (v > m ? v : m) : (v < m ? v : m); }             \
            tmp[(size_t)y * cols + x] = m; }                                                                   \
        _Pragma("omp parallel for") for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {            \
            T m = tmp[(size_t)y * cols + x];                                                                   \
            for (int i = y - ry < 0 ? 0 : y - ry; i <= (y + ry > rows - 1 ? rows - 1 : y + ry); i++) {         \
                T v = tmp[(size_t)i * cols + x]; m = is_max ? (v > m ? v : m) : (v < m ? v : m); }             \
            out[(size_t)y * cols + x] = m; }                                                                   \
        memcpy(dst, out, n * sizeof(T)); free(tmp); free(out);                                                 \
    }
MORPH_IMPL(morph_u8, uint8_t)
MORPH_IMPL(morph_u16, uint16_t)
MORPH_IMPL(morph_s16, int16_t)
MORPH_IMPL(morph_f32, float)
MORPH_IMPL(morph_f64, double)

int orc_morph(const void* src, void* dst, int rows, int cols, int depth, int kw, int kh, int is_max) {
    switch (depth) {
    case ORC_8U: morph_u8((const uint8_t*)src, (uint8_t*)dst, rows, cols, kw, kh, is_max); return 0;
    case ORC_16U: morph_u16((const uint16_t*)src, (uint16_t*)dst, rows, cols, kw, kh, is_max); return 0;
    case ORC_16S: morph_s16((const int16_t*)src, (int16_t*)dst, rows, cols, kw, kh, is_max); return 0;
    case ORC_32F: morph_f32((const float*)src, (float*)dst, rows, cols, kw, kh, is_max); return 0;
    case ORC_64F: morph_f64((const double*)src, (double*)dst, rows, cols, kw, kh, is_max); return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------------
 * blurRemoveMinMax_<T> (minmaxFilter.cpp:48-174): mx = dilate, mn = erode over (2r+1)^2; mind = |src-mn|,
 * maxd = |src-mx|, mask = min(mind,maxd); out = (mind == mask) ? mn : mx   (:83-89 / :162-173).
 * absdiff is cv::absdiff (saturating for 16S).  Multi-channel: per channel (:189-213).
 * ---------------------------------------------------------------------------------------------- */
#define BRM_IMPL(NAME, T, MORPH, ABSDIFF)                                                                      \
    static void NAME(const T* src, T* dst, int rows, int cols, int cn, int r) {                                \
        size_t n = (size_t)rows * cols; int k = 2 * r + 1;                                                     \
        T* pl = (T*)malloc(n * sizeof(T)); T* mx = (T*)malloc(n * sizeof(T)); T* mn = (T*)malloc(n * sizeof(T)); \
        for (int c = 0; c < cn; c++) {                                                                         \
            for (size_t i = 0; i < n; i++) pl[i] = src[i * cn + c];                                            \
            MORPH(pl, mx, rows, cols, k, k, 1); MORPH(pl, mn, rows, cols, k, k, 0);                            \
            for (size_t i = 0; i < n; i++) {                                                                   \
                T mind = ABSDIFF(pl[i], mn[i]), maxd = ABSDIFF(pl[i], mx[i]);                                  \
                T mask = maxd < mind ? maxd : mind;                                                            \
                dst[i * cn + c] = (mind == mask) ? mn[i] : mx[i]; } }                                          \
        free(pl); free(mx); free(mn);                                                                          \
    }
#define AD_U(a, b) ((a) > (b) ? (a) - (b) : (b) - (a))
static inline int16_t ad_s16(int16_t a, int16_t b) { int d = (int)a - (int)b; if (d < 0) d = -d; return sat_s16(d); }
static inline float ad_f32(float a, float b) { return fabsf(a - b); }
static inline double ad_f64(double a, double b) { return fabs(a - b); }
BRM_IMPL(brm_u8, uint8_t, morph_u8, AD_U)
BRM_IMPL(brm_u16, uint16_t, morph_u16, AD_U)
BRM_IMPL(brm_s16, int16_t, morph_s16, ad_s16)
BRM_IMPL(brm_f32, float, morph_f32, ad_f32)
BRM_IMPL(brm_f64, double, morph_f64, ad_f64)

int orc_blur_remove_minmax(const void* src, void* dst, int rows, int cols, int cvtype, int r) {
    int cn = CN(cvtype); size_t bytes = (size_t)rows * cols * cn * depth_size(DEPTH(cvtype));
    void* in = malloc(bytes); memcpy(in, src, bytes);          /* in-place safe, as the reference */
    int rc = 0;
    switch (DEPTH(cvtype)) {
    case ORC_8U: brm_u8((const uint8_t*)in, (uint8_t*)dst, rows, cols, cn, r); break;
    case ORC_16S: brm_s16((const int16_t*)in, (int16_t*)dst, rows, cols, cn, r); break;
    case ORC_16U: brm_u16((const uint16_t*)in, (uint16_t*)dst, rows, cols, cn, r); break;
    case ORC_32F: brm_f32((const float*)in, (float*)dst, rows, cols, cn, r); break;
    case ORC_64F: brm_f64((const double*)in, (double*)dst, rows, cols, cn, r); break;
    default: if (dst != src) memcpy(dst, in, bytes); rc = 1;   /* src.copyTo(dest) happened at :52, nothing else */
    }
    free(in); return rc;
}

/* ------------------------------------------------------------------------------------------------
 * maxFilter / minFilter (minmaxFilter.cpp:256-414): separable sliding max/min, BORDER_REPLICATE,
 * single channel 8U/16S/16U/32F.  Restated literally (including the seed values `maxval`: 0, SHRT_MIN,
 * 0, FLT_MIN for max -- :318-333 -- and 255, SHRT_MAX, USHRT_MAX, FLT_MAX for min -- :398-413), because
 * the float seed FLT_MIN leaks into the output wherever a window maximum is < FLT_MIN.
 * ---------------------------------------------------------------------------------------------- */
#define MMF_IMPL(NAME, T)                                                                                      \
    static void NAME##_sp(const T* src, T* dst, int rows, int cols, int width, T seed, int is_max) {          \
        if (width == 1) { memcpy(dst, src, (size_t)rows * cols * sizeof(T)); return; }                         \
        int rx = width / 2, st = width - 1, pc = cols + 2 * rx; T* sim = (T*)malloc((size_t)pc * sizeof(T));   \
        for (int i = 0; i < rows; i++) {                                                                       \
            const T* s = src + (size_t)i * cols; T* d = dst + (size_t)i * cols;                                \
            for (int x = 0; x < pc; x++) sim[x] = s[clampi(x - rx, 0, cols - 1)];                              \
            T prev = seed;                                                                                     \
            for (int k = 0; k < width; k++) prev = is_max ? (sim[k] > prev ? sim[k] : prev) : (sim[k] < prev ? sim[k] : prev); \
            d[0] = prev; T ed = sim[0];                                                                        \
            for (int j = 1; j < cols; j++) {                                                                   \
                if (is_max ? (prev <= sim[j + st]) : (prev >= sim[j + st])) { prev = sim[j + st]; d[j] = prev; } \
                else if (ed != prev) { d[j] = prev; ed = sim[j]; }                                             \
                else { T m = seed;                                                                             \
                    for (int k = 0; k < width; k++) m = is_max ? (sim[j + k] > m ? sim[j + k] : m) : (sim[j + k] < m ? sim[j + k] : m); \
                    d[j] = m; prev = m; ed = sim[j]; } } }                                                     \
        free(sim);                                                                                             \
    }                                                                                                          \
    static void NAME(const T* src, T* dst, int rows, int cols, int kw, int kh, T seed, int is_max) {          \
        size_t n = (size_t)rows * cols; T* a = (T*)malloc(n * sizeof(T)); T* b = (T*)malloc(n * sizeof(T));    \
        T* c = (T*)malloc(n * sizeof(T));                                                                      \
        NAME##_sp(src, a, rows, cols, kw, seed, is_max);                                                       \
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) b[(size_t)x * rows + y] = a[(size_t)y * cols + x]; \
        NAME##_sp(b, c, cols, rows, kh, seed, is_max);                                                         \
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) dst[(size_t)y * cols + x] = c[(size_t)x * rows + y]; \
        free(a); free(b); free(c);                                                                             \
    }
MMF_IMPL(mmf_u8, uint8_t)
MMF_IMPL(mmf_s16, int16_t)
MMF_IMPL(mmf_u16, uint16_t)
MMF_IMPL(mmf_f32, float)

static int minmax_filter(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh, int is_max) {
    switch (cvtype) {   /* `src.type()==CV_8U` etc.: single channel only; anything else is a silent no-op */
    case ORC_8U: mmf_u8((const uint8_t*)src, (uint8_t*)dst, rows, cols, kw, kh, is_max ? 0 : 255, is_max); return 0;
    case ORC_16S: mmf_s16((const int16_t*)src, (int16_t*)dst, rows, cols, kw, kh, is_max ? SHRT_MIN : SHRT_MAX, is_max); return 0;
    case ORC_16U: mmf_u16((const uint16_t*)src, (uint16_t*)dst, rows, cols, kw, kh, is_max ? 0 : USHRT_MAX, is_max); return 0;
    case ORC_32F: mmf_f32((const float*)src, (float*)dst, rows, cols, kw, kh, is_max ? FLT_MIN : FLT_MAX, is_max); return 0;
    }
    return 1;
}
int orc_max_filter(const void* s, void* d, int rows, int cols, int t, int kw, int kh) { return minmax_filter(s, d, rows, cols, t, kw, kh, 1); }
int orc_min_filter(const void* s, void* d, int rows, int cols, int t, int kw, int kh) { return minmax_filter(s, d, rows, cols, t, kw, kh, 0); }

/* ------------------------------------------------------------------------------------------------
 * Binary-weighted range filter.  Tap list: raster order over i in [-rV,rV], j in [-rH,rH], kept iff
 * sqrt(i^2+j^2) <= max(rV,rH)  (binalyWeightedRangeFilter.cpp:1066-1076 / :1013-1023).
 * Border: copyMakeBorder(BORDER_REPLICATE) (:1053 / :1000)  == clamped coordinates.
 * ---------------------------------------------------------------------------------------------- */
static int make_taps(int rH, int rV, int** di, int** dj) {
    int n = 0, cap = (2 * rH + 1) * (2 * rV + 1), rmax = rV > rH ? rV : rH;
    *di = (int*)malloc(cap * sizeof(int)); *dj = (int*)malloc(cap * sizeof(int));
    for (int i = -rV; i <= rV; i++) for (int j = -rH; j <= rH; j++) {
        double r = sqrt((double)i * i + (double)j * j);
        if (r > rmax) continue;
        (*di)[n] = i; (*dj)[n] = j; n++;
    }
    return n;
}

/* 8u kernels: SSE4.1 invoker, C1 :131-236 and C3 :237-462.  w = (|v-c| <= th) with unsigned-saturating
 * byte arithmetic (:169; C3: saturating sum of the three |d| :297-301); t += w*v and W += w in FP32
 * (:180-210); out = pack_saturate(cvtps_epi32(t / W)) (:212-216). */
static int bwrf_8u(const uint8_t* src, uint8_t* dst, int rows, int cols, int cn, int kw, int kh, uint8_t th) {
    size_t n = (size_t)rows * cols * cn;
    if (kw == 0 || kh == 0) { if (dst != src) memmove(dst, src, n); return 0; }   /* :1033 */
    if (cn != 1 && cn != 3) return -2;                                             /* CV_Assert :1038 */
    int rH = kw >> 1, rV = kh >> 1, *di, *dj, maxk = make_taps(rH, rV, &di, &dj);
    uint8_t* out = (uint8_t*)malloc(n);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        const uint8_t* c0 = src + ((size_t)y * cols + x) * cn;
        float t[3] = {0.f, 0.f, 0.f}, W = 0.f;
        for (int k = 0; k < maxk; k++) {
            const uint8_t* v = src + ((size_t)clampi(y + di[k], 0, rows - 1) * cols + clampi(x + dj[k], 0, cols - 1)) * cn;
            int d = 0;
            for (int c = 0; c < cn; c++) { d += abs((int)v[c] - (int)c0[c]); if (d > 255) d = 255; }
            float w = d <= th ? 1.f : 0.f;
            for (int c = 0; c < cn; c++) t[c] = t[c] + w * (float)v[c];
            W = W + w;
        }
        for (int c = 0; c < cn; c++) out[((size_t)y * cols + x) * cn + c] = sat_u8(sat_s16(cvround_f(t[c] / W)));
    }
    memcpy(dst, out, n); free(out); free(di); free(dj);
    return 0;
}

/* Joint (guided) range filter -- NOT in the reference (include/dmc_c.h dmc_joint_bwrf): bwrf_8u above with
 * the weight computed on `guide` (gcn = 1 or 3 channels) and the average taken over the single-channel `src`.  Every rule
 * (taps, border, saturated L1 distance, FP32 sums, division, rounding) is that of bwrf_8u, so guide == src reproduces it. */
int orc_joint_bwrf(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int rows, int cols, int gcn, int kw, int kh, float threshold) {
    size_t n = (size_t)rows * cols; uint8_t th = (uint8_t)(int)threshold;
    if (kw == 0 || kh == 0) { if (dst != src) memmove(dst, src, n); return 0; }
    if (gcn != 1 && gcn != 3) return -2;
    int rH = kw >> 1, rV = kh >> 1, *di, *dj, maxk = make_taps(rH, rV, &di, &dj);
    uint8_t* out = (uint8_t*)malloc(n);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        const uint8_t* g0 = guide + ((size_t)y * cols + x) * gcn;
        float t = 0.f, W = 0.f;
        for (int k = 0; k < maxk; k++) {
            size_t q = (size_t)clampi(y + di[k], 0, rows - 1) * cols + clampi(x + dj[k], 0, cols - 1);
            int d = 0;
            for (int c = 0; c < gcn; c++) { d += abs((int)guide[q * gcn + c] - (int)g0[c]); if (d > 255) d = 255; }
            float w = d <= th ? 1.f : 0.f;
            t = t + w * (float)src[q]; W = W + w;
        }
        out[(size_t)y * cols + x] = sat_u8(sat_s16(cvround_f(t / W)));
    }
    memcpy(dst, out, n); free(out); free(di); free(dj);
    return 0;
}

/* 32f kernels: SSE4.1 invoker, C1 :491-550 and C3 :551-656.  w = (|c-v| <= th) ? 1.f : 0.f (:521-523;
 * C3: (|d2|+|d1|)+|d0| :595-600); t += w*v (a multiply: 0*inf = NaN propagates, :525-526); W += w; t/W.
 *
 * Padding quirk (:993-997): dpad=(4-cols%4)%4; spad=dpad+(4-(2*rH)%16)%4 (C remainder, can be -2);
 * if(spad<4) spad+=4; lpad=4*(rH/4+1)-rH; rpad=spad-lpad.  For rH%8==5 and cols%4==0 this gives rpad=-1:
 * the padded row is ONE element too short on the right, so the tap (i, j=+rH) of the last column reads
 * the first element of the next line of `temp` instead of the replicated border.  C1: next line = padded
 * row y+i+1, column 0 (= src[clamp(y+i+1)][0]).  C3 (planes B,G,R line-interleaved, split.cpp:103-166):
 * plane c<2 reads plane c+1 of the same row, plane 2 reads plane 0 of the next row.  When that next line
 * does not exist (rV==0, last row) the reference reads past the buffer: undefined; we keep the replicated
 * value there and the tests exclude that single pixel. */
static int bwrf_32f(const float* src, float* dst, int rows, int cols, int cn, int kw, int kh, float th) {
    size_t n = (size_t)rows * cols * cn;
    if (kw == 0 || kh == 0) { if (dst != src) memmove(dst, src, n * sizeof(float)); return 0; }  /* :980 */
    if (cn != 1 && cn != 3) return -2;                                                            /* :985 */
    int rH = kw >> 1, rV = kh >> 1, *di, *dj, maxk = make_taps(rH, rV, &di, &dj);
    const int quirk = (rH % 8 == 5) && (cols % 4 == 0);
    float* out = (float*)malloc(n * sizeof(float));
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        const float* c0 = src + ((size_t)y * cols + x) * cn;
        float t[3] = {0.f, 0.f, 0.f}, W = 0.f;
        for (int k = 0; k < maxk; k++) {
            const float* vp = src + ((size_t)clampi(y + di[k], 0, rows - 1) * cols + clampi(x + dj[k], 0, cols - 1)) * cn;
            float v[3]; for (int c = 0; c < cn; c++) v[c] = vp[c];
            if (quirk && x == cols - 1 && dj[k] == rH) {
                int prow = y + di[k] + rV;                         /* padded-row index of this tap */
                const float* same = src + ((size_t)clampi(y + di[k], 0, rows - 1) * cols) * cn;       /* column 0, same row */
                int has_next = prow + 1 < rows + 2 * rV;
                const float* next = src + ((size_t)clampi(y + di[k] + 1, 0, rows - 1) * cols) * cn;   /* column 0, next row */
                if (cn == 1) { if (has_next) v[0] = next[0]; }
                else { v[0] = same[1]; v[1] = same[2]; if (has_next) v[2] = next[0]; }
            }
            float d;
            if (cn == 1) d = fabsf(c0[0] - v[0]);
            else d = (fabsf(c0[2] - v[2]) + fabsf(c0[1] - v[1])) + fabsf(c0[0] - v[0]);
            float w = d <= th ? 1.f : 0.f;
            for (int c = 0; c < cn; c++) t[c] = t[c] + w * v[c];
            W = W + w;
        }
        for (int c = 0; c < cn; c++) out[((size_t)y * cols + x) * cn + c] = t[c] / W;
    }
    memcpy(dst, out, n * sizeof(float)); free(out); free(di); free(dj);
    return 0;
}

/* Dispatcher binalyWeightedRangeFilter (:1106-1178).  dst must hold the caller's previous contents: the
 * unsupported (type, method) pairs leave it untouched (return 1).
 * FULL_KERNEL_PAIR (:1138-1166) has no deterministic reference output (racy scatter :772-776, unwritten
 * tail columns :693): 8U is a no-op as in the reference, 16S/16U/32F compute the
 * FULL_KERNEL result it approximates -- NO bit-parity claim for that method. */
int orc_bwrf(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh, float threshold, int method) {
    int depth = DEPTH(cvtype), cn = CN(cvtype); size_t n = (size_t)rows * cols * cn;
    if (method == ORC_FULL_KERNEL || method == ORC_FULL_KERNEL_PAIR) {
        if (depth == ORC_8U) { if (method == ORC_FULL_KERNEL_PAIR) return 1;
            return bwrf_8u((const uint8_t*)src, (uint8_t*)dst, rows, cols, cn, kw, kh, (uint8_t)(int)threshold); /* :1113 */ }
        if (depth == ORC_16S || depth == ORC_16U) {                                  /* :1115-1132 */
            float* f = (float*)malloc(n * sizeof(float));
            if (depth == ORC_16S) for (size_t i = 0; i < n; i++) f[i] = (float)((const int16_t*)src)[i];
            else for (size_t i = 0; i < n; i++) f[i] = (float)((const uint16_t*)src)[i];
            int rc = bwrf_32f(f, f, rows, cols, cn, kw, kh, threshold);
            if (!rc) { if (depth == ORC_16S) for (size_t i = 0; i < n; i++) ((int16_t*)dst)[i] = sat_s16(cvround_f(f[i]));
                       else for (size_t i = 0; i < n; i++) ((uint16_t*)dst)[i] = sat_u16(cvround_f(f[i])); }
            free(f); return rc; }
        if (depth == ORC_32F) return bwrf_32f((const float*)src, (float*)dst, rows, cols, cn, kw, kh, threshold);
        return 1;
    }
    if (method == ORC_SEPARABLE_KERNEL) {                                            /* :1084-1099, :1167-1177 */
        if (depth == ORC_8U) { uint8_t th = (uint8_t)(int)threshold;
            if (kw <= 1) { if (dst != src) memmove(dst, src, n); return 0; }          /* both guards test .width */
            int rc = bwrf_8u((const uint8_t*)src, (uint8_t*)dst, rows, cols, cn, kw, 1, th);
            return rc ? rc : bwrf_8u((const uint8_t*)dst, (uint8_t*)dst, rows, cols, cn, 1, kh, th); }
        if (depth == ORC_32F) {
            if (kw <= 1) { if (dst != src) memmove(dst, src, n * sizeof(float)); return 0; }
            int rc = bwrf_32f((const float*)src, (float*)dst, rows, cols, cn, kw, 1, threshold);
            return rc ? rc : bwrf_32f((const float*)dst, (float*)dst, rows, cols, cn, 1, kh, threshold); }
        return 1;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * boundaryReconstructionFilter_<T> (boundaryReconstructionFilter.cpp:12-131), single channel only.
 * ---------------------------------------------------------------------------------------------- */
#define BRF_IMPL(NAME, T, SUBEXPR, CASTT)                                                                      \
    static void NAME(const T* src, T* dst, int rows, int cols, int kw, int kh, float frec, float color, float space) { \
        int rw = kw / 2, rh = kh / 2, maxk = 0, cap = kw * kh > 0 ? (2 * rw + 1) * (2 * rh + 1) : 1;           \
        int* di = (int*)malloc(cap * sizeof(int)); int* dj = (int*)malloc(cap * sizeof(int));                  \
        float* sd = (float*)malloc(cap * sizeof(float));                                                       \
        for (int i = -rh; i <= rh; i++) for (int j = -rw; j <= rw; j++) {              /* :26-38 */            \
            double r = sqrt((double)i * i + (double)j * j); if (r > rw) continue;                              \
            sd[maxk] = (float)r; di[maxk] = i; dj[maxk] = j; maxk++; }                                         \
        size_t n = (size_t)rows * cols; T* out = (T*)malloc(n * sizeof(T));                                    \
        _Pragma("omp parallel")                                                                                \
        {                                                                                                      \
            T* val = (T*)malloc(cap * sizeof(T)); int* cnt = (int*)malloc(cap * sizeof(int));                  \
            float* dist = (float*)malloc(cap * sizeof(float)); float* sub = (float*)malloc(cap * sizeof(float)); \
            _Pragma("omp for schedule(dynamic, 4)")                                                            \
            for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {                                    \
                T val0 = src[(size_t)y * cols + x]; int nd = 0;                                                \
                for (int k = 0; k < maxk; k++) {                                        /* :54-78 */           \
                    T v = src[(size_t)reflect101(y + di[k], rows) * cols + reflect101(x + dj[k], cols)];       \
                    int f = 1;                                                                                 \
                    for (int q = 0; q < nd; q++) if (v == val[q]) { f = 0; cnt[q]++; dist[q] = dist[q] + sd[k]; break; } \
                    if (f) { val[nd] = v; cnt[nd] = 1; dist[nd] = sd[k]; nd++; } }                             \
                if (nd == 1) { out[(size_t)y * cols + x] = val[0]; continue; }           /* :80-84 */          \
                float maxDis = 0.f, minDis = FLT_MAX; int maxOcc = 0, minOcc = maxk; T maxDiff = 0, minDiff = (T)255; \
                for (int q = 0; q < nd; q++) {                                          /* :93-103 */          \
                    dist[q] = (float)(dist[q] / (double)cnt[q]);                                               \
                    sub[q] = SUBEXPR(val[q], val0);                                                            \
                    if (dist[q] > maxDis) maxDis = dist[q];                                                    \
                    if (dist[q] < minDis) minDis = dist[q];                                                    \
                    if (cnt[q] > maxOcc) maxOcc = cnt[q];                                                      \
                    if (cnt[q] < minOcc) minOcc = cnt[q];                                                      \
                    T s = CASTT(fabsf(sub[q]));                                                                \
                    if (s > maxDiff) maxDiff = s;                                                              \
                    if (s < minDiff) minDiff = s; }                                                            \
                float divOcc = (maxOcc == minOcc) ? 0.00000001f : 1.0f / (float)(maxOcc - minOcc);             \
                float divDiff = (maxDiff == minDiff) ? 0.00000001f : 1.0f / (float)(maxDiff - minDiff);        \
                float divDis = (maxDis == minDis) ? 0.00000001f : 1.0f / (float)(maxDis - minDis);             \
                float maxE = 0.f; T mind = val0;                                                               \
                for (int q = 0; q < nd; q++) {                                          /* :113-125 */         \
                    float J = frec * (float)(cnt[q] - minOcc) * divOcc;                                        \
                    J = J + color * ((float)maxDiff - sub[q]) * divDiff;                                       \
                    J = J + space * (maxDis - dist[q]) * divDis;                                               \
                    if (J > maxE) { maxE = J; mind = val[q]; } }                                               \
                out[(size_t)y * cols + x] = mind;                                                              \
            }                                                                                                  \
            free(val); free(cnt); free(dist); free(sub);                                                       \
        }                                                                                                      \
        memcpy(dst, out, n * sizeof(T)); free(out); free(di); free(dj); free(sd);                              \
    }
#define SUB_INT(v, v0) ((float)abs((int)(v) - (int)(v0)))
#define SUB_F32(v, v0) ((float)fabsf((v) - (v0)))
#define SUB_F64(v, v0) ((float)fabs((v) - (v0)))
#define CAST_U8(f) ((uint8_t)(int)(f))
#define CAST_S16(f) ((int16_t)(int)(f))
#define CAST_U16(f) ((uint16_t)(int)(f))
#define CAST_F32(f) ((float)(f))
#define CAST_F64(f) ((double)(f))
BRF_IMPL(brf_u8, uint8_t, SUB_INT, CAST_U8)
BRF_IMPL(brf_s16, int16_t, SUB_INT, CAST_S16)
BRF_IMPL(brf_u16, uint16_t, SUB_INT, CAST_U16)
BRF_IMPL(brf_f32, float, SUB_F32, CAST_F32)
BRF_IMPL(brf_f64, double, SUB_F64, CAST_F64)

int orc_brf(const void* src, void* dst, int rows, int cols, int cvtype, int kw, int kh, float frec, float color, float space) {
    switch (cvtype) {   /* dispatcher :133-155 matches single-channel types only */
    case ORC_8U: brf_u8((const uint8_t*)src, (uint8_t*)dst, rows, cols, kw, kh, frec, color, space); return 0;
    case ORC_16S: brf_s16((const int16_t*)src, (int16_t*)dst, rows, cols, kw, kh, frec, color, space); return 0;
    case ORC_16U: brf_u16((const uint16_t*)src, (uint16_t*)dst, rows, cols, kw, kh, frec, color, space); return 0;
    case ORC_32F: brf_f32((const float*)src, (float*)dst, rows, cols, kw, kh, frec, color, space); return 0;
    case ORC_64F: brf_f64((const double*)src, (double*)dst, rows, cols, kw, kh, frec, color, space); return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Disparity <-> depth converters (depthmapUtil.cpp).  The first (area/16)*16 elements go through the
 * SSE body (divide, [add b], cvtps_epi32, packs_epi32 [, packus_epi16]); the last area%16 through the
 * scalar tail, which truncates instead of saturating.
 * ---------------------------------------------------------------------------------------------- */
int orc_disp8u2depth32f(const uint8_t* src, float* dst, int rows, int cols, float fb, float a, float b) {   /* :923-1014 */
    long n = (long)rows * cols, sse = (n / 16) * 16; const float maf = a * fb;
    if (b == 0.f) {
        for (long i = 0; i < sse; i++) dst[i] = maf / (float)src[i];                   /* :939-968 */
        for (long i = sse; i < n; i++) dst[i] = a * fb / (float)src[i] + b;            /* :1003-1008 */
    } else {
        /* SSE body commented out (:970-1001), so s/d are never advanced: the scalar tail (:1003-1008) converts
         * the FIRST area%16 elements and everything else in dest is left as it was. */
        for (long i = 0; i < n - sse; i++) dst[i] = a * fb / (float)src[i] + b;
    }
    return 0;
}
int orc_depth32f2disp8u(const float* src, uint8_t* dst, int rows, int cols, float fb, float a, float b) {   /* :768-838 */
    long n = (long)rows * cols, sse = (n / 16) * 16; const float maf = a * fb;
    for (long i = 0; i < sse; i++) { float v = maf / src[i]; if (b != 0.f) { v = v + b; }
        dst[i] = sat_u8(sat_s16(cvround_f(v))); }
    for (long i = sse; i < n; i++) dst[i] = (uint8_t)cvround_f(a * fb / src[i] + b);   /* :825-830 truncating */
    return 0;
}
int orc_depth16u2disp8u(const uint16_t* src, uint8_t* dst, int rows, int cols, float fb, float a, float b) { /* :840-921 */
    long n = (long)rows * cols, sse = (n / 16) * 16; const float maf = a * fb;
    for (long i = 0; i < sse; i++) { float v = maf / (float)(int16_t)src[i];           /* sign-extending load :858-859 */
        if (b != 0.f) { v = v + b; }
        dst[i] = sat_u8(sat_s16(cvround_f(v))); }
    for (long i = sse; i < n; i++) dst[i] = (uint8_t)cvround_f(a * fb / (float)src[i] + b);   /* :908-913 */
    return 0;
}
int orc_disp16s2depth16u(const int16_t* src, uint16_t* dst, int rows, int cols, float fb, float a, float b) { /* :685-765 */
    long n = (long)rows * cols, sse = (n / 16) * 16; const float maf = a * fb;
    for (long i = 0; i < sse; i++) { float v = maf / (float)src[i]; if (b != 0.f) v = v + b;
        dst[i] = (uint16_t)sat_s16(cvround_f(v)); }                                    /* packs_epi32 only :716-717 */
    for (long i = sse; i < n; i++) dst[i] = (uint16_t)cvround_f(a * fb / (float)src[i] + b);
    return 0;
}

/* fillOcclusion_<T> / fillOcclusionInv_<T> (depthmapUtil.cpp:548-636), in place, row-serial. */
#define FILL_IMPL(NAME, T)                                                                                     \
    static int NAME(T* data, int rows, int cols, T invalid, T edge, int inv) {                                 \
        const int MAX_LENGTH = inv ? cols : (int)(cols * 0.5);                                                 \
        for (int j = 0; j < rows; j++) {                                                                       \
            T* s = data + (size_t)j * cols;                                                                    \
            s[0] = edge; s[cols - 1] = edge;                                                                   \
            for (int i = 1; i < cols - 1; i++) {                                                               \
                if (s[i] == invalid) {                                                                         \
                    int t = i;                                                                                 \
                    do { t++; if (t > cols - 1) break; } while (s[t] == invalid);                              \
                    /* t == cols happens once a row was blanked (:572-577) or if invalid == edge: the reference  \
                     * then reads s[cols] = first element of the next row; past the last row it is undefined. */ \
                    if (t > cols - 1 && j == rows - 1) return -3;                                              \
                    const T dd = inv ? (s[i - 1] > s[t] ? s[i - 1] : s[t]) : (s[t] < s[i - 1] ? s[t] : s[i - 1]); \
                    if (t - i > MAX_LENGTH) { for (int n = 0; n < cols; n++) s[n] = invalid; }                 \
                    else { for (; i < t; i++) s[i] = dd; }                                                     \
                }                                                                                              \
            }                                                                                                  \
            s[0] = s[1]; s[cols - 1] = s[cols - 2];                                                            \
        }                                                                                                      \
        return 0;                                                                                              \
    }
FILL_IMPL(fill_u8, uint8_t)
FILL_IMPL(fill_s16, int16_t)
FILL_IMPL(fill_u16, uint16_t)
FILL_IMPL(fill_f32, float)

int orc_fill_occlusion(void* data, int rows, int cols, int cvtype, int invalid, int disp_or_depth) {       /* :643-683 */
    int inv = disp_or_depth == 1;   /* FILL_DEPTH = 1 -> Inv_ (max of neighbours, edge = 0) */
    switch (cvtype) {
    case ORC_8U: return fill_u8((uint8_t*)data, rows, cols, (uint8_t)invalid, inv ? 0 : 255, inv);
    case ORC_16S: return fill_s16((int16_t*)data, rows, cols, (int16_t)invalid, inv ? 0 : SHRT_MAX, inv);
    case ORC_16U: return fill_u16((uint16_t*)data, rows, cols, (uint16_t)invalid, inv ? 0 : USHRT_MAX, inv);
    case ORC_32F: return fill_f32((float*)data, rows, cols, (float)invalid, inv ? 0.f : FLT_MAX, inv);
    }
    return 1;
}

/* reprojectXYZ_<T>(depth, xyz, f) (depthmapUtil.cpp:450-481): x is a running FP32 sum along the row. */
int orc_reproject_xyz(const void* depth, float* xyz, int rows, int cols, int cvtype, double f) {
    const float bigZ = 10000.f, fxinv = (float)(1.0 / f), fyinv = (float)(1.0 / f);
    const float cw = (cols - 1) * 0.5f, ch = (rows - 1) * 0.5f;
    if (cvtype != ORC_8U && cvtype != ORC_16S && cvtype != ORC_16U && cvtype != ORC_32F) return 1;
    size_t p = 0;
    for (int j = 0; j < rows; j++) {
        float b = j - ch; const float y = b * fyinv; float x = (-cw) * fxinv;
        for (int i = 0; i < cols; i++, p++) {
            float z;
            switch (cvtype) { case ORC_8U: z = (float)((const uint8_t*)depth)[p]; break; case ORC_16S: z = (float)((const int16_t*)depth)[p]; break;
                case ORC_16U: z = (float)((const uint16_t*)depth)[p]; break; default: z = ((const float*)depth)[p]; }
            xyz[3 * p + 0] = x * z; xyz[3 * p + 1] = y * z; xyz[3 * p + 2] = (z == 0) ? bigZ : z;
            x = x + fxinv;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Point-cloud render: projectPointsSimple (depthmapUtil.cpp:10-156), projectImagefromXYZ
 * (:285-448), fillSmallHole (:187-283); call sites main.cpp:341-373.
 *
 * PARITY: bit-exact, pinned to the reference build (tests/test_pointcloud.py: port == oracle/_ref on several views,
 * with and without isSub).  Two things make that possible:
 *  (1) the reference projects with _mm_rcp_ps (:78), a 12-bit reciprocal look-up whose table differs between
 *      CPU vendors.  The port executes the same instruction (so it equals the reference build on whatever CPU
 *      it runs on); the CUDA kernel carries Intel's table (csrc/dmc_rcp_intel.inc, verified against the
 *      instruction on all 2^32 operands by tools/gen_rcp_table.c), so GPU == port holds on Intel hosts.
 *      rcp = 0 selects true division instead, like the reference's scalar twin myProjectPoint_BF (:99-146).
 *  (2) orc_project_image_serial is the literal restatement of the reference's z-buffer splat: points are
 *      visited in raster order, a point tries its neighbour pixels only if it won its own pixel AT THAT
 *      MOMENT, and two of the neighbour writes colour a different pixel than the one they z-test (:366-379,
 *      :409-422).  The CUDA implementation reproduces this order-dependent result exactly with a parallel
 *      fixed-point iteration (csrc/dmc_render.cu).
 * ---------------------------------------------------------------------------------------------- */
static int f2i_x86(float v) {                       /* (int)v as cvttss2si does it: out of range / NaN -> INT_MIN */
    if (!(v > -2147483904.f && v < 2147483648.f)) return (int)0x80000000u;
    return (int)v;
}
static int sub_wrap(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

static void kr_float(const double* R, const double* K, float r[3][3]) {      /* Mat kr = K*R; (float)kr(i,j)  :12-23 */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += K[3 * i + k] * R[3 * k + j]; r[i][j] = (float)s; }
}

/* rcp != 0: the reference's build (CV_SSE4_1): the first 4*(n/4) points take the SSE body, whose reciprocal is the CPU's
 * own RCPPS instruction (:78) -- executed here as the same instruction, so the port follows whatever CPU it runs on --
 * and the last n%4 points the scalar tail with a true division (:88-97).  rcp == 0: true division everywhere
 * (myProjectPoint_BF :99-146). */
int orc_project_points(const float* xyz, long n, const double* R, const double* t, const double* K, float* pt, int rcp) {
    float r[3][3]; kr_float(R, K, r);
    const float tt[3] = {(float)t[0], (float)t[1], (float)t[2]};
    const long n_sse = rcp ? 4 * (n / 4) : 0;
    for (long i = 0; i < n; i++) {                                             /* :58-97 / :131-145 */
        const float x = xyz[3 * i] + tt[0], y = xyz[3 * i + 1] + tt[1], z = xyz[3 * i + 2] + tt[2];
        const float den = r[2][0] * x + r[2][1] * y + r[2][2] * z;
        const float div = i < n_sse ? _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(den))) : 1.f / den;
        pt[2 * i] = (r[0][0] * x + r[0][1] * y + r[0][2] * z) * div;
        pt[2 * i + 1] = (r[1][0] * x + r[1][1] * y + r[1][2] * z) * div;
    }
    return 0;
}

/* Neighbour fragments of a point that landed on (x, y): which pixel is z-tested (dq) and which is coloured (dc),
 * as offsets (dy, dx), in the reference's code order (:358-450).  Returns the count. */
static int sub_fragments(const float* pt, long p, int cols, int x, int y, int dq[6][2], int dc[6][2]) {
    int n = 0;
#define FRAG(qy, qx, cy, cx) do { dq[n][0] = qy; dq[n][1] = qx; dc[n][0] = cy; dc[n][1] = cx; n++; } while (0)
    const int down = sub_wrap(f2i_x86(pt[2 * (p + cols) + 1]), y) > 1, right = sub_wrap(f2i_x86(pt[2 * (p + 1)]), x) > 1;
    if (down && right) { FRAG(0, 1, 0, 1); FRAG(1, 1, 1, 0); FRAG(1, 0, 1, 1); }          /* (the last two colour each other's pixel) */
    else if (right) FRAG(0, 1, 0, 1);
    else if (down) FRAG(1, 0, 1, 0);
    const int up = sub_wrap(f2i_x86(pt[2 * (p - cols) + 1]), y) < -1, left = sub_wrap(f2i_x86(pt[2 * (p - 1)]), x) < -1;
    if (up && left) { FRAG(0, -1, 0, -1); FRAG(-1, -1, -1, 0); FRAG(-1, 0, -1, -1); }
    else if (left) FRAG(0, -1, 0, -1);
    else if (up) FRAG(-1, 0, -1, 0);
#undef FRAG
    return n;
}

/* (2) literal, order-dependent */
int orc_project_image_serial(const uint8_t* image, const float* xyz, int rows, int cols, const double* R, const double* t, const double* K,
                             int is_sub, int rcp, uint8_t* dest, float* depth) {
    const long n = (long)rows * cols;
    float* pt = (float*)malloc(sizeof(float) * 2 * n);
    orc_project_points(xyz, n, R, t, K, pt, rcp);
    memset(dest, 0, (size_t)n * 3);
    for (long i = 0; i < n; i++) depth[i] = 10000.f;
    for (int j = 1; j < rows - 1; j++) for (int i = 1; i < cols - 1; i++) {
        const long p = (long)j * cols + i;
        const int x = f2i_x86(pt[2 * p]), y = f2i_x86(pt[2 * p + 1]);
        if (!(x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1)) continue;
        const float z = xyz[3 * p + 2];
        float* zb = depth + (long)y * cols + x;
        if (!(*zb > z)) continue;
        uint8_t* d = dest + ((long)y * cols + x) * 3; const uint8_t* c = image + 3 * p;
        d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; *zb = z;
        if (!is_sub) continue;
        int dq[6][2], dc[6][2]; const int nf = sub_fragments(pt, p, cols, x, y, dq, dc);
        for (int k = 0; k < nf; k++) {
            float* zq = zb + (long)dq[k][0] * cols + dq[k][1];
            if (*zq > z) { uint8_t* dd = d + ((long)dc[k][0] * cols + dc[k][1]) * 3; dd[0] = c[0]; dd[1] = c[1]; dd[2] = c[2]; *zq = z; }
        }
    }
    free(pt);
    return 0;
}

/* fillSmallHole (:187-283): a pixel whose GREEN is 0 becomes the mean of those 8 neighbours whose BLUE is not 0
 * (the reference tests s[lstep+1-1]); cvRound of a double quotient; border pixels are never written. */
int orc_fill_small_hole(const uint8_t* src, uint8_t* dst, int rows, int cols) {
    uint8_t* tmp = NULL;
    if (src == dst) { tmp = (uint8_t*)malloc((size_t)rows * cols * 3); memcpy(tmp, src, (size_t)rows * cols * 3); src = tmp; }
    const long step = (long)cols * 3;
    for (int j = 1; j < rows - 1; j++) for (int i = 1; i < cols - 1; i++) {
        const uint8_t* s = src + j * step + 3 * i; uint8_t* d = dst + j * step + 3 * i;
        if (s[1] != 0) continue;
        int count = 0, b = 0, g = 0, r = 0;
        for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            if (!dy && !dx) continue;
            const uint8_t* q = s + dy * step + 3 * dx;
            if (q[0] != 0) { b += q[0]; g += q[1]; r += q[2]; count++; }
        }
        d[0] = count ? (uint8_t)nearbyint((double)b / (double)count) : 0;
        d[1] = count ? (uint8_t)nearbyint((double)g / (double)count) : 0;
        d[2] = count ? (uint8_t)nearbyint((double)r / (double)count) : 0;
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * PostFilterSet (postFilterSet.cpp:21-63)
 * ---------------------------------------------------------------------------------------------- */
static int chain_front(const uint8_t* src, uint8_t* buff, int rows, int cols, int mr, int gr, int mmr) {
    int rc = orc_median_blur_8u(src, buff, rows, cols, 2 * mr + 1);                          /* :23 */
    if (!rc) rc = orc_small_gaussian_8u(buff, buff, rows, cols, 2 * gr + 1, gr + 0.5);       /* :24 */
    if (!rc) rc = orc_blur_remove_minmax(buff, buff, rows, cols, ORC_8U, mmr);               /* :25 */
    return rc;
}
int orc_post_filter_set(const uint8_t* src, uint8_t* dst, int rows, int cols, int mr, int gr, int mmr, int br, int th, int method) {
    size_t n = (size_t)rows * cols; uint8_t* buff = (uint8_t*)malloc(n);
    int rc = chain_front(src, buff, rows, cols, mr, gr, mmr);
    if (!rc) rc = orc_bwrf(buff, dst, rows, cols, ORC_8U, 2 * br + 1, 2 * br + 1, (float)th, method);   /* :62 */
    free(buff); return rc;
}
int orc_filter_disp8u_depth32f(const uint8_t* src, float* dst, int rows, int cols, double focus, double baseline, double amp,
                               int mr, int gr, int mmr, int br, float th, int method) {
    size_t n = (size_t)rows * cols; uint8_t* buff = (uint8_t*)malloc(n); float* bufff = (float*)calloc(n, sizeof(float));
    int rc = chain_front(src, buff, rows, cols, mr, gr, mmr);
    if (!rc) rc = orc_disp8u2depth32f(buff, bufff, rows, cols, (float)(focus * baseline), (float)amp, 0.f);   /* :40 */
    if (!rc) rc = orc_bwrf(bufff, dst, rows, cols, ORC_32F, 2 * br + 1, 2 * br + 1, th, method);             /* :42 */
    free(buff); free(bufff); return rc;
}
int orc_filter_disp8u_depth16u(const uint8_t* src, uint16_t* dst, int rows, int cols, double focus, double baseline, double amp,
                               int mr, int gr, int mmr, int br, float th, int method) {
    size_t n = (size_t)rows * cols; uint8_t* buff = (uint8_t*)malloc(n); float* bufff = (float*)calloc(n, sizeof(float));
    int rc = chain_front(src, buff, rows, cols, mr, gr, mmr);
    if (!rc) rc = orc_disp8u2depth32f(buff, bufff, rows, cols, (float)(focus * baseline), (float)amp, 0.f);   /* :27 */
    if (!rc) { rc = orc_bwrf(bufff, bufff, rows, cols, ORC_32F, 2 * br + 1, 2 * br + 1, th, method); if (rc == 1) rc = 0; }   /* :29 in place */
    if (!rc) orc_convert_32f_to_16u(bufff, dst, (long)n);                                                      /* :31 */
    free(buff); free(bufff); return rc;
}
int orc_filter_disp8u_disp32f(const uint8_t* src, uint16_t* dst, int rows, int cols, int mr, int gr, int mmr, int br, float th, int method) {
    size_t n = (size_t)rows * cols; uint8_t* buff = (uint8_t*)malloc(n); float* bufff = (float*)malloc(n * sizeof(float));
    int rc = chain_front(src, buff, rows, cols, mr, gr, mmr);
    if (!rc) { for (size_t i = 0; i < n; i++) bufff[i] = (float)buff[i];                                       /* :51 */
        rc = orc_bwrf(bufff, bufff, rows, cols, ORC_32F, 2 * br + 1, 2 * br + 1, th, method); if (rc == 1) rc = 0; }   /* :52 */
    if (!rc) orc_convert_32f_to_16u(bufff, dst, (long)n);                                                      /* :54 */
    free(buff); free(bufff); return rc;
}
