// dmc_kernels_32f.cu -- the floating-point half of the path on sm_100a: the 32-bit binary-weighted range
// filter (with the disparity->depth / integer conversions of the PostFilterSet entry points fused into its
// tile load and store), the boundary reconstruction filter, the disparity<->depth converters, fillOcclusion
// and reprojectXYZ.  Every float operation is written with an explicit round-to-nearest intrinsic so that no
// FMA contraction can change a bit (SURVEY.md 8a "parity hazards").
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"

namespace dmc {

constexpr int kFX = 64, kFY = 16;   // output tile of the float range filter

template <int LOAD> __device__ __forceinline__ float load_as_f32(const void* p, size_t i, float maf);
template <> __device__ __forceinline__ float load_as_f32<LOAD_F32>(const void* p, size_t i, float) { return ((const float*)p)[i]; }
template <> __device__ __forceinline__ float load_as_f32<LOAD_U16>(const void* p, size_t i, float) { return (float)((const uint16_t*)p)[i]; }
template <> __device__ __forceinline__ float load_as_f32<LOAD_S16>(const void* p, size_t i, float) { return (float)((const int16_t*)p)[i]; }
template <> __device__ __forceinline__ float load_as_f32<LOAD_U8>(const void* p, size_t i, float) { return (float)((const uint8_t*)p)[i]; }
// disp8U2depth32F with b == 0 (depthmapUtil.cpp:935-968): depth = (a * focal_baseline) / disp, disp 0 -> +inf
template <> __device__ __forceinline__ float load_as_f32<LOAD_U8_DISP2DEPTH>(const void* p, size_t i, float maf) { return __fdiv_rn(maf, (float)((const uint8_t*)p)[i]); }

template <int STORE> __device__ __forceinline__ void store_from_f32(void* p, size_t i, float v);
template <> __device__ __forceinline__ void store_from_f32<STORE_F32>(void* p, size_t i, float v) { ((float*)p)[i] = v; }
template <> __device__ __forceinline__ void store_from_f32<STORE_U16>(void* p, size_t i, float v) { ((uint16_t*)p)[i] = sat_u16(cvround(v)); }   // convertTo(CV_16U)
template <> __device__ __forceinline__ void store_from_f32<STORE_S16>(void* p, size_t i, float v) { ((int16_t*)p)[i] = (int16_t)sat_s16(cvround(v)); }

// binalyWeightedRangeFilter_32f + BinalyWeightedRangeFilter_32f_InvokerSSE4 (binalyWeightedRangeFilter.cpp:471-663,
// :978-1029).  Taps in raster order, sequential FP32 accumulation  t += w*v, W += w  with w in {0.f, 1.f}; the
// product is a real multiply so that 0*inf = NaN propagates as in the SSE code (:525-526).
//
// `quirk`: for rH % 8 == 5 and cols % 4 == 0 the reference pads one column too few on the right (rpad = -1,
// :993-997), so the tap (i, +rH) of the LAST column reads the first element of the next line of its padded
// buffer.  That element is staged into the halo column W-1+rH here, which only that tap ever reads.
template <int CN, int LOAD, int STORE>
__global__ void __launch_bounds__(256) bwrf32f_kernel(const void* __restrict__ src, void* __restrict__ dst, int H, int W,
                                                      RowSpan rs, float th, float maf, int quirk) {
    extern __shared__ float smf[];
    const int rH = rs.rH, rV = rs.rV, TW = kFX + 2 * rH, TH = kFY + 2 * rV;
    const size_t fo = (size_t)blockIdx.z * H * W * CN;
    const int x0 = blockIdx.x * kFX, y0 = blockIdx.y * kFY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int idx = tid; idx < TW * TH; idx += 256) {
        int ty = idx / TW, tx = idx - ty * TW;
        int ux = x0 - rH + tx, uy = y0 - rV + ty;                  // unclamped (padded-buffer) coordinates
        int gx = clampi(ux, 0, W - 1), gy = clampi(uy, 0, H - 1);
        size_t base = fo + ((size_t)gy * W + gx) * CN;
#pragma unroll
        for (int c = 0; c < CN; c++) smf[idx * CN + c] = load_as_f32<LOAD>(src, base + c, maf);
        if (quirk && ux == W - 1 + rH) {
            bool has_next = uy + 1 <= H - 1 + rV;                  // next line of the padded buffer exists
            size_t same0 = fo + ((size_t)gy * W) * CN, next0 = fo + ((size_t)clampi(uy + 1, 0, H - 1) * W) * CN;
            if (CN == 1) { if (has_next) smf[idx] = load_as_f32<LOAD>(src, next0, maf); }
            else {
                smf[idx * CN + 0] = load_as_f32<LOAD>(src, same0 + 1, maf);
                smf[idx * CN + 1] = load_as_f32<LOAD>(src, same0 + 2, maf);
                if (has_next) smf[idx * CN + 2] = load_as_f32<LOAD>(src, next0, maf);
            }
        }
    }
    __syncthreads();
    for (int ly = threadIdx.y; ly < kFY; ly += blockDim.y)
        for (int lx = threadIdx.x; lx < kFX; lx += blockDim.x) {
            int x = x0 + lx, y = y0 + ly;
            if (x >= W || y >= H) continue;
            const float* pc = smf + ((ly + rV) * TW + lx + rH) * CN;
            float c0[CN], t[CN], wsum = 0.f;
#pragma unroll
            for (int c = 0; c < CN; c++) { c0[c] = pc[c]; t[c] = 0.f; }
            for (int i = -rV; i <= rV; i++) {
                int hw = rs.hw[i + rV];
                const float* pr = pc + i * TW * CN;
                for (int j = -hw; j <= hw; j++) {
                    float v[CN], d;
#pragma unroll
                    for (int c = 0; c < CN; c++) v[c] = pr[j * CN + c];
                    if (CN == 1) d = fabsf(__fsub_rn(c0[0], v[0]));
                    else d = __fadd_rn(__fadd_rn(fabsf(__fsub_rn(c0[CN - 1], v[CN - 1])), fabsf(__fsub_rn(c0[CN > 1 ? 1 : 0], v[CN > 1 ? 1 : 0]))),
                                       fabsf(__fsub_rn(c0[0], v[0])));
                    float w = d <= th ? 1.f : 0.f;
#pragma unroll
                    for (int c = 0; c < CN; c++) t[c] = __fmaf_rn(w, v[c], t[c]);      // exact product (w is 0 or 1): one rounding, as mul then add
                    wsum = __fadd_rn(wsum, w);
                }
            }
#pragma unroll
            for (int c = 0; c < CN; c++) store_from_f32<STORE>(dst, fo + ((size_t)y * W + x) * CN + c, __fdiv_rn(t[c], wsum));
        }
}

template <int CN, int LOAD>
static int launch_bwrf32f_ls(const void* src, void* dst, dim3 grid, size_t smem, int H, int W, const RowSpan& rs, float th, float maf, int quirk, int store_op, cudaStream_t s) {
    dim3 block(32, 8);
    switch (store_op) {
    case STORE_F32: bwrf32f_kernel<CN, LOAD, STORE_F32><<<grid, block, smem, s>>>(src, dst, H, W, rs, th, maf, quirk); return 1;
    case STORE_U16: bwrf32f_kernel<CN, LOAD, STORE_U16><<<grid, block, smem, s>>>(src, dst, H, W, rs, th, maf, quirk); return 1;
    case STORE_S16: bwrf32f_kernel<CN, LOAD, STORE_S16><<<grid, block, smem, s>>>(src, dst, H, W, rs, th, maf, quirk); return 1;
    }
    return 0;
}

int launch_bwrf32f(const void* src, void* dst, int n, int H, int W, int cn, const RowSpan& rs, float th,
                   int load_op, float maf, int store_op, cudaStream_t s) {
    if (cn == 1 && rs.rH == rs.rV && rs.rH >= 1 && rs.rH <= 7)
        if (int nk = launch_bwrf32f_tiled(src, dst, n, H, W, rs.rH, th, load_op, maf, store_op, s)) return nk;
    dim3 grid((W + kFX - 1) / kFX, (H + kFY - 1) / kFY, n);
    size_t smem = (size_t)(kFX + 2 * rs.rH) * (kFY + 2 * rs.rV) * cn * sizeof(float);
    int quirk = (rs.rH % 8 == 5) && (W % 4 == 0);
    if (cn == 1) {
        switch (load_op) {
        case LOAD_F32: return launch_bwrf32f_ls<1, LOAD_F32>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_U16: return launch_bwrf32f_ls<1, LOAD_U16>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_S16: return launch_bwrf32f_ls<1, LOAD_S16>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_U8: return launch_bwrf32f_ls<1, LOAD_U8>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_U8_DISP2DEPTH: return launch_bwrf32f_ls<1, LOAD_U8_DISP2DEPTH>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        }
    } else if (cn == 3) {
        switch (load_op) {
        case LOAD_F32: return launch_bwrf32f_ls<3, LOAD_F32>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_U16: return launch_bwrf32f_ls<3, LOAD_U16>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        case LOAD_S16: return launch_bwrf32f_ls<3, LOAD_S16>(src, dst, grid, smem, H, W, rs, th, maf, quirk, store_op, s);
        }
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// boundary reconstruction filter (boundaryReconstructionFilter.cpp:12-131)
// ----------------------------------------------------------------------------------------------------------
constexpr int kBrfMaxTaps = 320;    // circle of radius 10 has 317 taps
struct BrfTaps { int n; signed char di[kBrfMaxTaps], dj[kBrfMaxTaps]; float dist[kBrfMaxTaps]; };

template <typename T> struct BrfTraits;
template <> struct BrfTraits<uint8_t>  { static __device__ float sub(uint8_t a, uint8_t b) { return (float)abs((int)a - (int)b); } static __device__ uint8_t cast(float f) { return (uint8_t)(int)f; }
                                         static __device__ float rangef(uint8_t mx, uint8_t mn) { return (float)((int)mx - (int)mn); } };
template <> struct BrfTraits<int16_t>  { static __device__ float sub(int16_t a, int16_t b) { return (float)abs((int)a - (int)b); } static __device__ int16_t cast(float f) { return (int16_t)(int)f; }
                                         static __device__ float rangef(int16_t mx, int16_t mn) { return (float)((int)mx - (int)mn); } };
template <> struct BrfTraits<uint16_t> { static __device__ float sub(uint16_t a, uint16_t b) { return (float)abs((int)a - (int)b); } static __device__ uint16_t cast(float f) { return (uint16_t)(int)f; }
                                         static __device__ float rangef(uint16_t mx, uint16_t mn) { return (float)((int)mx - (int)mn); } };
template <> struct BrfTraits<float>    { static __device__ float sub(float a, float b) { return fabsf(__fsub_rn(a, b)); } static __device__ float cast(float f) { return f; }
                                         static __device__ float rangef(float mx, float mn) { return __fsub_rn(mx, mn); } };
template <> struct BrfTraits<double>   { static __device__ float sub(double a, double b) { return (float)fabs(__dsub_rn(a, b)); } static __device__ double cast(float f) { return (double)f; }
                                         static __device__ float rangef(double mx, double mn) { return (float)__dsub_rn(mx, mn); } };

constexpr int kRX = 32, kRY = 8;

template <typename T>
__global__ void __launch_bounds__(256) brf_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int rw, int rh,
                                                  BrfTaps taps, float frec, float color, float space) {
    extern __shared__ unsigned char smraw[];
    T* sm = (T*)smraw;
    const int TW = kRX + 2 * rw, TH = kRY + 2 * rh;
    const int x0 = blockIdx.x * kRX, y0 = blockIdx.y * kRY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int idx = tid; idx < TW * TH; idx += 256) {                // copyMakeBorder(BORDER_DEFAULT = REFLECT_101) :19
        int ty = idx / TW, tx = idx - ty * TW;
        sm[idx] = src[(size_t)reflect101(y0 - rh + ty, H) * W + reflect101(x0 - rw + tx, W)];
    }
    __syncthreads();
    int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const T* pc = sm + (threadIdx.y + rh) * TW + threadIdx.x + rw;
    const T val0 = pc[0];
    T val[kBrfMaxTaps]; short cnt[kBrfMaxTaps]; float dist[kBrfMaxTaps];
    // distinct values in first-encounter order :54-78.  Neighbouring taps usually repeat the previous value, so the entry of
    // the current run (index rq) lives in registers and is written back only when the value changes: the per-entry sums
    // still see their taps in tap order, one add per tap.
    int nd = 1, rq = 0, rcnt = 1; T rv = pc[taps.di[0] * TW + taps.dj[0]]; float rdist = taps.dist[0];
    val[0] = rv;
    for (int k = 1; k < taps.n; k++) {
        const T v = pc[taps.di[k] * TW + taps.dj[k]];
        if (v == rv) { rcnt++; rdist = __fadd_rn(rdist, taps.dist[k]); continue; }
        cnt[rq] = (short)rcnt; dist[rq] = rdist;                     // close the run
        int q = 0;
        for (; q < nd; q++) if (v == val[q]) break;
        if (q < nd) { rcnt = cnt[q] + 1; rdist = __fadd_rn(dist[q], taps.dist[k]); }
        else { val[nd] = v; nd++; rcnt = 1; rdist = taps.dist[k]; }
        rq = q; rv = v;
    }
    cnt[rq] = (short)rcnt; dist[rq] = rdist;
    if (nd == 1) { dst[(size_t)y * W + x] = val[0]; return; }      // :80-84
    float maxDis = 0.f, minDis = FLT_MAX; int maxOcc = 0, minOcc = taps.n; T maxDiff = (T)0, minDiff = (T)255;
    for (int q = 0; q < nd; q++) {                                  // :93-103
        // the reference divides in double and narrows (:96).  With a 24-bit dividend and a count < 2^9 the exact quotient is
        // either a float midpoint or at least 2^-20 ulp away from one, so the double rounding cannot change the result:
        // the correctly rounded FP32 quotient is the same number (and costs no FP64 division routine).
        float dq = __fdiv_rn(dist[q], (float)cnt[q]);
        dist[q] = dq;
        float sq = BrfTraits<T>::sub(val[q], val0);
        maxDis = fmaxf(dq, maxDis); minDis = fminf(dq, minDis);
        maxOcc = max((int)cnt[q], maxOcc); minOcc = min((int)cnt[q], minOcc);
        T s = BrfTraits<T>::cast(fabsf(sq));
        maxDiff = s > maxDiff ? s : maxDiff; minDiff = s < minDiff ? s : minDiff;
    }
    float divOcc = (maxOcc == minOcc) ? 0.00000001f : __fdiv_rn(1.0f, (float)(maxOcc - minOcc));
    float divDiff = (maxDiff == minDiff) ? 0.00000001f : __fdiv_rn(1.0f, BrfTraits<T>::rangef(maxDiff, minDiff));
    float divDis = (maxDis == minDis) ? 0.00000001f : __fdiv_rn(1.0f, __fsub_rn(maxDis, minDis));
    float maxE = 0.f; T mind = val0; const float fmaxDiff = (float)maxDiff;
    for (int q = 0; q < nd; q++) {                                  // :113-125
        float sq = BrfTraits<T>::sub(val[q], val0);
        float J = __fmul_rn(__fmul_rn(frec, (float)((int)cnt[q] - minOcc)), divOcc);
        J = __fadd_rn(J, __fmul_rn(__fmul_rn(color, __fsub_rn(fmaxDiff, sq)), divDiff));
        J = __fadd_rn(J, __fmul_rn(__fmul_rn(space, __fsub_rn(maxDis, dist[q])), divDis));
        if (J > maxE) { maxE = J; mind = val[q]; }
    }
    dst[(size_t)y * W + x] = mind;
}

// 8-bit fast path.  The divergent part of brf_kernel is the per-pixel search for "have I seen this value": the 32 lanes of
// a warp change value at different taps, so nearly every tap runs the scan.  Here the CTA first ranks the byte values
// that occur anywhere in its staged tile (256-bit presence set -> dense rank 0..M-1, a few dozen on decoded depth maps);
// a tap's value then indexes the thread's (count, distance) table directly -- no search.  The table lives in shared
// memory as [rank][thread] (conflict-free whatever the ranks are) when M <= kBrfCap, else in local memory.  order[]
// keeps the reference's first-encounter order for the scoring pass.
// measured on the Kinect fixture, 13x13 at 1080p: 32x4 threads / 64 ranks 0.62 ms, 32x4 / 32 0.71, 32x8 / 32 0.77, 32x2 / 128 0.78
// (the per-pixel scan this replaces: 1.15 ms)
constexpr int kBX = 32, kBY = 4, kBrfCap = 64;

#pragma nv_diag_suppress 549      // the local tables are zeroed for ranks 0..M-1 before use; the front end cannot see that
template <bool SMEM>
__device__ __forceinline__ void brf8u_pixel(const uint8_t* __restrict__ pc, int TW, const BrfTaps& taps, const uint8_t* __restrict__ lut,
                                            const uint8_t* __restrict__ vals, int M, uint8_t* __restrict__ cnt_s, float* __restrict__ dist_s,
                                            int tid, float frec, float color, float space, uint8_t* __restrict__ out) {
    constexpr int NT = kBX * kBY;
    short cnt_l[SMEM ? 1 : 256]; float dist_l[SMEM ? 1 : 256]; uint8_t order[256];
    auto get_cnt = [&](int r) -> int { if constexpr (SMEM) return (int)cnt_s[r * NT + tid]; else return (int)cnt_l[r]; };
    auto set_cnt = [&](int r, int c) { if constexpr (SMEM) cnt_s[r * NT + tid] = (uint8_t)c; else cnt_l[r] = (short)c; };
    auto get_dist = [&](int r) -> float { if constexpr (SMEM) return dist_s[r * NT + tid]; else return dist_l[r]; };
    auto set_dist = [&](int r, float d) { if constexpr (SMEM) dist_s[r * NT + tid] = d; else dist_l[r] = d; };
    for (int i = 0; i < M; i++) set_cnt(i, 0);
    const uint8_t val0 = pc[0];
    int nd = 0;
    for (int k = 0; k < taps.n; k++) {                              // :54-78, the same few instructions for every lane and tap
        const int r = lut[pc[taps.di[k] * TW + taps.dj[k]]];
        const int c = get_cnt(r);
        const float d = taps.dist[k];
        if (c == 0) order[nd++] = (uint8_t)r;
        set_dist(r, c == 0 ? d : __fadd_rn(get_dist(r), d));
        set_cnt(r, c + 1);
    }
    if (nd == 1) { *out = vals[order[0]]; return; }                 // :80-84
    float maxDis = 0.f, minDis = FLT_MAX; int maxOcc = 0, minOcc = taps.n; uint8_t maxDiff = 0, minDiff = 255;
    for (int q = 0; q < nd; q++) {                                  // :93-103
        const int r = order[q], c = get_cnt(r);
        const float dq = __fdiv_rn(get_dist(r), (float)c);          // == (float)((double)dist / cnt): see brf_kernel
        set_dist(r, dq);
        const float sq = (float)abs((int)vals[r] - (int)val0);
        maxDis = fmaxf(dq, maxDis); minDis = fminf(dq, minDis);
        maxOcc = max(c, maxOcc); minOcc = min(c, minOcc);
        const uint8_t sd = (uint8_t)(int)sq;
        maxDiff = sd > maxDiff ? sd : maxDiff; minDiff = sd < minDiff ? sd : minDiff;
    }
    const float divOcc = (maxOcc == minOcc) ? 0.00000001f : __fdiv_rn(1.0f, (float)(maxOcc - minOcc));
    const float divDiff = (maxDiff == minDiff) ? 0.00000001f : __fdiv_rn(1.0f, (float)((int)maxDiff - (int)minDiff));
    const float divDis = (maxDis == minDis) ? 0.00000001f : __fdiv_rn(1.0f, __fsub_rn(maxDis, minDis));
    float maxE = 0.f; uint8_t mind = val0; const float fmaxDiff = (float)maxDiff;
    for (int q = 0; q < nd; q++) {                                  // :113-125
        const int r = order[q];
        const float sq = (float)abs((int)vals[r] - (int)val0);
        float J = __fmul_rn(__fmul_rn(frec, (float)(get_cnt(r) - minOcc)), divOcc);
        J = __fadd_rn(J, __fmul_rn(__fmul_rn(color, __fsub_rn(fmaxDiff, sq)), divDiff));
        J = __fadd_rn(J, __fmul_rn(__fmul_rn(space, __fsub_rn(maxDis, get_dist(r))), divDis));
        if (J > maxE) { maxE = J; mind = vals[r]; }
    }
    *out = mind;
}

#pragma nv_diag_default 549

__global__ void __launch_bounds__(kBX * kBY) brf8u_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int rw, int rh,
                                                          BrfTaps taps, float frec, float color, float space) {
    constexpr int NT = kBX * kBY;
    extern __shared__ __align__(16) unsigned char smraw[];       // [dist table: kBrfCap*NT floats][count table: kBrfCap*NT bytes][tile]
    __shared__ uint32_t present[8];
    __shared__ uint8_t lut[256], vals[256];
    __shared__ int s_m;
    float* dist_s = (float*)smraw;
    uint8_t* cnt_s = smraw + (size_t)kBrfCap * NT * sizeof(float);
    uint8_t* sm = cnt_s + (size_t)kBrfCap * NT;
    const int TW = kBX + 2 * rw, TH = kBY + 2 * rh;
    const int x0 = blockIdx.x * kBX, y0 = blockIdx.y * kBY;
    const int tid = threadIdx.y * kBX + threadIdx.x;
    if (tid < 8) present[tid] = 0u;
    __syncthreads();
    for (int idx = tid; idx < TW * TH; idx += NT) {                 // copyMakeBorder(BORDER_DEFAULT = REFLECT_101) :19
        int ty = idx / TW, tx = idx - ty * TW;
        const uint8_t v = src[(size_t)reflect101(y0 - rh + ty, H) * W + reflect101(x0 - rw + tx, W)];
        sm[idx] = v;
        atomicOr(&present[v >> 5], 1u << (v & 31));
    }
    __syncthreads();
    for (int b = tid; b < 256; b += NT) {                           // rank of every byte value present in the tile
        int below = 0;
        for (int w = 0; w < (b >> 5); w++) below += __popc(present[w]);
        const uint32_t word = present[b >> 5];
        const int rank = below + __popc(word & ((1u << (b & 31)) - 1u));
        if ((word >> (b & 31)) & 1u) { lut[b] = (uint8_t)rank; vals[rank] = (uint8_t)b; }
        if (b == 255) s_m = rank + (int)((word >> 31) & 1u);
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int M = s_m;
    const uint8_t* pc = sm + (threadIdx.y + rh) * TW + threadIdx.x + rw;
    uint8_t* out = dst + (size_t)y * W + x;
    if (M <= kBrfCap && taps.n <= 255) brf8u_pixel<true>(pc, TW, taps, lut, vals, M, cnt_s, dist_s, tid, frec, color, space, out);
    else brf8u_pixel<false>(pc, TW, taps, lut, vals, M, cnt_s, dist_s, tid, frec, color, space, out);
}

template <typename T>
static int launch_brf_t(const void* src, void* dst, int H, int W, int rw, int rh, const BrfTaps& taps, float frec, float color, float space, cudaStream_t s) {
    dim3 grid((W + kRX - 1) / kRX, (H + kRY - 1) / kRY), block(kRX, kRY);
    size_t smem = (size_t)(kRX + 2 * rw) * (kRY + 2 * rh) * sizeof(T);
    brf_kernel<T><<<grid, block, smem, s>>>((const T*)src, (T*)dst, H, W, rw, rh, taps, frec, color, space);
    return 1;
}

int launch_brf(const void* src, void* dst, int H, int W, int depth, int kw, int kh, float frec, float color, float space, cudaStream_t s) {
    const int rw = kw / 2, rh = kh / 2;
    BrfTaps taps; taps.n = 0;
    for (int i = -rh; i <= rh; i++) for (int j = -rw; j <= rw; j++) {                      // :26-38
        double r = sqrt((double)i * i + (double)j * j);
        if (r > rw) continue;
        if (taps.n >= kBrfMaxTaps) return 0;
        taps.di[taps.n] = (signed char)i; taps.dj[taps.n] = (signed char)j; taps.dist[taps.n] = (float)r; taps.n++;
    }
    switch (depth) {
    case 0: {
        dim3 grid((W + kBX - 1) / kBX, (H + kBY - 1) / kBY), block(kBX, kBY);
        const size_t smem = (size_t)kBrfCap * kBX * kBY * 5 + (size_t)(kBX + 2 * rw) * (kBY + 2 * rh);
        brf8u_kernel<<<grid, block, smem, s>>>((const uint8_t*)src, (uint8_t*)dst, H, W, rw, rh, taps, frec, color, space);
        return 1;
    }
    case 2: return launch_brf_t<uint16_t>(src, dst, H, W, rw, rh, taps, frec, color, space, s);
    case 3: return launch_brf_t<int16_t>(src, dst, H, W, rw, rh, taps, frec, color, space, s);
    case 5: return launch_brf_t<float>(src, dst, H, W, rw, rh, taps, frec, color, space, s);
    case 6: return launch_brf_t<double>(src, dst, H, W, rw, rh, taps, frec, color, space, s);
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// converters (depthmapUtil.cpp:685-1014).  Elements [0, sse) follow the SSE body (saturating packs), the last
// n % 16 the scalar tail (truncating casts).
// ----------------------------------------------------------------------------------------------------------
__global__ void convert_kernel(int kind, const void* __restrict__ src, void* __restrict__ dst, long n, long sse, float fb, float a, float b) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float maf = __fmul_rn(a, fb);
    if (kind == 0) {                                           // disp8U2depth32F :923-1014
        float s = (float)((const uint8_t*)src)[i];
        if (b == 0.f) ((float*)dst)[i] = i < sse ? __fdiv_rn(maf, s) : __fadd_rn(__fdiv_rn(maf, s), b);
        else if (i < n - sse) ((float*)dst)[i] = __fadd_rn(__fdiv_rn(maf, s), b);   // SSE body commented out: only the first n%16
        return;
    }
    float s;
    if (kind == 1) s = ((const float*)src)[i];                 // depth32F2disp8U :768-838
    else if (kind == 2) s = i < sse ? (float)(int)(int16_t)((const uint16_t*)src)[i] : (float)(int)((const uint16_t*)src)[i];   // :858-859 sign extension
    else s = (float)(int)((const int16_t*)src)[i];             // disp16S2depth16U :685-765
    float v = __fdiv_rn(maf, s);
    if (i >= sse || b != 0.f) v = __fadd_rn(v, b);
    int iv = cvround(v);
    if (kind == 3) ((uint16_t*)dst)[i] = i < sse ? (uint16_t)(int16_t)sat_s16(iv) : (uint16_t)iv;
    else ((uint8_t*)dst)[i] = i < sse ? sat_u8(sat_s16(iv)) : (uint8_t)iv;
}

int launch_convert(int kind, const void* src, void* dst, long n, float fb, float a, float b, cudaStream_t s) {
    if (n <= 0) return 0;
    convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(kind, src, dst, n, (n / 16) * 16, fb, a, b);
    return 1;
}

__global__ void f32_to_u16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = sat_u16(cvround(src[i]));
}
int launch_f32_to_u16(const float* src, uint16_t* dst, long n, cudaStream_t s) {
    if (n <= 0) return 0;
    f32_to_u16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n);
    return 1;
}

// ----------------------------------------------------------------------------------------------------------
// fillOcclusion_<T> / fillOcclusionInv_<T> (depthmapUtil.cpp:548-636): row-serial run filling, one thread/row.
// ----------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void fill_occlusion_kernel(T* __restrict__ img, const T* __restrict__ pristine, int rows, int cols, T invalid, T edge, int inv) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    const int MAX_LENGTH = inv ? cols : (int)(cols * 0.5);
    T* s = img + (size_t)j * cols;
    s[0] = edge; s[cols - 1] = edge;
    for (int i = 1; i < cols - 1; i++) {
        if (s[i] == invalid) {
            int t = i;
            do { t++; if (t > cols - 1) break; } while (s[t] == invalid);
            // t == cols: the reference reads one element past the row = the (not yet processed) next row's first pixel
            T st = t <= cols - 1 ? s[t] : (j + 1 < rows ? pristine[(size_t)(j + 1) * cols] : invalid);
            const T dd = inv ? (s[i - 1] > st ? s[i - 1] : st) : (st < s[i - 1] ? st : s[i - 1]);
            if (t - i > MAX_LENGTH) { for (int n = 0; n < cols; n++) s[n] = invalid; }
            else { for (; i < t && i < cols; i++) s[i] = dd; }
        }
    }
    s[0] = s[1]; s[cols - 1] = s[cols - 2];
}

int launch_fill_occlusion(void* img, const void* pristine, int H, int W, int depth, double invalid, int inv, cudaStream_t s) {
    int blocks = (H + 63) / 64;
    switch (depth) {
    case 0: fill_occlusion_kernel<uint8_t><<<blocks, 64, 0, s>>>((uint8_t*)img, (const uint8_t*)pristine, H, W, (uint8_t)(int)invalid, inv ? 0 : 255, inv); return 1;
    case 3: fill_occlusion_kernel<int16_t><<<blocks, 64, 0, s>>>((int16_t*)img, (const int16_t*)pristine, H, W, (int16_t)(int)invalid, inv ? 0 : SHRT_MAX, inv); return 1;
    case 2: fill_occlusion_kernel<uint16_t><<<blocks, 64, 0, s>>>((uint16_t*)img, (const uint16_t*)pristine, H, W, (uint16_t)(int)invalid, inv ? 0 : USHRT_MAX, inv); return 1;
    case 5: fill_occlusion_kernel<float><<<blocks, 64, 0, s>>>((float*)img, (const float*)pristine, H, W, (float)(int)invalid, inv ? 0.f : FLT_MAX, inv); return 1;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// reprojectXYZ_<T> (depthmapUtil.cpp:450-481): xyz = (x*z, y*z, z==0 ? 10000 : z); x is the reference's running
// FP32 sum along the row (xtab, built by the host with the same serial adds), y = (j - ch) * fyinv.
// ----------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void reproject_kernel(const T* __restrict__ depth, float* __restrict__ xyz, const float* __restrict__ xtab, int H, int W, float fyinv, float ch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= W) return;
    float y = __fmul_rn(__fsub_rn((float)j, ch), fyinv);
    float z = (float)depth[(size_t)j * W + i];
    float* o = xyz + ((size_t)j * W + i) * 3;
    o[0] = __fmul_rn(xtab[i], z); o[1] = __fmul_rn(y, z); o[2] = (z == 0) ? 10000.f : z;
}

int launch_reproject(const void* depth, float* xyz, const float* xtab, int H, int W, int dtype, float fyinv, float ch, cudaStream_t s) {
    dim3 grid((W + 255) / 256, H), block(256);
    switch (dtype) {
    case 0: reproject_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)depth, xyz, xtab, H, W, fyinv, ch); return 1;
    case 3: reproject_kernel<int16_t><<<grid, block, 0, s>>>((const int16_t*)depth, xyz, xtab, H, W, fyinv, ch); return 1;
    case 2: reproject_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t*)depth, xyz, xtab, H, W, fyinv, ch); return 1;
    case 5: reproject_kernel<float><<<grid, block, 0, s>>>((const float*)depth, xyz, xtab, H, W, fyinv, ch); return 1;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// cv::transpose (main.cpp:258, :260 between the two fillOcclusion passes): 32x32 shared-memory tiles
// ----------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W) {
    __shared__ T t[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) if (x < W && y0 + j < H) t[j][threadIdx.x] = src[(size_t)(y0 + j) * W + x];
    __syncthreads();
    int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;          // dst is W rows x H cols
    for (int j = threadIdx.y; j < 32; j += blockDim.y) if (ox < H && oy0 + j < W) dst[(size_t)(oy0 + j) * H + ox] = t[threadIdx.x][j];
}

int launch_transpose(const void* src, void* dst, int H, int W, int elem_size, cudaStream_t s) {
    dim3 grid((W + 31) / 32, (H + 31) / 32), block(32, 8);
    switch (elem_size) {
    case 1: transpose_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)src, (uint8_t*)dst, H, W); return 1;
    case 2: transpose_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t*)src, (uint16_t*)dst, H, W); return 1;
    case 4: transpose_kernel<uint32_t><<<grid, block, 0, s>>>((const uint32_t*)src, (uint32_t*)dst, H, W); return 1;
    case 8: transpose_kernel<uint64_t><<<grid, block, 0, s>>>((const uint64_t*)src, (uint64_t*)dst, H, W); return 1;
    }
    return 0;
}

}  // namespace dmc
