// dmc_front8u.cu -- packed-SIMD kernels for the three 8-bit stages in front of the range filter, for the radii the
// reference's call sites use (median 3x3 / 5x5, Gaussian 3x3 / 5x5, min-max up to 21x21):
//
//   median   exchange networks on u16x2 lanes (VIMNMX.U16x2 issues at the full 4 warp-instructions/clk/SM on B200,
//            tools/ubench_pipes.cu), two pixels per instruction, rolling 5-row register window down a column strip
//   gauss    exact FP32 separable blur in OpenCV's operation order, byte<->float conversion by magic-number
//            permutes/adds instead of I2F/F2I (F2I.RN issues at 0.5/clk/SM)
//   min-max  separable dilate/erode on u16x2 lanes + branch-free "blur remove" select
//
// Every kernel stages its input tile (+halo, with the stage's own border rule) in shared memory once and then works
// out of registers: a thread owns a 2- or 4-pixel-wide column strip of R rows.
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"
#include <stdlib.h>

namespace dmc {

namespace {

constexpr int kTW = 128;     // output tile width  (2 warps x 32 lanes x 2 px)
constexpr int kHW = 4;       // staged horizontal halo for median (>= radius + 1, multiple of 4)

// Lanes hold pixels as fp16 with a +1024 bias (0x6400 | byte): normal numbers whose order equals the byte order, so
// HMNMX2 (full issue rate; VIMNMX.U16x2 measured at half rate in the median kernel's ncu profile) is an exact min/max.
__device__ __forceinline__ uint32_t pmin(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t pmax(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#define PCE(a, b) { uint32_t _lo = pmin(a, b); b = pmax(a, b); a = _lo; }

// The same exchange on the FMA pipe (HMNMX2 only runs on the half-rate ALU pipe: ncu shows pipe_alu ~100 % busy at
// 2 warp-instructions/clk/SM for a pure min/max network).  For UNBIASED fp16 lanes holding integers 0..255:
//   e = a - b;  m = b + e/2 = (a+b)/2;  hi = m + |e|/2;  lo = m - |e|/2      -- every intermediate is exact in fp16
// (multiples of 0.5 below 256).  Four HFMA2 instead of two HMNMX2; giving about a third of the exchanges to the FMA
// pipe balances the two pipes (2 x 2/3 ALU slots vs 4 x 1/3 FMA slots per exchange).
__device__ __forceinline__ void fce(uint32_t& a, uint32_t& b) {
    const __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b), hf = __float2half2_rn(0.5f);
    const __half2 e = __hsub2(x, y), m = __hfma2(e, hf, y), ae = __habs2(e);
    const __half2 hi = __hfma2(ae, hf, m), lo = __hfma2(__hneg2(ae), hf, m);
    a = *reinterpret_cast<const uint32_t*>(&lo); b = *reinterpret_cast<const uint32_t*>(&hi);
}
#define FCE(a, b) fce(a, b);

__device__ __forceinline__ uint32_t pmedian9(uint32_t p[9]) {
    PCE(p[1], p[2]) PCE(p[4], p[5]) PCE(p[7], p[8]) PCE(p[0], p[1]) PCE(p[3], p[4]) PCE(p[6], p[7])
    PCE(p[1], p[2]) PCE(p[4], p[5]) PCE(p[7], p[8]) PCE(p[0], p[3]) PCE(p[5], p[8]) PCE(p[4], p[7])
    PCE(p[3], p[6]) PCE(p[1], p[4]) PCE(p[2], p[5]) PCE(p[4], p[7]) PCE(p[4], p[2]) PCE(p[6], p[4])
    PCE(p[4], p[2])
    return p[4];
}
// VAR selects which exchanges go to the FMA pipe: exchange k uses HFMA2 when (k % FMOD) == FPH.
template <int FMOD, int FPH> __device__ __forceinline__ uint32_t pmedian25(uint32_t p[25]) {
#define XCE(k, a, b) { if ((k) % FMOD == FPH) fce(a, b); else PCE(a, b) }
    XCE(0, p[0], p[1]) XCE(1, p[3], p[4]) XCE(2, p[2], p[4]) XCE(3, p[2], p[3]) XCE(4, p[6], p[7]) XCE(5, p[5], p[7])
    XCE(6, p[5], p[6]) XCE(7, p[9], p[10]) XCE(8, p[8], p[10]) XCE(9, p[8], p[9]) XCE(10, p[12], p[13]) XCE(11, p[11], p[13])
    XCE(12, p[11], p[12]) XCE(13, p[15], p[16]) XCE(14, p[14], p[16]) XCE(15, p[14], p[15]) XCE(16, p[18], p[19]) XCE(17, p[17], p[19])
    XCE(18, p[17], p[18]) XCE(19, p[21], p[22]) XCE(20, p[20], p[22]) XCE(21, p[20], p[21]) XCE(22, p[23], p[24]) XCE(23, p[2], p[5])
    XCE(24, p[3], p[6]) XCE(25, p[0], p[6]) XCE(26, p[0], p[3]) XCE(27, p[4], p[7]) XCE(28, p[1], p[7]) XCE(29, p[1], p[4])
    XCE(30, p[11], p[14]) XCE(31, p[8], p[14]) XCE(32, p[8], p[11]) XCE(33, p[12], p[15]) XCE(34, p[9], p[15]) XCE(35, p[9], p[12])
    XCE(36, p[13], p[16]) XCE(37, p[10], p[16]) XCE(38, p[10], p[13]) XCE(39, p[20], p[23]) XCE(40, p[17], p[23]) XCE(41, p[17], p[20])
    XCE(42, p[21], p[24]) XCE(43, p[18], p[24]) XCE(44, p[18], p[21]) XCE(45, p[19], p[22]) XCE(46, p[8], p[17]) XCE(47, p[9], p[18])
    XCE(48, p[0], p[18]) XCE(49, p[0], p[9]) XCE(50, p[10], p[19]) XCE(51, p[1], p[19]) XCE(52, p[1], p[10]) XCE(53, p[11], p[20])
    XCE(54, p[2], p[20]) XCE(55, p[2], p[11]) XCE(56, p[12], p[21]) XCE(57, p[3], p[21]) XCE(58, p[3], p[12]) XCE(59, p[13], p[22])
    XCE(60, p[4], p[22]) XCE(61, p[4], p[13]) XCE(62, p[14], p[23]) XCE(63, p[5], p[23]) XCE(64, p[5], p[14]) XCE(65, p[15], p[24])
    XCE(66, p[6], p[24]) XCE(67, p[6], p[15]) XCE(68, p[7], p[16]) XCE(69, p[7], p[19]) XCE(70, p[13], p[21]) XCE(71, p[15], p[23])
    XCE(72, p[7], p[13]) XCE(73, p[7], p[15]) XCE(74, p[1], p[9]) XCE(75, p[3], p[11]) XCE(76, p[5], p[17]) XCE(77, p[11], p[17])
    XCE(78, p[9], p[17]) XCE(79, p[4], p[10]) XCE(80, p[6], p[12]) XCE(81, p[7], p[14]) XCE(82, p[4], p[6]) XCE(83, p[4], p[7])
    XCE(84, p[12], p[14]) XCE(85, p[10], p[14]) XCE(86, p[6], p[7]) XCE(87, p[10], p[12]) XCE(88, p[6], p[10]) XCE(89, p[6], p[17])
    XCE(90, p[12], p[17]) XCE(91, p[7], p[17]) XCE(92, p[7], p[10]) XCE(93, p[12], p[18]) XCE(94, p[7], p[12]) XCE(95, p[10], p[18])
    XCE(96, p[12], p[20]) XCE(97, p[10], p[20]) XCE(98, p[10], p[12])
    return p[12];
#undef XCE
}

// Loads 4 pixels starting at image column gx of row `row` (clamped = BORDER_REPLICATE) as one little-endian word.
__device__ __forceinline__ uint32_t load4_replicate(const uint8_t* __restrict__ row, int gx, int W, bool aligned_ok) {
    if (aligned_ok && gx >= 0 && gx + 3 < W) return *(const uint32_t*)(row + gx);
    return (uint32_t)row[clampi(gx, 0, W - 1)] | ((uint32_t)row[clampi(gx + 1, 0, W - 1)] << 8) |
           ((uint32_t)row[clampi(gx + 2, 0, W - 1)] << 16) | ((uint32_t)row[clampi(gx + 3, 0, W - 1)] << 24);
}

__device__ __forceinline__ void store_pair(uint8_t* __restrict__ dst, size_t off, int x, int W, uint32_t v, bool aligned_ok) {
    // v holds two 16-bit lanes whose low bytes are the pixels (the 0x64 bias byte is dropped)
    if (aligned_ok && x + 1 < W) *(uchar2*)(dst + off) = make_uchar2((uint8_t)(v & 0xFF), (uint8_t)((v >> 16) & 0xFF));
    else { dst[off] = (uint8_t)(v & 0xFF); if (x + 1 < W) dst[off + 1] = (uint8_t)((v >> 16) & 0xFF); }
}

// ------------------------------------------------------------------------------------------------------------------
// median (cv::medianBlur, 8UC1, BORDER_REPLICATE), RAD = 1 or 2
// ------------------------------------------------------------------------------------------------------------------
template <int RAD, int R, int FMOD, int FPH, int MINB>
__global__ void __launch_bounds__(256, MINB) median8u_p2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    constexpr int TILE_H = 4 * R, SW = kTW + 2 * kHW, SH = TILE_H + 2 * RAD, SWW = SW / 2, K = 2 * RAD + 1;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const uint8_t* fsrc = src + fo;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool al = (W & 3) == 0 && (fo & 3) == 0 && (reinterpret_cast<size_t>(src) & 3) == 0;
    for (int idx = tid; idx < SH * (SW / 4); idx += 256) {
        int ty = idx / (SW / 4), tq = idx - ty * (SW / 4);
        uint32_t w = load4_replicate(fsrc + (size_t)clampi(Y0 - RAD + ty, 0, H - 1) * W, X0 - kHW + 4 * tq, W, al);
        uint2 o; o.x = __byte_perm(w, 0x64646464u, 0x4140); o.y = __byte_perm(w, 0x64646464u, 0x4342);   // 0x6400 | byte = 1024 + byte
        const __half2 k1024 = __float2half2_rn(1024.f);                                                   // -> plain fp16 0..255
        __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&o.x), k1024), h1 = __hsub2(*reinterpret_cast<__half2*>(&o.y), k1024);
        o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
        *(uint2*)&sm[ty * SWW + 2 * tq] = o;
    }
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    const uint32_t* base = sm + (wy * R) * SWW + (xl + kHW - 2) / 2;      // word holding pixels (x-2, x-1)
    const int x = X0 + xl;
    const bool sal = (W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0 && (fo & 1) == 0;

    uint32_t win[K][K];     // rolling window: win[row][dx]
    auto load_row = [&](int yy, uint32_t (&v)[K]) {
        uint32_t w0 = base[yy * SWW], w1 = base[yy * SWW + 1], w2 = base[yy * SWW + 2];
        if (RAD == 2) { v[0] = w0; v[1] = __byte_perm(w0, w1, 0x5432); v[2] = w1; v[3] = __byte_perm(w1, w2, 0x5432); v[4] = w2; }
        else { v[0] = __byte_perm(w0, w1, 0x5432); v[1] = w1; v[2] = __byte_perm(w1, w2, 0x5432); }
    };
#pragma unroll
    for (int i = 0; i < K - 1; i++) load_row(i, win[i + 1]);
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int i = 0; i < K - 1; i++)
#pragma unroll
            for (int j = 0; j < K; j++) win[i][j] = win[i + 1][j];
        load_row(r + K - 1, win[K - 1]);
        uint32_t p[K * K];
#pragma unroll
        for (int i = 0; i < K; i++)
#pragma unroll
            for (int j = 0; j < K; j++) p[i * K + j] = win[i][j];
        uint32_t m = RAD == 1 ? pmedian9(p) : pmedian25<FMOD, FPH>(p);
        { __half2 mb = __hadd2(*reinterpret_cast<__half2*>(&m), __float2half2_rn(1024.f)); m = *reinterpret_cast<uint32_t*>(&mb); }   // low byte = pixel
        const int y = Y0 + wy * R + r;
        if (y < H && x < W) store_pair(dst, fo + (size_t)y * W + x, x, W, m, sal);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// small Gaussian, d = 3 or 5 (symmetric-pair form on both passes), BORDER_REFLECT_101
// ------------------------------------------------------------------------------------------------------------------
template <int GR> struct GaussK { float kx[GR + 1], ky[GR + 1]; };     // k[0] = centre tap, k[i] = tap at +-i

template <int GR, int R>
__global__ void __launch_bounds__(256) gauss8u_p4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, GaussK<GR> gk) {
    // thread: 4 pixels wide x R rows.  tile: 128 px (32 lanes x 4) x (8 warps x R) rows.
    constexpr int TILE_W = 128, TILE_H = 8 * R, HALO = 4, SW = TILE_W + 2 * HALO, SH = TILE_H + 2 * GR, SWW = SW / 4;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const uint8_t* fsrc = src + fo;
    const int X0 = blockIdx.x * TILE_W, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool al = (W & 3) == 0 && (fo & 3) == 0 && (reinterpret_cast<size_t>(src) & 3) == 0;
    for (int idx = tid; idx < SH * SWW; idx += 256) {
        int ty = idx / SWW, tq = idx - ty * SWW;
        const uint8_t* row = fsrc + (size_t)reflect101(Y0 - GR + ty, H) * W;
        int gx = X0 - HALO + 4 * tq;
        uint32_t w;
        if (al && gx >= 0 && gx + 3 < W) w = *(const uint32_t*)(row + gx);
        else w = (uint32_t)row[reflect101(gx, W)] | ((uint32_t)row[reflect101(gx + 1, W)] << 8) |
                 ((uint32_t)row[reflect101(gx + 2, W)] << 16) | ((uint32_t)row[reflect101(gx + 3, W)] << 24);
        sm[idx] = w;
    }
    __syncthreads();
    const int lane = threadIdx.x, wy = threadIdx.y;
    const uint32_t* base = sm + (wy * R) * SWW + lane;        // word left of this thread's 4 pixels (HALO = 4 px = 1 word)
    const int x = X0 + 4 * lane;
    constexpr uint32_t MAGIC = 0x4B000000u;                    // 2^23: (MAGIC | byte) as float = 8388608 + byte
    float rowv[2 * GR + 1][4];                                 // rolling window of row-pass results
    auto row_pass = [&](int yy, float (&o)[4]) {
        uint32_t w0 = base[yy * SWW], w1 = base[yy * SWW + 1], w2 = base[yy * SWW + 2];
        float f[12];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint32_t w = i < 4 ? w0 : (i < 8 ? w1 : w2);
            if (i >= 4 - GR && i < 8 + GR) f[i] = __uint_as_float(__byte_perm(w, MAGIC, 0x7440 + (i & 3))) - 8388608.f;   // exact
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float acc = f[4 + k] * gk.kx[0];
#pragma unroll
            for (int i = 1; i <= GR; i++) acc = acc + (f[4 + k - i] + f[4 + k + i]) * gk.kx[i];
            o[k] = acc;
        }
    };
#pragma unroll
    for (int i = 0; i < 2 * GR; i++) row_pass(i, rowv[i + 1]);
    const bool sal = (W & 3) == 0 && (reinterpret_cast<size_t>(dst) & 3) == 0 && (fo & 3) == 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int i = 0; i < 2 * GR; i++)
#pragma unroll
            for (int k = 0; k < 4; k++) rowv[i][k] = rowv[i + 1][k];
        row_pass(r + 2 * GR, rowv[2 * GR]);
        uint32_t ob[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float acc = rowv[GR][k] * gk.ky[0];
#pragma unroll
            for (int i = 1; i <= GR; i++) acc = acc + (rowv[GR - i][k] + rowv[GR + i][k]) * gk.ky[i];
            // RNE to integer: acc is in [0, 255.001], so acc + 1.5*2^23 holds RNE(acc) in its low mantissa bits
            ob[k] = __float_as_uint(acc + 12582912.f);
        }
        const int y = Y0 + wy * R + r;
        if (y >= H || x >= W) continue;
        uint32_t packed = __byte_perm(__byte_perm(ob[0], ob[1], 0x0040), __byte_perm(ob[2], ob[3], 0x0040), 0x5410);
        uint8_t* o = dst + fo + (size_t)y * W + x;
        if (sal && x + 3 < W) *(uint32_t*)o = packed;
        else for (int k = 0; k < 4 && x + k < W; k++) o[k] = (uint8_t)(packed >> (8 * k));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// min-max "blur remove" (minmaxFilter.cpp:48-174), 8UC1, radius RAD (compile time, 1..5)
// ------------------------------------------------------------------------------------------------------------------
template <int RAD, int R>
__global__ void __launch_bounds__(256) minmax8u_p2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    constexpr int HALO = 8, TILE_H = 4 * R, SW = kTW + 2 * HALO, SH = TILE_H + 2 * RAD, SWW = SW / 2;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const uint8_t* fsrc = src + fo;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool al = (W & 3) == 0 && (fo & 3) == 0 && (reinterpret_cast<size_t>(src) & 3) == 0;
    for (int idx = tid; idx < SH * (SW / 4); idx += 256) {
        int ty = idx / (SW / 4), tq = idx - ty * (SW / 4);
        uint32_t w = load4_replicate(fsrc + (size_t)clampi(Y0 - RAD + ty, 0, H - 1) * W, X0 - HALO + 4 * tq, W, al);
        uint2 o; o.x = __byte_perm(w, 0x64646464u, 0x4140); o.y = __byte_perm(w, 0x64646464u, 0x4342);   // 0x6400 | byte
        *(uint2*)&sm[ty * SWW + 2 * tq] = o;
    }
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    constexpr int E = (RAD + 1) & ~1;    // even offset >= RAD: words cover pixels x-E .. x+E+1
    const uint32_t* base = sm + (wy * R) * SWW + (xl + HALO - E) / 2;
    const int x = X0 + xl;
    const bool sal = (W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0 && (fo & 1) == 0;
    uint32_t rmx[2 * RAD + 1], rmn[2 * RAD + 1], ctr[2 * RAD + 1];     // rolling row max / min / centre pair
    auto row_pass = [&](int yy, uint32_t& mx, uint32_t& mn, uint32_t& c) {
        uint32_t w[E + 1];
#pragma unroll
        for (int i = 0; i <= E; i++) w[i] = base[yy * SWW + i];
        c = w[E / 2];
        mx = c; mn = c;
#pragma unroll
        for (int dx = -RAD; dx <= RAD; dx++) {
            if (dx == 0) continue;
            int o = dx + E;              // pixel offset from the first staged pixel of this thread
            uint32_t v = (o & 1) == 0 ? w[o / 2] : __byte_perm(w[(o - 1) / 2], w[(o + 1) / 2], 0x5432);
            mx = pmax(mx, v); mn = pmin(mn, v);
        }
    };
#pragma unroll
    for (int i = 0; i < 2 * RAD; i++) row_pass(i, rmx[i + 1], rmn[i + 1], ctr[i + 1]);
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int i = 0; i < 2 * RAD; i++) { rmx[i] = rmx[i + 1]; rmn[i] = rmn[i + 1]; ctr[i] = ctr[i + 1]; }
        row_pass(r + 2 * RAD, rmx[2 * RAD], rmn[2 * RAD], ctr[2 * RAD]);
        uint32_t mx = rmx[0], mn = rmn[0];
#pragma unroll
        for (int i = 1; i <= 2 * RAD; i++) { mx = pmax(mx, rmx[i]); mn = pmin(mn, rmn[i]); }
        const uint32_t v = ctr[RAD];
        // out = (v - mn <= mx - v) ? mn : mx  per 16-bit lane:  s = (mx - v) - (v - mn) + 0x8000 keeps bit 15 iff mn wins
        const uint32_t s = (mx + mn + 0x80008000u) - 2u * v;          // the 0x6400 biases cancel; lanes stay in 0x8000 +- 255
        uint32_t mask;                                                // replicate bit 15 of each lane over the lane
        asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(mask) : "r"(s), "r"(0u));   // (__byte_perm ignores the sign-replicate bit)
        const uint32_t out = (mn & mask) | (mx & ~mask);
        const int y = Y0 + wy * R + r;
        if (y < H && x < W) store_pair(dst, fo + (size_t)y * W + x, x, W, out, sal);
    }
}

// Few-frame launches (e.g. one 640x480 frame = 75 tiles of 128x32 on 148 SMs) use shorter tiles so that the grid fills the GPU.
inline bool small_launch(int W, int H, int n, int tile_h) { return (long)((W + kTW - 1) / kTW) * ((H + tile_h - 1) / tile_h) * n < 2 * 148; }

template <int RAD> int launch_minmax_rad(const uint8_t* src, uint8_t* dst, int n, int H, int W, cudaStream_t s) {
    dim3 block(32, 8);
    if (small_launch(W, H, n, 64)) { constexpr int R = 4; dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n); minmax8u_p2_kernel<RAD, R><<<grid, block, 0, s>>>(src, dst, H, W); }
    else { constexpr int R = 16; dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n); minmax8u_p2_kernel<RAD, R><<<grid, block, 0, s>>>(src, dst, H, W); }   // per 400 1080p frames: R=4 1.80 ms, R=8 1.50, R=16 1.37
    return 1;
}

}  // namespace

int launch_median8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s) {
    if (r != 1 && r != 2) return 0;
    dim3 block(32, 8);
    // exchange split between the pipes, measured per 400 1080p frames (5x5): every 5th exchange on the FMA pipe 4.11 ms,
    // every 3rd 4.22, none 5.02, every 2nd 5.17
    if (small_launch(W, H, n, 32)) {
        constexpr int R = 2; dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n);
        if (r == 1) median8u_p2_kernel<1, R, 3, 1, 3><<<grid, block, 0, s>>>(src, dst, H, W);
        else median8u_p2_kernel<2, R, 5, 1, 3><<<grid, block, 0, s>>>(src, dst, H, W);
    } else {
        constexpr int R = 8; dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n);
        if (r == 1) median8u_p2_kernel<1, R, 3, 1, 3><<<grid, block, 0, s>>>(src, dst, H, W);
        else median8u_p2_kernel<2, R, 5, 1, 3><<<grid, block, 0, s>>>(src, dst, H, W);
    }
    return 1;
}

template <int GR, int R> static void launch_gauss_gr(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s) {
    GaussK<GR> g; for (int i = 0; i <= GR; i++) { g.kx[i] = t.kx[GR + i]; g.ky[i] = t.ky[GR + i]; }
    dim3 grid((W + 127) / 128, (H + 8 * R - 1) / (8 * R), n), block(32, 8);
    gauss8u_p4_kernel<GR, R><<<grid, block, 0, s>>>(src, dst, H, W, g);
}

int launch_gauss8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s) {
    if (t.rx != t.ry || t.rx < 1 || t.rx > 2) return 0;        // 1-pixel-wide/high images and large kernels: generic kernel
    const bool small = small_launch(W, H, n, 64);              // rows per thread, per 400 1080p frames: R=2 1.29 ms, R=4 1.13, R=8 1.10
    if (t.rx == 1) { if (small) launch_gauss_gr<1, 2>(src, dst, n, H, W, t, s); else launch_gauss_gr<1, 8>(src, dst, n, H, W, t, s); }
    else { if (small) launch_gauss_gr<2, 2>(src, dst, n, H, W, t, s); else launch_gauss_gr<2, 4>(src, dst, n, H, W, t, s); }
    return 1;
}

int launch_minmax8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s) {
    switch (r) {
    case 1: return launch_minmax_rad<1>(src, dst, n, H, W, s);
    case 2: return launch_minmax_rad<2>(src, dst, n, H, W, s);
    case 3: return launch_minmax_rad<3>(src, dst, n, H, W, s);
    case 4: return launch_minmax_rad<4>(src, dst, n, H, W, s);
    case 5: return launch_minmax_rad<5>(src, dst, n, H, W, s);
    }
    return 0;
}

}  // namespace dmc
