// dmc_front8u.cu -- packed-SIMD kernels for the three 8-bit stages in front of the range filter, for the radii the
// reference's call sites use (median 3x3 / 5x5, Gaussian 3x3 / 5x5, min-max up to 21x21):
//
//   median   3x3 / 5x5: shared-sort min/max circuits on fp16x2 lanes (tools/median_circuit.py): every row window is sorted
//            once and serves five outputs, merged row pairs serve two output pairs; a third of the exchanges run on the
//            FMA pipe.  7x7 .. 21x21: bisection on the value with packed fp16 counting
//   gauss    exact FP32 separable blur in OpenCV's operation order, byte<->float conversion by magic-number
//            permutes/adds instead of I2F/F2I (F2I.RN issues at 0.5/clk/SM)
//   min-max  separable dilate/erode on u16x2 lanes + branch-free "blur remove" select
//
// Every kernel stages its input tile (+halo, with the stage's own border rule) in shared memory once and then works
// out of registers: a thread owns a 2- or 4-pixel-wide column strip of R rows.
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"
#include "dmc_stage.cuh"
#include <stdlib.h>
#include <type_traits>

#ifndef DMC_MED_FNUM
#define DMC_MED_FNUM 1      // exchanges (i % DMC_MED_FMOD) < DMC_MED_FNUM of every circuit run on the FMA pipe
#define DMC_MED_FMOD 3
#endif
#ifndef DMC_MM_R
#define DMC_MM_R 16     // rows per thread of the min-max kernel
#endif
#ifndef DMC_G1_R
#define DMC_G1_R 8      // rows per thread of the 3x3 Gaussian
#endif
#ifndef DMC_MED_R
#define DMC_MED_R 16
#endif
#ifndef DMC_MED_MINB
#define DMC_MED_MINB 3
#endif

namespace dmc {

namespace {


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

// The shared-sort median circuits (tools/median_circuit.py designs, verifies on every binary input, and generates them).
// Exchange i of a circuit runs on the FMA pipe when (i % FMOD) < FNUM.
template <bool FMA> __device__ __forceinline__ void med_xchg(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
    if (FMA) { fce(a, b); lo = a; hi = b; } else { lo = pmin(a, b); hi = pmax(a, b); }
}
template <int FNUM, int FMOD> struct MedCircuit {
#define DMC_MED_T uint32_t
#define DMC_MED_FN static __device__ __forceinline__
#define DMC_MED_MIN(a, b) pmin(a, b)
#define DMC_MED_MAX(a, b) pmax(a, b)
#define DMC_MED_XCHG(i, a, b, lo, hi) uint32_t lo, hi; med_xchg<(((i) % FMOD) < FNUM)>(a, b, lo, hi)
#include "dmc_median_gen.inc"
#undef DMC_MED_T
#undef DMC_MED_FN
#undef DMC_MED_MIN
#undef DMC_MED_MAX
#undef DMC_MED_XCHG
};

// ------------------------------------------------------------------------------------------------------------------
// median (cv::medianBlur, 8UC1, BORDER_REPLICATE), RAD = 1 or 2
// ------------------------------------------------------------------------------------------------------------------
// A thread produces R vertically adjacent outputs of one pixel pair.  The horizontal window of every input row is sorted
// once (it serves up to five outputs), pairs of sorted rows are merged once (each pair serves two output pairs), and only
// the ranks of the four shared rows that can still be the median (7..12 of 20) are kept before the last row of each of
// the two outputs is taken in: 71 min/max per output at R = 8 instead of the ~190 of a 25-input selection network.
template <int RAD, int R, int FNUM, int FMOD, int MINB, bool EVEN>     // EVEN: W, the frame size and dst allow 2-byte stores
__global__ void __launch_bounds__(256, MINB) median8u_p2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    static_assert(R % 2 == 0, "outputs are produced in vertical pairs");
    constexpr int HW = kHalo16, TILE_H = 4 * R, SH = TILE_H + 2 * RAD, SWW = kSW16 / 2, K = 2 * RAD + 1;
    using MC = MedCircuit<FNUM, FMOD>;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const uint8_t* fsrc = src + fo;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    stage_tile16<SH, 2, false>(sm, fsrc, X0, Y0 - RAD, H, W, aligned16(src, W), tid);      // plain fp16 0..255 (exact FMA-pipe exchanges)
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    const uint32_t* base = sm + (wy * R) * SWW + (xl + HW - 2) / 2;      // word holding pixels (x-2, x-1)
    const int x = X0 + xl;
    if (x >= W) return;                                                   // (no barrier below)
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);                                   // rows of this strip inside the image

    auto sorted_row = [&](int yy, uint32_t (&v)[K]) {      // the K horizontal neighbours of both lanes, sorted
        uint32_t w0 = base[yy * SWW], w1 = base[yy * SWW + 1], w2 = base[yy * SWW + 2];
        if constexpr (RAD == 2) { v[0] = w0; v[1] = __byte_perm(w0, w1, 0x5432); v[2] = w1; v[3] = __byte_perm(w1, w2, 0x5432); v[4] = w2; MC::med_sort5(v); }
        else { v[0] = __byte_perm(w0, w1, 0x5432); v[1] = w1; v[2] = __byte_perm(w1, w2, 0x5432); MC::med_sort3(v); }
    };
    auto put = [&](int r, uint32_t m) {
        { __half2 mb = __hadd2(*reinterpret_cast<__half2*>(&m), __float2half2_rn(1024.f)); m = *reinterpret_cast<uint32_t*>(&mb); }   // low byte = pixel
        if (r < yrem) {
            if (EVEN) *reinterpret_cast<uint16_t*>(op) = (uint16_t)__byte_perm(m, 0, 0x4420);
            else { op[0] = (uint8_t)m; if (x + 1 < W) op[1] = (uint8_t)(m >> 16); }
        }
        op += W;
    };

    uint32_t S[R + 2 * RAD][K];
    if constexpr (RAD == 1) {
#pragma unroll
        for (int i = 0; i < R + 2; i++) {
            sorted_row(i, S[i]);
            if (i >= 2) put(i - 2, MC::med_select9(S[i - 2], S[i - 1], S[i]));
        }
    } else {
        uint32_t P[R / 2 + 1][10];      // P[j] = merge(S[2j+1], S[2j+2])
#pragma unroll
        for (int i = 0; i < 6; i++) sorted_row(i, S[i]);
        MC::med_merge55(S[1], S[2], P[0]);
        MC::med_merge55(S[3], S[4], P[1]);
#pragma unroll
        for (int o = 0; o < R; o += 2) {
            if (o > 0) { sorted_row(o + 4, S[o + 4]); sorted_row(o + 5, S[o + 5]); MC::med_merge55(S[o + 3], S[o + 4], P[o / 2 + 1]); }
            uint32_t M[6];
            MC::med_mid6(P[o / 2], P[o / 2 + 1], M);
            put(o, MC::med_select(S[o], M));
            put(o + 1, MC::med_select(S[o + 5], M));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// median of larger windows (RAD = 3 .. 10, the rest of the GUI's trackbar range): bisection on the value, two pixels per
// instruction.  The median of k*k bytes is the smallest m with #{v <= m} > k*k/2; eight halvings of [0, 255] find it.
// Bounds, candidate and count are small integers in fp16 lanes (count <= 441): per tap HSET2.BF (v <= mid) + HADD2.
// ------------------------------------------------------------------------------------------------------------------
template <int RAD, int R, bool EVEN>
__global__ void __launch_bounds__(256) median8u_bisect_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    constexpr int TILE_H = 4 * R, SH = TILE_H + 2 * RAD, SWW = kSW16 / 2, K = 2 * RAD + 1, E = (RAD + 1) & ~1, NW = E + 1;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    stage_tile16<SH, 2, false>(sm, src + fo, X0, Y0 - RAD, H, W, aligned16(src, W), tid);      // plain fp16 0..255
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane, x = X0 + xl;
    if (x >= W) return;                                                    // (no barrier below)
    const uint32_t* base = sm + (wy * R) * SWW + (xl + kHalo16 - E) / 2;  // word holding pixels (x-E, x-E+1)
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);
    const __half2 half2v = __float2half2_rn((float)((K * K) / 2)), one = __float2half2_rn(1.f), hf = __float2half2_rn(0.5f);
#pragma unroll 1
    for (int r = 0; r < R; r++) {
        __half2 lo = __float2half2_rn(0.f), hi = __float2half2_rn(255.f);
#pragma unroll 1
        for (int pass = 0; pass < 8; pass++) {
            const __half2 mid = h2floor(__hmul2(__hadd2(lo, hi), hf));
            __half2 cnt = __float2half2_rn(0.f);
#pragma unroll
            for (int dy = 0; dy < K; dy++) {
                uint32_t wd[NW];
#pragma unroll
                for (int i = 0; i < NW; i++) {     // small windows: the compiler keeps the (pass-invariant) window in registers;
                    if constexpr (RAD >= 8) wd[i] = reinterpret_cast<const volatile uint32_t*>(base)[(r + dy) * SWW + i];   // large ones: reload, no spills
                    else wd[i] = base[(r + dy) * SWW + i];
                }
#pragma unroll
                for (int dx = -RAD; dx <= RAD; dx++) {
                    const int o = dx + E;
                    uint32_t vb = (o & 1) == 0 ? wd[o / 2] : __byte_perm(wd[(o - 1) / 2], wd[(o + 1) / 2], 0x5432);
                    cnt = __hadd2(cnt, __hle2(*reinterpret_cast<__half2*>(&vb), mid));
                }
            }
            const __half2 m = __hgt2(cnt, half2v);                        // 1: the median is <= mid
            hi = __hfma2(m, __hsub2(mid, hi), hi);                          // m ? mid : hi
            lo = __hfma2(__hsub2(one, m), __hsub2(__hadd2(mid, one), lo), lo);      // m ? lo : mid + 1
        }
        uint32_t res; { __half2 b = __hadd2(lo, __float2half2_rn(1024.f)); res = *reinterpret_cast<uint32_t*>(&b); }   // low bytes = pixels
        if (r < yrem) {
            if (EVEN) *reinterpret_cast<uint16_t*>(op) = (uint16_t)__byte_perm(res, 0, 0x4420);
            else { op[0] = (uint8_t)res; if (x + 1 < W) op[1] = (uint8_t)(res >> 16); }
        }
        op += W;
    }
}

template <int RAD> static void launch_median_bisect(const uint8_t* src, uint8_t* dst, int n, int H, int W, cudaStream_t s) {
    constexpr int R = 4;
    dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n), block(32, 8);
    if ((W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0) median8u_bisect_kernel<RAD, R, true><<<grid, block, 0, s>>>(src, dst, H, W);
    else median8u_bisect_kernel<RAD, R, false><<<grid, block, 0, s>>>(src, dst, H, W);
}

// ------------------------------------------------------------------------------------------------------------------
// small Gaussian, d = 3 or 5 (symmetric-pair form on both passes), BORDER_REFLECT_101
// ------------------------------------------------------------------------------------------------------------------
template <int GR> struct GaussK { float kx[GR + 1], ky[GR + 1]; };     // k[0] = centre tap, k[i] = tap at +-i

template <int GR, int R, bool QUAD>     // QUAD: W and dst allow 4-byte stores
__global__ void __launch_bounds__(256) gauss8u_p4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, GaussK<GR> gk) {
    // thread: 4 pixels wide x R rows.  tile: 128 px (32 lanes x 4) x (8 warps x R) rows.
    constexpr int TILE_H = 8 * R, SH = TILE_H + 2 * GR, SWW = kSW16 / 4;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    stage_tile16<SH, 0, true>(sm, src + fo, X0, Y0 - GR, H, W, aligned16(src, W), tid);      // raw bytes, REFLECT_101
    __syncthreads();
    const int lane = threadIdx.x, wy = threadIdx.y;
    const uint32_t* base = sm + (wy * R) * SWW + lane + (kHalo16 / 4 - 1);        // word left of this thread's 4 pixels
    const int x = X0 + 4 * lane;
    if (x >= W) return;                                        // (no barrier below)
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);
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
        const uint32_t packed = __byte_perm(__byte_perm(ob[0], ob[1], 0x0040), __byte_perm(ob[2], ob[3], 0x0040), 0x5410);
        if (r < yrem) {
            if (QUAD) *reinterpret_cast<uint32_t*>(op) = packed;
            else for (int k = 0; k < 4 && x + k < W; k++) op[k] = (uint8_t)(packed >> (8 * k));
        }
        op += W;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// min-max "blur remove" (minmaxFilter.cpp:48-174), 8UC1, radius RAD (compile time, 1..5)
// ------------------------------------------------------------------------------------------------------------------
template <int RAD, int R, bool EVEN>     // EVEN: W and dst allow 2-byte stores
__global__ void __launch_bounds__(256) minmax8u_p2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    constexpr int HALO = kHalo16, TILE_H = 4 * R, SH = TILE_H + 2 * RAD, SWW = kSW16 / 2;
    __shared__ __align__(16) uint32_t sm[SH * SWW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    stage_tile16<SH, 1, false>(sm, src + fo, X0, Y0 - RAD, H, W, aligned16(src, W), tid);      // 0x6400 | byte
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    constexpr int E = (RAD + 1) & ~1;    // even offset >= RAD: words cover pixels x-E .. x+E+1
    const uint32_t* base = sm + (wy * R) * SWW + (xl + HALO - E) / 2;
    const int x = X0 + xl;
    if (x >= W) return;                                                   // (no barrier below)
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);
    uint32_t rmx[2 * RAD + 1], rmn[2 * RAD + 1], ctr[2 * RAD + 1];     // rolling row max / min / centre pair
    auto row_pass = [&](int yy, uint32_t& mx, uint32_t& mn, uint32_t& c) {
        uint32_t w[E + 1];
#pragma unroll
        for (int i = 0; i <= E; i++) w[i] = base[yy * SWW + i];
        c = w[E / 2];
        if constexpr ((RAD & 1) == 1) {
            // odd radius: the two lanes' windows share the 2*RAD pixels x-RAD+1 .. x+RAD, which are RAD aligned words
            // w[1..RAD]: reduce them lane-wise (even | odd positions), fold the two halves together, then add each
            // lane's own end pixel (x-RAD for the low lane, x+RAD+1 for the high one) -- no shifted taps needed.
            mx = w[1]; mn = w[1];
#pragma unroll
            for (int i = 2; i <= RAD; i++) { mx = pmax(mx, w[i]); mn = pmin(mn, w[i]); }
            const uint32_t edge = __byte_perm(w[0], w[RAD + 1], 0x5432);        // (p[x-RAD], p[x+RAD+1])
            mx = pmax(pmax(mx, __byte_perm(mx, mx, 0x1032)), edge);
            mn = pmin(pmin(mn, __byte_perm(mn, mn, 0x1032)), edge);
        } else {
            mx = c; mn = c;
#pragma unroll
            for (int dx = -RAD; dx <= RAD; dx++) {
                if (dx == 0) continue;
                int o = dx + E;              // pixel offset from the first staged pixel of this thread
                uint32_t v = (o & 1) == 0 ? w[o / 2] : __byte_perm(w[(o - 1) / 2], w[(o + 1) / 2], 0x5432);
                mx = pmax(mx, v); mn = pmin(mn, v);
            }
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
        if (r < yrem) {
            if (EVEN) *reinterpret_cast<uint16_t*>(op) = (uint16_t)__byte_perm(out, 0, 0x4420);
            else { op[0] = (uint8_t)out; if (x + 1 < W) op[1] = (uint8_t)(out >> 16); }
        }
        op += W;
    }
}

// Few-frame launches (e.g. one 640x480 frame = 75 tiles of 128x32 on 148 SMs) use shorter tiles so that the grid fills the GPU.
inline bool small_launch(int W, int H, int n, int tile_h) { return (long)((W + kTW - 1) / kTW) * ((H + tile_h - 1) / tile_h) * n < 2 * 148; }

template <int RAD, int R> void launch_minmax_rr(const uint8_t* src, uint8_t* dst, int n, int H, int W, cudaStream_t s) {
    dim3 block(32, 8), grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n);
    if ((W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0) minmax8u_p2_kernel<RAD, R, true><<<grid, block, 0, s>>>(src, dst, H, W);
    else minmax8u_p2_kernel<RAD, R, false><<<grid, block, 0, s>>>(src, dst, H, W);
}
template <int RAD> int launch_minmax_rad(const uint8_t* src, uint8_t* dst, int n, int H, int W, cudaStream_t s) {
    if (small_launch(W, H, n, 64)) launch_minmax_rr<RAD, 4>(src, dst, n, H, W, s);
    else launch_minmax_rr<RAD, DMC_MM_R>(src, dst, n, H, W, s);      // per 400 1080p frames: R=4 1.80 ms, R=8 1.50, R=16 1.37
    return 1;
}

}  // namespace

int launch_median8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, int r, cudaStream_t s) {
    switch (r) {
    case 3: launch_median_bisect<3>(src, dst, n, H, W, s); return 1;
    case 4: launch_median_bisect<4>(src, dst, n, H, W, s); return 1;
    case 5: launch_median_bisect<5>(src, dst, n, H, W, s); return 1;
    case 6: launch_median_bisect<6>(src, dst, n, H, W, s); return 1;
    case 7: launch_median_bisect<7>(src, dst, n, H, W, s); return 1;
    case 8: launch_median_bisect<8>(src, dst, n, H, W, s); return 1;
    case 9: launch_median_bisect<9>(src, dst, n, H, W, s); return 1;
    case 10: launch_median_bisect<10>(src, dst, n, H, W, s); return 1;
    }
    if (r != 1 && r != 2) return 0;
    dim3 block(32, 8);
    // measured per 100 1080p frames (5x5, tools/quick_med.py): R = 8 0.45 ms, R = 16 0.40 ms; every 3rd exchange on the FMA pipe
    // 0.54 (R = 8, before the staging rework), every 4th 0.54, 2 of 5 0.55, none 0.63.  The 25-input network this replaces: 1.03 ms.
    const bool even = (W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0;      // then every frame offset n*H*W is even too
    auto go = [&](auto rad, auto rows, auto ev) {
        constexpr int RAD = decltype(rad)::value, R = decltype(rows)::value; constexpr bool EV = decltype(ev)::value;
        dim3 grid((W + kTW - 1) / kTW, (H + 4 * R - 1) / (4 * R), n);
        median8u_p2_kernel<RAD, R, DMC_MED_FNUM, DMC_MED_FMOD, DMC_MED_MINB, EV><<<grid, block, 0, s>>>(src, dst, H, W);
    };
    auto go_r = [&](auto rad, auto rows) { if (even) go(rad, rows, std::true_type()); else go(rad, rows, std::false_type()); };
    using std::integral_constant;
    if (small_launch(W, H, n, 32)) {
        if (r == 1) go_r(integral_constant<int, 1>(), integral_constant<int, 2>()); else go_r(integral_constant<int, 2>(), integral_constant<int, 2>());
    } else {
        if (r == 1) go_r(integral_constant<int, 1>(), integral_constant<int, DMC_MED_R>()); else go_r(integral_constant<int, 2>(), integral_constant<int, DMC_MED_R>());
    }
    return 1;
}

template <int GR, int R> static void launch_gauss_gr(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s) {
    GaussK<GR> g; for (int i = 0; i <= GR; i++) { g.kx[i] = t.kx[GR + i]; g.ky[i] = t.ky[GR + i]; }
    dim3 grid((W + 127) / 128, (H + 8 * R - 1) / (8 * R), n), block(32, 8);
    if ((W & 3) == 0 && (reinterpret_cast<size_t>(dst) & 3) == 0) gauss8u_p4_kernel<GR, R, true><<<grid, block, 0, s>>>(src, dst, H, W, g);
    else gauss8u_p4_kernel<GR, R, false><<<grid, block, 0, s>>>(src, dst, H, W, g);
}

int launch_gauss8u_fast(const uint8_t* src, uint8_t* dst, int n, int H, int W, const GaussTaps& t, cudaStream_t s) {
    if (t.rx != t.ry || t.rx < 1 || t.rx > 2) return 0;        // 1-pixel-wide/high images and large kernels: generic kernel
    const bool small = small_launch(W, H, n, 64);              // rows per thread, per 400 1080p frames: R=2 1.29 ms, R=4 1.13, R=8 1.10
    if (t.rx == 1) { if (small) launch_gauss_gr<1, 2>(src, dst, n, H, W, t, s); else launch_gauss_gr<1, DMC_G1_R>(src, dst, n, H, W, t, s); }
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
