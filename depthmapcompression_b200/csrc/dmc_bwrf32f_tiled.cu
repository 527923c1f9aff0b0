// dmc_bwrf32f_tiled.cu -- register-tiled fast path of the 32-bit binary-weighted range filter (single channel, square
// window of radius 1..10): the kernel behind filterDisp8U2Depth32F / Depth16U / Disp32F and the 16U/16S/32F
// binalyWeightedRangeFilter (binalyWeightedRangeFilter.cpp:471-550, :978-1029).
//
// Parity rules are those of the generic kernel in dmc_kernels_32f.cu: for every output pixel the taps are visited
// in raster order and accumulated sequentially in FP32 with individually rounded operations,
//     d = c - v;  w = (|d| <= th) ? 1.f : 0.f;  t = t + w*v;  W = W + w;  out = t / W
// (w*v is a real multiply so that 0*inf = NaN propagates; NaN compares false; since w is 0 or 1 the product is exact
// and t + w*v is computed with one FFMA -- same rounding, same NaN/inf/signed-zero behaviour as mul then add).  What changes is the schedule: a
// thread owns one pixel column and R output rows; it walks the input rows top to bottom, loads the 2*RAD+1 floats of
// a row once from shared memory and feeds them to every output row whose window contains that row.  For a fixed
// output row the order is still (dy ascending, dx ascending) = raster order.  ~0.4 shared loads per tap instead of 1.
//
// Integer sources (16U / 16S / 8U loads, i.e. every input of the reference's convertTo(CV_32F) front end): all values are
// integers below 2^16 and a window has at most 255 taps, so every partial sum of the reference's FP32 accumulation is an
// integer below 2^24 -- exact, hence independent of the order and equal to an integer sum.  The kernel then runs in
// integers with the count packed under the sum: a staged element is e = 256*v + 1, a tap is
//     u = e + (256*floor(th) - e_centre);  if (u <=u 512*floor(th)) acc += e        (3 instructions instead of 5)
// and sum = acc >> 8, count = acc & 255 (255 * (65535 * 256 + 1) < 2^32); the quotient is the same IEEE division of two
// exactly converted floats.  |c - v| <= th on integers is |c - v| <= floor(th) for th >= 0 (th < 0 or NaN keeps the float path).
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"

namespace dmc {

namespace {

constexpr int kTW = 128;      // output tile width (4 warps x 32 lanes)

__host__ __device__ constexpr int hw_of(int rad, int dy) {
    int lim = rad * rad - dy * dy, j = 0;
    while ((j + 1) * (j + 1) <= lim) j++;
    return j;
}

__device__ __forceinline__ float load_px(const void* __restrict__ p, size_t i, int load_op, float maf) {
    switch (load_op) {
    case LOAD_F32: return ((const float*)p)[i];
    case LOAD_U16: return (float)((const uint16_t*)p)[i];
    case LOAD_S16: return (float)((const int16_t*)p)[i];
    case LOAD_U8: return (float)((const uint8_t*)p)[i];
    default: return __fdiv_rn(maf, (float)((const uint8_t*)p)[i]);       // disp8U2depth32F, b == 0 (depthmapUtil.cpp:935-968)
    }
}

// MODE 0: FP32 accumulation; 1: packed integer accumulation, unsigned source; 2: the same, signed 16-bit source
template <int RAD, int R, int MODE>
__global__ void __launch_bounds__(256) bwrf32f_tiled_kernel(const void* __restrict__ src, void* __restrict__ dst, int H, int W,
                                                            float th, float maf, int load_op, int store_op, int quirk) {
    constexpr int TILE_H = 2 * R;                 // 8 warps = 4 (x) x 2 (y)
    constexpr int SW = kTW + 2 * RAD + 1, SH = TILE_H + 2 * RAD;      // +1: odd stride, conflict-free rows
    __shared__ float sm[SH * SW];
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    // disp8U2depth32F fused into the load (depthmapUtil.cpp:935-968): one IEEE division per byte VALUE, then a table look-up
    __shared__ float s_lut[256];
    if (load_op == LOAD_U8_DISP2DEPTH) { s_lut[tid] = __fdiv_rn(maf, (float)tid); __syncthreads(); }
    for (int ty = threadIdx.y; ty < SH; ty += 8) {                          // a warp per staged row: no index division
        const int uy = Y0 - RAD + ty, gy = clampi(uy, 0, H - 1);          // unclamped (padded-buffer) / clamped row
        const size_t rowi = fo + (size_t)gy * W;
        for (int tx = threadIdx.x; tx < kTW + 2 * RAD; tx += 32) {
            const int ux = X0 - RAD + tx;
            size_t gi = rowi + clampi(ux, 0, W - 1);
            // padding quirk of the reference (see dmc_kernels_32f.cu): halo column W-1+RAD holds the next padded line's first element
            if (quirk && ux == W - 1 + RAD && uy + 1 <= H - 1 + RAD) gi = fo + (size_t)clampi(uy + 1, 0, H - 1) * W;
            if constexpr (MODE == 0) sm[ty * SW + tx] = load_op == LOAD_U8_DISP2DEPTH ? s_lut[((const uint8_t*)src)[gi]] : load_px(src, gi, load_op, maf);
            else {
                const int iv = load_op == LOAD_U16 ? (int)((const uint16_t*)src)[gi] : load_op == LOAD_S16 ? (int)((const int16_t*)src)[gi] : (int)((const uint8_t*)src)[gi];
                sm[ty * SW + tx] = __int_as_float(iv * 256 + 1);
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 3, wy = threadIdx.y >> 2;
    const int xl = 32 * wx + lane;
    const float* base = sm + (wy * R) * SW + xl;                       // staged column of pixel x - RAD, first input row of the block
    const int x = X0 + xl;
    if constexpr (MODE != 0) {
        const uint32_t thi = (uint32_t)fminf(floorf(th), 131071.f) * 256u, lim = 2u * thi;
        uint32_t kk[R], acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) { kk[r] = thi - __float_as_uint(base[(r + RAD) * SW + RAD]); acc[r] = 0u; }
#pragma unroll
        for (int yy = 0; yy < R + 2 * RAD; yy++) {
            uint32_t v[2 * RAD + 1];
#pragma unroll
            for (int i = 0; i < 2 * RAD + 1; i++) v[i] = __float_as_uint(base[yy * SW + i]);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
                if (ady <= RAD) {
#pragma unroll
                    for (int dx = -hw_of(RAD, ady); dx <= hw_of(RAD, ady); dx++) {
                        const uint32_t e = v[dx + RAD];
// add, compare, predicated add.  The two integer pipes (ALU: ISETP / IADD3, FMA: IMAD) issue 2 warp-instructions
                        // per clock each, so the predicated add alternates between them: 1.5 instructions per tap on either pipe.
                        if ((dx + RAD + r) & 1) asm("{.reg .pred p; .reg .u32 u; add.u32 u, %1, %2; setp.le.u32 p, u, %3; @p add.u32 %0, %0, %1;}" : "+r"(acc[r]) : "r"(e), "r"(kk[r]), "r"(lim));
                        else asm("{.reg .pred p; .reg .u32 u; add.u32 u, %1, %2; setp.le.u32 p, u, %3; @p mad.lo.u32 %0, %1, 1, %0;}" : "+r"(acc[r]) : "r"(e), "r"(kk[r]), "r"(lim));
                    }
                }
            }
        }
        if (x >= W) return;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int y = Y0 + wy * R + r;
            if (y >= H) continue;
            const uint32_t n = acc[r] & 255u;
            const float fs = MODE == 2 ? (float)((int)(acc[r] - n) >> 8) : (float)(acc[r] >> 8);
            const float o = __fdiv_rn(fs, (float)n);
            const size_t oi = fo + (size_t)y * W + x;
            if (store_op == STORE_F32) ((float*)dst)[oi] = o;
            else if (store_op == STORE_U16) ((uint16_t*)dst)[oi] = sat_u16(cvround(o));
            else ((int16_t*)dst)[oi] = (int16_t)sat_s16(cvround(o));
        }
    } else {
    float c[R], t[R], wsum[R];
#pragma unroll
    for (int r = 0; r < R; r++) { c[r] = base[(r + RAD) * SW + RAD]; t[r] = 0.f; wsum[r] = 0.f; }
#pragma unroll
    for (int yy = 0; yy < R + 2 * RAD; yy++) {
        float v[2 * RAD + 1];
#pragma unroll
        for (int i = 0; i < 2 * RAD + 1; i++) v[i] = base[yy * SW + i];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
            if (ady <= RAD) {
#pragma unroll
                for (int dx = -hw_of(RAD, ady); dx <= hw_of(RAD, ady); dx++) {       // raster order within the row
                    const float vv = v[dx + RAD];
                    const float w = fabsf(__fsub_rn(c[r], vv)) <= th ? 1.f : 0.f;
                    t[r] = __fmaf_rn(w, vv, t[r]);        // w is 0 or 1: the product is exact, so one rounding = the reference's mul then add
                    wsum[r] = __fadd_rn(wsum[r], w);
                }
            }
        }
    }
    if (x >= W) return;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = Y0 + wy * R + r;
        if (y >= H) continue;
        const float o = __fdiv_rn(t[r], wsum[r]);
        const size_t oi = fo + (size_t)y * W + x;
        if (store_op == STORE_F32) ((float*)dst)[oi] = o;
        else if (store_op == STORE_U16) ((uint16_t*)dst)[oi] = sat_u16(cvround(o));
        else ((int16_t*)dst)[oi] = (int16_t)sat_s16(cvround(o));
    }
    }
}

template <int RAD, int MODE>
int launch_rad_mode(const void* src, void* dst, int n, int H, int W, float th, float maf, int load_op, int store_op, int quirk, cudaStream_t s) {
    // the unrolled body (ntaps * R * 5, or * 3 in integer mode, instructions) has to stay inside the instruction cache
    constexpr int R = MODE == 0 ? (RAD <= 3 ? 8 : (RAD <= 5 ? 4 : (RAD <= 7 ? 2 : 1))) : (RAD <= 5 ? 8 : (RAD <= 7 ? 4 : 2));
    dim3 block(32, 8);
    if (R > 2 && (long)((W + kTW - 1) / kTW) * ((H + 2 * R - 1) / (2 * R)) * n < 2 * 148) {    // few tiles (single small frame): shorter tiles fill the GPU
        constexpr int RS = 2; dim3 grid((W + kTW - 1) / kTW, (H + 2 * RS - 1) / (2 * RS), n);
        bwrf32f_tiled_kernel<RAD, RS, MODE><<<grid, block, 0, s>>>(src, dst, H, W, th, maf, load_op, store_op, quirk);
        return 1;
    }
    dim3 grid((W + kTW - 1) / kTW, (H + 2 * R - 1) / (2 * R), n);
    bwrf32f_tiled_kernel<RAD, R, MODE><<<grid, block, 0, s>>>(src, dst, H, W, th, maf, load_op, store_op, quirk);
    return 1;
}

template <int RAD>
int launch_rad(const void* src, void* dst, int n, int H, int W, float th, float maf, int load_op, int store_op, int quirk, cudaStream_t s) {
    // integer mode: integer source, th >= 0 (and not NaN), at most 255 taps (radius <= 9: 253 taps)
    const bool integer = (load_op == LOAD_U16 || load_op == LOAD_S16 || load_op == LOAD_U8) && th >= 0.f && RAD <= 9;
    if (integer) return load_op == LOAD_S16 ? launch_rad_mode<RAD, 2>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s)
                                            : launch_rad_mode<RAD, 1>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    return launch_rad_mode<RAD, 0>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
}

// Three channels (32FC3 / 16UC3 / 16SC3): same schedule, interleaved pixels in shared memory (a lane's stride of 3 floats
// is conflict-free), weight from the L1 distance over the channels in the reference's order of additions
// (binalyWeightedRangeFilter.cpp:555-663: |db| + |dg| first, then + |dr|, i.e. channel 2, 1, 0 here), one FFMA per channel.
template <int RAD, int R>
__global__ void __launch_bounds__(256) bwrf32f_c3_tiled_kernel(const void* __restrict__ src, void* __restrict__ dst, int H, int W,
                                                               float th, int load_op, int store_op, int quirk) {
    constexpr int TILE_H = 2 * R;
    constexpr int SWP = kTW + 2 * RAD, SW = SWP * 3 + 1, SH = TILE_H + 2 * RAD;
    __shared__ float sm[SH * SW];
    const size_t fo = (size_t)blockIdx.z * H * W * 3;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * TILE_H;
    for (int ty = threadIdx.y; ty < SH; ty += 8) {
        const int uy = Y0 - RAD + ty, gy = clampi(uy, 0, H - 1);
        const size_t rowi = fo + (size_t)gy * W * 3;
        for (int tx = threadIdx.x; tx < SWP; tx += 32) {
            const int ux = X0 - RAD + tx;
            const size_t gi = rowi + (size_t)clampi(ux, 0, W - 1) * 3;
            float* d = sm + ty * SW + tx * 3;
            d[0] = load_px(src, gi, load_op, 0.f); d[1] = load_px(src, gi + 1, load_op, 0.f); d[2] = load_px(src, gi + 2, load_op, 0.f);
            if (quirk && ux == W - 1 + RAD) {      // padding quirk, three planes (see bwrf32f_kernel in dmc_kernels_32f.cu)
                d[0] = load_px(src, rowi + 1, load_op, 0.f); d[1] = load_px(src, rowi + 2, load_op, 0.f);
                if (uy + 1 <= H - 1 + RAD) d[2] = load_px(src, fo + (size_t)clampi(uy + 1, 0, H - 1) * W * 3, load_op, 0.f);
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 3, wy = threadIdx.y >> 2;
    const int xl = 32 * wx + lane;
    const float* base = sm + (wy * R) * SW + xl * 3;
    float c[R][3], t[R][3], wsum[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { c[r][ch] = base[(r + RAD) * SW + RAD * 3 + ch]; t[r][ch] = 0.f; }
        wsum[r] = 0.f;
    }
#pragma unroll
    for (int yy = 0; yy < R + 2 * RAD; yy++) {
        float v[(2 * RAD + 1) * 3];
#pragma unroll
        for (int i = 0; i < (2 * RAD + 1) * 3; i++) v[i] = base[yy * SW + i];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
            if (ady <= RAD) {
#pragma unroll
                for (int dx = -hw_of(RAD, ady); dx <= hw_of(RAD, ady); dx++) {
                    const float v0 = v[(dx + RAD) * 3], v1 = v[(dx + RAD) * 3 + 1], v2 = v[(dx + RAD) * 3 + 2];
                    const float d = __fadd_rn(__fadd_rn(fabsf(__fsub_rn(c[r][2], v2)), fabsf(__fsub_rn(c[r][1], v1))), fabsf(__fsub_rn(c[r][0], v0)));
                    const float w = d <= th ? 1.f : 0.f;
                    t[r][0] = __fmaf_rn(w, v0, t[r][0]); t[r][1] = __fmaf_rn(w, v1, t[r][1]); t[r][2] = __fmaf_rn(w, v2, t[r][2]);
                    wsum[r] = __fadd_rn(wsum[r], w);
                }
            }
        }
    }
    const int x = X0 + xl;
    if (x >= W) return;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = Y0 + wy * R + r;
        if (y >= H) continue;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float o = __fdiv_rn(t[r][ch], wsum[r]);
            const size_t oi = fo + ((size_t)y * W + x) * 3 + ch;
            if (store_op == STORE_F32) ((float*)dst)[oi] = o;
            else if (store_op == STORE_U16) ((uint16_t*)dst)[oi] = sat_u16(cvround(o));
            else ((int16_t*)dst)[oi] = (int16_t)sat_s16(cvround(o));
        }
    }
}

template <int RAD>
int launch_rad_c3(const void* src, void* dst, int n, int H, int W, float th, int load_op, int store_op, int quirk, cudaStream_t s) {
    constexpr int R = RAD <= 3 ? 8 : (RAD <= 5 ? 4 : (RAD <= 7 ? 2 : 1));       // about 11 instructions per tap and output row in the unrolled body
    constexpr int RR = ((kTW + 2 * RAD) * 3 + 1) * (2 * R + 2 * RAD) * 4 <= 48 * 1024 ? R : R / 2;      // static shared memory limit
    dim3 block(32, 8), grid((W + kTW - 1) / kTW, (H + 2 * RR - 1) / (2 * RR), n);
    bwrf32f_c3_tiled_kernel<RAD, RR><<<grid, block, 0, s>>>(src, dst, H, W, th, load_op, store_op, quirk);
    return 1;
}

}  // namespace

int launch_bwrf32f_tiled(const void* src, void* dst, int n, int H, int W, int radius, float th, int load_op, float maf, int store_op, cudaStream_t s) {
    const int quirk = (radius % 8 == 5) && (W % 4 == 0);
    switch (radius) {
    case 1: return launch_rad<1>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 2: return launch_rad<2>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 3: return launch_rad<3>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 4: return launch_rad<4>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 5: return launch_rad<5>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 6: return launch_rad<6>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 7: return launch_rad<7>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 8: return launch_rad<8>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 9: return launch_rad<9>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    case 10: return launch_rad<10>(src, dst, n, H, W, th, maf, load_op, store_op, quirk, s);
    }
    return 0;
}

int launch_bwrf32f_c3_tiled(const void* src, void* dst, int n, int H, int W, int radius, float th, int load_op, int store_op, cudaStream_t s) {
    const int quirk = (radius % 8 == 5) && (W % 4 == 0);
    if (load_op != LOAD_F32 && load_op != LOAD_U16 && load_op != LOAD_S16) return 0;
    switch (radius) {
    case 1: return launch_rad_c3<1>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 2: return launch_rad_c3<2>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 3: return launch_rad_c3<3>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 4: return launch_rad_c3<4>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 5: return launch_rad_c3<5>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 6: return launch_rad_c3<6>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 7: return launch_rad_c3<7>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 8: return launch_rad_c3<8>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 9: return launch_rad_c3<9>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    case 10: return launch_rad_c3<10>(src, dst, n, H, W, th, load_op, store_op, quirk, s);
    }
    return 0;
}

}  // namespace dmc
