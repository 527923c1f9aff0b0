// dmc_bwrf32f_tiled.cu -- register-tiled fast path of the 32-bit binary-weighted range filter (single channel, square
// window of radius 1..7): the kernel behind filterDisp8U2Depth32F / Depth16U / Disp32F and the 16U/16S/32F
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

template <int RAD, int R>
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
            sm[ty * SW + tx] = load_op == LOAD_U8_DISP2DEPTH ? s_lut[((const uint8_t*)src)[gi]] : load_px(src, gi, load_op, maf);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x, wx = threadIdx.y & 3, wy = threadIdx.y >> 2;
    const int xl = 32 * wx + lane;
    const float* base = sm + (wy * R) * SW + xl;                       // staged column of pixel x - RAD, first input row of the block
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
    const int x = X0 + xl;
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

template <int RAD>
int launch_rad(const void* src, void* dst, int n, int H, int W, float th, float maf, int load_op, int store_op, int quirk, cudaStream_t s) {
    constexpr int R = RAD <= 3 ? 8 : (RAD <= 5 ? 4 : 2);       // unrolled body (ntaps * R * 5 instructions) stays inside the instruction cache
    dim3 block(32, 8);
    if (R > 2 && (long)((W + kTW - 1) / kTW) * ((H + 2 * R - 1) / (2 * R)) * n < 2 * 148) {    // few tiles (single small frame): shorter tiles fill the GPU
        constexpr int RS = 2; dim3 grid((W + kTW - 1) / kTW, (H + 2 * RS - 1) / (2 * RS), n);
        bwrf32f_tiled_kernel<RAD, RS><<<grid, block, 0, s>>>(src, dst, H, W, th, maf, load_op, store_op, quirk);
        return 1;
    }
    dim3 grid((W + kTW - 1) / kTW, (H + 2 * R - 1) / (2 * R), n);
    bwrf32f_tiled_kernel<RAD, R><<<grid, block, 0, s>>>(src, dst, H, W, th, maf, load_op, store_op, quirk);
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
    }
    return 0;
}

}  // namespace dmc
