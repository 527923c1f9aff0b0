// dmc_bwrf8u_h2.cu -- the hot kernel of the chain: 8-bit binary-weighted range filter, square window (circular tap
// set), in packed half-precision SIMD.
//
// Formulation.  For a centre c and a tap v (both 0..255):  w = [|v-c| <= th],  sum w*v = c * sum w + sum w*(v-c).
// With d = v-c, |w*d| <= th, so  S = sum w*d  and  N = sum w  stay integers of magnitude <= ntaps*th and <= ntaps.
// When ntaps*th <= 2048 every partial sum is exactly representable in fp16, so S and N can be accumulated with
// HFMA2 on two pixels at once -- 4 instructions per pixel pair per tap, spread over two pipes (measured on B200 with
// tools/ubench_pipes.cu: HFMA2/HADD2 issue at 2 warp-instructions/clk/SM on the FMA pipe, HSET2/LEA/PRMT at 2/clk/SM
// on the ALU pipe, so two of each per tap keeps both pipes and the 4/clk issue port equally busy):
//     d = v - c (HFMA2)   w = (|d| <= th) ? 1.0 : 0.0 (HSET2.BF)   S += w*d (HFMA2)   N15 += bits(w) >> 10 (LEA.HI)
// (bits(1.0h) >> 10 = 15 in each 16-bit lane, so N15 counts 15 per accepted tap: 15*317 < 65536, no lane overflow)
// and the result  RNE(float(c*N + S) / float(N))  is bit-identical to the reference's FP32 sums (which are exact
// integers as well) followed by _mm_div_ps / _mm_cvtps_epi32 (binalyWeightedRangeFilter.cpp:165-216).
//
// Data layout.  The CTA stages its input tile (+halo) in shared memory as fp16 with a +1024 bias (0x6400 | byte):
// the bias makes the byte->half conversion a byte permute and cancels in v - c.  A thread owns a 2-pixel-wide, R-row
// tall block of outputs; for every input row it loads 7 aligned half2 words, funnel-shifts the odd offsets, and
// reuses each tap vector for all the output rows whose window contains it (register tiling: ~0.5 LDS per 2x81 taps).
// R = 4 at radius >= 4 keeps the fully unrolled body inside the 32 KB instruction cache (R = 8 ran 5x slower).  A
// variant with a run-time loop over the horizontal offset, R = 8 and two shifted shared-memory copies was measured
// too: fewer loads per tap but more staging/loop instructions, 17.9 ms vs 16.0 ms per 1000 frames -- not kept.
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"
#include "dmc_stage.cuh"

namespace dmc {

namespace {

constexpr int kHalo = kHalo16;    // staged halo in pixels (>= radius + 1; 16 keeps the staging loads aligned)
constexpr int kTileW = kTW;       // output tile width: 2 warps x 32 lanes x 2 pixels

__host__ __device__ constexpr int hw_of(int rad, int dy) {      // circle_halfwidth as a constant expression
    int lim = rad * rad - dy * dy, j = 0;
    while ((j + 1) * (j + 1) <= lim) j++;
    return j;
}

// FLUSH: for thresholds with ntaps*th > 2048 the half accumulator S is folded into an FP32 accumulator after every
// input row (a row contributes at most 2*RAD+1 taps, so |S| <= (2*RAD+1)*th <= 2048 stays exact); +5 instructions
// per output row and input row.
template <int RAD, int R, bool FLUSH, bool EVEN>     // EVEN: W and dst allow 2-byte stores
__global__ void __launch_bounds__(256) bwrf8u_h2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int th) {
    constexpr int TILE_H = 4 * R;                         // 4 warp rows
    constexpr int SH = TILE_H + 2 * RAD;
    constexpr int SWW = kSW16 / 2;                        // row stride in 32-bit words
    __shared__ __align__(16) uint32_t sm[SH * SWW];

    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTileW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    stage_tile16<SH, 1, false>(sm, src + fo, X0, Y0 - RAD, H, W, aligned16(src, W), tid);      // biased half2 words (0x6400 | byte)
    __syncthreads();

    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;                    // first pixel of this thread's pair inside the tile
    constexpr int E = RAD <= 6 ? 6 : (RAD <= 8 ? 8 : 10), NW = E + 1;     // words (x-E, x-E+1) .. (x+E, x+E+1) cover offsets -E .. E+1
    const uint32_t* base = sm + (wy * R) * SWW + (xl + kHalo - E) / 2;      // word holding pixels (x-E, x-E+1)
    const __half2 th2 = __half2half2(__int2half_rn(th));

    __half2 c[R], S[R]; uint32_t N15[R]; float Sf0[FLUSH ? R : 1], Sf1[FLUSH ? R : 1];
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t cw = base[(r + RAD) * SWW + E / 2];      // pixels (x, x+1) of output row r
        c[r] = *reinterpret_cast<__half2*>(&cw);
        S[r] = __float2half2_rn(0.f); N15[r] = 0u;
        if (FLUSH) { Sf0[r] = 0.f; Sf1[r] = 0.f; }
    }

#pragma unroll
    for (int yy = 0; yy < R + 2 * RAD; yy++) {
        uint32_t wd[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) wd[i] = base[yy * SWW + i];
#pragma unroll
        for (int dx = -RAD; dx <= RAD; dx++) {
            // tap vector for offset dx: pixels (x+dx, x+dx+1)
            uint32_t vb;
            if ((dx & 1) == 0) vb = wd[(dx + E) / 2];
            else vb = __byte_perm(wd[(dx + E - 1) / 2], wd[(dx + E + 1) / 2], 0x5432);
            const __half2 v = *reinterpret_cast<__half2*>(&vb);
            const int adx = dx < 0 ? -dx : dx;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
                if (ady <= RAD && adx <= hw_of(RAD, ady)) {
                    const __half2 d = __hsub2(v, c[r]);
                    const __half2 w = __hle2(__habs2(d), th2);
                    S[r] = __hfma2(w, d, S[r]);
                    N15[r] += (*reinterpret_cast<const uint32_t*>(&w)) >> 10;
                }
            }
        }
        if (FLUSH) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD;
                if (dy >= -RAD && dy <= RAD) { const float2 f = __half22float2(S[r]); Sf0[r] += f.x; Sf1[r] += f.y; S[r] = __float2half2_rn(0.f); }
            }
        }
    }

    // ---- epilogue: out = RNE(float(c*N + S) / float(N)) ----
    const int x = X0 + xl;
    if (x >= W) return;
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);
#pragma unroll
    for (int r = 0; r < R; r++) {
        // 15*(c*N + S) / (15*N): the factor 15 of the packed counter cancels exactly in the IEEE division of two exactly
        // represented integers (< 2^24), so N15 is never divided by 15.  RNE by the 1.5*2^23 magic add (F2I is slow).
        const float2 cf = __half22float2(c[r]);
        const float2 sf = FLUSH ? make_float2(Sf0[r], Sf1[r]) : __half22float2(S[r]);
        const float n0 = (float)(N15[r] & 0xFFFFu), n1 = (float)(N15[r] >> 16);
        const float t0 = (cf.x - 1024.f) * n0 + 15.f * sf.x, t1 = (cf.y - 1024.f) * n1 + 15.f * sf.y;
        const uint32_t o0 = __float_as_uint(__fdiv_rn(t0, n0) + 12582912.f), o1 = __float_as_uint(__fdiv_rn(t1, n1) + 12582912.f);
        if (r < yrem) {
            if (EVEN) *reinterpret_cast<uint16_t*>(op) = (uint16_t)__byte_perm(o0, o1, 0x4440);
            else { op[0] = (uint8_t)o0; if (x + 1 < W) op[1] = (uint8_t)o1; }
        }
        op += W;
    }
}

template <int RAD, int R>
void launch_rr(const uint8_t* src, uint8_t* dst, int n, int H, int W, int th, bool flush, cudaStream_t s) {
    dim3 block(32, 8), grid((W + kTileW - 1) / kTileW, (H + 4 * R - 1) / (4 * R), n);
    const bool even = (W & 1) == 0 && (reinterpret_cast<size_t>(dst) & 1) == 0;
    if (flush) { if (even) bwrf8u_h2_kernel<RAD, R, true, true><<<grid, block, 0, s>>>(src, dst, H, W, th); else bwrf8u_h2_kernel<RAD, R, true, false><<<grid, block, 0, s>>>(src, dst, H, W, th); }
    else { if (even) bwrf8u_h2_kernel<RAD, R, false, true><<<grid, block, 0, s>>>(src, dst, H, W, th); else bwrf8u_h2_kernel<RAD, R, false, false><<<grid, block, 0, s>>>(src, dst, H, W, th); }
}

template <int RAD>
int launch_rad(const uint8_t* src, uint8_t* dst, int n, int H, int W, int th, bool flush, cudaStream_t s) {
    constexpr int R = RAD <= 3 ? 8 : (RAD <= 6 ? 4 : (RAD <= 8 ? 2 : 1));     // keeps the unrolled body (ntaps * R * 4 instructions) under the 32 KB instruction cache (R = 8 at RAD = 5 ran 5x slower)
    static_assert(RAD <= 10, "11 words per row cover offsets -10..11 only");
    if (R > 2 && (long)((W + kTileW - 1) / kTileW) * ((H + 4 * R - 1) / (4 * R)) * n < 2 * 148) launch_rr<RAD, 2>(src, dst, n, H, W, th, flush, s);   // few tiles (single small frame): shorter tiles fill the GPU
    else launch_rr<RAD, R>(src, dst, n, H, W, th, flush, s);
    return 1;
}

}  // namespace

int launch_bwrf8u_h2(const uint8_t* src, uint8_t* dst, int n, int H, int W, int radius, int th, int ntaps, cudaStream_t s) {
    if (radius < 1 || radius > 10 || th < 0 || (long)(2 * radius + 1) * th > 2048) return 0;
    const bool flush = (long)ntaps * th > 2048;
    switch (radius) {
    case 1: return launch_rad<1>(src, dst, n, H, W, th, flush, s);
    case 2: return launch_rad<2>(src, dst, n, H, W, th, flush, s);
    case 3: return launch_rad<3>(src, dst, n, H, W, th, flush, s);
    case 4: return launch_rad<4>(src, dst, n, H, W, th, flush, s);
    case 5: return launch_rad<5>(src, dst, n, H, W, th, flush, s);
    case 6: return launch_rad<6>(src, dst, n, H, W, th, flush, s);
    case 7: return launch_rad<7>(src, dst, n, H, W, th, flush, s);
    case 8: return launch_rad<8>(src, dst, n, H, W, th, flush, s);
    case 9: return launch_rad<9>(src, dst, n, H, W, th, flush, s);
    case 10: return launch_rad<10>(src, dst, n, H, W, th, flush, s);
    }
    return 0;
}

}  // namespace dmc
