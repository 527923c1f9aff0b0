// dmc_bwrf8u_c3_h2.cu -- packed-half fast path of the 8-bit, 3-channel binary-weighted range filter
// (binalyWeightedRangeFilter.cpp:237-462): the weight is common to the three channels,
//     w = [ min(255, |db| + |dg| + |dr|) <= th ]        (saturating byte adds, :297-301)
// and every channel is averaged with it.  Same exact-fp16 formulation as the single-channel kernel
// (dmc_bwrf8u_h2.cu): per channel S_c = sum w*(v_c - c_c) with |w*(v_c - c_c)| <= th, N = sum w, exact while
// ntaps*th <= 2048; two pixels per instruction.  Per pixel pair and tap: 3 HADD2 (differences), 2 HADD2 with |.|
// modifiers (L1 distance, <= 765: exact), 1 HSET2, 3 HFMA2, 1 LEA.HI = 10 instructions for 3 channels.
// The interleaved BGR bytes are de-interleaved into three biased-half planes while the tile is staged.
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"

namespace dmc {

namespace {

constexpr int kHalo = 8, kTileW = 128;

__host__ __device__ constexpr int hw_of(int rad, int dy) {
    int lim = rad * rad - dy * dy, j = 0;
    while ((j + 1) * (j + 1) <= lim) j++;
    return j;
}

template <int RAD, int R, bool FLUSH>     // FLUSH: see dmc_bwrf8u_h2.cu
__global__ void __launch_bounds__(256) bwrf8u_c3_h2_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int th) {
    constexpr int TILE_H = 4 * R, SW = kTileW + 2 * kHalo, SH = TILE_H + 2 * RAD, SWW = SW / 2, PL = SH * SWW;
    __shared__ __align__(16) uint32_t sm[3 * PL];          // plane c at sm + c*PL
    const size_t fo = (size_t)blockIdx.z * H * W * 3;
    const uint8_t* fsrc = src + fo;
    const int X0 = blockIdx.x * kTileW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool fast_rows = (W & 3) == 0 && (fo & 3) == 0 && (reinterpret_cast<size_t>(src) & 3) == 0;
    for (int idx = tid; idx < SH * (SW / 4); idx += 256) {
        int ty = idx / (SW / 4), tq = idx - ty * (SW / 4);
        int gy = clampi(Y0 - RAD + ty, 0, H - 1), gx = X0 - kHalo + 4 * tq;
        const uint8_t* row = fsrc + (size_t)gy * W * 3;
        uint32_t b[12];
        if (fast_rows && gx >= 0 && gx + 3 < W) {
            const uint32_t* p = (const uint32_t*)(row + (size_t)gx * 3);
            const uint32_t w[3] = {p[0], p[1], p[2]};
#pragma unroll
            for (int k = 0; k < 12; k++) b[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint8_t* q = row + (size_t)clampi(gx + k, 0, W - 1) * 3; b[3 * k] = q[0]; b[3 * k + 1] = q[1]; b[3 * k + 2] = q[2]; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            uint2 o;
            o.x = 0x64006400u | b[c] | (b[3 + c] << 16);
            o.y = 0x64006400u | b[6 + c] | (b[9 + c] << 16);
            *(uint2*)&sm[c * PL + ty * SWW + 2 * tq] = o;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    constexpr int E = RAD <= 6 ? 6 : 8, NW = E + 1;          // words (x-E, x-E+1) .. (x+E, x+E+1) cover offsets -E .. E+1
    const uint32_t* base = sm + (wy * R) * SWW + (xl + kHalo - E) / 2;
    const __half2 th2 = __half2half2(__int2half_rn(th >= 255 ? 765 : th));      // saturated distance <= 255 always passes th = 255

    __half2 c[R][3], S[R][3]; uint32_t N15[R]; float Sf0[FLUSH ? R : 1][3], Sf1[FLUSH ? R : 1][3];
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) { uint32_t cw = base[ch * PL + (r + RAD) * SWW + E / 2]; c[r][ch] = *reinterpret_cast<__half2*>(&cw); S[r][ch] = __float2half2_rn(0.f); if (FLUSH) { Sf0[r][ch] = 0.f; Sf1[r][ch] = 0.f; } }
        N15[r] = 0u;
    }
#pragma unroll
    for (int yy = 0; yy < R + 2 * RAD; yy++) {
        uint32_t wd[3][NW];
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
#pragma unroll
            for (int i = 0; i < NW; i++) wd[ch][i] = base[ch * PL + yy * SWW + i];
#pragma unroll
        for (int dx = -RAD; dx <= RAD; dx++) {
            __half2 v[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                uint32_t vb = (dx & 1) == 0 ? wd[ch][(dx + E) / 2] : __byte_perm(wd[ch][(dx + E - 1) / 2], wd[ch][(dx + E + 1) / 2], 0x5432);
                v[ch] = *reinterpret_cast<__half2*>(&vb);
            }
            const int adx = dx < 0 ? -dx : dx;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
                if (ady <= RAD && adx <= hw_of(RAD, ady)) {
                    const __half2 d0 = __hsub2(v[0], c[r][0]), d1 = __hsub2(v[1], c[r][1]), d2 = __hsub2(v[2], c[r][2]);
                    const __half2 l1 = __hadd2(__hadd2(__habs2(d0), __habs2(d1)), __habs2(d2));
                    const __half2 w = __hle2(l1, th2);
                    S[r][0] = __hfma2(w, d0, S[r][0]); S[r][1] = __hfma2(w, d1, S[r][1]); S[r][2] = __hfma2(w, d2, S[r][2]);
                    N15[r] += (*reinterpret_cast<const uint32_t*>(&w)) >> 10;
                }
            }
        }
        if (FLUSH) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD;
                if (dy >= -RAD && dy <= RAD) {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) { const float2 f = __half22float2(S[r][ch]); Sf0[r][ch] += f.x; Sf1[r][ch] += f.y; S[r][ch] = __float2half2_rn(0.f); }
                }
            }
        }
    }
    const int x = X0 + xl;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int y = Y0 + wy * R + r;
        if (y >= H || x >= W) continue;
        const float n0 = (float)(N15[r] & 0xFFFFu), n1 = (float)(N15[r] >> 16);
        uint8_t* o = dst + fo + ((size_t)y * W + x) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float2 cf = __half22float2(c[r][ch]);
            const float2 sf = FLUSH ? make_float2(Sf0[r][ch], Sf1[r][ch]) : __half22float2(S[r][ch]);
            const float t0 = (cf.x - 1024.f) * n0 + 15.f * sf.x, t1 = (cf.y - 1024.f) * n1 + 15.f * sf.y;      // 15*(c*N + S), exact
            o[ch] = (uint8_t)__float_as_uint(__fdiv_rn(t0, n0) + 12582912.f);
            if (x + 1 < W) o[3 + ch] = (uint8_t)__float_as_uint(__fdiv_rn(t1, n1) + 12582912.f);
        }
    }
}

template <int RAD>
int launch_rad(const uint8_t* src, uint8_t* dst, int n, int H, int W, int th, bool flush, cudaStream_t s) {
    constexpr int R = RAD <= 2 ? 4 : (RAD <= 5 ? 2 : 1);          // ntaps * R * 10 instructions must stay inside the instruction cache
    dim3 grid((W + kTileW - 1) / kTileW, (H + 4 * R - 1) / (4 * R), n), block(32, 8);
    if (flush) bwrf8u_c3_h2_kernel<RAD, R, true><<<grid, block, 0, s>>>(src, dst, H, W, th);
    else bwrf8u_c3_h2_kernel<RAD, R, false><<<grid, block, 0, s>>>(src, dst, H, W, th);
    return 1;
}

}  // namespace

int launch_bwrf8u_c3_h2(const uint8_t* src, uint8_t* dst, int n, int H, int W, int radius, int th, int ntaps, cudaStream_t s) {
    if (radius < 1 || radius > 7 || th < 0 || (long)(2 * radius + 1) * th > 2048) return 0;
    const bool flush = (long)ntaps * th > 2048;
    switch (radius) {
    case 1: return launch_rad<1>(src, dst, n, H, W, th, flush, s);
    case 2: return launch_rad<2>(src, dst, n, H, W, th, flush, s);
    case 3: return launch_rad<3>(src, dst, n, H, W, th, flush, s);
    case 4: return launch_rad<4>(src, dst, n, H, W, th, flush, s);
    case 5: return launch_rad<5>(src, dst, n, H, W, th, flush, s);
    case 6: return launch_rad<6>(src, dst, n, H, W, th, flush, s);
    case 7: return launch_rad<7>(src, dst, n, H, W, th, flush, s);
    }
    return 0;
}

}  // namespace dmc
