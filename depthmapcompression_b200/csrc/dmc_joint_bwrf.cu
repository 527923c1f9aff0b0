// dmc_joint_bwrf.cu -- joint ("guided") binary-weighted range filter: an 8-bit single-channel image (depth / disparity)
// is averaged over the circular window with weights taken from a GUIDE image (8UC3 colour or 8UC1) of the same size:
//     w(q) = [ dist(guide(q), guide(p)) <= th ],   dist = min(255, |db| + |dg| + |dr|)  (C3)   or   |dg|  (C1)
//     out(p) = RNE( float(sum w*src(q)) / float(sum w) )
// This is row (f)-4 of SURVEY.md section 8 and what BASELINE.json's north_star calls "binary weighted range filter guided
// by the colour image".  The reference does NOT implement it (its binalyWeightedRangeFilter weights an image by itself),
// so it lives under a new name and has no reference parity; every rule is borrowed from the reference's own filter:
// window and taps of binalyWeightedRangeFilter.cpp:1066-1076, BORDER_REPLICATE, the saturated L1 colour distance of the
// 8UC3 filter (:297-301), the FP32 division and cvtps rounding of the 8UC1 filter (:165-216).  With guide == src it IS
// binalyWeightedRangeFilter on 8UC1, which is how tests/ pin it to the reference.
//
// Fast path (square window, radius 1..5): packed half2 like dmc_bwrf8u_c3_h2.cu -- per pixel pair and tap 3 HADD2
// (guide differences), 2 HADD2 (L1), HSET2 (weight), HADD2 + HFMA2 (S += w*(v - c) on the filtered image), LEA.HI (count).
// |v - c| is not bounded by th here, so S is folded into FP32 after every 8 taps (8 * 255 <= 2048 keeps fp16 exact).
#include "dmc_common.cuh"
#include "dmc_kernels.cuh"

namespace dmc {

namespace {

constexpr int kGX = 64, kGY = 16;   // output tile of the generic kernel

template <int GCN>
__global__ void __launch_bounds__(256) joint_bwrf8u_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ guide, uint8_t* __restrict__ dst,
                                                           int H, int W, RowSpan rs, int th) {
    extern __shared__ unsigned char smraw[];
    const int rH = rs.rH, rV = rs.rV, TW = kGX + 2 * rH, TH = kGY + 2 * rV;
    uint8_t* sv = smraw;                       // TH x TW filtered image
    uint8_t* sg = smraw + (size_t)TW * TH;     // TH x TW x GCN guide, interleaved
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * kGX, y0 = blockIdx.y * kGY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int idx = tid; idx < TW * TH; idx += 256) {
        int ty = idx / TW, tx = idx - ty * TW;
        const size_t gi = fo + (size_t)clampi(y0 - rV + ty, 0, H - 1) * W + clampi(x0 - rH + tx, 0, W - 1);
        sv[idx] = src[gi];
#pragma unroll
        for (int c = 0; c < GCN; c++) sg[idx * GCN + c] = guide[gi * GCN + c];
    }
    __syncthreads();
    for (int ly = threadIdx.y; ly < kGY; ly += blockDim.y)
        for (int lx = threadIdx.x; lx < kGX; lx += blockDim.x) {
            const int x = x0 + lx, y = y0 + ly;
            if (x >= W || y >= H) continue;
            const int ci = (ly + rV) * TW + lx + rH;
            int g0[GCN], sum = 0, cnt = 0;
#pragma unroll
            for (int c = 0; c < GCN; c++) g0[c] = sg[ci * GCN + c];
            for (int i = -rV; i <= rV; i++) {
                const int hw = rs.hw[i + rV];
                for (int j = -hw; j <= hw; j++) {
                    const int ti = ci + i * TW + j;
                    int d = 0;
#pragma unroll
                    for (int c = 0; c < GCN; c++) d += abs((int)sg[ti * GCN + c] - g0[c]);
                    if (GCN > 1) d = min(d, 255);
                    const int w = d <= th;
                    sum += w ? (int)sv[ti] : 0; cnt += w;
                }
            }
            dst[fo + (size_t)y * W + x] = sat_u8(sat_s16(cvround(__fdiv_rn((float)sum, (float)cnt))));
        }
}

// ---- packed fast path ---------------------------------------------------------------------------------------------
constexpr int kHalo = 8, kTileW = 128;

__host__ __device__ constexpr int hw_of(int rad, int dy) {
    int lim = rad * rad - dy * dy, j = 0;
    while ((j + 1) * (j + 1) <= lim) j++;
    return j;
}

template <int RAD, int R, int GCN>
__global__ void __launch_bounds__(256) joint_bwrf8u_h2_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ guide, uint8_t* __restrict__ dst,
                                                              int H, int W, int th) {
    constexpr int TILE_H = 4 * R, SW = kTileW + 2 * kHalo, SH = TILE_H + 2 * RAD, SWW = SW / 2, PL = SH * SWW, NP = GCN + 1;
    __shared__ __align__(16) uint32_t sm[NP * PL];          // planes 0..GCN-1: guide channels, plane GCN: the filtered image
    const size_t fo = (size_t)blockIdx.z * H * W;
    const int X0 = blockIdx.x * kTileW, Y0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int idx = tid; idx < SH * (SW / 4); idx += 256) {
        const int ty = idx / (SW / 4), tq = idx - ty * (SW / 4);
        const size_t rowi = fo + (size_t)clampi(Y0 - RAD + ty, 0, H - 1) * W;
        const int gx = X0 - kHalo + 4 * tq;
        uint32_t b[NP][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t gi = rowi + clampi(gx + k, 0, W - 1);
#pragma unroll
            for (int c = 0; c < GCN; c++) b[c][k] = guide[gi * GCN + c];
            b[GCN][k] = src[gi];
        }
#pragma unroll
        for (int c = 0; c < NP; c++) {
            uint2 o;
            o.x = 0x64006400u | b[c][0] | (b[c][1] << 16);      // 0x6400 | byte = fp16 1024 + byte
            o.y = 0x64006400u | b[c][2] | (b[c][3] << 16);
            *(uint2*)&sm[c * PL + ty * SWW + 2 * tq] = o;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x, wx = threadIdx.y & 1, wy = threadIdx.y >> 1;
    const int xl = 64 * wx + 2 * lane;
    const int x = X0 + xl;
    if (x >= W) return;                                      // (no barrier below)
    const uint32_t* base = sm + (wy * R) * SWW + (xl + kHalo - 6) / 2;
    const __half2 th2 = __half2half2(__int2half_rn((GCN > 1 && th >= 255) ? 765 : th));      // saturated distance <= 255 always passes th = 255

    __half2 c[R][NP], S[R]; uint32_t N15[R]; float Sf0[R], Sf1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int ch = 0; ch < NP; ch++) { uint32_t cw = base[ch * PL + (r + RAD) * SWW + 3]; c[r][ch] = *reinterpret_cast<__half2*>(&cw); }
        S[r] = __float2half2_rn(0.f); N15[r] = 0u; Sf0[r] = 0.f; Sf1[r] = 0.f;
    }
    auto fold = [&](int r) { const float2 f = __half22float2(S[r]); Sf0[r] += f.x; Sf1[r] += f.y; S[r] = __float2half2_rn(0.f); };
#pragma unroll
    for (int yy = 0; yy < R + 2 * RAD; yy++) {
        uint32_t wd[NP][7];
#pragma unroll
        for (int ch = 0; ch < NP; ch++)
#pragma unroll
            for (int i = 0; i < 7; i++) wd[ch][i] = base[ch * PL + yy * SWW + i];
#pragma unroll
        for (int dx = -RAD; dx <= RAD; dx++) {
            __half2 v[NP];
#pragma unroll
            for (int ch = 0; ch < NP; ch++) {
                uint32_t vb = (dx & 1) == 0 ? wd[ch][(dx + 6) / 2] : __byte_perm(wd[ch][(dx + 5) / 2], wd[ch][(dx + 7) / 2], 0x5432);
                v[ch] = *reinterpret_cast<__half2*>(&vb);
            }
            const int adx = dx < 0 ? -dx : dx;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int dy = yy - r - RAD, ady = dy < 0 ? -dy : dy;
                if (ady <= RAD && adx <= hw_of(RAD, ady)) {
                    __half2 l1 = __habs2(__hsub2(v[0], c[r][0]));
#pragma unroll
                    for (int ch = 1; ch < GCN; ch++) l1 = __hadd2(l1, __habs2(__hsub2(v[ch], c[r][ch])));
                    const __half2 w = __hle2(l1, th2);
                    S[r] = __hfma2(w, __hsub2(v[GCN], c[r][GCN]), S[r]);
                    N15[r] += (*reinterpret_cast<const uint32_t*>(&w)) >> 10;
                    if (((dx + hw_of(RAD, ady)) & 7) == 7) fold(r);      // every 8th tap of this window row: |S| <= 8 * 255
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) { const int dy = yy - r - RAD; if (dy >= -RAD && dy <= RAD) fold(r); }
    }
    uint8_t* op = dst + fo + (size_t)(Y0 + wy * R) * W + x;
    const int yrem = H - (Y0 + wy * R);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float2 cf = __half22float2(c[r][GCN]);
        const float n0 = (float)(N15[r] & 0xFFFFu), n1 = (float)(N15[r] >> 16);
        const float t0 = (cf.x - 1024.f) * n0 + 15.f * Sf0[r], t1 = (cf.y - 1024.f) * n1 + 15.f * Sf1[r];      // 15*(c*N + S), exact
        if (r < yrem) {
            op[0] = (uint8_t)__float_as_uint(__fdiv_rn(t0, n0) + 12582912.f);
            if (x + 1 < W) op[1] = (uint8_t)__float_as_uint(__fdiv_rn(t1, n1) + 12582912.f);
        }
        op += W;
    }
}

template <int RAD, int GCN>
void launch_h2(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int n, int H, int W, int th, cudaStream_t s) {
    constexpr int R = RAD <= 2 ? 4 : 2;          // ntaps * R * (6 + 2*GCN) instructions must stay inside the instruction cache
    dim3 grid((W + kTileW - 1) / kTileW, (H + 4 * R - 1) / (4 * R), n), block(32, 8);
    joint_bwrf8u_h2_kernel<RAD, R, GCN><<<grid, block, 0, s>>>(src, guide, dst, H, W, th);
}

template <int GCN>
int launch_h2_rad(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int n, int H, int W, int radius, int th, cudaStream_t s) {
    switch (radius) {
    case 1: launch_h2<1, GCN>(src, guide, dst, n, H, W, th, s); return 1;
    case 2: launch_h2<2, GCN>(src, guide, dst, n, H, W, th, s); return 1;
    case 3: launch_h2<3, GCN>(src, guide, dst, n, H, W, th, s); return 1;
    case 4: launch_h2<4, GCN>(src, guide, dst, n, H, W, th, s); return 1;
    case 5: launch_h2<5, GCN>(src, guide, dst, n, H, W, th, s); return 1;
    }
    return 0;
}

}  // namespace

int launch_joint_bwrf8u(const uint8_t* src, const uint8_t* guide, uint8_t* dst, int n, int H, int W, int gcn, const RowSpan& rs, int th, cudaStream_t s) {
    if (gcn != 1 && gcn != 3) return 0;
    if (rs.rH == rs.rV && rs.rH >= 1 && rs.rH <= 5 && th >= 0) {
        const int nk = gcn == 3 ? launch_h2_rad<3>(src, guide, dst, n, H, W, rs.rH, th, s) : launch_h2_rad<1>(src, guide, dst, n, H, W, rs.rH, th, s);
        if (nk) return nk;
    }
    dim3 grid((W + kGX - 1) / kGX, (H + kGY - 1) / kGY, n), block(32, 8);
    const size_t smem = (size_t)(kGX + 2 * rs.rH) * (kGY + 2 * rs.rV) * (1 + gcn);
    if (gcn == 3) joint_bwrf8u_kernel<3><<<grid, block, smem, s>>>(src, guide, dst, H, W, rs, th);
    else joint_bwrf8u_kernel<1><<<grid, block, smem, s>>>(src, guide, dst, H, W, rs, th);
    return 1;
}

}  // namespace dmc
