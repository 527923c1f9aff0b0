// dmc_jpeg.cu -- frame-parallel baseline-JPEG decoder on sm_100a (row (f)-1 of SURVEY.md section 8: the decode that
// feeds the post filter set, done by libjpeg in the reference: main.cpp:284,521, jpegTurboDemo.cpp:217-271).
//
//   host   : parse the markers of every frame (DQT / SOF0 / DHT / DRI / SOS), derive and de-duplicate the tables
//   kernel 1: entropy decoding -- inherently serial inside a frame (no restart markers in cv2/libjpeg output), so
//             the parallelism is ACROSS frames: one warp (one active lane) per frame, thousands of frames in flight,
//             each walking its own bitstream; writes raw coefficients (int16, natural order) of every 8x8 block
//   kernel 2: dequantisation + IJG "islow" integer inverse DCT + range limit, one thread per block
//
// Output is bit-identical to libjpeg / libjpeg-turbo's default decoder (JDCT_ISLOW), checked against cv2.imdecode.
#include "dmc_common.cuh"
#include "dmc_jpeg_core.h"
#include "dmc_kernels.cuh"
#include <string.h>
#include <string>
#include <vector>

namespace dmc {

using namespace dmcjpeg;

// ---- host: marker parsing ----------------------------------------------------------------------------------------
// Derives the decoding tables of one DHT entry.  Returns false for tables libjpeg's jpeg_make_d_derived_tbl rejects with
// JERR_BAD_HUFF_TABLE: more than 256 symbols, or code lengths that do not form a prefix code (a code of length l must be
// < 2^l; an over-subscribed BITS array would otherwise index past look[]).
static bool derive_table(const uint8_t bits[17], const uint8_t* vals, int nvals, HuffTable* t) {
    memset(t, 0, sizeof *t);
    int huffsize[257], huffcode[257], p = 0;
    for (int l = 1; l <= 16; l++) {
        if (p + bits[l] > 256) return false;
        for (int i = 0; i < bits[l]; i++) huffsize[p++] = l;
    }
    if (p != nvals) return false;
    huffsize[p] = 0;
    int code = 0, si = huffsize[0]; p = 0;
    while (huffsize[p]) {
        while (huffsize[p] == si) huffcode[p++] = code++;
        if (code > (1 << si)) return false;                   // jdhuff.c: "code is now 1 more than the last code used for codelength si"
        code <<= 1; si++;
    }
    p = 0;
    for (int l = 1; l <= 16; l++) {
        if (bits[l]) { t->valoffset[l] = p - huffcode[p]; p += bits[l]; t->maxcode[l] = huffcode[p - 1]; }
        else t->maxcode[l] = -1;
    }
    t->maxcode[17] = 0xFFFFF;
    for (int i = 0; i < nvals; i++) t->huffval[i] = vals[i];
    p = 0;
    for (int l = 1; l <= 9; l++)
        for (int i = 0; i < bits[l]; i++, p++) {
            const int lookbits = huffcode[p] << (9 - l), span = 1 << (9 - l);
            if (lookbits + span > 512) return false;          // (cannot happen once the prefix-code check passed; belt and braces)
            for (int c = 0; c < span; c++) t->look[lookbits + c] = (uint16_t)((l << 8) | vals[p]);
        }
    return true;
}

template <class T> static int intern(std::vector<T>& pool, const T& v) {
    for (size_t i = 0; i < pool.size(); i++) if (memcmp(&pool[i], &v, sizeof(T)) == 0) return (int)i;
    pool.push_back(v); return (int)pool.size() - 1;
}

// Parses one stream [p, p+len).  Returns an empty string on success, else the reason it is not supported.
std::string jpeg_parse_frame(const uint8_t* p, uint64_t len, uint64_t blob_offset, int rows, int cols,
                             std::vector<QuantTable>& qpool, std::vector<HuffTable>& hpool, FrameDesc* d) {
    QuantTable qt[4]; bool have_q[4] = {false, false, false, false};
    HuffTable dc[4], ac[4]; bool have_dc[4] = {false, false, false, false}, have_ac[4] = {false, false, false, false};
    int comp_tq = -1, restart = 0; bool have_sof = false;
    if (len < 4 || p[0] != 0xFF || p[1] != 0xD8) return "no SOI marker";
    uint64_t i = 2;
    while (i + 4 <= len) {
        if (p[i] != 0xFF) return "marker expected";
        uint8_t m = p[i + 1];
        if (m == 0xFF) { i++; continue; }                       // fill byte
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { i += 2; continue; }     // TEM / RSTn: stand-alone markers without a length
        uint64_t seg = ((uint64_t)p[i + 2] << 8) | p[i + 3];    // includes the two length bytes
        if (seg < 2) return "bad segment length";
        if (i + 2 + seg > len) return "truncated segment";
        const uint8_t* s = p + i + 4; uint64_t sl = seg - 2;
        if (m == 0xDB) {                                        // DQT
            uint64_t k = 0;
            while (k < sl) {
                int pq = s[k] >> 4, tq = s[k] & 15; k++;
                if (tq > 3 || pq > 1) return "bad DQT";
                if (k + (pq ? 128u : 64u) > sl) return "truncated DQT";
                for (int z = 0; z < 64; z++) {
                    int v = pq ? ((s[k] << 8) | s[k + 1]) : s[k]; k += pq ? 2 : 1;
                    qt[tq].q[zigzag_to_natural(z)] = (uint16_t)v;
                }
                have_q[tq] = true;
            }
        } else if (m == 0xC4) {                                 // DHT
            uint64_t k = 0;
            while (k < sl) {
                if (k + 17 > sl) return "truncated DHT";
                int tc = s[k] >> 4, th = s[k] & 15; k++;
                if (th > 3 || tc > 1) return "bad DHT";
                uint8_t bits[17]; bits[0] = 0; int n = 0;
                for (int l = 1; l <= 16; l++) { bits[l] = s[k++]; n += bits[l]; }
                if (n > 256 || k + n > sl) return "bad DHT";
                if (tc == 0) for (int v = 0; v < n; v++) if (s[k + v] > 15) return "bad DHT (DC category > 15)";      // jdhuff.c rejects these too
                if (!derive_table(bits, s + k, n, tc ? &ac[th] : &dc[th])) return "bad DHT (code lengths do not form a prefix code)";
                k += n;
                (tc ? have_ac : have_dc)[th] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {                    // SOF0 / SOF1 (sequential Huffman)
            if (sl < 9 || s[0] != 8) return "only 8-bit precision is supported";
            int h = (s[1] << 8) | s[2], w = (s[3] << 8) | s[4];
            if (s[5] != 1) return "only single-component (grayscale) JPEG is supported";
            if (s[7] != 0x11) return "unexpected sampling factors";
            if (h != rows || w != cols) return "frame size differs from the batch size";
            comp_tq = s[8]; have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return "progressive / lossless / arithmetic JPEG is not supported";
        } else if (m == 0xDD) {                                 // DRI
            if (sl < 2) return "truncated DRI";
            restart = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {                                 // SOS: entropy-coded data follows
            if (!have_sof) return "SOS before SOF";
            if (sl < 6) return "truncated SOS";
            if (s[0] != 1) return "only single-component scans are supported";
            int td = s[2] >> 4, ta = s[2] & 15;
            if (td > 3 || ta > 3 || comp_tq < 0 || comp_tq > 3 || !have_dc[td] || !have_ac[ta] || !have_q[comp_tq]) return "scan refers to a missing table";
            d->scan_offset = blob_offset + i + 2 + seg; d->scan_end = blob_offset + len; d->restart_interval = restart;
            d->qt = intern(qpool, qt[comp_tq]);
            d->dc = intern(hpool, dc[td]); d->ac = intern(hpool, ac[ta]);
            return "";
        }
        i += 2 + seg;
    }
    return "no SOS marker";
}

// ---- kernel 1: entropy decoding, one warp (lane 0) per frame ----------------------------------------------------------
__global__ void __launch_bounds__(128) jpeg_huff_kernel(const uint8_t* __restrict__ blob, const FrameDesc* __restrict__ desc,
                                                        const HuffTable* __restrict__ ht, int16_t* __restrict__ coefs, int n, int blocks_per_frame) {
    // The decode loop is one dependent chain per frame (peek -> table look-up -> shift): ncu shows ~60 instructions per
    // symbol at ~5 cycles each, i.e. latency of dependent instructions, not memory.  The warp's 32 lanes first copy the
    // frame's two tables into shared memory (shorter look-up latency), then lane 0 decodes.
    __shared__ HuffTable s_tab[4][2];
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (f >= n) return;                              // uniform per warp
    const FrameDesc d = desc[f];
    {
        const uint32_t* g0 = reinterpret_cast<const uint32_t*>(&ht[d.dc]); const uint32_t* g1 = reinterpret_cast<const uint32_t*>(&ht[d.ac]);
        uint32_t* s0 = reinterpret_cast<uint32_t*>(&s_tab[w][0]); uint32_t* s1 = reinterpret_cast<uint32_t*>(&s_tab[w][1]);
        for (int i = lane; i < (int)(sizeof(HuffTable) / 4); i += 32) { s0[i] = g0[i]; s1[i] = g1[i]; }
    }
    __syncwarp();
    if (lane != 0) return;
    const HuffTable& dc = s_tab[w][0]; const HuffTable& ac = s_tab[w][1];
    BitReader br; br_init(br, blob, d.scan_offset, d.scan_end);
    int pred = 0, until_restart = d.restart_interval;
    int16_t* c = coefs + (size_t)f * blocks_per_frame * 64;
    for (int b = 0; b < blocks_per_frame; b++, c += 64) {
        if (d.restart_interval) {
            if (until_restart == 0) {       // byte-align, consume RSTn, reset the predictor (jdhuff.c process_restart)
                if (!br.marker) { br.nbits = 0; br_fill(br); }       // (all data bits are consumed; make the feeder meet the marker)
                if (br.marker >= 0xD0 && br.marker <= 0xD7) br.pos += 2;
                br.acc = 0; br.nbits = 0; br.marker = 0; pred = 0; until_restart = d.restart_interval;
            }
            until_restart--;
        }
        decode_block(br, dc, ac, pred, c);
    }
}

// ---- kernel 2: dequantise + islow IDCT + range limit, one thread per 8x8 block ---------------------------------------
__global__ void __launch_bounds__(128) jpeg_idct_kernel(const int16_t* __restrict__ coefs, const FrameDesc* __restrict__ desc, const QuantTable* __restrict__ qts,
                                                        uint8_t* __restrict__ dst, int n, int H, int W, int bw, int bh) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    __shared__ uint16_t sq[64];
    // (every thread of the CTA works on frame f: stage its quantisation table once)
    if (threadIdx.x < 64) sq[threadIdx.x] = qts[desc[f].qt].q[threadIdx.x];
    __syncthreads();
    if (b >= bw * bh) return;
    int16_t cf[64];
    const uint4* cp = (const uint4*)(coefs + ((size_t)f * bw * bh + b) * 64);
#pragma unroll
    for (int i = 0; i < 8; i++) { uint4 v = cp[i]; *(uint4*)&cf[i * 8] = v; }
    uint8_t px[64];
    idct_islow_block(cf, sq, px);
    const int bx = b % bw, by = b / bw, x0 = bx * 8, y0 = by * 8;
    uint8_t* o = dst + (size_t)f * H * W;
    const bool wide = x0 + 8 <= W && (W & 7) == 0 && (reinterpret_cast<size_t>(dst) & 7) == 0 && ((size_t)H * W & 7) == 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (y0 + r >= H) break;
        uint8_t* row = o + (size_t)(y0 + r) * W + x0;
        if (wide) *(uint2*)row = *(uint2*)&px[r * 8];
        else for (int c = 0; c < 8 && x0 + c < W; c++) row[c] = px[r * 8 + c];
    }
}

int launch_jpeg_decode(const uint8_t* blob, const void* desc, const void* hts, const void* qts, int16_t* coefs, uint8_t* dst,
                       int n, int H, int W, cudaStream_t s) {
    const int bw = (W + 7) / 8, bh = (H + 7) / 8;
    cudaMemsetAsync(coefs, 0, (size_t)n * bw * bh * 64 * sizeof(int16_t), s);
    jpeg_huff_kernel<<<(n * 32 + 127) / 128, 128, 0, s>>>(blob, (const FrameDesc*)desc, (const HuffTable*)hts, coefs, n, bw * bh);
    dim3 grid((bw * bh + 127) / 128, n);
    jpeg_idct_kernel<<<grid, 128, 0, s>>>(coefs, (const FrameDesc*)desc, (const QuantTable*)qts, dst, n, H, W, bw, bh);
    return 2;
}

}  // namespace dmc
