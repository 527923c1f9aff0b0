// dmc_jpeg_parse.h -- host-side marker parsing of baseline grayscale JPEG streams (DQT / SOF0 / DHT / DRI / SOS), Huffman
// table derivation and validation.  Host only; shared by dmc_jpeg.cu and by the CPU emulation harness of the parallel
// decoder (tests/cpp/jpeg_emul.cpp).  Reference counterpart: libjpeg's jdmarker.c / jdhuff.c as reached through
// cv::imdecode(buf, 0) (main.cpp:284, :521) and jpeg_decode() (jpegTurboDemo.cpp:217-271).
#pragma once
#include "dmc_jpeg_core.h"
#include <string.h>
#include <string>
#include <vector>

namespace dmc {

using namespace dmcjpeg;

// ---- host: marker parsing ----------------------------------------------------------------------------------------
// Derives the decoding tables of one DHT entry.  Returns false for tables libjpeg's jpeg_make_d_derived_tbl rejects with
// JERR_BAD_HUFF_TABLE: more than 256 symbols, or code lengths that do not form a prefix code (a code of length l must be
// < 2^l; an over-subscribed BITS array would otherwise index past look[]).
static inline bool derive_table(const uint8_t bits[17], const uint8_t* vals, int nvals, HuffTable* t) {
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

template <class T> static inline int intern(std::vector<T>& pool, const T& v) {
    for (size_t i = 0; i < pool.size(); i++) if (memcmp(&pool[i], &v, sizeof(T)) == 0) return (int)i;
    pool.push_back(v); return (int)pool.size() - 1;
}

// Parses one stream [p, p+len).  Returns an empty string on success, else the reason it is not supported.
// rows < 0: any frame size is accepted (reported through out_rows / out_cols).
inline std::string jpeg_parse_frame(const uint8_t* p, uint64_t len, uint64_t blob_offset, int rows, int cols,
                             std::vector<QuantTable>& qpool, std::vector<HuffTable>& hpool, FrameDesc* d, int* out_rows = nullptr, int* out_cols = nullptr) {
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
            if (h == 0 || w == 0) return "empty frame (DNL-defined height is not supported)";
            if (out_rows) *out_rows = h;
            if (out_cols) *out_cols = w;
            if (rows >= 0 && (h != rows || w != cols)) return "frame size differs from the batch size";
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

}  // namespace dmc
