// jpeg_emul.cpp -- CPU emulation of jpeg_frame_kernel (csrc/dmc_jpeg.cu): the same phases, built from the same host/device
// primitives of csrc/dmc_jpeg_core.h (scan_count, scan_symbol, idct_islow_inplace, scan_subseq_bits), with the CTA's lanes
// executed one after the other and the barriers turned into loop boundaries.  Test infrastructure: it lets the CPU test
// suite check the self-synchronising decode (records, rounds, prefix sums, block ownership) against cv2.imdecode without a
// GPU, and reports how many rounds the synchronisation took.  Built by tests/test_jpeg_emul.py with g++.
#include "../../depthmapcompression_b200/csrc/dmc_jpeg_parse.h"

#include <algorithm>

using namespace dmcjpeg;

extern "C" int jpeg_emul_decode(const uint8_t* stream, size_t len, int rows, int cols, uint8_t* out, int lanes, int* rounds_out, char* err, size_t err_len) {
    std::vector<QuantTable> qp; std::vector<HuffTable> hp; FrameDesc d;
    std::string why = dmc::jpeg_parse_frame(stream, len, 0, rows, cols, qp, hp, &d);
    if (!why.empty()) { if (err && err_len) { strncpy(err, why.c_str(), err_len - 1); err[err_len - 1] = 0; } return -1; }
    if (d.restart_interval) { if (err && err_len) strncpy(err, "restart intervals take the legacy path", err_len - 1); return -2; }
    const HuffTable& dc = hp[d.dc]; const HuffTable& ac = hp[d.ac];
    int quant[64]; for (int i = 0; i < 64; i++) quant[i] = qp[d.qt].q[i];
    // phase 0: de-stuff up to the first marker, zero padding, big-endian words
    const uint8_t* raw = stream + d.scan_offset; const uint64_t raw_len = d.scan_end - d.scan_offset;
    std::vector<uint8_t> ds;
    for (uint64_t i = 0; i < raw_len; i++) {
        if (raw[i] == 0xFF) { const uint8_t nx = i + 1 < raw_len ? raw[i + 1] : 0xD9; if (nx != 0) break; ds.push_back(0xFF); i++; }
        else ds.push_back(raw[i]);
    }
    const uint32_t L = (uint32_t)ds.size(), last_word = (L + 3) / 4;
    ds.resize(4 * (last_word + 2), 0);
    std::vector<uint32_t> words(last_word + 2);
    for (uint32_t i = 0; i < last_word + 2; i++) words[i] = ((uint32_t)ds[4 * i] << 24) | ((uint32_t)ds[4 * i + 1] << 16) | ((uint32_t)ds[4 * i + 2] << 8) | ds[4 * i + 3];
    // phases 1 and 2
    const uint32_t total_bits = L * 8, S = scan_subseq_bits(total_bits, (uint32_t)lanes);
    const uint32_t nsub = total_bits ? (total_bits + S - 1) / S : 1;
    std::vector<uint32_t> rec_p(nsub), rec_k(nsub), rec_nb(nsub); std::vector<int32_t> rec_dc(nsub);
    std::vector<ScanState> st(nsub); std::vector<char> active(nsub, 0);
    for (uint32_t j = 0; j < nsub; j++) {
        st[j].p = j * S; st[j].k = 0;
        uint32_t nb; int32_t dcs;
        scan_count(words.data(), last_word, dc, ac, st[j], std::min((j + 1) * S, total_bits), &nb, &dcs);
        rec_p[j] = st[j].p; rec_k[j] = (uint32_t)st[j].k; rec_nb[j] = nb; rec_dc[j] = dcs;
        active[j] = j + 1 < nsub;
    }
    int rounds = 0;
    for (uint32_t r = 1; std::any_of(active.begin(), active.end(), [](char a) { return a != 0; }); r++, rounds++) {
        // (within a round every lane touches only record j + r: order between lanes does not matter)
        for (uint32_t j = 0; j < nsub; j++) {
            if (!active[j]) continue;
            const uint32_t tgt = j + r; uint32_t nb; int32_t dcs;
            scan_count(words.data(), last_word, dc, ac, st[j], std::min((tgt + 1) * S, total_bits), &nb, &dcs);
            const bool met = rec_p[tgt] == st[j].p && rec_k[tgt] == (uint32_t)st[j].k;
            rec_p[tgt] = st[j].p; rec_k[tgt] = (uint32_t)st[j].k; rec_nb[tgt] = nb; rec_dc[tgt] = dcs;
            if (met || tgt + 1 >= nsub) active[j] = 0;
        }
    }
    if (rounds_out) *rounds_out = rounds;
    // phase 3 + 4
    const int bw = (cols + 7) / 8, nblocks = bw * ((rows + 7) / 8);
    uint32_t blk0 = 0; int32_t pred0 = 0;
    for (uint32_t j = 0; j < nsub; j++) {
        ScanState s; if (j == 0) { s.p = 0; s.k = 0; } else { s.p = rec_p[j - 1]; s.k = (int)rec_k[j - 1]; }
        const uint32_t hi = j + 1 == nsub ? 0xFFFFFFFFu : (j + 1) * S;
        uint32_t blk = blk0; int pred = pred0; int kz;
        int cf[64]; for (int i = 0; i < 64; i++) cf[i] = 0;
        while (s.k != 0 && s.p < hi) scan_symbol(words.data(), last_word, dc, ac, s, &kz);
        while (s.k == 0 && s.p < hi && blk < (uint32_t)nblocks) {
            pred += scan_symbol(words.data(), last_word, dc, ac, s, &kz);
            cf[0] = (int)(int16_t)pred;
            uint64_t mask = 0;
            while (s.k != 0) {
                const int v = scan_symbol(words.data(), last_word, dc, ac, s, &kz);
                if (kz > 0) { const int nat = zigzag_to_natural(kz); cf[nat] = v; mask |= 1ull << nat; }
            }
            const int x0 = (int)(blk % (uint32_t)bw) * 8, y0 = (int)(blk / (uint32_t)bw) * 8;
            idct_islow_inplace([&cf](int i) -> int& { return cf[i]; }, quant, mask, [&](int r, const uint8_t* row) {
                if (y0 + r >= rows) return;
                for (int c = 0; c < 8 && x0 + c < cols; c++) out[(size_t)(y0 + r) * cols + x0 + c] = row[c];
            });
            if (mask & ~1ull) for (int i = 1; i < 64; i++) cf[i] = 0;
            blk++;
        }
        blk0 += rec_nb[j]; pred0 += rec_dc[j];
    }
    return 0;
}

extern "C" int jpeg_emul_probe(const uint8_t* stream, size_t len, int* rows, int* cols, char* err, size_t err_len) {
    std::vector<QuantTable> qp; std::vector<HuffTable> hp; FrameDesc d; int r = 0, c = 0;
    std::string why = dmc::jpeg_parse_frame(stream, len, 0, -1, -1, qp, hp, &d, &r, &c);
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (err && err_len) { strncpy(err, why.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return why.empty() ? 0 : -1;
}
