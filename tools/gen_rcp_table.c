/* gen_rcp_table.c -- the table behind the reference renderer's _mm_rcp_ps (depthmapUtil.cpp:78).
 *
 *   gcc -O2 -fopenmp tools/gen_rcp_table.c -o /tmp/gen_rcp && /tmp/gen_rcp > depthmapcompression_b200/csrc/dmc_rcp_intel.inc
 *
 * On Intel CPUs RCPPS / RCPSS are a pure look-up: the result depends on the sign, the exponent and the top 11 mantissa bits
 * of the operand only; its mantissa has 12 significant bits.  This program reads the 2048 entries off the instruction,
 * prints them, and then checks the emulation (the same few lines as rcp_intel() in csrc/dmc_render.cu) against the
 * instruction on ALL 2^32 operands (about 2 s on 8 threads); exit status 1 and a message on stderr if this CPU disagrees
 * (AMD parts use a different table).  Verified on Intel Xeon family 6 model 207.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <xmmintrin.h>

static uint16_t T[2048];
static uint32_t hw(uint32_t b) { float x; memcpy(&x, &b, 4); __m128 v = _mm_rcp_ps(_mm_set1_ps(x)); float o[4]; _mm_storeu_ps(o, v); uint32_t r; memcpy(&r, &o[1], 4); return r; }
static uint32_t emu(uint32_t b) {
    const uint32_t s = b & 0x80000000u, E = (b >> 23) & 0xff, m = b & 0x7fffff;
    if (E == 0xff) return m ? (b | 0x400000u) : s;      /* NaN -> quiet NaN, inf -> signed zero */
    if (E == 0) return s | 0x7f800000u;                 /* zero / denormal -> signed infinity */
    if (E >= 253) return s;                             /* the result would be denormal -> signed zero */
    return s | ((253u - E) << 23) | ((uint32_t)T[m >> 12] << 11);
}
int main(void) {
    int fmt = 1;
    printf("// Intel RCPPS / RCPSS mantissa table: entry i = (mantissa of rcp(1 + i/2048)) >> 11; the result's biased exponent is 253 - E(x).\n");
    printf("// Generated and verified on all 2^32 operands by tools/gen_rcp_table.c (Intel Xeon, family 6 model 207).\n");
    for (uint32_t i = 0; i < 2048; i++) {
        const uint32_t r = hw(0x3f800000u | (i << 12));
        if ((r >> 23) != 126 || (r & 0x7ff)) fmt = 0;
        T[i] = (uint16_t)((r & 0x7fffff) >> 11);
        printf("%u,%s", T[i], (i % 16 == 15) ? "\n" : " ");
    }
    unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad)
    for (long long hi = 0; hi < 65536; hi++)
        for (uint32_t lo = 0; lo < 65536; lo++) { const uint32_t b = ((uint32_t)hi << 16) | lo; if (hw(b) != emu(b)) bad++; }
    if (!fmt || bad) { fprintf(stderr, "this CPU's rcpps is not the Intel table (format ok %d, %llu mismatches)\n", fmt, bad); return 1; }
    fprintf(stderr, "table verified on all 2^32 operands\n");
    return 0;
}
