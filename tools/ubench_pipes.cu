// Micro-benchmark of the sm_100a instruction rates that decide the SIMD-in-register design of the filters
// (packed half2 vs packed u16x2 vs byte SWAR).  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
template <int OP> __device__ __forceinline__ void step(unsigned (&a)[CH], unsigned b, unsigned c) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
        unsigned x = a[i];
        if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == 1) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == 2) asm volatile("set.le.f16x2.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == 3) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == 4) x = __vminu2(x, b);
        if (OP == 5) x = __vimax3_u16x2(x, b, c);
        if (OP == 6) x = __byte_perm(x, b, 0x5432);
        if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(__uint_as_float(b)));
        if (OP == 11) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(__uint_as_float(b)));
        if (OP == 12) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
        if (OP == 13) x = __vabsdiffu4(x, b);
        if (OP == 14) x = __dp4a(x, b, c);
        if (OP == 15) x = __vmaxu2(x, b);
        if (OP == 16) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
        if (OP == 17) x = (x >> 15) + b;                                   // LEA.HI
        if (OP == 18) asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(x) : "f"(__uint_as_float(x)));
        if (OP == 19) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(*(float*)&x) : "r"(x));
        // mixes: alternate two op kinds over the chains
        if (OP == 20) { if (i & 1) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b)); else x = __vminu2(x, b); }
        if (OP == 21) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); }
        if (OP == 22) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("set.le.f16x2.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b)); }
        if (OP == 23) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); }
        if (OP == 24) { if (i & 1) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b)); else x = __byte_perm(x, b, 0x5432); }
        if (OP == 25) { if (i & 1) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(b)); else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        if (OP == 26) { if (i & 1) x = __vminu2(x, b); else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        if (OP == 27) { if (i & 1) x = __vminu2(x, b); else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(__uint_as_float(b))); }
        a[i] = x;
    }
}
template <int OP> __global__ void k(unsigned* out, unsigned b, unsigned c, long long* cyc) {
    unsigned a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 7 + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) { step<OP>(a, b, c); step<OP>(a, b, c); step<OP>(a, b, c); step<OP>(a, b, c); }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// packed fp32x2 (FADD2 / FMUL2 / FFMA2)
template <int OP> __global__ void k2(float2* out, float2 b, float2 c, long long* cyc) {
    float2 a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS * 4; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("{.reg .b64 x,y; mov.b64 x,{%0,%1}; mov.b64 y,{%2,%3}; add.rn.f32x2 x,x,y; mov.b64 {%0,%1},x;}" : "+f"(a[i].x), "+f"(a[i].y) : "f"(b.x), "f"(b.y));
            if (OP == 1) asm volatile("{.reg .b64 x,y; mov.b64 x,{%0,%1}; mov.b64 y,{%2,%3}; mul.rn.f32x2 x,x,y; mov.b64 {%0,%1},x;}" : "+f"(a[i].x), "+f"(a[i].y) : "f"(b.x), "f"(b.y));
            if (OP == 2) asm volatile("{.reg .b64 x,y,z; mov.b64 x,{%0,%1}; mov.b64 y,{%2,%3}; mov.b64 z,{%4,%5}; fma.rn.f32x2 x,x,y,z; mov.b64 {%0,%1},x;}" : "+f"(a[i].x), "+f"(a[i].y) : "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < CH; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, unsigned* out, long long* cyc) {
    const int threads = 1024;   // 32 warps: 8 per SMSP
    k<OP><<<148, threads>>>(out, 0x3c003c00u, 0x00010001u, cyc);
    k<OP><<<148, threads>>>(out, 0x3c003c00u, 0x00010001u, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    double winst = (double)ITERS * 4 * CH * (threads / 32);
    printf("%-34s %6.3f warp-inst/clk/SM\n", name, winst / c);
}
template <int OP> void run2(const char* name, float2* out, long long* cyc) {
    const int threads = 1024;
    k2<OP><<<148, threads>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), cyc);
    k2<OP><<<148, threads>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    double winst = (double)ITERS * 4 * CH * (threads / 32);
    printf("%-34s %6.3f warp-inst/clk/SM\n", name, winst / c);
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    run<0>("HFMA2", out, cyc); run<1>("HADD2", out, cyc); run<2>("HSET2.LE", out, cyc); run<3>("HMNMX2 (min.f16x2)", out, cyc); run<16>("HMNMX2 (max.f16x2)", out, cyc);
    run<4>("VIMNMX.U16x2 (vminu2)", out, cyc); run<15>("VIMNMX.U16x2 (vmaxu2)", out, cyc); run<5>("VIMNMX3.U16x2", out, cyc);
    run<6>("PRMT", out, cyc); run<7>("LOP3", out, cyc); run<8>("IADD", out, cyc); run<9>("IMAD", out, cyc); run<17>("LEA.HI (shift-add)", out, cyc);
    run<10>("FADD", out, cyc); run<11>("FMUL", out, cyc); run<12>("FFMA", out, cyc); run<13>("VABSDIFF4", out, cyc); run<14>("IDP.4A", out, cyc);
    run<18>("F2I.RN", out, cyc); run<19>("I2F", out, cyc);
    run<20>("mix HMNMX2 + VIMNMX.U16x2", out, cyc); run<21>("mix HFMA2 + LOP3", out, cyc); run<22>("mix HFMA2 + HSET2", out, cyc);
    run<23>("mix HFMA2 + FFMA", out, cyc); run<24>("mix HADD2 + PRMT", out, cyc); run<25>("mix HMNMX2 + HFMA2", out, cyc); run<26>("mix VIMNMX + HFMA2", out, cyc); run<27>("mix VIMNMX + FADD", out, cyc);
    run2<0>("FADD2 (f32x2)", (float2*)out, cyc); run2<1>("FMUL2 (f32x2)", (float2*)out, cyc); run2<2>("FFMA2 (f32x2)", (float2*)out, cyc);
    return 0;
}
