// Second micro-benchmark: the same instructions with ALL operands in distinct registers (the form real kernels use),
// to expose register-file port / bank limits that the uniform-operand form in ubench_pipes.cu hides.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
template <int OP> __global__ void k(unsigned* out, const unsigned* in, long long* cyc) {
    unsigned a[CH], b[CH], c[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 1024]; c[i] = in[threadIdx.x + 32 * i + 2048]; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int i = 0; i < CH; i++) {
            unsigned x = a[i], y = b[(i + rep) % CH], z = c[(i + 2 * rep + 1) % CH];
            if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
            if (OP == 1) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            if (OP == 2) asm volatile("set.le.f16x2.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            if (OP == 3) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            if (OP == 4) asm volatile("{.reg .u32 t; vmin2.u32.u32.u32 %0, %0, %1, %2;}" : "+r"(x) : "r"(y), "r"(0u));
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(x) : "r"(y));
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
            if (OP == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(__uint_as_float(y)));
            if (OP == 8) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x) : "f"(__uint_as_float(y)), "f"(__uint_as_float(z)));
            if (OP == 9) asm volatile("{.reg .u32 t; shr.u32 t, %1, 10; add.u32 %0, %0, t;}" : "+r"(x) : "r"(y));
            if (OP == 10) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
            if (OP == 11) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            if (OP == 12) { unsigned lo, hi; asm volatile("min.f16x2 %0, %2, %3; max.f16x2 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y)); x = lo; b[(i + rep) % CH] = hi; }   // a compare-exchange
            if (OP == 13) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(__uint_as_float(y)));
            if (OP == 14) asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(x) : "r"(y), "r"(z));      // no dependence on x: pure throughput
            if (OP == 15) asm volatile("min.f16x2 %0, %1, %2;" : "=r"(x) : "r"(y), "r"(z));
            a[i] = x;
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, unsigned* out, unsigned* in, long long* cyc, int per = 1) {
    const int threads = 1024;
    k<OP><<<148, threads>>>(out, in, cyc); k<OP><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    printf("%-44s %6.3f warp-inst/clk/SM\n", name, (double)ITERS * 4 * CH * per * (threads / 32) / c);
}
int main() {
    unsigned *out, *in; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0x3c, 4096 * 4);
    run<0>("HFMA2 r,r,r", out, in, cyc); run<1>("HADD2 r,r", out, in, cyc); run<14>("HSUB2 r,r (independent)", out, in, cyc); run<2>("HSET2.LE r,r", out, in, cyc);
    run<3>("HMNMX2 min r,r", out, in, cyc); run<11>("HMNMX2 max r,r", out, in, cyc); run<15>("HMNMX2 min r,r (independent)", out, in, cyc); run<12>("HMNMX2 compare-exchange (2 inst)", out, in, cyc, 2);
    run<4>("VMIN2 r,r", out, in, cyc); run<5>("PRMT r,r", out, in, cyc); run<6>("LOP3 r,r,r", out, in, cyc);
    run<7>("FADD r,r", out, in, cyc); run<13>("FMUL r,r", out, in, cyc); run<8>("FFMA r,r,r", out, in, cyc); run<9>("LEA.HI (x + (y>>10))", out, in, cyc); run<10>("IADD r,r", out, in, cyc);
    return 0;
}
