// Micro-benchmark: issue rate of the 25-input median exchange network written with different min/max instructions.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int M> __device__ __forceinline__ void ce(unsigned& a, unsigned& b, int k) {
    unsigned lo, hi;
    if (M <= 2) {
        bool h = (M == 0) || (M == 2 && (k & 1));
        if (h) { asm volatile("min.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b)); asm volatile("max.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b)); }
        else { lo = __vminu2(a, b); hi = __vmaxu2(a, b); }
    } else {
        asm volatile("min.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        int mode = M == 3 ? 0 : M == 4 ? 1 : M == 5 ? (k % 3) : M == 6 ? (k & 1) : (k & 1) * 2;
        if (mode == 0) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; sub.u32 %0, t, %3;}" : "=r"(hi) : "r"(a), "r"(b), "r"(lo));   // hi = a + b - lo
        else if (mode == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(hi) : "r"(a), "r"(b), "r"(lo));                      // hi = a ^ b ^ lo
        else asm volatile("max.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    }
    a = lo; b = hi;
}
#define CE(i, j) ce<M>(p[i], p[j], kk++);
template <int M> __global__ void k(unsigned* out, const unsigned* in, long long* cyc) {
    unsigned p[25];
#pragma unroll
    for (int i = 0; i < 25; i++) p[i] = in[threadIdx.x + 32 * i];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        int kk = 0;
        CE(0,1) CE(3,4) CE(2,4) CE(2,3) CE(6,7) CE(5,7) CE(5,6) CE(9,10) CE(8,10) CE(8,9) CE(12,13) CE(11,13) CE(11,12) CE(15,16) CE(14,16) CE(14,15) CE(18,19) CE(17,19)
        CE(17,18) CE(21,22) CE(20,22) CE(20,21) CE(23,24) CE(2,5) CE(3,6) CE(0,6) CE(0,3) CE(4,7) CE(1,7) CE(1,4) CE(11,14) CE(8,14) CE(8,11) CE(12,15) CE(9,15) CE(9,12)
        CE(13,16) CE(10,16) CE(10,13) CE(20,23) CE(17,23) CE(17,20) CE(21,24) CE(18,24) CE(18,21) CE(19,22) CE(8,17) CE(9,18) CE(0,18) CE(0,9) CE(10,19) CE(1,19) CE(1,10) CE(11,20)
        CE(2,20) CE(2,11) CE(12,21) CE(3,21) CE(3,12) CE(13,22) CE(4,22) CE(4,13) CE(14,23) CE(5,23) CE(5,14) CE(15,24) CE(6,24) CE(6,15) CE(7,16) CE(7,19) CE(13,21) CE(15,23)
        CE(7,13) CE(7,15) CE(1,9) CE(3,11) CE(5,17) CE(11,17) CE(9,17) CE(4,10) CE(6,12) CE(7,14) CE(4,6) CE(4,7) CE(12,14) CE(10,14) CE(6,7) CE(10,12) CE(6,10) CE(6,17)
        CE(12,17) CE(7,17) CE(7,10) CE(12,18) CE(7,12) CE(10,18) CE(12,20) CE(10,20) CE(10,12)
#pragma unroll
        for (int i = 0; i < 25; i++) p[i] ^= (unsigned)it;        // keep every wire live and changing (cheap vs 198 ops)
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) s ^= p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int M> void run(const char* name, int threads, unsigned* out, unsigned* in, long long* cyc) {
    k<M><<<148, threads>>>(out, in, cyc); k<M><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    printf("%-40s threads=%4d  %6.3f warp-inst/clk/SM (198+25 per iter)\n", name, threads, (double)ITERS * 223 * (threads / 32) / c);
}
int main() {
    unsigned *out, *in; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0x3c, 4096 * 4);
    for (int t : {512, 1024}) { run<0>("median25 HMNMX2 min+max", t, out, in, cyc); run<3>("median25 min + IADD3(a+b-lo)", t, out, in, cyc); run<4>("median25 min + LOP3(a^b^lo)", t, out, in, cyc);
        run<5>("median25 min + {IADD3,LOP3,max} rotating", t, out, in, cyc); run<6>("median25 min + {IADD3,LOP3} alternating", t, out, in, cyc); run<7>("median25 min + {IADD3,max} alternating", t, out, in, cyc); }
    return 0;
}
