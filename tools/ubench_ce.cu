// Micro-benchmark: why does an exchange network issue at 2 warp-inst/clk/SM when HMNMX2 alone reaches 4?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define NP 12
__device__ __forceinline__ unsigned hmin(unsigned a, unsigned b) { unsigned r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned hmax(unsigned a, unsigned b) { unsigned r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int T> __global__ void k(unsigned* out, const unsigned* in, long long* cyc) {
    unsigned a[NP], b[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 1024]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            if (T == 0) {          // independent CEs on fixed pairs (a_i, b_i): 2 inst each
#pragma unroll
                for (int i = 0; i < NP; i++) { unsigned lo = hmin(a[i], b[i]), hi = hmax(a[i], b[i]); a[i] = lo; b[i] = hi; }
            } else if (T == 1) {   // CEs with rotating partners (a_i, b_{i+rep+1})
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 1) % NP; unsigned lo = hmin(a[i], b[j]), hi = hmax(a[i], b[j]); a[i] = lo; b[j] = hi; }
            } else if (T == 2) {   // only min, two fresh sources, separate destination, then rotate
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 1) % NP; unsigned lo = hmin(a[i], b[j]); a[i] = b[j]; b[j] = lo; }
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 2) % NP; unsigned lo = hmin(a[i], b[j]); a[i] = b[j]; b[j] = lo; }
            } else if (T == 3) {   // min on (a_i,b_j) and max on a different pair (a_j, b_i): no shared operands between the two
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 1) % NP; unsigned lo = hmin(a[i], b[j]); unsigned hi = hmax(a[j], b[i]); a[i] = lo; b[i] = hi; }
            } else if (T == 4) {   // accumulate-style: x = min(x, y) (destination == source)
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 1) % NP; a[i] = hmin(a[i], b[j]); }
#pragma unroll
                for (int i = 0; i < NP; i++) { int j = (i + rep + 2) % NP; b[i] = hmax(b[i], a[j]); }
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int T> void run(const char* name, unsigned* out, unsigned* in, long long* cyc) {
    const int threads = 512;
    k<T><<<148, threads>>>(out, in, cyc); k<T><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    printf("%-64s %6.3f warp-inst/clk/SM\n", name, (double)ITERS * 2 * 2 * NP * (threads / 32) / c);
}
int main() {
    unsigned *out, *in; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0x3c, 4096 * 4);
    run<0>("CE fixed pairs: lo=min(a,b) hi=max(a,b)", out, in, cyc);
    run<1>("CE rotating partners", out, in, cyc);
    run<2>("min only, fresh sources, separate dest", out, in, cyc);
    run<3>("min and max on disjoint operand pairs", out, in, cyc);
    run<4>("x=min(x,y) / y=max(y,x) accumulate style", out, in, cyc);
    return 0;
}
