// ffma2bw.cu - bring-up microbenchmark: issue rate of packed FFMA2 / FADD2 vs scalar FFMA / FADD on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1;} } while (0)
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
    float a[8], b = seed, c = seed * 0.5f;
    unsigned long long p[8], pb, pc;
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(c));
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.f));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(pb), "l"(pc));
            if (MODE == 2) a[i] = a[i] + c;
            if (MODE == 3) asm volatile("add.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(pc));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi; }
    if (s == 12345.678f) out[0] = s;
}
template <int MODE>
double run(float* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(d, iters, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, 512>>>(d, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; CK(cudaMalloc(&d, 64));
    const int iters = 200000;
    const char* names[4] = {"FFMA  scalar", "FFMA2 packed", "FADD  scalar", "FADD2 packed"};
    double ms[4] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters)};
    for (int m = 0; m < 4; ++m) {
        const double inst = 148.0 * 16 * 8.0 * iters;  // warp instructions
        printf("%s: %.3f ms, %.2f warp-inst/clk/SM (at 1.965 GHz), %.1f Gflop-lanes/s per SM\n", names[m], ms[m],
               inst / 148 / (ms[m] * 1e-3 * 1.965e9), (m & 1 ? 2.0 : 1.0) * 32 * inst / 148 / (ms[m] * 1e-3) * 1e-9);
    }
    return 0;
}
