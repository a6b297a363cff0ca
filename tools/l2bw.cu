// l2bw.cu - bring-up microbenchmark: L2-hit read bandwidth vs HBM read bandwidth vs copy, persistent grid.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void k_read(const uint4* __restrict__ p, size_t nvec, int reps, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride * 4) {
            uint4 a = make_uint4(0,0,0,0), b = a, c = a, d = a;
            a = __ldcg(p + i);
            if (i + stride < nvec) b = __ldcg(p + i + stride);
            if (i + 2 * stride < nvec) c = __ldcg(p + i + 2 * stride);
            if (i + 3 * stride < nvec) d = __ldcg(p + i + 3 * stride);
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}
// read src (maybe L2 resident) and write dst (streaming): models a normalise pass served from L2
__global__ void k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec, int reps) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride * 2) {
            uint4 a = __ldcg(src + i);
            uint4 b = make_uint4(0,0,0,0);
            if (i + stride < nvec) b = __ldcg(src + i + stride);
            a.x += 1; b.x += 1;
            __stcs(dst + (size_t)r * 0 + i, a);
            if (i + stride < nvec) __stcs(dst + i + stride, b);
        }
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    uint4 *a, *b; const size_t big = (size_t)1 << 30;
    CK(cudaMalloc(&a, big)); CK(cudaMalloc(&b, big)); CK(cudaMemset(a, 1, big)); CK(cudaMemset(b, 2, big));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t sizes[] = {16u << 20, 32u << 20, 48u << 20, 64u << 20, 85u << 20, 100u << 20, 512u << 20};
    for (int bps = 1; bps <= 2; ++bps)
    for (size_t sz : sizes) {
        const size_t nvec = sz / 16;
        const int reps = (int)(((size_t)4 << 30) / sz);
        k_read<<<sms * bps, 1024 / bps * 1>>>(a, nvec, 2, sink);  // warm L2
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventRecord(e0));
        k_read<<<sms * bps, 1024 / bps>>>(a, nvec, reps, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("read  %4zu MB x%3d  ctas/sm %d: %8.1f GB/s\n", sz >> 20, reps, bps, (double)sz * reps / ms * 1e-6);
        const int creps = reps / 2 > 0 ? reps / 2 : 1;
        CK(cudaEventRecord(e0));
        k_copy<<<sms * bps, 1024 / bps>>>(a, b, nvec, creps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("copy  %4zu MB x%3d  ctas/sm %d: %8.1f GB/s (read+write bytes; src re-read each rep)\n", sz >> 20, creps, bps,
               2.0 * sz * creps / ms * 1e-6);
    }
    return 0;
}
