// streambw.cu - bring-up microbenchmark: which per-SM load mechanism streams an HBM-cold tensor fastest from a
// persistent one-CTA-per-SM grid?  (a) LDG.128 into registers, U in flight per thread; (b) LDGSTS (cp.async
// 16 B) into a private per-thread ring, D groups in flight; (c) 1-D TMA bulk copies into a K-slot ring.
// Each variant reads `bytes` once (checksum keeps the loads alive) and, for the copy variants, writes them back.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ldg_na(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_na(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// ---------------------------------------------------------------- (a) LDG into registers, chunked like the flat path:
// CTA b owns chunks b, b+G, ... of `chunk` vectors; threads stride over the chunk U vectors at a time
template <int U, bool COPY>
__global__ void __launch_bounds__(1024, 1) k_ldg(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec,
                                                  unsigned chunk, unsigned* sink) {
    unsigned acc = 0;
    const size_t nchunks = nvec / chunk;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint4* s = src + c * chunk;
        uint4* d = dst + c * chunk;
        for (unsigned v0 = threadIdx.x; v0 < chunk; v0 += U * blockDim.x) {
            uint4 q[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * blockDim.x < chunk) q[u] = ldg_na(s + v0 + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * blockDim.x < chunk) {
                    if (COPY) { q[u].x += 1; stg_na(d + v0 + u * blockDim.x, q[u]); }
                    else acc += q[u].x ^ q[u].w;
                }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- (b) LDGSTS private rings, one group per sweep
template <int D, bool HINT>
__global__ void __launch_bounds__(1024, 1) k_ldgsts(const uint4* __restrict__ src, size_t nvec, unsigned chunk, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned T = blockDim.x;
    const uint32_t my = smem_u32(smem) + threadIdx.x * 16, stage = T * 16, end = my + (D + 1) * stage;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    unsigned acc = 0;
    const size_t nchunks = nvec / chunk;
    const unsigned spc = chunk / T;  // sweeps per chunk (chunk is a multiple of T)
    size_t ci = blockIdx.x;
    unsigned si = 0;
    uint32_t wr = my, rd = my;
    auto issue = [&]() {
        if (ci < nchunks) {
            const uint4* p = src + ci * chunk + (size_t)si * T + threadIdx.x;
            if (HINT) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(wr), "l"(p), "l"(pol) : "memory");
            else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wr), "l"(p) : "memory");
            if (++si == spc) { si = 0; ci += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        wr += stage;
        if (wr == end) wr = my;
    };
    for (int d = 0; d < D; ++d) issue();
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x)
        for (unsigned s = 0; s < spc; ++s) {
            asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
            const uint4 q = lds128(rd);
            rd += stage;
            if (rd == end) rd = my;
            issue();
            acc += q.x ^ q.w;
        }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- (c) TMA bulk ring: K slots of `chunk` vectors
__global__ void __launch_bounds__(544, 1) k_tma(const uint4* __restrict__ src, size_t nvec, unsigned chunk, int K, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t data0 = smem_u32(smem);
    const uint32_t slot_bytes = chunk * 16;
    const uint32_t full0 = data0 + K * slot_bytes, empty0 = full0 + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < K; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * i), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * i), "r"(16));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = nvec / chunk;
    auto wait = [](uint32_t bar, uint32_t par) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    };
    if (threadIdx.x >= 512) {
        if (threadIdx.x == 512) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            unsigned i = 0, ph = 0, t = 0;
            for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++t) {
                if (t >= (unsigned)K) wait(empty0 + 8 * i, ph ^ 1);
                const uint32_t bar = full0 + 8 * i;
                for (uint32_t off = 0; off < slot_bytes; off += 32768) {
                    const uint32_t n = slot_bytes - off < 32768 ? slot_bytes - off : 32768;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(data0 + i * slot_bytes + off), "l"((const char*)(src + c * chunk) + off), "r"(n), "r"(bar), "l"(pol) : "memory");
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
                if (++i == (unsigned)K) { i = 0; ph ^= 1; }
            }
        }
        return;
    }
    unsigned acc = 0, i = 0, ph = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        wait(full0 + 8 * i, ph);
        for (unsigned v = threadIdx.x; v < chunk; v += 512) {
            const uint4 q = lds128(data0 + i * slot_bytes + v * 16);
            acc += q.x ^ q.w;
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * i) : "memory");
        if (++i == (unsigned)K) { i = 0; ph ^= 1; }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static double time_it(F f, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / iters;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    const size_t bytes = (size_t)1 << 30, nvec = bytes / 16;
    uint4 *a, *b;
    unsigned* sink;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes));
    CK(cudaMemset(b, 2, bytes));
    const int it = 5;
#define REPORT(name, ms, factor) printf("%-44s %8.1f GB/s\n", name, factor * bytes / (ms) * 1e-6)
    char nm[128];
    for (unsigned chunk : {1536u, 3072u, 6144u, 12288u}) {
        for (int T : {512, 768, 1024}) {
            if (chunk % T) continue;
            snprintf(nm, sizeof nm, "ldg read  T=%d U=4 chunk=%uKB", T, chunk / 64);
            REPORT(nm, time_it([&] { k_ldg<4, false><<<G, T>>>(a, b, nvec, chunk, sink); }, it), 1.0);
            snprintf(nm, sizeof nm, "ldg read  T=%d U=8 chunk=%uKB", T, chunk / 64);
            REPORT(nm, time_it([&] { k_ldg<8, false><<<G, T>>>(a, b, nvec, chunk, sink); }, it), 1.0);
            snprintf(nm, sizeof nm, "ldg copy  T=%d U=4 chunk=%uKB", T, chunk / 64);
            REPORT(nm, time_it([&] { k_ldg<4, true><<<G, T>>>(a, b, nvec, chunk, sink); }, it), 2.0);
        }
    }
    CK(cudaFuncSetAttribute(k_ldgsts<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_ldgsts<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_ldgsts<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_ldgsts<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int T : {512, 1024}) {
        const unsigned chunk = 3072;
        snprintf(nm, sizeof nm, "ldgsts read T=%d D=4 hint", T);
        REPORT(nm, time_it([&] { k_ldgsts<4, true><<<G, T, 5 * T * 16>>>(a, nvec, chunk, sink); }, it), 1.0);
        snprintf(nm, sizeof nm, "ldgsts read T=%d D=8 hint", T);
        REPORT(nm, time_it([&] { k_ldgsts<8, true><<<G, T, 9 * T * 16>>>(a, nvec, chunk, sink); }, it), 1.0);
        snprintf(nm, sizeof nm, "ldgsts read T=%d D=8 nohint", T);
        REPORT(nm, time_it([&] { k_ldgsts<8, false><<<G, T, 9 * T * 16>>>(a, nvec, chunk, sink); }, it), 1.0);
        snprintf(nm, sizeof nm, "ldgsts read T=%d D=12 hint", T);
        REPORT(nm, time_it([&] { k_ldgsts<12, true><<<G, T, 13 * T * 16>>>(a, nvec, chunk, sink); }, it), 1.0);
    }
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    for (unsigned chunk : {512u, 1024u, 1536u, 2816u}) {
        for (int K : {3, 4, 6, 8, 12, 16}) {
            const size_t sm = (size_t)K * chunk * 16 + 256;
            if (sm > 225 * 1024) continue;
            snprintf(nm, sizeof nm, "tma read  K=%d slot=%uKB (%zu KB in flight)", K, chunk / 64, (size_t)K * chunk / 64);
            REPORT(nm, time_it([&] { k_tma<<<G, 544, sm>>>(a, nvec, chunk, K, sink); }, it), 1.0);
        }
    }
    return 0;
}
