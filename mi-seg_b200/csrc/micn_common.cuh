// micn_common.cuh - device building blocks shared by the instance_cond kernels (sm_100a only).
//
//  * 16-byte vector traits for fp32 / bf16 / fp16 slabs (128-bit coalesced HBM traffic)
//  * Welford/Chan partial statistics (count, mean, M2) and their fixed-order merges
//  * thin PTX wrappers: mbarrier (local + remote/cluster), 1-D TMA bulk copies
//    (cp.async.bulk global->shared with an mbarrier transaction count), L2 eviction policies,
//    DSMEM stores (st.shared::cluster), named barriers.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/micn.h"

namespace micn {

constexpr int kMaxStyles = MICN_MAX_STYLES;
constexpr int kMaxPeers = MICN_MAX_PEERS;

// ---------------------------------------------------------------------------------------------
// kernel parameter blocks (passed by value; < 1 KB)
// ---------------------------------------------------------------------------------------------
struct FwdParams {
    const void* x;
    void* y;
    const void* res;
    const float* gamma[kMaxStyles];
    const float* beta[kMaxStyles];
    const long long* styles;
    float* save_mean;
    float* save_rstd;
    int* status;  // workspace status word (sticky bit 0: style out of range)
    long long N, C, M;
    long long x_sN, x_sC;
    int num_styles;
    int affine;  // 0 -> gamma = 1, beta = 0
    float eps, slope;
    const float* slope_dev;  // non-null: the activation slope lives in device memory (nn.PReLU weight, one element)
    // MICN_EPI_NORM_ADD_LRELU: the second normalised tensor (dense [N,C,M]) and its norm's parameters / statistics
    const void* x2;
    const float* gamma2[kMaxStyles];
    const float* beta2[kMaxStyles];
    float* save_mean2;
    float* save_rstd2;
};

struct BwdParams {
    const void* dy;
    const void* x;
    const void* act_out;
    const float* gamma[kMaxStyles];
    const float* beta[kMaxStyles];
    const long long* styles;
    const float* save_mean;
    const float* save_rstd;
    void* dx;
    void* dres;
    float* dgamma;  // [S*C] or null
    float* dbeta;
    float* ws_sum_dy;    // [N*C] per-slab sum(g)
    float* ws_sum_dyxh;  // [N*C] per-slab sum(g*xhat)
    unsigned int* ws_counter;
    unsigned int* ws_chan_cnt;  // [C] samples of a channel that have delivered their sums (small path, N > 1; self-resetting)
    int* status;
    long long N, C, M;
    long long x_sN, x_sC;
    int num_styles;
    int affine;
    float slope;
    const float* slope_dev;
    float* dslope;  // PReLU (EPI_LRELU with slope_dev): partial sums of dy * pre over pre <= 0, one entry per CTA
                    // (flat path) or per slab (small path); the caller zero-fills the buffer and adds it up
    // MICN_EPI_NORM_ADD_LRELU: the second normalised tensor, its parameters / statistics and its gradients
    const void* x2;
    const float* gamma2[kMaxStyles];
    const float* beta2[kMaxStyles];
    const float* save_mean2;
    const float* save_rstd2;
    void* dx2;
    float* dgamma2;  // [S*C] or null (dbeta2 goes with it)
    float* dbeta2;
    float* ws_sum_dyxh2;  // [N*C] per-slab sum(g*xhat2)
    // micn_bwd_allreduce: exchange of d(gamma)/d(beta) with the other GPUs of the box, fused into the backward kernel.
    // xchg_peers[r] = rank r's exchange buffer as mapped in this process (NVLink peer memory; [xchg_rank] is local)
    void* xchg_peers[kMaxPeers];
    int xchg_rank, xchg_world;  // world <= 1: no exchange
    int xchg_mode;              // 1: fold this call's exchange at the kernel's end; 2: fold the PREVIOUS call's at its start
};

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, two fp32 operations per issued instruction).
// The hot loops are bound by instruction issue, not by the fp32 pipe, so halving the arithmetic
// instruction count is what counts.  A pair lives in one 64-bit register.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 f2_bits(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ f32x2 f2_splat(float v) { return f2_make(v, v); }
__device__ __forceinline__ void f2_split(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float f2_hsum(f32x2 v) {
    float lo, hi;
    f2_split(v, lo, hi);
    return lo + hi;
}

// ---------------------------------------------------------------------------------------------
// element / vector traits: one 16-byte vector per thread per access
// ---------------------------------------------------------------------------------------------
template <typename T>
struct VecT;

template <>
struct VecT<float> {
    static constexpr int N = 4;
    static constexpr bool kWideExponent = true;
    __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x);
        f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z);
        f[3] = __uint_as_float(v.w);
    }
    __device__ __forceinline__ static uint4 pack(const float* f) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    __device__ __forceinline__ static void unpack2(const uint4& v, f32x2* f) {
        f[0] = f2_bits(v.x, v.y);
        f[1] = f2_bits(v.z, v.w);
    }
    __device__ __forceinline__ static uint4 pack2v(const f32x2* f) {
        float a, b, c, d;
        f2_split(f[0], a, b);
        f2_split(f[1], c, d);
        return make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
    }
    __device__ __forceinline__ static float load1(const float* p) { return *p; }
    __device__ __forceinline__ static void store1(float* p, float v) { *p = v; }
};

template <>
struct VecT<__nv_bfloat16> {
    static constexpr int N = 8;
    static constexpr bool kWideExponent = true;  // same exponent range as fp32
    __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
        // bf16 -> fp32 is a 16-bit shift: no conversion instruction needed
        f[0] = __uint_as_float(v.x << 16);
        f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16);
        f[3] = __uint_as_float(v.y & 0xffff0000u);
        f[4] = __uint_as_float(v.z << 16);
        f[5] = __uint_as_float(v.z & 0xffff0000u);
        f[6] = __uint_as_float(v.w << 16);
        f[7] = __uint_as_float(v.w & 0xffff0000u);
    }
    __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static uint4 pack(const float* f) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    __device__ __forceinline__ static void unpack2(const uint4& v, f32x2* f) {
        f[0] = f2_bits(v.x << 16, v.x & 0xffff0000u);
        f[1] = f2_bits(v.y << 16, v.y & 0xffff0000u);
        f[2] = f2_bits(v.z << 16, v.z & 0xffff0000u);
        f[3] = f2_bits(v.w << 16, v.w & 0xffff0000u);
    }
    __device__ __forceinline__ static uint32_t packp(f32x2 p) {
        float lo, hi;
        f2_split(p, lo, hi);
        return pack2(lo, hi);
    }
    __device__ __forceinline__ static uint4 pack2v(const f32x2* f) {
        return make_uint4(packp(f[0]), packp(f[1]), packp(f[2]), packp(f[3]));
    }
    __device__ __forceinline__ static float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ __forceinline__ static void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <>
struct VecT<__half> {
    static constexpr int N = 8;
    static constexpr bool kWideExponent = false;
    __device__ __forceinline__ static void unpack2(uint32_t u, float& lo, float& hi) {
        float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
        lo = f.x;
        hi = f.y;
    }
    __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
        unpack2(v.x, f[0], f[1]);
        unpack2(v.y, f[2], f[3]);
        unpack2(v.z, f[4], f[5]);
        unpack2(v.w, f[6], f[7]);
    }
    __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static uint4 pack(const float* f) {
        return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
    __device__ __forceinline__ static f32x2 unpackp(uint32_t u) {
        float lo, hi;
        unpack2(u, lo, hi);
        return f2_make(lo, hi);
    }
    __device__ __forceinline__ static void unpack2(const uint4& v, f32x2* f) {
        f[0] = unpackp(v.x);
        f[1] = unpackp(v.y);
        f[2] = unpackp(v.z);
        f[3] = unpackp(v.w);
    }
    __device__ __forceinline__ static uint32_t packp(f32x2 p) {
        float lo, hi;
        f2_split(p, lo, hi);
        return pack2(lo, hi);
    }
    __device__ __forceinline__ static uint4 pack2v(const f32x2* f) {
        return make_uint4(packp(f[0]), packp(f[1]), packp(f[2]), packp(f[3]));
    }
    __device__ __forceinline__ static float load1(const __half* p) { return __half2float(*p); }
    __device__ __forceinline__ static void store1(__half* p, float v) { *p = __float2half_rn(v); }
};

// ---------------------------------------------------------------------------------------------
// streaming global accesses (each voxel is touched once per pass: keep it out of L1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Welford / Chan partial statistics
// ---------------------------------------------------------------------------------------------
struct Stat {
    float n, mean, m2;
};

__device__ __forceinline__ Stat stat_merge(const Stat& a, const Stat& b) {
    Stat r;
    r.n = a.n + b.n;
    const float inv = r.n > 0.f ? __fdividef(1.f, r.n) : 0.f;
    const float d = b.mean - a.mean;
    const float wb = b.n * inv;
    r.mean = fmaf(d, wb, a.mean);
    r.m2 = a.m2 + b.m2 + d * d * a.n * wb;
    return r;
}

// shifted sums (sum(x-K), sum((x-K)^2), n) -> (n, mean, M2)
__device__ __forceinline__ Stat stat_from_shifted(float K, float s1, float s2, float n) {
    Stat r;
    r.n = n;
    if (n > 0.f) {
        const float m = s1 / n;
        r.mean = K + m;
        r.m2 = fmaxf(s2 - s1 * m, 0.f);
    } else {
        r.mean = 0.f;
        r.m2 = 0.f;
    }
    return r;
}

__device__ __forceinline__ Stat stat_warp_reduce(Stat s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Stat t;
        t.n = __shfl_xor_sync(0xffffffffu, s.n, o);
        t.mean = __shfl_xor_sync(0xffffffffu, s.mean, o);
        t.m2 = __shfl_xor_sync(0xffffffffu, s.m2, o);
        // fixed pairing -> every lane ends with the same bits (xor butterfly merges (lo,hi) in the
        // same order on both sides only if the merge is symmetric; make it so by ordering on lane)
        const bool lo = ((threadIdx.x & o) == 0);
        s = lo ? stat_merge(s, t) : stat_merge(t, s);
    }
    return s;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// activation slope: a launch parameter (LeakyReLU) or a device scalar (PReLU, acti_norm.py:104-110)
template <typename P>
__device__ __forceinline__ float load_slope(const P& p) {
    return p.slope_dev ? __ldg(p.slope_dev) : p.slope;
}

// ---------------------------------------------------------------------------------------------
// style / affine lookup
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int load_style(const long long* styles, long long n, int num_styles, int* status) {
    long long s = styles ? styles[n] : 0;
    if (s < 0) s += num_styles;  // python indexing of the ModuleList (conditional_instance_norm.py:60)
    if (s < 0 || s >= num_styles) {
        if (status) atomicOr(status, 1);
        s = s < 0 ? 0 : num_styles - 1;
    }
    return (int)s;
}

template <typename P>
__device__ __forceinline__ void load_affine(const P& p, int s, long long c, float& g, float& b) {
    if (p.affine) {
        // pointer tables live in the kernel parameter block; a dynamic index would force a local
        // copy, so select with a short unrolled scan
        const float* gp = p.gamma[0];
        const float* bp = p.beta[0];
#pragma unroll
        for (int i = 1; i < kMaxStyles; ++i) {
            if (i == s) {
                gp = p.gamma[i];
                bp = p.beta[i];
            }
        }
        g = __ldg(gp + c);
        b = __ldg(bp + c);
    } else {
        g = 1.f;
        b = 0.f;
    }
}

// the same for a __grid_constant__ parameter block: the pointer tables stay in the constant bank and are
// indexed there (no local copy, no scan)
template <typename P>
__device__ __forceinline__ void load_affine_gc(const P& p, int s, long long c, float& g, float& b) {
    if (p.affine) {
        g = __ldg(p.gamma[s] + c);
        b = __ldg(p.beta[s] + c);
    } else {
        g = 1.f;
        b = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// PTX: shared addresses, mbarrier, TMA bulk copies, cluster
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the warp for a system-defined time; a warp choosing between two
// rings must not)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// A lost arrival must surface as a trapped kernel, never as a hung GPU: bound every wait (~4 s).
#ifndef MICN_WAIT_TIMEOUT_NS
#define MICN_WAIT_TIMEOUT_NS 4000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 0x3ffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (((++spins) & 0x3ffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
    }
}

// Waits of helper warps and of consumers that have nothing else to do: try_wait with a suspend-time hint parks
// the warp in hardware until the phase completes (or the hint expires) instead of spinning on issue slots the
// working warps need.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_park_t(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_hint(bar, parity, 1000u)) return;
    uint64_t t0 = 0;  // the clock is read only once a wait has lasted ~256 naps: ordinary waits never touch it
    for (uint32_t spins = 1;; ++spins) {
        __nanosleep(SLEEP_NS);  // the suspend hint alone still lets the warp re-issue every few hundred cycles
        if (mbar_try_wait_hint(bar, parity, 10000u)) return;
        if ((spins & 0xffu) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
        }
    }
}
// consumers (on the critical path: short naps) / helper warps (producers, publish, gather: longer naps)
#ifndef MICN_PARK_NS
#define MICN_PARK_NS 32
#endif
#ifndef MICN_IDLE_NS
#define MICN_IDLE_NS 200
#endif
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity) { mbar_wait_park_t<MICN_PARK_NS>(bar, parity); }
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity) { mbar_wait_park_t<MICN_IDLE_NS>(bar, parity); }

// x / d for x < 2^31 with a precomputed multiplier (host: fastdiv_make)
struct FastDiv {
    unsigned mul, shr;
};
__device__ __forceinline__ unsigned fastdiv(unsigned x, const FastDiv& d) {
    return d.mul ? __umulhi(x, d.mul) >> d.shr : x;
}
inline FastDiv fastdiv_make(unsigned d) {
    FastDiv f{0u, 0u};
    if (d <= 1u) return f;  // mul == 0 marks the identity
    unsigned lg = 0;
    while ((1ull << lg) < d) ++lg;
    const unsigned p = 31u + lg;
    f.mul = (unsigned)(((1ull << p) + d - 1ull) / d);
    f.shr = p - 32u;
    return f;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// 1-D TMA bulk copy global -> this CTA's shared memory; completion is signalled on `bar` as
// `bytes` of transaction count.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// map a shared::cta address of this CTA to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running; `pdl_wait` blocks until the predecessor grid has COMPLETED and its
// memory is visible (a no-op for a normal launch), `pdl_launch_dependents` lets the successor's CTAs be placed as soon as
// SMs free up.  The kernels here put only their shared-memory set-up before the wait: what overlaps is the launch latency
// and the barrier initialisation, never a memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// named barrier over a subset of the CTA's warps (id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace micn
