// micn_small.cuh - instance_cond forward / backward for SMALL or UNALIGNED slabs (sm_100a).
//
// One thread group (a warp, a 256-thread CTA or a 1024-thread CTA, chosen on the host from the slab
// size) owns one (n, c) slab.  The deep levels of the MI-Seg encoders produce thousands of tiny slabs
// ([1,192,12^3], [1,384,6^3], [1,768,3^3]) and the ViT/PatchMerging calls produce odd lengths
// (27, 216 tokens): the cluster/TMA machinery does not pay there, parallelism comes from the number
// of slabs instead.  Pass 2 re-reads the slab from L1/L2 (<= 64 KB per group, 8 groups per SM), or - REG
// instantiation, 8-32 KB slabs with a CTA each - takes it from registers.  Parameter gradients: one sample writes
// them directly; several samples fold per channel when its last sample arrives (no tail in one CTA).
//
// Any alignment is accepted: when source and destination slabs share the same misalignment modulo
// 16 bytes the body is peeled into scalar head / 128-bit body / scalar tail, otherwise all-scalar.
#pragma once

#include "micn_common.cuh"

namespace micn {

constexpr int kSmallRegVecs = 8;  // 16-byte vectors of a slab one thread keeps in registers (see the forward kernel)

template <int TPS>
struct SmallCfg {
    static constexpr int BLOCK = TPS < 256 ? 256 : TPS;
    static constexpr int GROUPS = BLOCK / TPS;
    static constexpr int WARPS = TPS / 32;
};

// Peeled iteration space of one slab: [0,head) scalar, nvec vectors, [tail0, M) scalar.
struct Peel {
    long long head, nvec, tail0;
};

template <typename T>
__device__ __forceinline__ Peel make_peel(long long M, const void* a, const void* b, const void* c, const void* d,
                                          const void* e) {
    constexpr int VN = VecT<T>::N;
    const uintptr_t m = reinterpret_cast<uintptr_t>(a) & 15u;
    bool same = true;
    if (b) same = same && ((reinterpret_cast<uintptr_t>(b) & 15u) == m);
    if (c) same = same && ((reinterpret_cast<uintptr_t>(c) & 15u) == m);
    if (d) same = same && ((reinterpret_cast<uintptr_t>(d) & 15u) == m);
    if (e) same = same && ((reinterpret_cast<uintptr_t>(e) & 15u) == m);
    Peel p;
    if (!same || (m % sizeof(T)) != 0) {
        p.head = M;
        p.nvec = 0;
        p.tail0 = M;
        return p;
    }
    long long head = (long long)(((16u - m) & 15u) / sizeof(T));
    if (head > M) head = M;
    p.head = head;
    p.nvec = (M - head) / VN;
    p.tail0 = head + p.nvec * VN;
    return p;
}

// group-wide reductions.  TPS == 32: shuffles only.  Otherwise warp shuffle + shared memory; every
// thread then folds the warp partials in warp order (same bits everywhere, one barrier).
template <int TPS>
__device__ __forceinline__ Stat group_reduce_stat(Stat s, float* scratch /* [WARPS*4] */) {
    s = stat_warp_reduce(s);
    if (TPS == 32) return s;
    constexpr int W = TPS / 32;
    const int warp = (threadIdx.x % TPS) >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // scratch reuse across calls
    if (lane == 0) {
        scratch[warp * 4 + 0] = s.n;
        scratch[warp * 4 + 1] = s.mean;
        scratch[warp * 4 + 2] = s.m2;
    }
    __syncthreads();
    Stat t{scratch[0], scratch[1], scratch[2]};
    for (int w = 1; w < W; ++w) t = stat_merge(t, Stat{scratch[w * 4 + 0], scratch[w * 4 + 1], scratch[w * 4 + 2]});
    return t;
}

template <int TPS>
__device__ __forceinline__ void group_reduce_sum2(float& a, float& b, float* scratch) {
    a = warp_sum(a);
    b = warp_sum(b);
    if (TPS == 32) return;
    constexpr int W = TPS / 32;
    const int warp = (threadIdx.x % TPS) >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
        scratch[warp * 4 + 0] = a;
        scratch[warp * 4 + 1] = b;
    }
    __syncthreads();
    float ta = scratch[0], tb = scratch[1];
    for (int w = 1; w < W; ++w) {
        ta += scratch[w * 4 + 0];
        tb += scratch[w * 4 + 1];
    }
    a = ta;
    b = tb;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int EPI, int TPS, bool REG = false>
__global__ void __launch_bounds__(SmallCfg<TPS>::BLOCK) micn_fwd_small_kernel(const FwdParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    using V = VecT<T>;
    constexpr int VN = V::N;
    __shared__ float scratch[32 * 4];
    const int t = threadIdx.x % TPS;
    const long long slab = (long long)blockIdx.x * SmallCfg<TPS>::GROUPS + threadIdx.x / TPS;
    if (slab >= p.N * p.C) return;  // only possible when TPS == 32 (no CTA barriers on that path)
    const long long n = slab / p.C, ch = slab - n * p.C;
    const int style = load_style(p.styles, n, p.num_styles, p.status);
    float gamma, beta;
    load_affine(p, style, ch, gamma, beta);

    const T* xs = reinterpret_cast<const T*>(p.x) + n * p.x_sN + ch * p.x_sC;
    T* ys = reinterpret_cast<T*>(p.y) + slab * p.M;
    const T* rs = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<const T*>(p.res) + slab * p.M : nullptr;
    const Peel pl = make_peel<T>(p.M, xs, ys, rs, nullptr, nullptr);

    // ---- pass 1: per-thread shifted sums
    float K = 0.f, s1 = 0.f, s2 = 0.f, cnt = 0.f;
    bool haveK = false;
    auto acc = [&](float v) {
        if (!haveK) {
            K = v;
            haveK = true;
        }
        const float d = v - K;
        s1 += d;
        s2 = fmaf(d, d, s2);
        cnt += 1.f;
    };
    // REG instantiation (the host picks it for slabs of more than two vectors per thread, up to kSmallRegVecs: 1-4 KB
    // per warp, 8-32 KB per 256-thread CTA): the slab stays in registers, all its loads are in flight at once and the
    // second pass never goes back to memory ([1,96,24^3] bf16 fwd+bwd 12.4 -> 10.4 us).  Smaller slabs keep the lean
    // two-pass instantiation: there parallelism comes from resident warps, and registers cost occupancy.
    constexpr bool kRegPath = REG && TPS <= 256;
    const bool in_regs = kRegPath && pl.nvec <= (long long)kSmallRegVecs * TPS;
    uint4 q[kRegPath ? kSmallRegVecs : 1];
    for (long long i = t; i < pl.head; i += TPS) acc(V::load1(xs + i));
    if (in_regs) {
        const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
#pragma unroll
        for (int k = 0; k < kSmallRegVecs; ++k) {
            const long long i = t + (long long)k * TPS;
            q[k] = i < pl.nvec ? __ldg(xv + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < kSmallRegVecs; ++k) {
            if (t + (long long)k * TPS < pl.nvec) {
                float f[VN];
                V::unpack(q[k], f);
#pragma unroll
                for (int e = 0; e < VN; ++e) acc(f[e]);
            }
        }
    } else {
        const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
        long long i = t;
        for (; i + TPS < pl.nvec; i += 2 * TPS) {  // two loads in flight per thread
            const uint4 q0 = __ldg(xv + i), q1 = __ldg(xv + i + TPS);
            float f[VN];
            V::unpack(q0, f);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc(f[e]);
            V::unpack(q1, f);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc(f[e]);
        }
        if (i < pl.nvec) {
            float f[VN];
            V::unpack(__ldg(xv + i), f);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc(f[e]);
        }
    }
    for (long long i = pl.tail0 + t; i < p.M; i += TPS) acc(V::load1(xs + i));

    const Stat tot = group_reduce_stat<TPS>(stat_from_shifted(K, s1, s2, cnt), scratch);
    const float mean = tot.mean;
    const float rstd = 1.f / sqrtf(tot.m2 / tot.n + p.eps);
    if (t == 0 && p.save_mean) {
        p.save_mean[slab] = mean;
        p.save_rstd[slab] = rstd;
    }
    const float a = rstd * gamma;

    // ---- pass 2
    auto apply = [&](float x, float r) {
        float v = fmaf(x - mean, a, beta);
        if (EPI == MICN_EPI_ADD_LRELU) v += r;
        if (EPI != MICN_EPI_NONE) v = v > 0.f ? v : v * load_slope(p);
        return v;
    };
    for (long long i = t; i < pl.head; i += TPS)
        V::store1(ys + i, apply(V::load1(xs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(rs + i) : 0.f));
    if (in_regs) {
        const uint4* rv = reinterpret_cast<const uint4*>(rs + pl.head);
        uint4* yv = reinterpret_cast<uint4*>(ys + pl.head);
#pragma unroll
        for (int k = 0; k < kSmallRegVecs; ++k) {
            const long long i = t + (long long)k * TPS;
            if (i < pl.nvec) {
                float f[VN], r[VN];
                V::unpack(q[k], f);
                if (EPI == MICN_EPI_ADD_LRELU) V::unpack(ldg_stream(rv + i), r);
#pragma unroll
                for (int e = 0; e < VN; ++e) f[e] = apply(f[e], EPI == MICN_EPI_ADD_LRELU ? r[e] : 0.f);
                stg_stream(yv + i, V::pack(f));
            }
        }
    } else {
        const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
        const uint4* rv = reinterpret_cast<const uint4*>(rs + pl.head);
        uint4* yv = reinterpret_cast<uint4*>(ys + pl.head);
        for (long long i = t; i < pl.nvec; i += TPS) {
            float f[VN], r[VN];
            V::unpack(__ldg(xv + i), f);
            if (EPI == MICN_EPI_ADD_LRELU) V::unpack(ldg_stream(rv + i), r);
#pragma unroll
            for (int e = 0; e < VN; ++e) f[e] = apply(f[e], EPI == MICN_EPI_ADD_LRELU ? r[e] : 0.f);
            stg_stream(yv + i, V::pack(f));
        }
    }
    for (long long i = pl.tail0 + t; i < p.M; i += TPS)
        V::store1(ys + i, apply(V::load1(xs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(rs + i) : 0.f));
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// d(gamma)/d(beta) of one channel: fold the per-slab sums of its N samples per style, samples in order (deterministic);
// four samples' loads in flight, non-matching styles add 0
__device__ __forceinline__ void small_fold_channel(const BwdParams& p, long long ch) {
    for (int s = 0; s < p.num_styles; ++s) {
        float acc_b = 0.f, acc_g = 0.f;
        long long k = 0;
        for (; k + 4 <= p.N; k += 4) {
            float vb[4], vg[4];
            int st[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                st[q] = load_style(p.styles, k + q, p.num_styles, nullptr);
                vb[q] = __ldcg(p.ws_sum_dy + (k + q) * p.C + ch);
                vg[q] = __ldcg(p.ws_sum_dyxh + (k + q) * p.C + ch);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc_b += st[q] == s ? vb[q] : 0.f;
                acc_g += st[q] == s ? vg[q] : 0.f;
            }
        }
        for (; k < p.N; ++k) {
            if (load_style(p.styles, k, p.num_styles, nullptr) == s) {
                acc_b += __ldcg(p.ws_sum_dy + k * p.C + ch);
                acc_g += __ldcg(p.ws_sum_dyxh + k * p.C + ch);
            }
        }
        p.dbeta[(long long)s * p.C + ch] = acc_b;
        p.dgamma[(long long)s * p.C + ch] = acc_g;
        if (p.dgamma2) {  // second norm of the dual epilogue: same sum(g), its own sum(g * xhat2)
            float acc_g2 = 0.f;
            for (long long kk = 0; kk < p.N; ++kk)
                if (load_style(p.styles, kk, p.num_styles, nullptr) == s) acc_g2 += __ldcg(p.ws_sum_dyxh2 + kk * p.C + ch);
            p.dbeta2[(long long)s * p.C + ch] = acc_b;
            p.dgamma2[(long long)s * p.C + ch] = acc_g2;
        }
    }
}

template <typename T, int EPI, int TPS, bool REG = false>
__global__ void __launch_bounds__(SmallCfg<TPS>::BLOCK) micn_bwd_small_kernel(const BwdParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    using V = VecT<T>;
    constexpr int VN = V::N;
    __shared__ float scratch[32 * 4];
    __shared__ int is_last;
    const int t = threadIdx.x % TPS;
    const long long slab = (long long)blockIdx.x * SmallCfg<TPS>::GROUPS + threadIdx.x / TPS;
    const bool active = slab < p.N * p.C;  // false only for trailing warps when TPS == 32

    if (active) {
        const long long n = slab / p.C, ch = slab - n * p.C;
        const int style = load_style(p.styles, n, p.num_styles, p.status);
        float gamma, beta;
        load_affine(p, style, ch, gamma, beta);
        const float mean = __ldg(p.save_mean + slab), rstd = __ldg(p.save_rstd + slab);
        const float a = rstd * gamma;

        const T* xs = reinterpret_cast<const T*>(p.x) + n * p.x_sN + ch * p.x_sC;
        const T* gs = reinterpret_cast<const T*>(p.dy) + slab * p.M;
        const T* os = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<const T*>(p.act_out) + slab * p.M : nullptr;
        T* dxs = reinterpret_cast<T*>(p.dx) + slab * p.M;
        T* drs = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<T*>(p.dres) + slab * p.M : nullptr;
        const Peel pl = make_peel<T>(p.M, xs, gs, os, dxs, drs);

        auto masked = [&](float d, float g, float o) {
            if (EPI == MICN_EPI_LRELU) g = fmaf(d, a, beta) > 0.f ? g : g * load_slope(p);
            if (EPI == MICN_EPI_ADD_LRELU) g = o > 0.f ? g : g * load_slope(p);
            return g;
        };

        // ---- pass 1
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const bool want_ds = EPI == MICN_EPI_LRELU && p.dslope != nullptr;
        auto acc = [&](float x, float g, float o) {
            const float d = x - mean;
            if (want_ds) {  // d prelu / d slope = pre on the negative side
                const float pre = fmaf(d, a, beta);
                s3 += pre > 0.f ? 0.f : g * pre;
            }
            g = masked(d, g, o);
            s1 += g;
            s2 = fmaf(g, d * rstd, s2);
        };
        for (long long i = t; i < pl.head; i += TPS)
            acc(V::load1(xs + i), V::load1(gs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(os + i) : 0.f);
        // register-resident slab (x and dy: 2 x kSmallRegVecs vectors per thread), as in the forward kernel; the
        // residual variant carries a third stream and stays on the two-pass loop
        constexpr bool kRegPath = REG && TPS <= 256 && EPI != MICN_EPI_ADD_LRELU;
        const bool in_regs = kRegPath && pl.nvec <= (long long)kSmallRegVecs * TPS;
        uint4 qx[kRegPath ? kSmallRegVecs : 1], qg[kRegPath ? kSmallRegVecs : 1];
        if (in_regs) {
            const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
            const uint4* gv = reinterpret_cast<const uint4*>(gs + pl.head);
#pragma unroll
            for (int k = 0; k < kSmallRegVecs; ++k) {
                const long long i = t + (long long)k * TPS;
                qx[k] = i < pl.nvec ? __ldg(xv + i) : make_uint4(0u, 0u, 0u, 0u);
                qg[k] = i < pl.nvec ? __ldg(gv + i) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < kSmallRegVecs; ++k) {
                if (t + (long long)k * TPS < pl.nvec) {
                    float xf[VN], gf[VN];
                    V::unpack(qx[k], xf);
                    V::unpack(qg[k], gf);
#pragma unroll
                    for (int e = 0; e < VN; ++e) acc(xf[e], gf[e], 0.f);
                }
            }
        } else {
            const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
            const uint4* gv = reinterpret_cast<const uint4*>(gs + pl.head);
            const uint4* ov = reinterpret_cast<const uint4*>(os + pl.head);
            for (long long i = t; i < pl.nvec; i += TPS) {
                float xf[VN], gf[VN], of[VN];
                V::unpack(__ldg(xv + i), xf);
                V::unpack(__ldg(gv + i), gf);
                if (EPI == MICN_EPI_ADD_LRELU) V::unpack(__ldg(ov + i), of);
#pragma unroll
                for (int e = 0; e < VN; ++e) acc(xf[e], gf[e], EPI == MICN_EPI_ADD_LRELU ? of[e] : 0.f);
            }
        }
        for (long long i = pl.tail0 + t; i < p.M; i += TPS)
            acc(V::load1(xs + i), V::load1(gs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(os + i) : 0.f);

        group_reduce_sum2<TPS>(s1, s2, scratch);
        if (want_ds) {
            float zero = 0.f;
            group_reduce_sum2<TPS>(s3, zero, scratch);
            if (t == 0) p.dslope[slab] = s3;
        }
        if (t == 0 && p.dgamma) {
            if (p.N == 1) {
                // one sample: this slab's sums ARE the gradients of its style's row (no cross-CTA fold, no tail)
                for (int s = 0; s < p.num_styles; ++s) {
                    p.dbeta[(long long)s * p.C + ch] = s == style ? s1 : 0.f;
                    p.dgamma[(long long)s * p.C + ch] = s == style ? s2 : 0.f;
                }
            } else {
                // the last sample of a channel to deliver its sums folds the channel, samples in order (deterministic,
                // and spread over the grid: no serial tail in one CTA)
                p.ws_sum_dy[slab] = s1;
                p.ws_sum_dyxh[slab] = s2;
                __threadfence();
                if (p.ws_chan_cnt && atomicAdd(p.ws_chan_cnt + ch, 1u) == (unsigned)p.N - 1u) {
                    p.ws_chan_cnt[ch] = 0u;  // reusable by the next launch
                    __threadfence();
                    small_fold_channel(p, ch);
                }
            }
        }
        const float invM = 1.f / (float)p.M;
        const float B0 = -a * s1 * invM;
        const float B1 = -a * s2 * invM * rstd;

        // ---- pass 2
        auto grad = [&](float x, float g, float o, float& gout) {
            const float d = x - mean;
            g = masked(d, g, o);
            gout = g;
            return fmaf(a, g, fmaf(B1, d, B0));
        };
        for (long long i = t; i < pl.head; i += TPS) {
            float go;
            const float v = grad(V::load1(xs + i), V::load1(gs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(os + i) : 0.f, go);
            V::store1(dxs + i, v);
            if (EPI == MICN_EPI_ADD_LRELU) V::store1(drs + i, go);
        }
        if (in_regs) {
            uint4* dxv = reinterpret_cast<uint4*>(dxs + pl.head);
#pragma unroll
            for (int k = 0; k < kSmallRegVecs; ++k) {
                const long long i = t + (long long)k * TPS;
                if (i < pl.nvec) {
                    float xf[VN], gf[VN];
                    V::unpack(qx[k], xf);
                    V::unpack(qg[k], gf);
#pragma unroll
                    for (int e = 0; e < VN; ++e) xf[e] = grad(xf[e], gf[e], 0.f, gf[e]);
                    stg_stream(dxv + i, V::pack(xf));
                }
            }
        } else {
            const uint4* xv = reinterpret_cast<const uint4*>(xs + pl.head);
            const uint4* gv = reinterpret_cast<const uint4*>(gs + pl.head);
            const uint4* ov = reinterpret_cast<const uint4*>(os + pl.head);
            uint4* dxv = reinterpret_cast<uint4*>(dxs + pl.head);
            uint4* drv = reinterpret_cast<uint4*>(drs + pl.head);
            for (long long i = t; i < pl.nvec; i += TPS) {
                float xf[VN], gf[VN], of[VN];
                V::unpack(__ldg(xv + i), xf);
                V::unpack(__ldg(gv + i), gf);
                if (EPI == MICN_EPI_ADD_LRELU) V::unpack(__ldg(ov + i), of);
#pragma unroll
                for (int e = 0; e < VN; ++e) xf[e] = grad(xf[e], gf[e], EPI == MICN_EPI_ADD_LRELU ? of[e] : 0.f, gf[e]);
                stg_stream(dxv + i, V::pack(xf));
                if (EPI == MICN_EPI_ADD_LRELU) stg_stream(drv + i, V::pack(gf));
            }
        }
        for (long long i = pl.tail0 + t; i < p.M; i += TPS) {
            float go;
            const float v = grad(V::load1(xs + i), V::load1(gs + i), EPI == MICN_EPI_ADD_LRELU ? V::load1(os + i) : 0.f, go);
            V::store1(dxs + i, v);
            if (EPI == MICN_EPI_ADD_LRELU) V::store1(drs + i, go);
        }
    }

    // ---- more channels than the workspace has arrival counters for: the last CTA folds every channel
    if (p.dgamma && p.N > 1 && !p.ws_chan_cnt) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(p.ws_counter, 1u);
            is_last = (prev == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            for (long long ch = threadIdx.x; ch < p.C; ch += blockDim.x) small_fold_channel(p, ch);
            if (threadIdx.x == 0) *p.ws_counter = 0u;
        }
    }
}

}  // namespace micn
