// micn_cl.cuh - channels-last (token-major) instance_cond forward / backward (sm_100a): SURVEY.md 8(f) row 2.
//
// PatchMerging's norms ([1,384,24^3] ... [1,3072,3^3], networks/blocks/patch_merging.py:136-141) and the 25 ViT token
// norms of C-UNETR ([B,768,216], transformer_block.py:87-92, vit.py:188-193) reach the norm as a permuted view of a
// [N, *spatial, C] tensor: stride_C = 1, the "slab" of one (n, c) is a strided column.  The reference (and the flat /
// small kernels here) pay a transposing copy in and another one out; these kernels reduce the columns in place
// and write the result in the SAME layout, so the `rearrange(...)` that follows the norm in the model is a free view.
//
// Layout: x[n][m][c], c fastest, C even.  A CTA owns a tile of 64 channels (two per lane: one 32-bit load for
// 16-bit types, one 64-bit load for fp32, so a warp reads 128 / 256 contiguous bytes of a row) and one of MS row
// ranges ("splits") of one sample; its 8 warps interleave the rows of the range, four rows in flight per lane.
//
//   stats kernel   per-CTA partial of every channel of the tile -> workspace[n][split][c]
//                  (forward: count, mean, M2 from fp32 shifted sums; backward: sum g, sum g*(x - mean))
//   apply kernel   warp 0 folds the MS partials of the tile in a fixed order (same reference-mean fold as the flat
//                  path: bit-identical in every CTA), then all warps stream their rows again (L2) and write y / dx
//   param kernel   (backward) d(gamma)/d(beta)[s][c] = sum over the samples of style s of the per-slab sums
//
// The tensors on this route are small (<= 20 MB in the models), so parallelism (MS is chosen for >= 2 CTAs per SM)
// and launch count matter more than the last percent of bandwidth.
#pragma once

#include "micn_common.cuh"

namespace micn {

constexpr int kClThreads = 256;
constexpr int kClWarps = kClThreads / 32;
constexpr int kClTile = 64;  // channels per CTA, two per lane
constexpr int kClMaxSplits = 64;

struct ClParams {
    const void* x;
    void* y;         // forward output / backward dx
    const void* dy;  // backward only
    const float* gamma[kMaxStyles];
    const float* beta[kMaxStyles];
    const long long* styles;
    float* save_mean;  // [N*C]  written by the forward, read by the backward
    float* save_rstd;
    float* dgamma;     // [S*C] or null
    float* dbeta;
    int* status;
    float4* ws_part;   // [N][MS][C]
    float2* ws_slab;   // [N][C]  backward: (sum g, sum g*xhat)
    unsigned* tile_cnt;  // [channel tiles] samples of a tile that delivered their sums (workspace prefix, self-resetting)
    long long N, C, M, rows_per_split;
    int MS, num_styles, affine;
    float eps;
};

// two adjacent channels of one row
template <typename T>
struct ClPair;
template <>
struct ClPair<float> {
    __device__ __forceinline__ static float2 load(const void* base, size_t elem) {
        return __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + elem));
    }
    __device__ __forceinline__ static void store(void* base, size_t elem, float a, float b) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + elem) = make_float2(a, b);
    }
};
template <>
struct ClPair<__nv_bfloat16> {
    __device__ __forceinline__ static float2 load(const void* base, size_t elem) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem));
        return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
    __device__ __forceinline__ static void store(void* base, size_t elem, float a, float b) {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = VecT<__nv_bfloat16>::pack2(a, b);
    }
};
template <>
struct ClPair<__half> {
    __device__ __forceinline__ static float2 load(const void* base, size_t elem) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const __half*>(base) + elem));
        float lo, hi;
        VecT<__half>::unpack2(u, lo, hi);
        return make_float2(lo, hi);
    }
    __device__ __forceinline__ static void store(void* base, size_t elem, float a, float b) {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(base) + elem) = VecT<__half>::pack2(a, b);
    }
};

struct ClTile {
    long long n, r0, r1;  // sample, row range of this CTA
    long long c;          // first of this lane's two channels
    bool valid;
    int warp, lane, split;
};
__device__ __forceinline__ ClTile cl_tile(const ClParams& p) {
    ClTile t;
    t.warp = threadIdx.x >> 5;
    t.lane = threadIdx.x & 31;
    t.n = blockIdx.z;
    t.split = blockIdx.y;
    t.c = (long long)blockIdx.x * kClTile + 2 * t.lane;
    t.valid = t.c < p.C;  // C is even: c + 1 is a channel too
    t.r0 = (long long)t.split * p.rows_per_split;
    t.r1 = t.r0 + p.rows_per_split < p.M ? t.r0 + p.rows_per_split : p.M;
    return t;
}

// rows r0 + warp, + W, ... of the CTA's range (W warps), kClInFlight rows in flight per lane (the loads are 4-8 bytes:
// memory-level parallelism, not load width, is what keeps the SM fed); f(x pair, second pair or dummy)
constexpr int kClInFlight = 8;
template <typename T, bool TWO, int W = kClWarps, int INF = kClInFlight, typename F>
__device__ __forceinline__ void cl_rows(const ClParams& p, const ClTile& t, const void* a, const void* b, F f) {
    constexpr int kClInFlight = INF;  // (shadows the default inside this function)
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    long long r = t.r0 + t.warp;
    for (; r + (kClInFlight - 1) * W < t.r1; r += kClInFlight * W) {
        float2 va[kClInFlight], vb[kClInFlight];
#pragma unroll
        for (int i = 0; i < kClInFlight; ++i) {
            const size_t e = base + (size_t)(r + i * W) * (size_t)p.C;
            va[i] = ClPair<T>::load(a, e);
            if (TWO) vb[i] = ClPair<T>::load(b, e);
        }
#pragma unroll
        for (int i = 0; i < kClInFlight; ++i) f(r + i * W, va[i], TWO ? vb[i] : make_float2(0.f, 0.f));
    }
    for (; r + W < t.r1; r += 2 * W) {  // remainder, two rows at a time
        const size_t e0 = base + (size_t)r * (size_t)p.C, e1 = base + (size_t)(r + W) * (size_t)p.C;
        const float2 a0 = ClPair<T>::load(a, e0), a1 = ClPair<T>::load(a, e1);
        const float2 b0 = TWO ? ClPair<T>::load(b, e0) : make_float2(0.f, 0.f);
        const float2 b1 = TWO ? ClPair<T>::load(b, e1) : make_float2(0.f, 0.f);
        f(r, a0, b0);
        f(r + W, a1, b1);
    }
    for (; r < t.r1; r += W) {
        const size_t e = base + (size_t)r * (size_t)p.C;
        const float2 va = ClPair<T>::load(a, e);
        const float2 vb = TWO ? ClPair<T>::load(b, e) : make_float2(0.f, 0.f);
        f(r, va, vb);
    }
}

// sum the W warps' (a0, b0, a1, b1) of every lane in a fixed order; the result is valid in warp 0
template <int W = kClWarps>
__device__ __forceinline__ float4 cl_block_sum(float4 v, float4* sm, int warp, int lane) {
    sm[warp * 32 + lane] = v;
    __syncthreads();
    float4 t = sm[lane];
    if (warp == 0) {
#pragma unroll
        for (int w = 1; w < W; ++w) {
            const float4 o = sm[w * 32 + lane];
            t.x += o.x;
            t.y += o.y;
            t.z += o.z;
            t.w += o.w;
        }
    }
    return t;
}

// ---------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_fwd_stats_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 sm[kClThreads];
    const ClTile t = cl_tile(p);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // (s0, q0, s1, q1) about the shift
    float2 K = make_float2(0.f, 0.f);
    if (t.valid && t.r0 < t.r1) {
        K = ClPair<T>::load(p.x, (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.r0 * (size_t)p.C + (size_t)t.c);
        cl_rows<T, false>(p, t, p.x, nullptr, [&](long long, float2 v, float2) {
            const float d0 = v.x - K.x, d1 = v.y - K.y;
            acc.x += d0;
            acc.y = fmaf(d0, d0, acc.y);
            acc.z += d1;
            acc.w = fmaf(d1, d1, acc.w);
        });
    }
    acc = cl_block_sum(acc, sm, t.warp, t.lane);
    if (t.warp == 0 && t.valid) {
        const float cnt = (float)(t.r1 > t.r0 ? t.r1 - t.r0 : 0);
        float4* out = p.ws_part + ((size_t)t.n * p.MS + t.split) * (size_t)p.C + (size_t)t.c;
        const float m0 = cnt > 0.f ? acc.x / cnt : 0.f, m1 = cnt > 0.f ? acc.z / cnt : 0.f;
        out[0] = make_float4(cnt, K.x + m0, fmaxf(acc.y - acc.x * m0, 0.f), 0.f);
        out[1] = make_float4(cnt, K.y + m1, fmaxf(acc.w - acc.z * m1, 0.f), 0.f);
    }
}

// fold of the MS partials (count, mean, M2) of one channel about split 0's mean: same algebra as the flat path
__device__ __forceinline__ float2 cl_fold_stats(const float4* part, int MS, size_t stride, float M) {
    const float ref = part[0].y;
    float A = 0.f, B = 0.f;
    for (int s = 0; s < MS; ++s) {
        const float4 q = part[(size_t)s * stride];
        const float d = q.y - ref;
        A = fmaf(q.x, d, A);
        B += fmaf(q.x * d, d, q.z);
    }
    const float m = A / M;
    return make_float2(ref + m, fmaxf(B - A * m, 0.f));  // (mean, M2)
}

template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_fwd_apply_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 coef[32];  // (a0, b0, a1, b1) per lane
    const ClTile t = cl_tile(p);
    if (t.warp == 0) {
        float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t.valid) {
            const int style = load_style(p.styles, t.n, p.num_styles, p.status);
            const float4* part = p.ws_part + (size_t)t.n * p.MS * (size_t)p.C + (size_t)t.c;
            float ab[4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float2 st = cl_fold_stats(part + k, p.MS, (size_t)p.C, (float)p.M);
                const float rstd = 1.f / sqrtf(st.y / (float)p.M + p.eps);  // biased variance, eps inside the sqrt
                float gamma, beta;
                load_affine(p, style, t.c + k, gamma, beta);
                const float a = rstd * gamma;
                ab[2 * k] = a;
                ab[2 * k + 1] = fmaf(-st.x, a, beta);
                if (t.split == 0 && p.save_mean) {
                    p.save_mean[t.n * p.C + t.c + k] = st.x;
                    p.save_rstd[t.n * p.C + t.c + k] = rstd;
                }
            }
            cf = make_float4(ab[0], ab[1], ab[2], ab[3]);
        }
        coef[t.lane] = cf;
    }
    __syncthreads();
    if (!t.valid) return;
    const float4 cf = coef[t.lane];
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    cl_rows<T, false>(p, t, p.x, nullptr, [&](long long r, float2 v, float2) {
        ClPair<T>::store(p.y, base + (size_t)r * (size_t)p.C, fmaf(v.x, cf.x, cf.y), fmaf(v.y, cf.z, cf.w));
    });
}

// Parameter gradients from warp 0 of the CTA that owns (sample, channel tile).  One sample: the slab sums are the
// gradients of its style's rows.  More: every sample delivers its sums, the last one of the tile to arrive (arrival
// counter in the workspace prefix) folds the tile, samples in order - deterministic, no extra launch.  Without a
// counter (more tiles than the prefix holds) micn_cl_param_grads_kernel does the fold.
__device__ __forceinline__ void cl_param_grads(const ClParams& p, const ClTile& t, const float* S1, const float* S2r, int style) {
    if (!p.dgamma) return;
    if (p.N == 1) {
        if (t.valid)
            for (int k = 0; k < 2; ++k)
                for (int s = 0; s < p.num_styles; ++s) {
                    p.dbeta[(long long)s * p.C + t.c + k] = s == style ? S1[k] : 0.f;
                    p.dgamma[(long long)s * p.C + t.c + k] = s == style ? S2r[k] : 0.f;
                }
        return;
    }
    if (t.valid)
        for (int k = 0; k < 2; ++k) p.ws_slab[t.n * p.C + t.c + k] = make_float2(S1[k], S2r[k]);
    if (!p.tile_cnt) return;
    __threadfence();
    __syncwarp();
    unsigned prev = 0u;
    if (t.lane == 0) prev = atomicAdd(p.tile_cnt + blockIdx.x, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != (unsigned)p.N - 1u) return;
    if (t.lane == 0) p.tile_cnt[blockIdx.x] = 0u;  // reusable by the next launch
    __threadfence();
    if (!t.valid) return;
    for (int k = 0; k < 2; ++k)
        for (int s = 0; s < p.num_styles; ++s) {
            float db = 0.f, dg = 0.f;
            for (long long n = 0; n < p.N; ++n) {
                const float2 v = __ldcg(p.ws_slab + n * p.C + t.c + k);
                const bool mine = load_style(p.styles, n, p.num_styles, nullptr) == s;
                db += mine ? v.x : 0.f;
                dg += mine ? v.y : 0.f;
            }
            p.dbeta[(long long)s * p.C + t.c + k] = db;
            p.dgamma[(long long)s * p.C + t.c + k] = dg;
        }
}

// ---------------------------------------------------------------- backward
template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_bwd_stats_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 sm[kClThreads];
    const ClTile t = cl_tile(p);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // (S1_0, S2_0, S1_1, S2_1)
    if (t.valid && t.r0 < t.r1) {
        const float m0 = __ldg(p.save_mean + t.n * p.C + t.c), m1 = __ldg(p.save_mean + t.n * p.C + t.c + 1);
        cl_rows<T, true>(p, t, p.x, p.dy, [&](long long, float2 v, float2 g) {
            acc.x += g.x;
            acc.y = fmaf(g.x, v.x - m0, acc.y);
            acc.z += g.y;
            acc.w = fmaf(g.y, v.y - m1, acc.w);
        });
    }
    acc = cl_block_sum(acc, sm, t.warp, t.lane);
    if (t.warp == 0 && t.valid) {
        float4* out = p.ws_part + ((size_t)t.n * p.MS + t.split) * (size_t)p.C + (size_t)t.c;
        out[0] = make_float4(acc.x, acc.y, 0.f, 0.f);
        out[1] = make_float4(acc.z, acc.w, 0.f, 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_bwd_apply_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 coefA[32];  // (A, B1, B0', mean) per channel 0
    __shared__ float4 coefB[32];  //                  ... channel 1
    const ClTile t = cl_tile(p);
    if (t.warp == 0) {
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        float S1v[2] = {0.f, 0.f}, S2rv[2] = {0.f, 0.f};
        int style = 0;
        if (t.valid) {
            style = load_style(p.styles, t.n, p.num_styles, p.status);
            const float invM = 1.f / (float)p.M;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float4* part = p.ws_part + (size_t)t.n * p.MS * (size_t)p.C + (size_t)t.c + k;
                float S1 = 0.f, S2 = 0.f;
                for (int s = 0; s < p.MS; ++s) {  // fixed order
                    const float4 q = part[(size_t)s * (size_t)p.C];
                    S1 += q.x;
                    S2 += q.y;
                }
                const float mean = __ldg(p.save_mean + t.n * p.C + t.c + k), rstd = __ldg(p.save_rstd + t.n * p.C + t.c + k);
                float gamma, beta;
                load_affine(p, style, t.c + k, gamma, beta);
                const float a = rstd * gamma, S2r = S2 * rstd;
                // dx = a*g - a*S1/M - a*rstd*(S2r/M)*(x - mean)
                const float B1 = -a * S2r * invM * rstd, B0 = fmaf(-B1, mean, -a * S1 * invM);
                (k == 0 ? c0 : c1) = make_float4(a, B1, B0, 0.f);
                S1v[k] = S1;
                S2rv[k] = S2r;
            }
        }
        coefA[t.lane] = c0;
        coefB[t.lane] = c1;
        if (t.split == 0) cl_param_grads(p, t, S1v, S2rv, style);
    }
    __syncthreads();
    if (!t.valid) return;
    const float4 c0 = coefA[t.lane], c1 = coefB[t.lane];
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    cl_rows<T, true>(p, t, p.x, p.dy, [&](long long r, float2 v, float2 g) {
        ClPair<T>::store(p.y, base + (size_t)r * (size_t)p.C, fmaf(c0.x, g.x, fmaf(c0.y, v.x, c0.z)),
                         fmaf(c1.x, g.y, fmaf(c1.y, v.y, c1.z)));
    });
}

// ---------------------------------------------------------------- wide variant of the two-kernel path
// Same tiles, same partials, same folds, but the rows are streamed with 16-byte loads: a lane owns CPL = 8 (16-bit) or
// 4 (fp32) adjacent channels, LPR = 64 / CPL lanes cover a row of the tile and a warp instruction covers 32 / LPR
// rows, four instructions in flight (64 bytes per lane instead of 32 in eight 4-byte loads).  Needs C % CPL == 0
// and 16-byte aligned tensors; the host falls back to the pair kernels above otherwise.  Forward only: measured on
// [8,384,24^3] bf16 the forward gains 14 % (74 -> 64 us), the two-stream backward loses (95 -> 113 us) and keeps the
// pair kernels.
template <typename T>
struct ClWide {
    static constexpr int CPL = 16 / (int)sizeof(T);
    static constexpr int LPR = kClTile / CPL;
    static constexpr int RPI = 32 / LPR;
};

struct ClWideTile {
    long long n, r0, r1, c;  // sample, row range, first of this lane's CPL channels
    bool valid;
    int warp, rsub, cg, split;
};
template <typename T>
__device__ __forceinline__ ClWideTile cl_wide_tile(const ClParams& p) {
    using W = ClWide<T>;
    ClWideTile t;
    const int lane = threadIdx.x & 31;
    t.warp = threadIdx.x >> 5;
    t.cg = lane % W::LPR;
    t.rsub = lane / W::LPR;
    t.n = blockIdx.z;
    t.split = blockIdx.y;
    t.c = (long long)blockIdx.x * kClTile + (long long)t.cg * W::CPL;
    t.valid = t.c < p.C;  // C % CPL == 0: the whole vector is inside
    t.r0 = (long long)t.split * p.rows_per_split;
    t.r1 = t.r0 + p.rows_per_split < p.M ? t.r0 + p.rows_per_split : p.M;
    return t;
}

template <typename T, bool TWO, typename F>
__device__ __forceinline__ void cl_rows_wide(const ClParams& p, const ClWideTile& t, const void* a, const void* b, F f) {
    using W = ClWide<T>;
    constexpr long long step = (long long)kClWarps * W::RPI;
    const T* pa = reinterpret_cast<const T*>(a);
    const T* pb = reinterpret_cast<const T*>(b);
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    long long r = t.r0 + (long long)t.warp * W::RPI + t.rsub;
    for (; r + 3 * step < t.r1; r += 4 * step) {
        uint4 va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t e = base + (size_t)(r + i * step) * (size_t)p.C;
            va[i] = __ldg(reinterpret_cast<const uint4*>(pa + e));
            if (TWO) vb[i] = __ldg(reinterpret_cast<const uint4*>(pb + e));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) f(r + i * step, va[i], TWO ? vb[i] : make_uint4(0u, 0u, 0u, 0u));
    }
    for (; r < t.r1; r += step) {
        const size_t e = base + (size_t)r * (size_t)p.C;
        const uint4 va = __ldg(reinterpret_cast<const uint4*>(pa + e));
        const uint4 vb = TWO ? __ldg(reinterpret_cast<const uint4*>(pb + e)) : make_uint4(0u, 0u, 0u, 0u);
        f(r, va, vb);
    }
}

// per-lane (u[j], v[j]) of CPL channels -> per-channel sums of the CTA in a fixed order; thread ch < 64 gets (U, V)
template <typename T>
__device__ __forceinline__ float2 cl_wide_block_sum(float* u, float* v, float2 (*sm)[kClTile], const ClWideTile& t) {
    using W = ClWide<T>;
#pragma unroll
    for (int off = W::LPR; off < 32; off <<= 1)
#pragma unroll
        for (int j = 0; j < W::CPL; ++j) {
            u[j] += __shfl_xor_sync(0xffffffffu, u[j], off);
            v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
        }
    if (t.rsub == 0)
#pragma unroll
        for (int j = 0; j < W::CPL; ++j) sm[t.warp][t.cg * W::CPL + j] = make_float2(u[j], v[j]);
    __syncthreads();
    float2 tot = make_float2(0.f, 0.f);
    if (threadIdx.x < kClTile) {
#pragma unroll
        for (int w = 0; w < kClWarps; ++w) {
            tot.x += sm[w][threadIdx.x].x;
            tot.y += sm[w][threadIdx.x].y;
        }
    }
    return tot;
}

template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_fwd_stats_wide_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    using W = ClWide<T>;
    using V = VecT<T>;
    __shared__ float2 sm[kClWarps][kClTile];
    const ClWideTile t = cl_wide_tile<T>(p);
    float s[W::CPL], q[W::CPL], K[W::CPL];
#pragma unroll
    for (int j = 0; j < W::CPL; ++j) s[j] = q[j] = K[j] = 0.f;
    if (t.valid && t.r0 < t.r1) {
        const T* px = reinterpret_cast<const T*>(p.x);
        V::unpack(__ldg(reinterpret_cast<const uint4*>(px + (size_t)t.n * (size_t)p.M * (size_t)p.C +
                                                       (size_t)t.r0 * (size_t)p.C + (size_t)t.c)), K);
        cl_rows_wide<T, false>(p, t, p.x, nullptr, [&](long long, const uint4& xv, const uint4&) {
            float f[W::CPL];
            V::unpack(xv, f);
#pragma unroll
            for (int j = 0; j < W::CPL; ++j) {
                const float d = f[j] - K[j];
                s[j] += d;
                q[j] = fmaf(d, d, q[j]);
            }
        });
    }
    const float2 tot = cl_wide_block_sum<T>(s, q, sm, t);
    const long long ch = (long long)blockIdx.x * kClTile + threadIdx.x;
    if (threadIdx.x < kClTile && ch < p.C) {
        const float cnt = (float)(t.r1 > t.r0 ? t.r1 - t.r0 : 0);
        float Kc = 0.f;
        if (cnt > 0.f)
            Kc = V::load1(reinterpret_cast<const T*>(p.x) + (size_t)t.n * (size_t)p.M * (size_t)p.C +
                          (size_t)t.r0 * (size_t)p.C + (size_t)ch);
        const float m = cnt > 0.f ? tot.x / cnt : 0.f;
        p.ws_part[((size_t)t.n * p.MS + t.split) * (size_t)p.C + (size_t)ch] =
            make_float4(cnt, Kc + m, fmaxf(tot.y - tot.x * m, 0.f), 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads) micn_cl_fwd_apply_wide_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    using W = ClWide<T>;
    using V = VecT<T>;
    __shared__ float4 coef[32];  // (a0, b0, a1, b1) per channel pair
    const ClTile t2 = cl_tile(p);  // the pair mapping, for the fold by warp 0
    if (t2.warp == 0) {
        float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t2.valid) {
            const int style = load_style(p.styles, t2.n, p.num_styles, p.status);
            const float4* part = p.ws_part + (size_t)t2.n * p.MS * (size_t)p.C + (size_t)t2.c;
            float ab[4];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float2 st = cl_fold_stats(part + k, p.MS, (size_t)p.C, (float)p.M);
                const float rstd = 1.f / sqrtf(st.y / (float)p.M + p.eps);
                float gamma, beta;
                load_affine(p, style, t2.c + k, gamma, beta);
                const float a = rstd * gamma;
                ab[2 * k] = a;
                ab[2 * k + 1] = fmaf(-st.x, a, beta);
                if (t2.split == 0 && p.save_mean) {
                    p.save_mean[t2.n * p.C + t2.c + k] = st.x;
                    p.save_rstd[t2.n * p.C + t2.c + k] = rstd;
                }
            }
            cf = make_float4(ab[0], ab[1], ab[2], ab[3]);
        }
        coef[t2.lane] = cf;
    }
    __syncthreads();
    const ClWideTile t = cl_wide_tile<T>(p);
    if (!t.valid) return;
    float ca[W::CPL], cb[W::CPL];
#pragma unroll
    for (int j = 0; j < W::CPL; ++j) {
        const int ci = t.cg * W::CPL + j;
        const float4 c4 = coef[ci >> 1];
        ca[j] = (ci & 1) ? c4.z : c4.x;
        cb[j] = (ci & 1) ? c4.w : c4.y;
    }
    T* py = reinterpret_cast<T*>(p.y);
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    cl_rows_wide<T, false>(p, t, p.x, nullptr, [&](long long r, const uint4& xv, const uint4&) {
        float f[W::CPL];
        V::unpack(xv, f);
#pragma unroll
        for (int j = 0; j < W::CPL; ++j) f[j] = fmaf(f[j], ca[j], cb[j]);
        *reinterpret_cast<uint4*>(py + base + (size_t)r * (size_t)p.C) = V::pack(f);
    });
}

// ---------------------------------------------------------------- fused kernels: short columns (M <= kClFusedMaxRows)
// The 25 ViT token norms of C-UNETR ([B,768,216]) and the deep PatchMerging norms ([1,1536,6^3], [1,3072,3^3]) are
// launch-bound: one 32-warp CTA per (sample, 64-channel tile) reduces its columns and applies the result in the same
// launch (second pass from L1/L2); with one sample the parameter gradients are the slab sums themselves, so a
// forward + backward pair is 2 launches instead of 5.
constexpr int kClFusedThreads = 1024;
constexpr int kClFusedWarps = kClFusedThreads / 32;
constexpr int kClFusedMaxRows = 1024;

template <typename T>
__global__ void __launch_bounds__(kClFusedThreads) micn_cl_fwd_fused_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 sm[kClFusedThreads];
    __shared__ float4 coef[32];
    const ClTile t = cl_tile(p);  // MS == 1: the CTA's row range is the whole column
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 K = make_float2(0.f, 0.f);
    if (t.valid) {
        K = ClPair<T>::load(p.x, (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c);
        cl_rows<T, false, kClFusedWarps, 4>(p, t, p.x, nullptr, [&](long long, float2 v, float2) {
            const float d0 = v.x - K.x, d1 = v.y - K.y;
            acc.x += d0;
            acc.y = fmaf(d0, d0, acc.y);
            acc.z += d1;
            acc.w = fmaf(d1, d1, acc.w);
        });
    }
    acc = cl_block_sum<kClFusedWarps>(acc, sm, t.warp, t.lane);
    if (t.warp == 0) {
        float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t.valid) {
            const int style = load_style(p.styles, t.n, p.num_styles, p.status);
            const float cnt = (float)p.M;
            const float s[2] = {acc.x, acc.z}, q[2] = {acc.y, acc.w}, k[2] = {K.x, K.y};
            float ab[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float m = s[j] / cnt, mean = k[j] + m, M2 = fmaxf(q[j] - s[j] * m, 0.f);
                const float rstd = 1.f / sqrtf(M2 / cnt + p.eps);  // biased variance, eps inside the sqrt
                float gamma, beta;
                load_affine(p, style, t.c + j, gamma, beta);
                const float a = rstd * gamma;
                ab[2 * j] = a;
                ab[2 * j + 1] = fmaf(-mean, a, beta);
                if (p.save_mean) {
                    p.save_mean[t.n * p.C + t.c + j] = mean;
                    p.save_rstd[t.n * p.C + t.c + j] = rstd;
                }
            }
            cf = make_float4(ab[0], ab[1], ab[2], ab[3]);
        }
        coef[t.lane] = cf;
    }
    __syncthreads();
    if (!t.valid) return;
    const float4 cf = coef[t.lane];
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    cl_rows<T, false, kClFusedWarps, 4>(p, t, p.x, nullptr, [&](long long r, float2 v, float2) {
        ClPair<T>::store(p.y, base + (size_t)r * (size_t)p.C, fmaf(v.x, cf.x, cf.y), fmaf(v.y, cf.z, cf.w));
    });
}

template <typename T>
__global__ void __launch_bounds__(kClFusedThreads) micn_cl_bwd_fused_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    __shared__ float4 sm[kClFusedThreads];
    __shared__ float4 coefA[32];
    __shared__ float4 coefB[32];
    const ClTile t = cl_tile(p);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // (S1_0, S2_0, S1_1, S2_1)
    float m[2] = {0.f, 0.f};
    if (t.valid) {
        m[0] = __ldg(p.save_mean + t.n * p.C + t.c);
        m[1] = __ldg(p.save_mean + t.n * p.C + t.c + 1);
        cl_rows<T, true, kClFusedWarps, 4>(p, t, p.x, p.dy, [&](long long, float2 v, float2 g) {
            acc.x += g.x;
            acc.y = fmaf(g.x, v.x - m[0], acc.y);
            acc.z += g.y;
            acc.w = fmaf(g.y, v.y - m[1], acc.w);
        });
    }
    acc = cl_block_sum<kClFusedWarps>(acc, sm, t.warp, t.lane);
    if (t.warp == 0) {
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        float S2rv[2] = {0.f, 0.f};
        const float S1v[2] = {acc.x, acc.z}, S2v[2] = {acc.y, acc.w};
        int style = 0;
        if (t.valid) {
            style = load_style(p.styles, t.n, p.num_styles, p.status);
            const float invM = 1.f / (float)p.M;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float rstd = __ldg(p.save_rstd + t.n * p.C + t.c + k);
                float gamma, beta;
                load_affine(p, style, t.c + k, gamma, beta);
                const float a = rstd * gamma, S1 = S1v[k], S2r = S2v[k] * rstd;
                const float B1 = -a * S2r * invM * rstd, B0 = fmaf(-B1, m[k], -a * S1 * invM);
                (k == 0 ? c0 : c1) = make_float4(a, B1, B0, 0.f);
                S2rv[k] = S2r;
            }
        }
        coefA[t.lane] = c0;
        coefB[t.lane] = c1;
        cl_param_grads(p, t, S1v, S2rv, style);
    }
    __syncthreads();
    if (!t.valid) return;
    const float4 c0 = coefA[t.lane], c1 = coefB[t.lane];
    const size_t base = (size_t)t.n * (size_t)p.M * (size_t)p.C + (size_t)t.c;
    cl_rows<T, true, kClFusedWarps, 4>(p, t, p.x, p.dy, [&](long long r, float2 v, float2 g) {
        ClPair<T>::store(p.y, base + (size_t)r * (size_t)p.C, fmaf(c0.x, g.x, fmaf(c0.y, v.x, c0.z)),
                         fmaf(c1.x, g.y, fmaf(c1.y, v.y, c1.z)));
    });
}

// d(gamma)/d(beta)[s][c] = sum over the samples of style s, in sample order (deterministic)
__global__ void micn_cl_param_grads_kernel(const ClParams p) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.num_styles * p.C) return;
    const int s = (int)(idx / p.C);
    const long long c = idx - (long long)s * p.C;
    float db = 0.f, dg = 0.f;
    for (long long n = 0; n < p.N; ++n) {
        if (load_style(p.styles, n, p.num_styles, nullptr) == s) {
            const float2 v = p.ws_slab[n * p.C + c];
            db += v.x;
            dg += v.y;
        }
    }
    p.dbeta[idx] = db;
    p.dgamma[idx] = dg;
}

}  // namespace micn
