// micn_res.cuh - shared-memory RESIDENT instance_cond forward / backward (sm_100a): the path for tensors small enough
// that every slab can sit in the shared memory of the SMs at the same time (one wave): the mid-size calls of the
// model - [1,48,48^3], [1,96,32^3], [1,192,24^3] ... - where the flat path's global-memory record exchange, poll delays
// and two launches' worth of pipeline fill cost more than the data movement itself.
//
//   * one thread-block CLUSTER of CS CTAs (1..8) per (n, c) slab, CS chosen on the host so that slabs * CS fills the
//     148 SMs in ONE wave; CTA `rank` owns a contiguous 1/CS share of the slab
//   * the CTA's whole share (every input stream of it) is fetched by 1-D TMA bulk copies issued up front - the entire
//     tensor is in flight within the first microsecond - in 8 KB chunks with one mbarrier each, so the statistics pass
//     consumes chunk k while chunks k+1.. are still landing
//   * pass 1 reduces thread -> warp (shuffle) -> CTA (shared memory) -> cluster: every CTA pushes its partial into every
//     peer's shared memory (st.shared::cluster) and arrives on the peer's mbarrier (release / acquire at cluster scope):
//     ~0.3 us instead of the ~2 us a record needs to cross L2 and be polled; all CTAs fold the CS partials in rank order
//     (bit-identical everywhere, no atomics)
//   * pass 2 normalises straight out of shared memory: each voxel crosses HBM once per tensor and is never re-read, not
//     even from L2
//
// Epilogues: none / lrelu / add_lrelu as the other paths, plus MICN_EPI_NORM_ADD_LRELU - the downsample branch of
// UnetResBlock, y = lrelu(norm2(a) + norm3(b)) (dynunet_block.py:113-125 with conv3 / norm3, :82-98) - which normalises
// TWO tensors in one pass (statistics of both, one output).  Math expressions are the flat path's, so a forward on one
// path and a backward on the other recompute the same LeakyReLU mask.
#pragma once

#include "micn_common.cuh"
#include "micn_small.cuh"  // small_fold_channel

namespace micn {
namespace res {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkVecs = kThreads;        // chunk granule: one 16-byte vector per thread
constexpr int kChunkBytes = kChunkVecs * 16;  // 8 KB
constexpr int kMaxChunks = 27;              // granules of ALL streams together: 27 * 8 KB = 216 KB
constexpr int kMaxCopies = 8;               // bulk copies (= mbarriers) per stream: a copy costs ~0.15 us to issue, so a
                                            // share travels as a few large copies, not as one per granule
constexpr int kMaxCluster = 8;
constexpr int kRec = 4;                     // floats per exchanged record
constexpr int kCtlBytes = kMaxChunks * 8 + 16 + kMaxCluster * kRec * 4 + kWarps * kRec * 4 + 64;

struct Geom {
    unsigned V;        // 16-byte vectors per slab
    unsigned CS;       // CTAs per cluster = per slab
    unsigned nv_base;  // V / CS
    unsigned nv_rem;   // V % CS: the first nv_rem ranks own one vector more
    unsigned nch_max;  // 8 KB granules of the largest share (stream stride in shared memory)
    unsigned cv;       // vectors per bulk copy (a multiple of kChunkVecs)
    long long* trace;  // bring-up only ("res_trace" option): [grid][8] %globaltimer stamps per CTA, or null
};

enum { TR_ENTRY = 0, TR_SETUP, TR_FIRST, TR_P1, TR_CTASUM, TR_XCHG, TR_P2 };
__device__ __forceinline__ void trace(const Geom& g, int ev) {
    if (g.trace && threadIdx.x == 0) g.trace[(size_t)blockIdx.x * 8 + ev] = (long long)globaltimer_ns();
}

__host__ __device__ constexpr int smem_bytes(int NS, int nch) { return NS * nch * kChunkBytes + kCtlBytes; }

struct Ctx {
    uint32_t data0, full0, xbar;
    float* peer;   // [kMaxCluster][kRec]
    float* wpart;  // [kWarps][kRec]
    float* bcast;  // [16]
    unsigned rank, CS, v0, nv, nch, cv, stream_stride;  // nch: bulk copies of this CTA's share, cv vectors each
};

template <typename T>
__device__ __forceinline__ float first_elem_of(uint32_t smem_addr) {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(smem_addr));
    if (sizeof(T) == 4) return __uint_as_float(w);
    float f[VecT<T>::N];
    VecT<T>::unpack(make_uint4(w, 0u, 0u, 0u), f);
    return f[0];
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Barriers + loads: the CTA's whole share is requested up front as a few large bulk copies (one mbarrier each).
// (Measured: initialising the barriers and issuing the copies from the lanes of one warp BEFORE the CTA-wide barrier
// delayed the first data by 0.7 us against this order; the number of copies per share, 1 to 10, made no difference.)
template <int NS>
__device__ __forceinline__ Ctx setup(unsigned char* smem, const Geom& g, const char* s0, const char* s1, const char* s2) {
    Ctx c;
    c.CS = g.CS;
    c.rank = g.CS > 1 ? cluster_ctarank() : 0u;
    c.v0 = c.rank * g.nv_base + (c.rank < g.nv_rem ? c.rank : g.nv_rem);
    c.nv = g.nv_base + (c.rank < g.nv_rem ? 1u : 0u);
    c.cv = g.cv;
    c.nch = (c.nv + g.cv - 1) / g.cv;
    c.stream_stride = g.nch_max * kChunkBytes;
    c.data0 = smem_u32(smem);
    unsigned char* ctl = smem + (size_t)NS * c.stream_stride;
    c.full0 = smem_u32(ctl);
    c.xbar = c.full0 + kMaxChunks * 8;
    c.peer = reinterpret_cast<float*>(ctl + kMaxChunks * 8 + 16);
    c.wpart = c.peer + kMaxCluster * kRec;
    c.bcast = c.wpart + kWarps * kRec;
    if (threadIdx.x < c.nch) mbar_init(c.full0 + 8 * threadIdx.x, 1);
    if (threadIdx.x == kMaxChunks) mbar_init(c.xbar, c.CS);
    if (threadIdx.x <= kMaxChunks) fence_mbar_init();
    __syncthreads();
    // peers may only touch this CTA's barrier / record slots once they exist: split cluster barrier, the wait half sits
    // right before the exchange, ~2 us of loads later
    if (g.CS > 1) cluster_arrive();
    // programmatic dependent launch: the launch latency and the set-up above overlap the tail of the kernel before this
    // one in the stream; global memory is touched only from here on
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        const uint64_t pol = l2_policy_evict_first();
        const char* const src[3] = {s0, s1, s2};
        for (unsigned k = 0; k < c.nch; ++k) {
            const unsigned vecs = c.nv - k * c.cv < c.cv ? c.nv - k * c.cv : c.cv;
            const uint32_t bytes = vecs * 16u, bar = c.full0 + 8 * k;
            mbar_arrive_expect_tx(bar, bytes * NS);
#pragma unroll
            for (int s = 0; s < NS; ++s)
                tma_load_1d(c.data0 + s * c.stream_stride + k * c.cv * 16u, src[s] + ((size_t)c.v0 + (size_t)k * c.cv) * 16, bytes,
                            bar, pol);
        }
    }
    return c;
}

// CTA-wide sum of K values per thread -> every thread gets the totals (warp order: same bits everywhere)
template <int K>
__device__ __forceinline__ void cta_sum(const Ctx& c, float (&v)[K]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // wpart reuse
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) c.wpart[warp * kRec + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = 0.f;
        for (int w = 0; w < kWarps; ++w) t += c.wpart[w * kRec + k];
        v[k] = t;
    }
}

// push this CTA's record (K floats) into slot [rank] of every CTA of the cluster, then wait for all CS records
template <int K>
__device__ __forceinline__ void exchange(const Ctx& c, const float (&rec)[K]) {
    if (c.CS == 1) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) c.peer[k] = rec[k];
        }
        __syncthreads();
        return;
    }
    cluster_wait();  // second half of setup()'s split barrier: every peer's xbar and record slots exist
    if (threadIdx.x < c.CS) {
        const uint32_t dst = mapa(smem_u32(c.peer + c.rank * kRec), threadIdx.x);
#pragma unroll
        for (int k = 0; k < K; ++k) st_cluster_f32(dst + 4 * k, rec[k]);
        mbar_arrive_remote(mapa(c.xbar, threadIdx.x));
    }
    mbar_wait_cluster(c.xbar, 0u);
}

__device__ __forceinline__ unsigned share_vecs(const Geom& g, unsigned r) { return g.nv_base + (r < g.nv_rem ? 1u : 0u); }

// statistics of one stream of this CTA's share about K = its first element; returns (S1, S2) summed over the thread's vectors
template <typename T>
__device__ __forceinline__ void stats_chunk(const uint4& q, f32x2 K2, f32x2& s, f32x2& qq) {
    constexpr int VN = VecT<T>::N;
    f32x2 f[VN / 2];
    VecT<T>::unpack2(q, f);
#pragma unroll
    for (int k = 0; k < VN / 2; ++k) {
        const f32x2 d = f2_sub(f[k], K2);
        s = f2_add(s, d);
        qq = f2_fma(d, d, qq);
    }
}

// cluster-wide (mean, rstd) of one tensor's slab from the exchanged per-CTA (mean, M2) records at c.peer[r*kRec + off]
template <typename T>
__device__ __forceinline__ void fold_stats(const Ctx& c, const Geom& g, int off, float M, float eps, float& mean, float& rstd) {
    constexpr int VN = VecT<T>::N;
    const float ref = c.peer[off];
    float A = 0.f, B = 0.f;
    for (unsigned r = 0; r < c.CS; ++r) {
        const float nq = (float)(share_vecs(g, r) * VN), d = c.peer[r * kRec + off] - ref;
        A = fmaf(nq, d, A);
        B += fmaf(nq * d, d, c.peer[r * kRec + off + 1]);
    }
    const float invM = 1.f / M, m = A * invM;
    mean = ref + m;
    rstd = 1.f / sqrtf(fmaxf(B - A * m, 0.f) * invM + eps);  // biased variance, eps inside the sqrt
}

// =================================================================================================
// forward
// =================================================================================================
template <typename T, int EPI>
__global__ void __launch_bounds__(kThreads, 2) micn_fwd_res_kernel(const __grid_constant__ FwdParams p, const Geom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VN = VecT<T>::N;
    constexpr bool DUAL = EPI == MICN_EPI_NORM_ADD_LRELU;
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU || DUAL) ? 2 : 1;
    trace(g, TR_ENTRY);
    const unsigned tid = threadIdx.x;
    const unsigned slab = blockIdx.x / g.CS;
    const unsigned n = slab / (unsigned)p.C, ch = slab - n * (unsigned)p.C;
    const size_t dense = (size_t)slab * (size_t)p.M * sizeof(T);
    const char* xsrc = reinterpret_cast<const char*>(p.x) + ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T);
    const Ctx c = setup<NS>(smem, g, xsrc, NS == 2 ? reinterpret_cast<const char*>(DUAL ? p.x2 : p.res) + dense : nullptr, nullptr);
    trace(g, TR_SETUP);
    // parameter loads: their latency hides behind the data
    const int style = load_style(p.styles, n, p.num_styles, p.status);
    float gamma, beta, gamma2 = 1.f, beta2 = 0.f;
    load_affine_gc(p, style, ch, gamma, beta);
    if (DUAL && p.affine) {
        gamma2 = __ldg(p.gamma2[style] + ch);
        beta2 = __ldg(p.beta2[style] + ch);
    }
    const float slope = load_slope(p);

    // ---- pass 1: shifted sums of this thread's vectors (stream 0, and stream 1 when it is normalised too)
    mbar_wait(c.full0, 0u);
    trace(g, TR_FIRST);
    const float Ka = first_elem_of<T>(c.data0), Kb = DUAL ? first_elem_of<T>(c.data0 + c.stream_stride) : 0.f;
    const f32x2 Ka2 = f2_splat(Ka), Kb2 = f2_splat(Kb);
    f32x2 sa = f2_splat(0.f), qa = sa, sb = sa, qb = sa;
    for (unsigned k = 0; k < c.nch; ++k) {
        if (k) mbar_wait(c.full0 + 8 * k, 0u);
        const unsigned vend = (k + 1) * c.cv < c.nv ? (k + 1) * c.cv : c.nv;
        for (unsigned v = k * c.cv + tid; v < vend; v += kThreads) {
            stats_chunk<T>(lds128(c.data0 + v * 16), Ka2, sa, qa);
            if (DUAL) stats_chunk<T>(lds128(c.data0 + c.stream_stride + v * 16), Kb2, sb, qb);
        }
    }
    trace(g, TR_P1);
    float part[DUAL ? 4 : 2];
    part[0] = f2_hsum(sa);
    part[1] = f2_hsum(qa);
    if (DUAL) {
        part[2] = f2_hsum(sb);
        part[3] = f2_hsum(qb);
    }
    cta_sum(c, part);
    trace(g, TR_CTASUM);
    const float ncta = (float)(c.nv * VN);
    float rec[DUAL ? 4 : 2];
    {
        const float m = part[0] / ncta;
        rec[0] = Ka + m;
        rec[1] = fmaxf(part[1] - part[0] * m, 0.f);
        if (DUAL) {
            const float mb = part[2] / ncta;
            rec[2] = Kb + mb;
            rec[3] = fmaxf(part[3] - part[2] * mb, 0.f);
        }
    }
    exchange(c, rec);
    trace(g, TR_XCHG);
    float mean, rstd, mean2 = 0.f, rstd2 = 0.f;
    fold_stats<T>(c, g, 0, (float)p.M, p.eps, mean, rstd);
    if (DUAL) fold_stats<T>(c, g, 2, (float)p.M, p.eps, mean2, rstd2);
    if (tid == 0 && c.rank == 0) {
        if (p.save_mean) {
            p.save_mean[slab] = mean;
            p.save_rstd[slab] = rstd;
        }
        if (DUAL && p.save_mean2) {
            p.save_mean2[slab] = mean2;
            p.save_rstd2[slab] = rstd2;
        }
    }
    // fp32: (x - mean) * a + beta.  16-bit: x * a + (beta - mean * a): one FMA per element (as the flat path)
    const float a = rstd * gamma, a2 = rstd2 * gamma2;
    const f32x2 sub2 = f2_splat(sizeof(T) == 4 ? mean : 0.f), ca2 = f2_splat(a),
                cb2 = f2_splat(sizeof(T) == 4 ? beta : fmaf(-mean, a, beta));
    const f32x2 subB = f2_splat(sizeof(T) == 4 ? mean2 : 0.f), caB = f2_splat(a2),
                cbB = f2_splat(sizeof(T) == 4 ? beta2 : fmaf(-mean2, a2, beta2));

    // ---- pass 2: normalise + epilogue out of shared memory
    char* ydst = reinterpret_cast<char*>(p.y) + dense + (size_t)c.v0 * 16;
    {
        for (unsigned v = tid; v < c.nv; v += kThreads) {
            f32x2 f[VN / 2], r[VN / 2];
            VecT<T>::unpack2(lds128(c.data0 + v * 16), f);
            if (NS == 2) VecT<T>::unpack2(lds128(c.data0 + c.stream_stride + v * 16), r);
#pragma unroll
            for (int e = 0; e < VN / 2; ++e) {
                f32x2 o = sizeof(T) == 4 ? f2_fma(f2_sub(f[e], sub2), ca2, cb2) : f2_fma(f[e], ca2, cb2);
                if (DUAL) o = f2_add(o, sizeof(T) == 4 ? f2_fma(f2_sub(r[e], subB), caB, cbB) : f2_fma(r[e], caB, cbB));
                if (EPI == MICN_EPI_ADD_LRELU) o = f2_add(o, r[e]);
                if (EPI != MICN_EPI_NONE) {
                    float lo, hi;
                    f2_split(o, lo, hi);
                    lo = lo > 0.f ? lo : lo * slope;
                    hi = hi > 0.f ? hi : hi * slope;
                    o = f2_make(lo, hi);
                }
                f[e] = o;
            }
            stg_stream(ydst + (size_t)v * 16, VecT<T>::pack2v(f));
        }
    }
    trace(g, TR_P2);
}

// =================================================================================================
// backward:  g = dy * act'(.) ; S1 = sum g ; S2 = sum g*(x-mean) ;
//            dx = a*(g - S1/M - xhat*rstd*S2/M) ; dresidual = g ; dgamma/dbeta from rstd*S2 / S1
// DUAL: the same g feeds both norms: S2 and dx per input tensor, S1 shared
// =================================================================================================
template <typename T, int EPI, bool DS = false>
__global__ void __launch_bounds__(kThreads, 2) micn_bwd_res_kernel(const __grid_constant__ BwdParams p, const Geom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VN = VecT<T>::N;
    constexpr bool DUAL = EPI == MICN_EPI_NORM_ADD_LRELU;
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU || DUAL) ? 3 : 2;  // x, dy [, act_out | x2]
    trace(g, TR_ENTRY);
    const unsigned tid = threadIdx.x;
    const unsigned slab = blockIdx.x / g.CS;
    const unsigned n = slab / (unsigned)p.C, ch = slab - n * (unsigned)p.C;
    const size_t dense = (size_t)slab * (size_t)p.M * sizeof(T);
    const char* xsrc = reinterpret_cast<const char*>(p.x) + ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T);
    const Ctx c = setup<NS>(smem, g, xsrc, reinterpret_cast<const char*>(p.dy) + dense,
                            NS == 3 ? reinterpret_cast<const char*>(DUAL ? p.x2 : p.act_out) + dense : nullptr);
    trace(g, TR_SETUP);
    const int style = load_style(p.styles, n, p.num_styles, p.status);
    float gamma, beta, gamma2 = 1.f, beta2 = 0.f;
    load_affine_gc(p, style, ch, gamma, beta);
    if (DUAL && p.affine) {
        gamma2 = __ldg(p.gamma2[style] + ch);
        beta2 = __ldg(p.beta2[style] + ch);
    }
    const float mean = __ldg(p.save_mean + slab), rstd = __ldg(p.save_rstd + slab);
    const float meanB = DUAL ? __ldg(p.save_mean2 + slab) : 0.f, rstdB = DUAL ? __ldg(p.save_rstd2 + slab) : 0.f;
    const float slope = load_slope(p);
    const float a = rstd * gamma, aB = rstdB * gamma2;
    const float bq = sizeof(T) == 4 ? beta : fmaf(-mean, a, beta);
    const float bqB = sizeof(T) == 4 ? beta2 : fmaf(-meanB, aB, beta2);
    const f32x2 mean2 = f2_splat(mean), ca2 = f2_splat(a), bq2 = f2_splat(bq);
    const f32x2 meanB2 = f2_splat(meanB), caB2 = f2_splat(aB), bqB2 = f2_splat(bqB);
    // the reference masks on the ROUNDED output (in-place LeakyReLU on the 16-bit result): an fp16 value in (0, 2^-25]
    // rounds to +0 and counts as "not positive"; bf16 and fp32 share fp32's exponent range (as micn_flat.cuh)
    const float zero = (EPI != MICN_EPI_ADD_LRELU && EPI != MICN_EPI_NONE && sizeof(T) == 2 && !VecT<T>::kWideExponent)
                           ? 2.98023224e-8f : 0.f;
    const uint32_t sx = c.data0, sg = c.data0 + c.stream_stride, so = c.data0 + 2 * c.stream_stride;

    // g of one packed pair; `pre` out for the PReLU slope gradient
    auto masked = [&](f32x2 x, f32x2 gy, f32x2 o, f32x2& pre) -> f32x2 {
        if (EPI == MICN_EPI_NONE) return gy;
        if (EPI == MICN_EPI_ADD_LRELU) pre = o;
        else {
            pre = sizeof(T) == 4 ? f2_fma(f2_sub(x, mean2), ca2, bq2) : f2_fma(x, ca2, bq2);
            if (DUAL) pre = f2_add(pre, sizeof(T) == 4 ? f2_fma(f2_sub(o, meanB2), caB2, bqB2) : f2_fma(o, caB2, bqB2));
        }
        float p0, p1, g0, g1;
        f2_split(pre, p0, p1);
        f2_split(gy, g0, g1);
        return f2_make(p0 > zero ? g0 : g0 * slope, p1 > zero ? g1 : g1 * slope);
    };

    // ---- pass 1
    f32x2 s1 = f2_splat(0.f), s2 = s1, s2b = s1, ds2 = s1;
    for (unsigned k = 0; k < c.nch; ++k) {
        mbar_wait(c.full0 + 8 * k, 0u);
        if (k == 0) trace(g, TR_FIRST);
        const unsigned vend = (k + 1) * c.cv < c.nv ? (k + 1) * c.cv : c.nv;
        for (unsigned v = k * c.cv + tid; v < vend; v += kThreads) {
            f32x2 xf[VN / 2], gf[VN / 2], of[VN / 2];
            VecT<T>::unpack2(lds128(sx + v * 16), xf);
            VecT<T>::unpack2(lds128(sg + v * 16), gf);
            if (NS == 3) VecT<T>::unpack2(lds128(so + v * 16), of);
#pragma unroll
            for (int e = 0; e < VN / 2; ++e) {
                f32x2 pre = 0ull;
                const f32x2 gg = masked(xf[e], gf[e], NS == 3 ? of[e] : 0ull, pre);
                if (DS) {
                    float p0, p1;
                    f2_split(pre, p0, p1);
                    ds2 = f2_fma(gf[e], f2_make(p0 > 0.f ? 0.f : p0, p1 > 0.f ? 0.f : p1), ds2);
                }
                s1 = f2_add(s1, gg);
                s2 = f2_fma(gg, f2_sub(xf[e], mean2), s2);
                if (DUAL) s2b = f2_fma(gg, f2_sub(of[e], meanB2), s2b);
            }
        }
    }
    trace(g, TR_P1);
    float part[4];
    part[0] = f2_hsum(s1);
    part[1] = f2_hsum(s2);
    part[2] = DUAL ? f2_hsum(s2b) : 0.f;
    part[3] = DS ? f2_hsum(ds2) : 0.f;
    cta_sum(c, part);
    trace(g, TR_CTASUM);
    exchange(c, part);
    trace(g, TR_XCHG);
    float S1 = 0.f, S2 = 0.f, S2B = 0.f, DSL = 0.f;
    for (unsigned r = 0; r < c.CS; ++r) {
        S1 += c.peer[r * kRec + 0];
        S2 += c.peer[r * kRec + 1];
        if (DUAL) S2B += c.peer[r * kRec + 2];
        if (DS) DSL += c.peer[r * kRec + 3];
    }
    const float S2r = S2 * rstd, S2Br = S2B * rstdB;  // sum g * xhat
    if (tid == 0 && c.rank == 0) {
        if (DS) p.dslope[slab] = DSL;
        if (p.dgamma) {
            const unsigned C = (unsigned)p.C;
            if (p.N == 1) {  // one sample: this slab's sums ARE the gradients of its style's row
                for (int s = 0; s < p.num_styles; ++s) {
                    p.dbeta[(size_t)s * C + ch] = s == style ? S1 : 0.f;
                    p.dgamma[(size_t)s * C + ch] = s == style ? S2r : 0.f;
                    if (DUAL && p.dgamma2) {
                        p.dbeta2[(size_t)s * C + ch] = s == style ? S1 : 0.f;
                        p.dgamma2[(size_t)s * C + ch] = s == style ? S2Br : 0.f;
                    }
                }
            } else {  // the last sample of a channel to deliver its sums folds the channel, samples in order (micn_small.cuh)
                p.ws_sum_dy[slab] = S1;
                p.ws_sum_dyxh[slab] = S2r;
                if (DUAL && p.ws_sum_dyxh2) p.ws_sum_dyxh2[slab] = S2Br;
                __threadfence();
                if (atomicAdd(p.ws_chan_cnt + ch, 1u) == (unsigned)p.N - 1u) {
                    p.ws_chan_cnt[ch] = 0u;
                    __threadfence();
                    small_fold_channel(p, ch);
                }
            }
        }
    }
    // dx = a*g - a*S1/M - a*rstd*(S2r/M)*(x - mean)
    const float invM = 1.f / (float)p.M;
    const float B1 = -a * S2r * invM * rstd, B0c = -a * S1 * invM;
    const f32x2 A2 = ca2, B12 = f2_splat(B1), B02 = f2_splat(sizeof(T) == 4 ? B0c : fmaf(-B1, mean, B0c));
    const float B1b = -aB * S2Br * invM * rstdB, B0cb = -aB * S1 * invM;
    const f32x2 AB2 = caB2, B1B2 = f2_splat(B1b), B0B2 = f2_splat(sizeof(T) == 4 ? B0cb : fmaf(-B1b, meanB, B0cb));

    // ---- pass 2
    char* dxdst = reinterpret_cast<char*>(p.dx) + dense + (size_t)c.v0 * 16;
    char* d2dst = NS == 3 ? reinterpret_cast<char*>(DUAL ? p.dx2 : p.dres) + dense + (size_t)c.v0 * 16 : nullptr;
    {
        for (unsigned v = tid; v < c.nv; v += kThreads) {
            f32x2 xf[VN / 2], gf[VN / 2], of[VN / 2];
            VecT<T>::unpack2(lds128(sx + v * 16), xf);
            VecT<T>::unpack2(lds128(sg + v * 16), gf);
            if (NS == 3) VecT<T>::unpack2(lds128(so + v * 16), of);
#pragma unroll
            for (int e = 0; e < VN / 2; ++e) {
                f32x2 pre = 0ull;
                const f32x2 gg = masked(xf[e], gf[e], NS == 3 ? of[e] : 0ull, pre);
                xf[e] = f2_fma(A2, gg, f2_fma(B12, sizeof(T) == 4 ? f2_sub(xf[e], mean2) : xf[e], B02));
                if (DUAL) of[e] = f2_fma(AB2, gg, f2_fma(B1B2, sizeof(T) == 4 ? f2_sub(of[e], meanB2) : of[e], B0B2));
                gf[e] = gg;
            }
            stg_stream(dxdst + (size_t)v * 16, VecT<T>::pack2v(xf));
            if (NS == 3) stg_stream(d2dst + (size_t)v * 16, VecT<T>::pack2v(DUAL ? of : gf));
        }
    }
    trace(g, TR_P2);
}

}  // namespace res
}  // namespace micn
