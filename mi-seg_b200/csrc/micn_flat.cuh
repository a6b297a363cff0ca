// micn_flat.cuh - flat-partition instance_cond forward / backward (sm_100a): the default large-slab path.
//
// Why: the copy roofline of a B200 is set by the L2<->SM fabric as much as by HBM, so a voxel must
// cross it exactly once per tensor, and every one of the 148 SMs has to carry an equal share.  Binding
// whole slabs to clusters cannot do both (48 slabs of 1.7 MB on 148 SMs: either the slab does not fit
// the cluster's shared memory and is re-fetched from L2, or most SMs idle in the last wave).
//
// How: every (n, c) slab is cut into P chunks ("pieces") of <= PV 16-byte vectors; piece g = slab*P + k
// belongs to CTA g % G at its local round g / G (G = one persistent CTA per SM, launched cooperatively so
// all are co-resident).  Inside a CTA three roles run decoupled over a ring of K shared-memory slots:
//
//   producer warp (1 lane)  1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of the next piece
//                           into a free slot, L2 evict_first (each voxel is read once); runs as far ahead
//                           as the ring allows, which is what keeps >= 100 KB per SM in flight.
//   16 consumer warps       in NG groups; group q owns the CTA's pieces j = q mod NG.  P1(j): statistics of
//                           the piece out of shared memory (fp32 shifted sums, warp shuffle) -> one partial
//                           per warp.  P2(j): normalise / epilogue / backward formula out of the SAME
//                           shared-memory copy, 128-bit streaming stores, then the slot returns to the
//                           producer.  Per group P2 trails P1 by Lg of its pieces: that lag (Lg*NG pieces,
//                           several microseconds) is what hides the cross-CTA exchange.
//   2 publish warps         slot s -> warp s mod 2: merge the piece's warp partials (Chan) and write the piece
//                           record to the workspace as soon as P1 is done; never wait on another CTA.
//   8 gather warps          slot s -> warp s mod 8: poll the P records of the piece's slab (a batch of loads in
//                           flight at once, first poll delayed so it usually hits), merge them in a fixed
//                           order (bit-identical in every CTA, no atomics) and hands P2(j) its per-slab
//                           coefficients; backward: also emits the per-slab sums and, for the last sample
//                           of a channel, d(gamma)/d(beta) per style in a fixed order.
//
// Cross-CTA exchange is by 16-byte self-validating records {a, tag, b, tag} (tag = per-launch epoch): no
// counters to reset, an aborted launch cannot poison the next one.  Under load one exchange (store, L2,
// poll) costs 3-5 us, hence small pieces, a deep ring and a long lag rather than few large pieces.
// Deadlock freedom: a slab of P pieces spans R <= ceil((P-1)/G)+1 rounds; the planner keeps
// R - 1 <= Lg*NG and K >= (Lg+1)*NG + 1, so every P1 a record depends on can run before anyone blocks
// in a P2.  Every wait is bounded and traps instead of hanging.
//
// HBM / L2 traffic: forward reads x once, writes y once (2*E*s); backward reads x, dy [, act_out]
// once and writes dx [, dresidual] once (3*E*s / 5*E*s) - the algorithmic minimum.
//
// Reference semantics: networks/norms/conditional_instance_norm.py:59-60 (+ ATen instance_norm:
// biased variance, eps inside the sqrt), epilogues networks/blocks/dynunet_block.py:107-125.
#pragma once

#include "micn_common.cuh"

namespace micn {

constexpr int kFlatConsumerWarps = 16;
constexpr int kFlatConsumerThreads = kFlatConsumerWarps * 32;  // 512
constexpr int kFlatProducerWarp = kFlatConsumerWarps;
constexpr int kFlatPublishWarp0 = kFlatConsumerWarps + 1;
constexpr int kFlatPublishWarps = 2;  // slot s -> publish warp s mod 2; a publish never waits on another CTA
constexpr int kFlatGatherWarp0 = kFlatPublishWarp0 + kFlatPublishWarps;
constexpr int kFlatGatherWarps = 8;  // slot s -> gather warp s mod 8; a gather is a multi-microsecond latency chain
constexpr int kFlatThreads = (kFlatConsumerWarps + 1 + kFlatPublishWarps + kFlatGatherWarps) * 32;  // 864
constexpr int kFlatMaxSlots = 24;
constexpr int kFlatMaxPieces = 1024;  // pieces per slab (workspace sizing); the planner enforces the round bound
constexpr int kFlatMinPieceVecs = 128;
constexpr uint32_t kFlatTmaChunk = 32768;

struct FlatGeom {
    unsigned long long V;  // 16-byte vectors per slab
    unsigned T;            // total pieces = num_slabs * P
    unsigned P;            // pieces per slab
    unsigned PV;           // vectors per piece (the last piece of a slab may be shorter)
    unsigned K;            // ring slots, a multiple of NG: a slot is always used by the same consumer group and
                           // the same gather warp, so every mbarrier's phases are observed in order
    unsigned NG;           // consumer groups (1, 2, 4, 8 or 16); group q owns pieces j = q mod NG
    unsigned Lg;           // per group, P2 trails P1 by Lg of the group's pieces
    unsigned slot_vecs;    // vectors reserved per stream per slot (>= PV, multiple of 8)
    unsigned epoch;        // per-launch tag of the workspace records (never 0)
    unsigned poll_delay_ns, poll_backoff_ns;
    uint4* ws_piece;       // [T] piece records
    uint4* ws_slab;        // [num_slabs] per-slab records (backward parameter gradients)
    long long* trace;      // bring-up only: [grid][kFlatTraceSteps][16] %globaltimer stamps (ns) per piece, or null
};

constexpr int kFlatTraceSteps = 64;
enum { TR_LOAD = 0, TR_P1_BEGIN, TR_P1_END, TR_PUB_BEGIN, TR_PUB_END, TR_GA_BEGIN, TR_GA_POLLED, TR_GA_END,
       TR_P2_WAIT, TR_P2_BEGIN, TR_P2_END, TR_LOAD_WAIT };
__device__ __forceinline__ void flat_trace(const FlatGeom& g, unsigned j, int ev) {
    if (g.trace && j < kFlatTraceSteps) g.trace[((size_t)blockIdx.x * kFlatTraceSteps + j) * 16 + ev] = (long long)globaltimer_ns();
}

// per-slot control block: 4 mbarriers, 16 warp partials, P2 coefficients, slab constants
__host__ __device__ constexpr int flat_ctl_bytes() {
    return kFlatMaxSlots * (4 * 8 + kFlatConsumerWarps * 16 + 32 + 16);
}

struct FlatCtx {
    uint32_t data0, full0, empty0, p1d0, coef0;  // shared::cta addresses
    float* warp_part;                            // [K][16][4]
    float* coefv;                                // [K][8]
    float* prec;                                 // [K][4]
    uint32_t stream_bytes, slot_bytes;
};

template <int NS>
__device__ __forceinline__ FlatCtx flat_setup(unsigned char* smem, const FlatGeom& g) {
    FlatCtx c;
    c.stream_bytes = g.slot_vecs * 16u;
    c.slot_bytes = c.stream_bytes * NS;
    c.data0 = smem_u32(smem);
    unsigned char* ctl = smem + (size_t)g.K * c.slot_bytes;
    c.full0 = smem_u32(ctl);
    c.empty0 = c.full0 + kFlatMaxSlots * 8;
    c.p1d0 = c.empty0 + kFlatMaxSlots * 8;
    c.coef0 = c.p1d0 + kFlatMaxSlots * 8;
    c.warp_part = reinterpret_cast<float*>(ctl + kFlatMaxSlots * 32);
    c.coefv = c.warp_part + kFlatMaxSlots * kFlatConsumerWarps * 4;
    c.prec = c.coefv + kFlatMaxSlots * 8;
    if (threadIdx.x == 0) {
        const unsigned wg = kFlatConsumerWarps / g.NG;  // warps per consumer group
        for (unsigned i = 0; i < g.K; ++i) {
            mbar_init(c.full0 + 8 * i, 1);
            mbar_init(c.empty0 + 8 * i, wg);
            mbar_init(c.p1d0 + 8 * i, wg);
            mbar_init(c.coef0 + 8 * i, 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    return c;
}

struct PieceId {
    unsigned slab, k, pv;  // slab index, piece index inside the slab, vectors in this piece
};
__device__ __forceinline__ unsigned piece_vecs(const FlatGeom& g, unsigned k) {
    const unsigned long long left = g.V - (unsigned long long)k * g.PV;
    return left < g.PV ? (unsigned)left : g.PV;
}
__device__ __forceinline__ PieceId piece_of(const FlatGeom& g, unsigned gidx) {
    PieceId p;
    p.slab = gidx / g.P;
    p.k = gidx - p.slab * g.P;
    p.pv = piece_vecs(g, p.k);
    return p;
}

// ring position of the CTA's j-th piece
struct Ring {
    unsigned i, ph;
};
__device__ __forceinline__ Ring ring_of(const FlatGeom& g, unsigned j) {
    const unsigned u = j / g.K;
    return Ring{j - u * g.K, u & 1u};
}

// ---- self-validating workspace records {a, tag, b, tag}: each 8-byte half carries its own tag, so a
//      torn 16-byte access can never be mistaken for a complete record
__device__ __forceinline__ void ll_store(uint4* p, float a, float b, unsigned tag) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag),
                 "r"(__float_as_uint(b)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4* p, unsigned tag, float& a, float& b) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    a = __uint_as_float(v.x);
    b = __uint_as_float(v.z);
    return v.y == tag && v.w == tag;
}
__device__ __forceinline__ void ll_wait(const uint4* p, unsigned tag, unsigned backoff_ns, float& a, float& b) {
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    do {
        __nanosleep(backoff_ns);
        if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
    } while (!ll_try(p, tag, a, b));
}

// Poll the `count` records at `recs` (lane q handles records q, q+32, ...) and fold them with `fold(q, a, b)`
// in ascending q per lane.  Up to four loads per lane are issued before the first is examined, so a
// slab's whole record set costs one L2 round trip in the common (already published) case.
template <typename Fold>
__device__ __forceinline__ void ll_gather(const uint4* recs, unsigned count, unsigned tag, unsigned backoff_ns, int lane,
                                          Fold fold) {
    for (unsigned q0 = 0; q0 < count; q0 += 128) {
        float a[4], b[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned q = q0 + lane + 32 * i;
            ok[i] = q < count ? ll_try(recs + q, tag, a[i], b[i]) : true;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned q = q0 + lane + 32 * i;
            if (q < count) {
                if (!ok[i]) ll_wait(recs + q, tag, backoff_ns, a[i], b[i]);
                fold(q, a[i], b[i]);
            }
        }
    }
}

// vectors of a piece (pv vectors, strided over the tg threads of its consumer group) that land in warp w of the group
__device__ __forceinline__ unsigned warp_vecs(unsigned pv, unsigned tg, unsigned w) {
    const unsigned full = pv / tg, rem = pv - full * tg;
    int r = (int)rem - 32 * (int)w;
    r = r < 0 ? 0 : (r > 32 ? 32 : r);
    return 32u * full + (unsigned)r;
}

template <typename T>
__device__ __forceinline__ float first_elem(uint32_t smem_addr) {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(smem_addr));
    if (sizeof(T) == 4) return __uint_as_float(w);
    float f[VecT<T>::N];
    VecT<T>::unpack(make_uint4(w, 0u, 0u, 0u), f);
    return f[0];
}

// producer: one bulk copy per <= 32 KB chunk of each stream of the piece, all on the slot's barrier
__device__ __forceinline__ void flat_issue(uint32_t dst, const char* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    for (uint32_t off = 0; off < bytes; off += kFlatTmaChunk) {
        const uint32_t n = bytes - off < kFlatTmaChunk ? bytes - off : kFlatTmaChunk;
        tma_load_1d(dst + off, src + off, n, bar, pol);
    }
}

// consumer-group geometry of the calling warp
struct Group {
    unsigned wg, tg, q, w, t;  // warps / threads per group, group id, warp and thread index inside the group
};
__device__ __forceinline__ Group group_of(const FlatGeom& g) {
    Group gr;
    gr.wg = kFlatConsumerWarps / g.NG;
    gr.tg = gr.wg * 32;
    const unsigned warp = threadIdx.x >> 5;
    gr.q = warp / gr.wg;
    gr.w = warp - gr.q * gr.wg;
    gr.t = threadIdx.x - gr.q * gr.tg;
    return gr;
}

// =================================================================================================
// forward
// =================================================================================================
template <typename T, int EPI>
__global__ void __launch_bounds__(kFlatThreads, 1) micn_fwd_flat_kernel(const FwdParams p, const FlatGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VN = VecT<T>::N;
    const FlatCtx c = flat_setup<1>(smem, g);
    const unsigned cta = blockIdx.x, G = gridDim.x;
    const unsigned nj = cta < g.T ? (g.T - cta + G - 1) / G : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned C = (unsigned)p.C;

    if (warp == kFlatProducerWarp) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            for (unsigned j = 0; j < nj; ++j) {
                const Ring r = ring_of(g, j);
                const PieceId pc = piece_of(g, j * G + cta);
                const unsigned n = pc.slab / C, ch = pc.slab - n * C;
                const char* src = reinterpret_cast<const char*>(p.x) +
                                  ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) +
                                  (size_t)pc.k * g.PV * 16;
                const uint32_t bytes = pc.pv * 16u, bar = c.full0 + 8 * r.i;
                flat_trace(g, j, TR_LOAD_WAIT);
                if (j >= g.K) mbar_wait(c.empty0 + 8 * r.i, r.ph ^ 1u);
                flat_trace(g, j, TR_LOAD);
                flat_issue(c.data0 + r.i * c.slot_bytes, src, bytes, bar, pol);
                mbar_arrive_expect_tx(bar, bytes);
            }
        }
    } else if (warp > kFlatProducerWarp && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish: warp partials -> piece record
        const unsigned wg = kFlatConsumerWarps / g.NG, tg = wg * 32;
        for (unsigned j = 0; j < nj; ++j) {
            const Ring r = ring_of(g, j);
            if (r.i % kFlatPublishWarps != (unsigned)(warp - kFlatPublishWarp0)) continue;
            const unsigned gidx = j * G + cta;
            const unsigned pv = piece_vecs(g, gidx % g.P);
            const float nw = (float)(warp_vecs(pv, tg, lane & 15) * VN);
            mbar_wait(c.p1d0 + 8 * r.i, r.ph);
            if (lane == 0) flat_trace(g, j, TR_PUB_BEGIN);
            Stat st{0.f, 0.f, 0.f};
            if ((unsigned)lane < wg) {
                const float4 w = *reinterpret_cast<const float4*>(c.warp_part + (r.i * kFlatConsumerWarps + lane) * 4);
                st = stat_from_shifted(w.z, w.x, w.y, nw);
            }
            st = stat_warp_reduce(st);
            if (lane == 0) ll_store(g.ws_piece + gidx, st.mean, st.m2, g.epoch);
            if (lane == 0) flat_trace(g, j, TR_PUB_END);
        }
    } else if (warp >= kFlatGatherWarp0) {
        // ------------------------------------------------------------------ gather: slab records -> coefficients for P2
        for (unsigned j = 0; j < nj; ++j) {
            // a slot always goes to the same gather warp, so its barrier phases are observed in order
            const Ring r = ring_of(g, j);
            if (r.i % kFlatGatherWarps != (unsigned)(warp - kFlatGatherWarp0)) continue;
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = pc.slab / C, ch = pc.slab - n * C;
            // parameter loads first: their latency hides behind everything below
            const int style = load_style(p.styles, n, p.num_styles, p.status);
            float gamma, beta;
            load_affine(p, style, ch, gamma, beta);
            // no polling before this CTA's own piece is through P1: the other CTAs are at the same point
            mbar_wait(c.p1d0 + 8 * r.i, r.ph);
            __nanosleep(g.poll_delay_ns);  // ... and let their record stores land
            if (lane == 0) flat_trace(g, j, TR_GA_BEGIN);
            Stat acc{0.f, 0.f, 0.f};
            ll_gather(g.ws_piece + (size_t)pc.slab * g.P, g.P, g.epoch, g.poll_backoff_ns, lane,
                      [&](unsigned q, float a, float b) {
                          acc = stat_merge(acc, Stat{(float)(piece_vecs(g, q) * VN), a, b});
                      });
            if (lane == 0) flat_trace(g, j, TR_GA_POLLED);
            acc = stat_warp_reduce(acc);
            if (lane == 0) {
                const float mean = acc.mean;
                const float rstd = 1.f / sqrtf(acc.m2 / (float)p.M + p.eps);  // biased variance, eps inside the sqrt
                const float a = rstd * gamma;
                // fp32: (x - mean) * a + beta.  16-bit: x * a + (beta - mean * a): one FMA per element.
                // The slot's previous coefficients were consumed before this piece was even loaded.
                *reinterpret_cast<float4*>(c.coefv + r.i * 8) =
                    make_float4(sizeof(T) == 4 ? mean : 0.f, a, sizeof(T) == 4 ? beta : fmaf(-mean, a, beta), 0.f);
                if (pc.k == 0 && p.save_mean) {
                    p.save_mean[pc.slab] = mean;
                    p.save_rstd[pc.slab] = rstd;
                }
                mbar_arrive(c.coef0 + 8 * r.i);
                flat_trace(g, j, TR_GA_END);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ consumers (group gr.q owns pieces j = q mod NG)
        const Group gr = group_of(g);
        const unsigned ni = nj > gr.q ? (nj - gr.q + g.NG - 1) / g.NG : 0;
        for (unsigned i = 0; i < ni + g.Lg; ++i) {
            if (i < ni) {  // P1: statistics
                const unsigned j = gr.q + g.NG * i;
                const Ring r = ring_of(g, j);
                const unsigned pv = piece_vecs(g, (j * G + cta) % g.P);
                mbar_wait(c.full0 + 8 * r.i, r.ph);
                if (gr.t == 0) flat_trace(g, j, TR_P1_BEGIN);
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f, Kw = 0.f;
                if (gr.w * 32u < pv) {
                    Kw = first_elem<T>(base + gr.w * 512);  // shift = the warp's first element of the piece
#pragma unroll 4
                    for (unsigned v = gr.t; v < pv; v += gr.tg) {
                        float f[VN];
                        VecT<T>::unpack(lds128(base + v * 16), f);
#pragma unroll
                        for (int e = 0; e < VN; e += 2) {
                            const float d0 = f[e] - Kw, d1 = f[e + 1] - Kw;
                            sa += d0;
                            sb += d1;
                            qa = fmaf(d0, d0, qa);
                            qb = fmaf(d1, d1, qb);
                        }
                    }
                }
                const float s1 = warp_sum(sa + sb), s2 = warp_sum(qa + qb);
                if (lane == 0) {
                    *reinterpret_cast<float4*>(c.warp_part + (r.i * kFlatConsumerWarps + gr.w) * 4) =
                        make_float4(s1, s2, Kw, 0.f);
                    mbar_arrive(c.p1d0 + 8 * r.i);
                }
                if (gr.t == 0) flat_trace(g, j, TR_P1_END);
            }
            if (i >= g.Lg) {  // P2: normalise + epilogue out of the same shared-memory copy
                const unsigned j = gr.q + g.NG * (i - g.Lg);
                const Ring r = ring_of(g, j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* ydst = reinterpret_cast<char*>(p.y) + goff;
                const char* rsrc = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<const char*>(p.res) + goff : nullptr;
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                uint4 rv0 = make_uint4(0u, 0u, 0u, 0u);
                if (EPI == MICN_EPI_ADD_LRELU && gr.t < pc.pv) rv0 = ldg_stream(rsrc + (size_t)gr.t * 16);
                if (gr.t == 0) flat_trace(g, j, TR_P2_WAIT);
                mbar_wait(c.coef0 + 8 * r.i, r.ph);
                if (gr.t == 0) flat_trace(g, j, TR_P2_BEGIN);
                const float4 cf = *reinterpret_cast<const float4*>(c.coefv + r.i * 8);
                const float sub = cf.x, a = cf.y, b = cf.z;
#pragma unroll 2
                for (unsigned v = gr.t; v < pc.pv; v += gr.tg) {
                    uint4 rv = rv0;
                    if (EPI == MICN_EPI_ADD_LRELU) {  // residual straight from HBM, next one in flight
                        const unsigned vn = v + gr.tg;
                        if (vn < pc.pv) rv0 = ldg_stream(rsrc + (size_t)vn * 16);
                    }
                    float f[VN], rr[VN];
                    VecT<T>::unpack(lds128(base + v * 16), f);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(rv, rr);
#pragma unroll
                    for (int e = 0; e < VN; ++e) {
                        float o = sizeof(T) == 4 ? fmaf(f[e] - sub, a, b) : fmaf(f[e], a, b);
                        if (EPI == MICN_EPI_ADD_LRELU) o += rr[e];
                        if (EPI != MICN_EPI_NONE) o = o > 0.f ? o : o * p.slope;
                        f[e] = o;
                    }
                    stg_stream(ydst + (size_t)v * 16, VecT<T>::pack(f));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
                if (gr.t == 0) flat_trace(g, j, TR_P2_END);
            }
        }
    }
}

// =================================================================================================
// backward:  g = dy * act'(.) ; S1 = sum g ; S2 = sum g*(x-mean) ;
//            dx = a*(g - S1/M - xhat*rstd*S2/M) ; dresidual = g ; dgamma/dbeta from rstd*S2 / S1
// =================================================================================================
template <typename T, int EPI>
__device__ __forceinline__ float bwd_masked(float x, float gy, float o, float mean, float a, float bq, float slope) {
    // the LeakyReLU mask is the sign of the SAME expression the forward evaluated
    if (EPI == MICN_EPI_LRELU) {
        const float pre = sizeof(T) == 4 ? fmaf(x - mean, a, bq) : fmaf(x, a, bq);
        return pre > 0.f ? gy : gy * slope;
    }
    if (EPI == MICN_EPI_ADD_LRELU) return o > 0.f ? gy : gy * slope;
    return gy;
}

template <typename T, int EPI>
__global__ void __launch_bounds__(kFlatThreads, 1) micn_bwd_flat_kernel(const BwdParams p, const FlatGeom g) {
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU) ? 3 : 2;  // x, dy [, act_out]
    constexpr int VN = VecT<T>::N;
    extern __shared__ __align__(128) unsigned char smem[];
    const FlatCtx c = flat_setup<NS>(smem, g);
    const unsigned cta = blockIdx.x, G = gridDim.x;
    const unsigned nj = cta < g.T ? (g.T - cta + G - 1) / G : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned C = (unsigned)p.C;

    if (warp == kFlatProducerWarp) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            for (unsigned j = 0; j < nj; ++j) {
                const Ring r = ring_of(g, j);
                const PieceId pc = piece_of(g, j * G + cta);
                const unsigned n = pc.slab / C, ch = pc.slab - n * C;
                // per-slab constants for P1 / gather: loads issued before the slot wait
                const int style = load_style(p.styles, n, p.num_styles, p.status);
                float gamma, beta;
                load_affine(p, style, ch, gamma, beta);
                const float mean = __ldg(p.save_mean + pc.slab), rstd = __ldg(p.save_rstd + pc.slab);
                const size_t poff = (size_t)pc.k * g.PV * 16;
                const size_t doff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + poff;
                const char* xsrc = reinterpret_cast<const char*>(p.x) +
                                   ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) + poff;
                const uint32_t bytes = pc.pv * 16u, bar = c.full0 + 8 * r.i;
                const uint32_t dst = c.data0 + r.i * c.slot_bytes;
                if (j >= g.K) mbar_wait(c.empty0 + 8 * r.i, r.ph ^ 1u);
                flat_issue(dst, xsrc, bytes, bar, pol);
                flat_issue(dst + c.stream_bytes, reinterpret_cast<const char*>(p.dy) + doff, bytes, bar, pol);
                if (NS == 3) flat_issue(dst + 2 * c.stream_bytes, reinterpret_cast<const char*>(p.act_out) + doff, bytes, bar, pol);
                // visible to the consumers (and, through their p1done arrival, to the gather warp) via the barrier
                *reinterpret_cast<float4*>(c.prec + r.i * 4) = make_float4(mean, rstd, gamma, beta);
                mbar_arrive_expect_tx(bar, bytes * NS);
            }
        }
    } else if (warp > kFlatProducerWarp && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish
        const unsigned wg = kFlatConsumerWarps / g.NG;
        for (unsigned j = 0; j < nj; ++j) {
            const Ring r = ring_of(g, j);
            if (r.i % kFlatPublishWarps != (unsigned)(warp - kFlatPublishWarp0)) continue;
            mbar_wait(c.p1d0 + 8 * r.i, r.ph);
            float s1 = 0.f, s2 = 0.f;
            if ((unsigned)lane < wg) {
                const float2 w = *reinterpret_cast<const float2*>(c.warp_part + (r.i * kFlatConsumerWarps + lane) * 4);
                s1 = w.x;
                s2 = w.y;
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) ll_store(g.ws_piece + (j * G + cta), s1, s2, g.epoch);
        }
    } else if (warp >= kFlatGatherWarp0) {
        // ------------------------------------------------------------------ gather
        const float invM = 1.f / (float)p.M;
        for (unsigned j = 0; j < nj; ++j) {
            // a slot always goes to the same gather warp, so its barrier phases are observed in order
            const Ring r = ring_of(g, j);
            if (r.i % kFlatGatherWarps != (unsigned)(warp - kFlatGatherWarp0)) continue;
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = pc.slab / C, ch = pc.slab - n * C;
            // no polling before this CTA's own piece is through P1 (the others are at the same point); the wait
            // also makes the producer's constants for the slot visible (full -> consumers -> p1done)
            mbar_wait(c.p1d0 + 8 * r.i, r.ph);
            const volatile float* pr = c.prec + r.i * 4;
            const float mean = pr[0], rstd = pr[1], gamma = pr[2], beta = pr[3];
            __nanosleep(g.poll_delay_ns);  // let the record stores land
            float S1 = 0.f, S2 = 0.f;
            ll_gather(g.ws_piece + (size_t)pc.slab * g.P, g.P, g.epoch, g.poll_backoff_ns, lane,
                      [&](unsigned, float a, float b) {
                          S1 += a;
                          S2 += b;
                      });
            S1 = warp_sum(S1);
            S2 = warp_sum(S2);
            const float a = rstd * gamma;
            const float S2r = S2 * rstd;  // sum g * xhat
            if (lane == 0) {
                // dx = a*g - a*S1/M - a*rstd*(S2r/M)*(x - mean)
                const float B1 = -a * S2r * invM * rstd, B0c = -a * S1 * invM;
                float* cf = c.coefv + r.i * 8;
                *reinterpret_cast<float4*>(cf) = make_float4(a, B1, sizeof(T) == 4 ? B0c : fmaf(-B1, mean, B0c), mean);
                cf[4] = sizeof(T) == 4 ? beta : fmaf(-mean, a, beta);
                mbar_arrive(c.coef0 + 8 * r.i);
            }
            if (pc.k == 0 && p.dgamma) {
                if (p.N == 1) {
                    // one sample: this slab's sums ARE the gradients of its style's row
                    const int style = load_style(p.styles, 0, p.num_styles, nullptr);
                    for (int s = lane; s < p.num_styles; s += 32) {
                        p.dbeta[(size_t)s * C + ch] = s == style ? S1 : 0.f;
                        p.dgamma[(size_t)s * C + ch] = s == style ? S2r : 0.f;
                    }
                } else {
                    if (lane == 0) ll_store(g.ws_slab + pc.slab, S1, S2r, g.epoch);
                    if (n == (unsigned)p.N - 1) {
                        // last sample of this channel: fold every sample's record per style, fixed order
                        for (int s = 0; s < p.num_styles; ++s) {
                            float ab = 0.f, ag = 0.f;
                            for (unsigned nn = lane; nn < (unsigned)p.N; nn += 32) {
                                float ra, rb;
                                if (!ll_try(g.ws_slab + (size_t)nn * C + ch, g.epoch, ra, rb))
                                    ll_wait(g.ws_slab + (size_t)nn * C + ch, g.epoch, g.poll_backoff_ns, ra, rb);
                                if (load_style(p.styles, nn, p.num_styles, nullptr) == s) {
                                    ab += ra;
                                    ag += rb;
                                }
                            }
                            ab = warp_sum(ab);
                            ag = warp_sum(ag);
                            if (lane == 0) {
                                p.dbeta[(size_t)s * C + ch] = ab;
                                p.dgamma[(size_t)s * C + ch] = ag;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const Group gr = group_of(g);
        const unsigned ni = nj > gr.q ? (nj - gr.q + g.NG - 1) / g.NG : 0;
        const uint32_t sb = c.stream_bytes;
        for (unsigned i = 0; i < ni + g.Lg; ++i) {
            if (i < ni) {  // P1
                const unsigned j = gr.q + g.NG * i;
                const Ring r = ring_of(g, j);
                const unsigned pv = piece_vecs(g, (j * G + cta) % g.P);
                mbar_wait(c.full0 + 8 * r.i, r.ph);
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                const float4 pr = *reinterpret_cast<const float4*>(c.prec + r.i * 4);
                const float mean = pr.x, a = pr.y * pr.z;
                const float bq = sizeof(T) == 4 ? pr.w : fmaf(-mean, a, pr.w);
                float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 2
                for (unsigned v = gr.t; v < pv; v += gr.tg) {
                    float xf[VN], gf[VN], of[VN];
                    VecT<T>::unpack(lds128(base + v * 16), xf);
                    VecT<T>::unpack(lds128(base + sb + v * 16), gf);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(lds128(base + 2 * sb + v * 16), of);
#pragma unroll
                    for (int e = 0; e < VN; e += 2) {
                        const float g0 = bwd_masked<T, EPI>(xf[e], gf[e], EPI == MICN_EPI_ADD_LRELU ? of[e] : 0.f, mean, a, bq, p.slope);
                        const float g1 = bwd_masked<T, EPI>(xf[e + 1], gf[e + 1], EPI == MICN_EPI_ADD_LRELU ? of[e + 1] : 0.f, mean, a, bq, p.slope);
                        s1a += g0;
                        s1b += g1;
                        s2a = fmaf(g0, xf[e] - mean, s2a);
                        s2b = fmaf(g1, xf[e + 1] - mean, s2b);
                    }
                }
                const float s1 = warp_sum(s1a + s1b), s2 = warp_sum(s2a + s2b);
                if (lane == 0) {
                    *reinterpret_cast<float2*>(c.warp_part + (r.i * kFlatConsumerWarps + gr.w) * 4) = make_float2(s1, s2);
                    mbar_arrive(c.p1d0 + 8 * r.i);
                }
            }
            if (i >= g.Lg) {  // P2
                const unsigned j = gr.q + g.NG * (i - g.Lg);
                const Ring r = ring_of(g, j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* dxdst = reinterpret_cast<char*>(p.dx) + goff;
                char* drdst = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<char*>(p.dres) + goff : nullptr;
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                mbar_wait(c.coef0 + 8 * r.i, r.ph);
                const float* cf = c.coefv + r.i * 8;
                const float4 cq = *reinterpret_cast<const float4*>(cf);
                const float A = cq.x, B1 = cq.y, B0 = cq.z, mean = cq.w, bq = cf[4];
#pragma unroll 2
                for (unsigned v = gr.t; v < pc.pv; v += gr.tg) {
                    float xf[VN], gf[VN], of[VN];
                    VecT<T>::unpack(lds128(base + v * 16), xf);
                    VecT<T>::unpack(lds128(base + sb + v * 16), gf);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(lds128(base + 2 * sb + v * 16), of);
#pragma unroll
                    for (int e = 0; e < VN; ++e) {
                        const float gg = bwd_masked<T, EPI>(xf[e], gf[e], EPI == MICN_EPI_ADD_LRELU ? of[e] : 0.f, mean, A, bq, p.slope);
                        gf[e] = gg;
                        xf[e] = sizeof(T) == 4 ? fmaf(A, gg, fmaf(B1, xf[e] - mean, B0)) : fmaf(A, gg, fmaf(B1, xf[e], B0));
                    }
                    stg_stream(dxdst + (size_t)v * 16, VecT<T>::pack(xf));
                    if (EPI == MICN_EPI_ADD_LRELU) stg_stream(drdst + (size_t)v * 16, VecT<T>::pack(gf));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
            }
        }
    }
}

}  // namespace micn
