// micn_flat.cuh - flat-partition instance_cond forward / backward (sm_100a): the default large-slab path.
//
// Measured on B200 (tools/l2bw.cu, tools/streambw.cu, profiles/): HBM copy 6.5 TB/s, HBM read 7.1-7.3 TB/s, reads
// that hit the 126 MB L2 18-19 TB/s.  So the budget is HBM bytes, not L2<->SM bytes: a voxel should cross HBM
// once per tensor, while a second look at it is cheap as long as it is still in L2.  What the earlier designs
// of this file taught: cluster-per-slab cannot balance 48 slabs on 148 SMs; shared-memory-resident
// pieces cannot hide the cross-CTA exchange; whenever the 16 warps of a CTA work on ONE piece in lock step
// (TMA ring + mbarriers) every per-piece latency chain - barrier wake-up, shuffles, the partial hand-off - is
// paid by the whole SM at once, ~1 us per piece; helper warps that spin steal the issue slots the bf16 math
// needs; a kernel that outgrows the instruction cache loses everything.
//
// How: every (n, c) slab is cut into P pieces of <= PV 16-byte vectors and every WARP of the persistent grid
// (one 16-warp CTA per SM, launched cooperatively so all are co-resident) is an independent worker: piece
// g = slab*P + k belongs to worker g % W in its step g / W, W = 16*G workers, worker = warp*G + cta so that
// neighbouring pieces sit on different SMs.  No shared memory, no block-wide barrier, no TMA: a piece is ONE
// batch of 128-bit loads per lane, issued a task ahead of its use, so a worker always has its next first-touch
// batch (HBM) or this step's second-touch batch (L2) in flight while it computes on the other (~64 KB in flight
// per SM, which tools/streambw.cu shows is enough for the full HBM read rate), and the 16 workers of an SM are
// out of phase by construction, so one worker's reduction / exchange latency is hidden by the streaming of the
// others.  Step s of a worker:
//
//     issue      second-touch batch of piece s-L (L2, evict_first)
//     P1(s)      statistics of its piece of step s (fp32 shifted sums, warp shuffle) -> piece record
//                                                 (first touch: HBM -> L2 -> SM, L2 evict_last)
//     issue      first-touch batch of piece s+1 (HBM)
//     fold       if it owns piece 0 of a slab whose pieces are at step s-1: poll the slab's P piece records
//                (batches of loads in flight), fold them in a fixed order (bit-identical whoever folds, no
//                atomics) -> slab record; backward: also d(gamma)/d(beta) for the last sample of a channel
//     P2(s - L)  poll the slab record (published a whole step ago; if its owner is late, fold it oneself: same
//                records, same order, same bits), then normalise / epilogue / backward formula of its piece of
//                step s-L with 128-bit streaming stores
//                                                 (second touch L steps later: served by L2, evict_first)
//
// Between the two touches the piece lives in L2 (L*W pieces, a few tens of MB).  At launch the workers
// therefore read at full HBM speed for L steps while the first statistics are exchanged, and at the end the
// backlog of L P2 steps hides the last exchange.
//
// Statistics are folded without a chain of divisions: partials (n, mean, M2) are summed about a common
// reference `ref` (the first partial's mean - itself a mean, never an outlier) as
//     A = sum n_q (mean_q - ref),  B = sum [M2_q + n_q (mean_q - ref)^2]   ->   mean = ref + A/N,  M2 = B - A^2/N
// which is exact algebra, well conditioned because |mean_q - ref| is of the order of the spread, and uses one
// division per fold.
//
// Cross-worker exchange is by 16-byte self-validating records {a, tag, b, tag} (tag = per-launch epoch): no
// counters to reset, an aborted launch cannot poison the next one.  Deadlock freedom (L >= 2): P1 never waits;
// a fold at step s needs P1 records of steps <= s (a slab may straddle two steps), all written before any fold
// of step s starts; a P2 at step s needs a fold of step s-L+1 <= s-1 (or folds for itself).  Every wait is bounded and traps instead of hanging.
//
// HBM traffic: forward reads x once and writes y once (2*E*s); backward reads x, dy [, act_out] once and
// writes dx [, dresidual] once (3*E*s / 5*E*s) - the algorithmic minimum; L2<->SM carries
// one extra read of the inputs.
//
// Reference semantics: networks/norms/conditional_instance_norm.py:59-60 (+ ATen instance_norm:
// biased variance, eps inside the sqrt), epilogues networks/blocks/dynunet_block.py:107-125.
#pragma once

#include "micn_common.cuh"

namespace micn {

constexpr int kFlatWarps = 16;                 // workers per CTA
constexpr int kFlatThreads = kFlatWarps * 32;  // 512: one CTA per SM, up to 128 registers per thread
constexpr int kFlatMaxLag = 8;
constexpr int kFlatMaxPieces = 1024;  // pieces per slab (workspace sizing)
constexpr int kFlatMinPieceVecs = 128;
#ifndef MICN_FLAT_LPL
#define MICN_FLAT_LPL 8
#endif
constexpr int kFlatLoadsPerLane = MICN_FLAT_LPL;  // 128-bit loads a lane has in flight per batch, over all streams

struct FlatGeom {
    unsigned long long V;  // 16-byte vectors per slab
    unsigned T;            // total pieces = num_slabs * P
    unsigned P;            // pieces per slab
    unsigned PV;           // vectors per piece ...
    unsigned PVlast;       // ... except the last piece of a slab
    unsigned L;            // steps P2 trails P1 (>= 2)
    unsigned epoch;        // per-launch tag of the workspace records (never 0)
    unsigned poll_backoff_ns;
    FastDiv divP, divC;    // piece index -> slab, slab -> sample
    uint4* ws_piece;       // [T] piece records
    uint4* ws_slab;        // [num_slabs] slab records (forward: mean, rstd; backward: sum g, sum g*xhat)
    long long* trace;      // bring-up only: [grid][kFlatTraceSteps][16] %globaltimer stamps (ns) of warp 0, or null
};

constexpr int kFlatTraceSteps = 64;
enum { TR_LOAD = 0, TR_P1_BEGIN, TR_P1_END, TR_PUB_BEGIN, TR_PUB_END, TR_GA_BEGIN, TR_GA_POLLED, TR_GA_END,
       TR_P2_WAIT, TR_P2_BEGIN, TR_P2_END, TR_LOAD2 };
__device__ __forceinline__ void flat_trace(const FlatGeom& g, unsigned j, int ev) {
    if (g.trace && j < kFlatTraceSteps)
        g.trace[((size_t)blockIdx.x * kFlatTraceSteps + j) * 16 + ev] = (long long)globaltimer_ns();
}

struct PieceId {
    unsigned slab, k, pv;  // slab index, piece index inside the slab, vectors in this piece
};
__device__ __forceinline__ unsigned piece_vecs(const FlatGeom& g, unsigned k) { return k + 1 == g.P ? g.PVlast : g.PV; }
__device__ __forceinline__ PieceId piece_of(const FlatGeom& g, unsigned gidx) {
    PieceId p;
    p.slab = fastdiv(gidx, g.divP);
    p.k = gidx - p.slab * g.P;
    p.pv = piece_vecs(g, p.k);
    return p;
}

// 128-bit streaming loads with an L2 eviction-priority hint (first touch: keep; second touch: done with it)
__device__ __forceinline__ uint4 ldg_hint(const void* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
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

// Poll the `count` records at `recs` (lane q handles records q, q+32, ...) and fold them with `fold(q, a, b)`
// in ascending q per lane; `first(a0, b0)` sees record 0 (the common reference) before any fold.  A batch of up to four loads per lane is in flight at once and the missing ones
// are re-polled together, so a slab's record set costs one L2 round trip once everything is published.
template <typename First, typename Fold>
__device__ __forceinline__ void ll_gather(const uint4* recs, unsigned count, unsigned tag, unsigned backoff_ns, int lane,
                                          First first, Fold fold) {
    constexpr int R = 8;  // records per lane in flight: a slab of <= 256 pieces costs one L2 round trip
    for (unsigned q0 = 0; q0 < count; q0 += 32 * R) {
        float a[R], b[R];
        bool ok[R], all = true;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const unsigned q = q0 + lane + 32 * i;
            ok[i] = q < count ? ll_try(recs + q, tag, a[i], b[i]) : true;
            all = all && ok[i];
        }
        if (!all) {
            const uint64_t t0 = globaltimer_ns();
            uint32_t spins = 0;
            do {
                __nanosleep(backoff_ns);
                all = true;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    if (!ok[i]) ok[i] = ll_try(recs + q0 + lane + 32 * i, tag, a[i], b[i]);
                    all = all && ok[i];
                }
                if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
            } while (!all);
        }
        if (q0 == 0) first(__shfl_sync(0xffffffffu, a[0], 0), __shfl_sync(0xffffffffu, b[0], 0));  // record 0
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const unsigned q = q0 + lane + 32 * i;
            if (q < count) fold(q, a[i], b[i]);
        }
    }
}
__device__ __forceinline__ void ll_wait1(const uint4* p, unsigned tag, unsigned backoff_ns, float& a, float& b) {
    if (ll_try(p, tag, a, b)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    do {
        __nanosleep(backoff_ns);
        if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
    } while (!ll_try(p, tag, a, b));
}


// ---- a worker's walk over its piece: lane handles vectors lane + 32*(b*U + i).  A whole batch of U vectors per
//      lane per stream is in flight at once; the planner sizes pieces to ONE batch (U*32 vectors), and the step
//      loop below issues a task's batch well before it is used, so a worker always has an HBM batch (next P1)
//      or an L2 batch (this step's P2) in flight while it computes on the other.
template <int NS, int U>
struct Batch {
    uint4 q[NS][U];
};

// one task of a worker: a piece and where its lane-0.. vectors live
struct FlatTask {
    unsigned gidx, slab, pv;
    unsigned long long xoff;  // bytes: this lane's vector 0 of the piece inside x (strided) ...
    unsigned long long doff;  // ... and inside the dense [N, C, M] tensors
    unsigned n, ch;
};
__device__ __forceinline__ FlatTask flat_task(const FlatGeom& g, unsigned gidx, unsigned C, long long x_sN, long long x_sC,
                                              long long M, unsigned es, int lane) {
    FlatTask t;
    const PieceId pc = piece_of(g, gidx);
    t.gidx = gidx;
    t.slab = pc.slab;
    t.pv = pc.pv;
    t.n = fastdiv(pc.slab, g.divC);
    t.ch = pc.slab - t.n * C;
    const unsigned long long poff = (unsigned long long)pc.k * g.PV * 16ull + (unsigned long long)lane * 16ull;
    t.xoff = (unsigned long long)(((long long)t.n * x_sN + (long long)t.ch * x_sC) * (long long)es) + poff;
    t.doff = (unsigned long long)pc.slab * (unsigned long long)M * es + poff;
    return t;
}

// fold of a slab's piece statistics about record 0's mean (see the file header); every lane returns the totals
struct SlabStat {
    float mean, m2;
};
template <int VN>
__device__ __noinline__ SlabStat fold_slab_stats(const FlatGeom& g, unsigned slab, float M, int lane) {
    float ref = 0.f, A = 0.f, B = 0.f;
    ll_gather(g.ws_piece + (size_t)slab * g.P, g.P, g.epoch, g.poll_backoff_ns, lane, [&](float a0, float) { ref = a0; },
              [&](unsigned q, float a, float b) {
                  const float nq = (float)(piece_vecs(g, q) * VN), d = a - ref;
                  A = fmaf(nq, d, A);
                  B += fmaf(nq * d, d, b);
              });
    A = warp_sum(A);
    B = warp_sum(B);
    const float m = A / M;
    return SlabStat{ref + m, fmaxf(B - A * m, 0.f)};
}

// a few polls of a slab record; false if its owner has not folded yet (the caller then folds for itself: same
// records, same order, same bits)
__device__ __forceinline__ bool ll_poll_short(const uint4* p, unsigned tag, unsigned backoff_ns, float& a, float& b) {
    for (int i = 0; i < 6; ++i) {
        if (ll_try(p, tag, a, b)) return true;
        __nanosleep(backoff_ns);
    }
    return ll_try(p, tag, a, b);
}

// =================================================================================================
// forward
// =================================================================================================
template <typename T, int EPI>
__global__ void __launch_bounds__(kFlatThreads, 1) micn_fwd_flat_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ FlatGeom g) {
    constexpr int VN = VecT<T>::N;
    constexpr int NS = EPI == MICN_EPI_ADD_LRELU ? 2 : 1;  // P2: x [, residual]
    constexpr int U = kFlatLoadsPerLane / NS;              // vectors per lane per stream per batch (P1 and P2 alike)
    const unsigned G = gridDim.x, W = G * kFlatWarps;
    const int lane = threadIdx.x & 31;
    const unsigned worker = (threadIdx.x >> 5) * G + blockIdx.x;
    const unsigned nj = worker < g.T ? (g.T - worker + W - 1) / W : 0;
    const unsigned C = (unsigned)p.C;
    const bool tr = threadIdx.x == 0;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
    const float Mf = (float)p.M;
    const char* xb = reinterpret_cast<const char*>(p.x);
    const char* rb = reinterpret_cast<const char*>(p.res);
    char* yb = reinterpret_cast<char*>(p.y);

    auto loadA = [&](const FlatTask& t, unsigned b, Batch<1, U>& A) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const unsigned v = lane + 32u * (b * U + i);
            A.q[0][i] = make_uint4(0u, 0u, 0u, 0u);
            if (v < t.pv) A.q[0][i] = ldg_hint(xb + t.xoff + (size_t)(b * U + i) * 512, pol_keep);
        }
    };
    auto loadB = [&](const FlatTask& t, unsigned b, Batch<NS, U>& B) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const unsigned v = lane + 32u * (b * U + i);
            if (v < t.pv) {
                B.q[0][i] = ldg_hint(xb + t.xoff + (size_t)(b * U + i) * 512, pol_done);
                if (NS == 2) B.q[NS - 1][i] = ldg_hint(rb + t.doff + (size_t)(b * U + i) * 512, pol_done);
            }
        }
    };

    FlatTask ta, tb;
    Batch<1, U> A;
    Batch<NS, U> B;
    if (nj) {
        ta = flat_task(g, worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
        loadA(ta, 0, A);
    }
    for (unsigned s = 0; s < nj + g.L; ++s) {
        const bool hasB = s >= g.L;
        float gamma = 1.f, beta = 0.f;
        if (hasB) {
            // ---- second touch of piece s-L: its batch (L2) and parameters go in flight now, used at the end of the step
            tb = flat_task(g, (s - g.L) * W + worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
            loadB(tb, 0, B);
            const int style = load_style(p.styles, tb.n, p.num_styles, p.status);
            load_affine_gc(p, style, tb.ch, gamma, beta);
        }
        if (s < nj) {
            // ---- P1(s): statistics of this worker's piece (its first batch was issued a step ago) -> piece record
            if (tr) flat_trace(g, s, TR_P1_BEGIN);
            float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f, K = 0.f;
            const unsigned pv = ta.pv, nb = (pv + 32 * U - 1) / (32 * U);
            for (unsigned b = 0; b < nb; ++b) {
                if (b) loadA(ta, b, A);
                if (b == 0) {  // shift = the piece's first element
                    float f0[VN];
                    VecT<T>::unpack(A.q[0][0], f0);
                    K = __shfl_sync(0xffffffffu, f0[0], 0);
                }
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    const unsigned v = lane + 32u * (b * U + i);
                    if (v < pv) {
                        float f[VN];
                        VecT<T>::unpack(A.q[0][i], f);
#pragma unroll
                        for (int k = 0; k < VN; k += 2) {
                            const float d0 = f[k] - K, d1 = f[k + 1] - K;
                            sa += d0;
                            sb += d1;
                            qa = fmaf(d0, d0, qa);
                            qb = fmaf(d1, d1, qb);
                        }
                    }
                }
            }
            const float s1 = warp_sum(sa + sb), s2 = warp_sum(qa + qb);
            if (lane == 0) {
                const float m = s1 / (float)(pv * VN);
                ll_store(g.ws_piece + ta.gidx, K + m, fmaxf(s2 - s1 * m, 0.f), g.epoch);
            }
            if (tr) flat_trace(g, s, TR_P1_END);
            // ---- first touch of piece s+1 goes in flight (HBM) while this step folds and normalises
            if (s + 1 < nj) {
                ta = flat_task(g, (s + 1) * W + worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
                loadA(ta, 0, A);
            }
        }
        if (s >= 1 && s - 1 < nj) {
            // ---- fold: the owner of a slab's piece 0 turns the slab's piece records (all written a step ago)
            //      into the slab record, L - 1 steps before anyone needs it
            const unsigned gidx = (s - 1) * W + worker;
            const unsigned slab = fastdiv(gidx, g.divP);
            if (gidx == slab * g.P) {
                if (tr) flat_trace(g, s - 1, TR_GA_BEGIN);
                const SlabStat st = fold_slab_stats<VN>(g, slab, Mf, lane);
                if (lane == 0) {
                    const float rstd = 1.f / sqrtf(st.m2 / Mf + p.eps);  // biased variance, eps inside the sqrt
                    ll_store(g.ws_slab + slab, st.mean, rstd, g.epoch);
                    if (p.save_mean) {
                        p.save_mean[slab] = st.mean;
                        p.save_rstd[slab] = rstd;
                    }
                }
                if (tr) flat_trace(g, s - 1, TR_GA_END);
            }
        }
        if (hasB) {
            // ---- P2(s - L): normalise + epilogue from the piece's second (L2-served) copy
            const unsigned j = s - g.L;
            if (tr) flat_trace(g, j, TR_P2_WAIT);
            float mean, rstd;
            if (!ll_poll_short(g.ws_slab + tb.slab, g.epoch, g.poll_backoff_ns, mean, rstd)) {
                const SlabStat st = fold_slab_stats<VN>(g, tb.slab, Mf, lane);  // the owner is late: same fold, same bits
                mean = st.mean;
                rstd = 1.f / sqrtf(st.m2 / Mf + p.eps);
            }
            const float ca = rstd * gamma;
            // fp32: (x - mean) * a + beta.  16-bit: x * a + (beta - mean * a): one FMA per element.
            const float sub = sizeof(T) == 4 ? mean : 0.f, cb = sizeof(T) == 4 ? beta : fmaf(-mean, ca, beta);
            if (tr) flat_trace(g, j, TR_P2_BEGIN);
            const unsigned pv = tb.pv, nb = (pv + 32 * U - 1) / (32 * U);
            for (unsigned b = 0; b < nb; ++b) {
                if (b) loadB(tb, b, B);
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    const unsigned v = lane + 32u * (b * U + i);
                    if (v < pv) {
                        float f[VN], rr[VN];
                        VecT<T>::unpack(B.q[0][i], f);
                        if (NS == 2) VecT<T>::unpack(B.q[NS - 1][i], rr);
#pragma unroll
                        for (int k = 0; k < VN; ++k) {
                            float o = sizeof(T) == 4 ? fmaf(f[k] - sub, ca, cb) : fmaf(f[k], ca, cb);
                            if (NS == 2) o += rr[k];
                            if (EPI != MICN_EPI_NONE) o = o > 0.f ? o : o * p.slope;
                            f[k] = o;
                        }
                        stg_stream(yb + tb.doff + (size_t)(b * U + i) * 512, VecT<T>::pack(f));
                    }
                }
            }
            if (tr) flat_trace(g, j, TR_P2_END);
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

// (mean, rstd, gamma, beta) of a slab
template <typename P>
__device__ __forceinline__ float4 slab_consts(const P& p, unsigned slab, unsigned n, unsigned ch) {
    const int style = load_style(p.styles, n, p.num_styles, p.status);
    float gamma, beta;
    load_affine_gc(p, style, ch, gamma, beta);
    return make_float4(__ldg(p.save_mean + slab), __ldg(p.save_rstd + slab), gamma, beta);
}

// sum of a slab's piece records (sum g, sum g*(x-mean)); every lane returns the totals
__device__ __noinline__ float2 fold_slab_sums(const FlatGeom& g, unsigned slab, int lane) {
    float S1 = 0.f, S2 = 0.f;
    ll_gather(g.ws_piece + (size_t)slab * g.P, g.P, g.epoch, g.poll_backoff_ns, lane, [](float, float) {},
              [&](unsigned, float a, float b) {
                  S1 += a;
                  S2 += b;
              });
    return make_float2(warp_sum(S1), warp_sum(S2));
}

template <typename T, int EPI>
__global__ void __launch_bounds__(kFlatThreads, 1) micn_bwd_flat_kernel(const __grid_constant__ BwdParams p, const __grid_constant__ FlatGeom g) {
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU) ? 3 : 2;  // x, dy [, act_out]
    constexpr int VN = VecT<T>::N;
    constexpr int U = kFlatLoadsPerLane / NS;  // vectors per lane per stream per batch
    const unsigned G = gridDim.x, W = G * kFlatWarps;
    const int lane = threadIdx.x & 31;
    const unsigned worker = (threadIdx.x >> 5) * G + blockIdx.x;
    const unsigned nj = worker < g.T ? (g.T - worker + W - 1) / W : 0;
    const unsigned C = (unsigned)p.C;
    const bool tr = threadIdx.x == 0;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
    const float invM = 1.f / (float)p.M;
    const char* xb = reinterpret_cast<const char*>(p.x);
    const char* gb = reinterpret_cast<const char*>(p.dy);
    const char* ob = reinterpret_cast<const char*>(p.act_out);
    char* dxb = reinterpret_cast<char*>(p.dx);
    char* drb = reinterpret_cast<char*>(p.dres);

    auto load = [&](const FlatTask& t, unsigned b, Batch<NS, U>& Q, uint64_t pol) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const unsigned v = lane + 32u * (b * U + i);
            if (v < t.pv) {
                const size_t o = (size_t)(b * U + i) * 512;
                Q.q[0][i] = ldg_hint(xb + t.xoff + o, pol);
                Q.q[1][i] = ldg_hint(gb + t.doff + o, pol);
                if (NS == 3) Q.q[NS - 1][i] = ldg_hint(ob + t.doff + o, pol);
            }
        }
    };

    FlatTask ta, tb;
    Batch<NS, U> A, B;
    float4 pra = make_float4(0.f, 0.f, 0.f, 0.f), prb = pra;
    if (nj) {
        ta = flat_task(g, worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
        load(ta, 0, A, pol_keep);
        pra = slab_consts(p, ta.slab, ta.n, ta.ch);
    }
    for (unsigned s = 0; s < nj + g.L; ++s) {
        const bool hasB = s >= g.L;
        if (hasB) {
            // ---- second touch of piece s-L: its batch (L2) and constants go in flight now
            tb = flat_task(g, (s - g.L) * W + worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
            load(tb, 0, B, pol_done);
            prb = slab_consts(p, tb.slab, tb.n, tb.ch);
        }
        if (s < nj) {
            // ---- P1(s): sum g, sum g*(x - mean) of this worker's piece -> piece record
            if (tr) flat_trace(g, s, TR_P1_BEGIN);
            const float mean = pra.x, ca = pra.y * pra.z;
            const float bq = sizeof(T) == 4 ? pra.w : fmaf(-mean, ca, pra.w);
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            const unsigned pv = ta.pv, nb = (pv + 32 * U - 1) / (32 * U);
            for (unsigned b = 0; b < nb; ++b) {
                if (b) load(ta, b, A, pol_keep);
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    const unsigned v = lane + 32u * (b * U + i);
                    if (v < pv) {
                        float xf[VN], gf[VN], of[VN];
                        VecT<T>::unpack(A.q[0][i], xf);
                        VecT<T>::unpack(A.q[1][i], gf);
                        if (NS == 3) VecT<T>::unpack(A.q[NS - 1][i], of);
#pragma unroll
                        for (int k = 0; k < VN; k += 2) {
                            const float g0 = bwd_masked<T, EPI>(xf[k], gf[k], NS == 3 ? of[k] : 0.f, mean, ca, bq, p.slope);
                            const float g1 = bwd_masked<T, EPI>(xf[k + 1], gf[k + 1], NS == 3 ? of[k + 1] : 0.f, mean, ca, bq, p.slope);
                            s1a += g0;
                            s1b += g1;
                            s2a = fmaf(g0, xf[k] - mean, s2a);
                            s2b = fmaf(g1, xf[k + 1] - mean, s2b);
                        }
                    }
                }
            }
            const float s1 = warp_sum(s1a + s1b), s2 = warp_sum(s2a + s2b);
            if (lane == 0) ll_store(g.ws_piece + ta.gidx, s1, s2, g.epoch);
            if (tr) flat_trace(g, s, TR_P1_END);
            // ---- first touch of piece s+1 goes in flight (HBM) while this step folds and writes dx
            if (s + 1 < nj) {
                ta = flat_task(g, (s + 1) * W + worker, C, p.x_sN, p.x_sC, p.M, sizeof(T), lane);
                load(ta, 0, A, pol_keep);
                pra = slab_consts(p, ta.slab, ta.n, ta.ch);
            }
        }
        if (s >= 1 && s - 1 < nj) {
            // ---- fold: the owner of a slab's piece 0 sums the slab's piece records (all written a step ago)
            //      into the slab record, L - 1 steps before anyone needs it
            const unsigned gidx = (s - 1) * W + worker;
            const unsigned slab = fastdiv(gidx, g.divP);
            if (gidx == slab * g.P) {
                const unsigned n = fastdiv(slab, g.divC), ch = slab - n * C;
                const float rstd = __ldg(p.save_rstd + slab);
                if (tr) flat_trace(g, s - 1, TR_GA_BEGIN);
                const float2 S = fold_slab_sums(g, slab, lane);
                const float S1 = S.x, S2r = S.y * rstd;  // sum g, sum g * xhat
                if (lane == 0) ll_store(g.ws_slab + slab, S1, S2r, g.epoch);
                if (p.dgamma) {
                    if (p.N == 1) {
                        // one sample: this slab's sums ARE the gradients of its style's row
                        const int style = load_style(p.styles, 0, p.num_styles, nullptr);
                        for (int st = lane; st < p.num_styles; st += 32) {
                            p.dbeta[(size_t)st * C + ch] = st == style ? S1 : 0.f;
                            p.dgamma[(size_t)st * C + ch] = st == style ? S2r : 0.f;
                        }
                    } else if (n == (unsigned)p.N - 1) {
                        // last sample of this channel: fold every sample's slab record per style, fixed order
                        for (int st = 0; st < p.num_styles; ++st) {
                            float ab = 0.f, ag = 0.f;
                            for (unsigned nn = lane; nn < (unsigned)p.N; nn += 32) {
                                float ra, rb;
                                ll_wait1(g.ws_slab + (size_t)nn * C + ch, g.epoch, g.poll_backoff_ns, ra, rb);
                                if (load_style(p.styles, nn, p.num_styles, nullptr) == st) {
                                    ab += ra;
                                    ag += rb;
                                }
                            }
                            ab = warp_sum(ab);
                            ag = warp_sum(ag);
                            if (lane == 0) {
                                p.dbeta[(size_t)st * C + ch] = ab;
                                p.dgamma[(size_t)st * C + ch] = ag;
                            }
                        }
                    }
                }
                if (tr) flat_trace(g, s - 1, TR_GA_END);
            }
        }
        if (hasB) {
            // ---- P2(s - L)
            const unsigned j = s - g.L;
            if (tr) flat_trace(g, j, TR_P2_WAIT);
            const float mean = prb.x, rstd = prb.y, Ac = rstd * prb.z;
            float S1, S2r;
            if (!ll_poll_short(g.ws_slab + tb.slab, g.epoch, g.poll_backoff_ns, S1, S2r)) {
                const float2 S = fold_slab_sums(g, tb.slab, lane);  // the owner is late: same fold, same bits
                S1 = S.x;
                S2r = S.y * rstd;
            }
            // dx = a*g - a*S1/M - a*rstd*(S2r/M)*(x - mean)
            const float B1 = -Ac * S2r * invM * rstd, B0c = -Ac * S1 * invM;
            const float B0 = sizeof(T) == 4 ? B0c : fmaf(-B1, mean, B0c);
            const float bq = sizeof(T) == 4 ? prb.w : fmaf(-mean, Ac, prb.w);
            if (tr) flat_trace(g, j, TR_P2_BEGIN);
            const unsigned pv = tb.pv, nb = (pv + 32 * U - 1) / (32 * U);
            for (unsigned b = 0; b < nb; ++b) {
                if (b) load(tb, b, B, pol_done);
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    const unsigned v = lane + 32u * (b * U + i);
                    if (v < pv) {
                        float xf[VN], gf[VN], of[VN];
                        VecT<T>::unpack(B.q[0][i], xf);
                        VecT<T>::unpack(B.q[1][i], gf);
                        if (NS == 3) VecT<T>::unpack(B.q[NS - 1][i], of);
#pragma unroll
                        for (int k = 0; k < VN; ++k) {
                            const float gg = bwd_masked<T, EPI>(xf[k], gf[k], NS == 3 ? of[k] : 0.f, mean, Ac, bq, p.slope);
                            gf[k] = gg;
                            xf[k] = sizeof(T) == 4 ? fmaf(Ac, gg, fmaf(B1, xf[k] - mean, B0)) : fmaf(Ac, gg, fmaf(B1, xf[k], B0));
                        }
                        const size_t o = (size_t)(b * U + i) * 512;
                        stg_stream(dxb + tb.doff + o, VecT<T>::pack(xf));
                        if (NS == 3) stg_stream(drb + tb.doff + o, VecT<T>::pack(gf));
                    }
                }
            }
            if (tr) flat_trace(g, j, TR_P2_END);
        }
    }
}

}  // namespace micn
