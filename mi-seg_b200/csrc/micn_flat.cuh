// micn_flat.cuh - flat-partition instance_cond forward / backward (sm_100a): the default large-slab path.
//
// Measured on B200 (tools/l2bw.cu, profiles/): HBM copy 6.5 TB/s, HBM read 7.3 TB/s, but reads that hit the
// 126 MB L2 stream at 18-19 TB/s.  So the budget is HBM bytes, not L2<->SM bytes: a voxel should cross HBM
// once per tensor, while a second look at it is cheap as long as it is still in L2.  Two designs were
// built and measured before this one: cluster-per-slab (cannot balance 48 slabs on 148 SMs)
// and a shared-memory-resident flat partition (the ~5-10 us cross-CTA exchange of the statistics cannot be
// hidden behind 220 KB of shared memory per SM).
//
// How: every (n, c) slab is cut into P pieces of <= PV 16-byte vectors; piece g = slab*P + k belongs to
// CTA g % G in its round g / G (G = one persistent CTA per SM, launched cooperatively so all are
// co-resident): every SM carries the same share whatever N*C is.  A CTA walks its pieces j = 0, 1, ... in
// steps; step s runs two TASKS, each a TMA load of a piece into a shared-memory slot plus a pass over it:
//
//     P1(s)      statistics of piece s            (first touch: HBM -> L2 -> SM, L2 evict_last)
//     P2(s - L)  normalise / epilogue / backward formula of piece s - L, 128-bit streaming stores
//                                                 (second touch L steps later: served by L2, evict_first)
//
// Between the two, the piece lives in L2 (L*G pieces, a few tens of MB), not in shared memory, so the lag L
// can be as long as the exchange needs while all K slots keep prefetching.  Roles inside a CTA:
//
//   producer warp (1 lane)  issues the tasks' 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) in
//                           order, up to K tasks ahead of the consumers.
//   16 consumer warps       run the tasks in order out of shared memory (fp32 shifted sums + warp shuffle in
//                           P1; FMA + pack + st.global.v4 in P2) and hand each slot straight back.
//   2 publish warps         merge the 16 warp partials of a piece (Chan) and write the piece record to the
//                           workspace as soon as its P1 is done; never wait on another CTA.
//   8 gather warps          poll the P records of the piece's slab (batches of loads in flight, re-polled in
//                           parallel), merge them in a fixed order (bit-identical in every CTA, no atomics)
//                           and publish the per-slab coefficients P2 needs; backward: also the per-slab
//                           sums and, for the last sample of a channel, d(gamma)/d(beta) per style.
//
// Cross-CTA exchange is by 16-byte self-validating records {a, tag, b, tag} (tag = per-launch epoch): no
// counters to reset, an aborted launch cannot poison the next one.  Deadlock freedom: a slab of P pieces
// spans R <= ceil((P-1)/G)+1 rounds and the planner keeps L >= R - 1, so every P1 a record depends on runs
// before anyone can block in a P2; publishes never block.  Every wait is bounded and traps instead of hanging.
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

constexpr int kFlatConsumerWarps = 16;
constexpr int kFlatConsumerThreads = kFlatConsumerWarps * 32;  // 512
constexpr int kFlatProducerWarp = kFlatConsumerWarps;
constexpr int kFlatPublishWarp0 = kFlatConsumerWarps + 1;
constexpr int kFlatPublishWarps = 2;
constexpr int kFlatGatherWarp0 = kFlatPublishWarp0 + kFlatPublishWarps;
constexpr int kFlatGatherWarps = 8;  // a gather is a multi-microsecond latency chain: keep several in flight
constexpr int kFlatThreads = (kFlatConsumerWarps + 1 + kFlatPublishWarps + kFlatGatherWarps) * 32;  // 864
constexpr int kFlatMaxSlots = 8;
constexpr int kFlatNB = 16;          // per-piece control ring (partials, coefficients): piece j -> entry j % 16
constexpr int kFlatMaxLag = 12;      // L <= kFlatNB - 4 (entry reuse needs NB > L plus the producer's run-ahead)
constexpr int kFlatMaxPieces = 1024; // pieces per slab (workspace sizing); the planner enforces the round bound
constexpr int kFlatMinPieceVecs = 128;
constexpr uint32_t kFlatTmaChunk = 32768;

struct FlatGeom {
    unsigned long long V;  // 16-byte vectors per slab
    unsigned T;            // total pieces = num_slabs * P
    unsigned P;            // pieces per slab
    unsigned PV;           // vectors per piece (the last piece of a slab may be shorter)
    unsigned K;            // shared-memory slots (TMA landing buffers)
    unsigned L;            // steps P2 trails P1
    unsigned slot_vecs;    // vectors reserved per stream per slot (>= PV, multiple of 8)
    unsigned epoch;        // per-launch tag of the workspace records (never 0)
    unsigned poll_delay_ns, poll_backoff_ns;
    uint4* ws_piece;       // [T] piece records
    uint4* ws_slab;        // [num_slabs] per-slab records (backward parameter gradients)
    long long* trace;      // bring-up only: [grid][kFlatTraceSteps][16] %globaltimer stamps (ns) per piece, or null
};

constexpr int kFlatTraceSteps = 64;
enum { TR_LOAD = 0, TR_P1_BEGIN, TR_P1_END, TR_PUB_BEGIN, TR_PUB_END, TR_GA_BEGIN, TR_GA_POLLED, TR_GA_END,
       TR_P2_WAIT, TR_P2_BEGIN, TR_P2_END, TR_LOAD2 };
__device__ __forceinline__ void flat_trace(const FlatGeom& g, unsigned j, int ev) {
    if (g.trace && j < kFlatTraceSteps)
        g.trace[((size_t)blockIdx.x * kFlatTraceSteps + j) * 16 + ev] = (long long)globaltimer_ns();
}

// control block: slot barriers + the per-piece ring
__host__ __device__ constexpr int flat_ctl_bytes() {
    return kFlatMaxSlots * (2 * 8 + 16) + kFlatNB * (2 * 8 + kFlatConsumerWarps * 16 + 32 + 16);
}

struct FlatCtx {
    uint32_t data0, full0, empty0, p1d0, coef0;  // shared::cta addresses
    float* slot_prec;                            // [K][4]   slab constants of the task in the slot (backward)
    float* warp_part;                            // [NB][16][4]
    float* coefv;                                // [NB][8]
    float* prec;                                 // [NB][4]  slab constants of the piece (backward, for the gather)
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
    c.coef0 = c.p1d0 + kFlatNB * 8;
    float* f = reinterpret_cast<float*>(ctl + kFlatMaxSlots * 16 + kFlatNB * 16);
    c.slot_prec = f;
    c.warp_part = c.slot_prec + kFlatMaxSlots * 4;
    c.coefv = c.warp_part + kFlatNB * kFlatConsumerWarps * 4;
    c.prec = c.coefv + kFlatNB * 8;
    if (threadIdx.x == 0) {
        for (unsigned i = 0; i < g.K; ++i) {
            mbar_init(c.full0 + 8 * i, 1);
            mbar_init(c.empty0 + 8 * i, kFlatConsumerWarps);
        }
        for (unsigned i = 0; i < kFlatNB; ++i) {
            mbar_init(c.p1d0 + 8 * i, kFlatConsumerWarps);
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

// ring cursors with phase parity: slots advance once per TASK, piece-ring entries once per PIECE
struct Ring {
    unsigned i, ph;
    __device__ __forceinline__ void next(unsigned n) {
        if (++i == n) {
            i = 0;
            ph ^= 1u;
        }
    }
};
__device__ __forceinline__ Ring entry_of(unsigned j) { return Ring{j % kFlatNB, (j / kFlatNB) & 1u}; }

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
// in ascending q per lane.  A batch of up to four loads per lane is in flight at once and the missing ones
// are re-polled together, so a slab's record set costs one L2 round trip once everything is published.
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
        if (!(ok[0] && ok[1] && ok[2] && ok[3])) {
            const uint64_t t0 = globaltimer_ns();
            uint32_t spins = 0;
            do {
                __nanosleep(backoff_ns);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (!ok[i]) ok[i] = ll_try(recs + q0 + lane + 32 * i, tag, a[i], b[i]);
                if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
            } while (!(ok[0] && ok[1] && ok[2] && ok[3]));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
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

// vectors of a piece (pv vectors, strided over the 512 consumer threads) that land in consumer warp w
__device__ __forceinline__ unsigned warp_vecs(unsigned pv, unsigned w) {
    const unsigned full = pv / kFlatConsumerThreads, rem = pv % kFlatConsumerThreads;
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
        // ------------------------------------------------------------------ producer: tasks P1(s), P2(s-L) in order
        if (lane == 0) {
            const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
            Ring r{0u, 0u};
            unsigned t = 0;  // task counter
            for (unsigned s = 0; s < nj + g.L; ++s) {
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 0 ? s >= nj : s < g.L) continue;
                    const unsigned j = pass == 0 ? s : s - g.L;
                    const PieceId pc = piece_of(g, j * G + cta);
                    const unsigned n = pc.slab / C, ch = pc.slab - n * C;
                    const char* src = reinterpret_cast<const char*>(p.x) +
                                      ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) +
                                      (size_t)pc.k * g.PV * 16;
                    const uint32_t bytes = pc.pv * 16u, bar = c.full0 + 8 * r.i;
                    if (t >= g.K) mbar_wait(c.empty0 + 8 * r.i, r.ph ^ 1u);
                    flat_trace(g, j, pass == 0 ? TR_LOAD : TR_LOAD2);
                    flat_issue(c.data0 + r.i * c.slot_bytes, src, bytes, bar, pass == 0 ? pol_keep : pol_done);
                    mbar_arrive_expect_tx(bar, bytes);
                    r.next(g.K);
                    ++t;
                }
            }
        }
    } else if (warp > kFlatProducerWarp && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish: warp partials -> piece record
        for (unsigned j = warp - kFlatPublishWarp0; j < nj; j += kFlatPublishWarps) {
            const Ring e = entry_of(j);
            const unsigned gidx = j * G + cta;
            const unsigned pv = piece_vecs(g, gidx % g.P);
            const float nw = (float)(warp_vecs(pv, lane & 15) * VN);
            mbar_wait(c.p1d0 + 8 * e.i, e.ph);
            if (lane == 0) flat_trace(g, j, TR_PUB_BEGIN);
            Stat st{0.f, 0.f, 0.f};
            if (lane < kFlatConsumerWarps) {
                const float4 w = *reinterpret_cast<const float4*>(c.warp_part + (e.i * kFlatConsumerWarps + lane) * 4);
                st = stat_from_shifted(w.z, w.x, w.y, nw);
            }
            st = stat_warp_reduce(st);
            if (lane == 0) ll_store(g.ws_piece + gidx, st.mean, st.m2, g.epoch);
            if (lane == 0) flat_trace(g, j, TR_PUB_END);
        }
    } else if (warp >= kFlatGatherWarp0) {
        // ------------------------------------------------------------------ gather: slab records -> coefficients for P2
        for (unsigned j = warp - kFlatGatherWarp0; j < nj; j += kFlatGatherWarps) {
            const Ring e = entry_of(j);  // kFlatNB is a multiple of both warp counts: an entry keeps its warps
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = pc.slab / C, ch = pc.slab - n * C;
            // parameter loads first: their latency hides behind everything below
            const int style = load_style(p.styles, n, p.num_styles, p.status);
            float gamma, beta;
            load_affine(p, style, ch, gamma, beta);
            // no polling before this CTA's own piece is through P1: the other CTAs are at the same point
            mbar_wait(c.p1d0 + 8 * e.i, e.ph);
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
                *reinterpret_cast<float4*>(c.coefv + e.i * 8) =
                    make_float4(sizeof(T) == 4 ? mean : 0.f, a, sizeof(T) == 4 ? beta : fmaf(-mean, a, beta), 0.f);
                if (pc.k == 0 && p.save_mean) {
                    p.save_mean[pc.slab] = mean;
                    p.save_rstd[pc.slab] = rstd;
                }
                mbar_arrive(c.coef0 + 8 * e.i);
                flat_trace(g, j, TR_GA_END);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ consumers: the tasks, in the producer's order
        Ring r{0u, 0u};
        for (unsigned s = 0; s < nj + g.L; ++s) {
            if (s < nj) {  // P1(s): statistics; the slot goes straight back, the piece stays in L2
                const Ring e = entry_of(s);
                const unsigned pv = piece_vecs(g, (s * G + cta) % g.P);
                mbar_wait(c.full0 + 8 * r.i, r.ph);
                if (tid == 0) flat_trace(g, s, TR_P1_BEGIN);
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f, Kw = 0.f;
                if ((unsigned)warp * 32u < pv) {
                    Kw = first_elem<T>(base + warp * 512);  // shift = the warp's first element of the piece
#pragma unroll 2
                    for (unsigned v = tid; v < pv; v += kFlatConsumerThreads) {
                        float f[VN];
                        VecT<T>::unpack(lds128(base + v * 16), f);
#pragma unroll
                        for (int k = 0; k < VN; k += 2) {
                            const float d0 = f[k] - Kw, d1 = f[k + 1] - Kw;
                            sa += d0;
                            sb += d1;
                            qa = fmaf(d0, d0, qa);
                            qb = fmaf(d1, d1, qb);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
                const float s1 = warp_sum(sa + sb), s2 = warp_sum(qa + qb);
                if (lane == 0) {
                    *reinterpret_cast<float4*>(c.warp_part + (e.i * kFlatConsumerWarps + warp) * 4) =
                        make_float4(s1, s2, Kw, 0.f);
                    mbar_arrive(c.p1d0 + 8 * e.i);
                }
                if (tid == 0) flat_trace(g, s, TR_P1_END);
                r.next(g.K);
            }
            if (s >= g.L) {  // P2(s - L): normalise + epilogue from the piece's second (L2-served) copy
                const unsigned j = s - g.L;
                const Ring e = entry_of(j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* ydst = reinterpret_cast<char*>(p.y) + goff;
                const char* rsrc = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<const char*>(p.res) + goff : nullptr;
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                uint4 rv0 = make_uint4(0u, 0u, 0u, 0u);
                if (EPI == MICN_EPI_ADD_LRELU && (unsigned)tid < pc.pv) rv0 = ldg_stream(rsrc + (size_t)tid * 16);
                if (tid == 0) flat_trace(g, j, TR_P2_WAIT);
                mbar_wait(c.coef0 + 8 * e.i, e.ph);
                const float4 cf = *reinterpret_cast<const float4*>(c.coefv + e.i * 8);
                const float sub = cf.x, a = cf.y, b = cf.z;
                mbar_wait(c.full0 + 8 * r.i, r.ph);
                if (tid == 0) flat_trace(g, j, TR_P2_BEGIN);
#pragma unroll 2
                for (unsigned v = tid; v < pc.pv; v += kFlatConsumerThreads) {
                    uint4 rv = rv0;
                    if (EPI == MICN_EPI_ADD_LRELU) {  // residual straight from HBM, next one in flight
                        const unsigned vn = v + kFlatConsumerThreads;
                        if (vn < pc.pv) rv0 = ldg_stream(rsrc + (size_t)vn * 16);
                    }
                    float f[VN], rr[VN];
                    VecT<T>::unpack(lds128(base + v * 16), f);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(rv, rr);
#pragma unroll
                    for (int k = 0; k < VN; ++k) {
                        float o = sizeof(T) == 4 ? fmaf(f[k] - sub, a, b) : fmaf(f[k], a, b);
                        if (EPI == MICN_EPI_ADD_LRELU) o += rr[k];
                        if (EPI != MICN_EPI_NONE) o = o > 0.f ? o : o * p.slope;
                        f[k] = o;
                    }
                    stg_stream(ydst + (size_t)v * 16, VecT<T>::pack(f));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
                if (tid == 0) flat_trace(g, j, TR_P2_END);
                r.next(g.K);
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
            const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
            Ring r{0u, 0u};
            unsigned t = 0;
            for (unsigned s = 0; s < nj + g.L; ++s) {
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 0 ? s >= nj : s < g.L) continue;
                    const unsigned j = pass == 0 ? s : s - g.L;
                    const PieceId pc = piece_of(g, j * G + cta);
                    const unsigned n = pc.slab / C, ch = pc.slab - n * C;
                    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pass == 0) {  // per-slab constants for P1 and the gather: loads issued before the slot wait
                        const int style = load_style(p.styles, n, p.num_styles, p.status);
                        float gamma, beta;
                        load_affine(p, style, ch, gamma, beta);
                        pr = make_float4(__ldg(p.save_mean + pc.slab), __ldg(p.save_rstd + pc.slab), gamma, beta);
                    }
                    const size_t poff = (size_t)pc.k * g.PV * 16;
                    const size_t doff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + poff;
                    const char* xsrc = reinterpret_cast<const char*>(p.x) +
                                       ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) + poff;
                    const uint32_t bytes = pc.pv * 16u, bar = c.full0 + 8 * r.i;
                    const uint32_t dst = c.data0 + r.i * c.slot_bytes;
                    const uint64_t pol = pass == 0 ? pol_keep : pol_done;
                    if (t >= g.K) mbar_wait(c.empty0 + 8 * r.i, r.ph ^ 1u);
                    flat_issue(dst, xsrc, bytes, bar, pol);
                    flat_issue(dst + c.stream_bytes, reinterpret_cast<const char*>(p.dy) + doff, bytes, bar, pol);
                    if (NS == 3) flat_issue(dst + 2 * c.stream_bytes, reinterpret_cast<const char*>(p.act_out) + doff, bytes, bar, pol);
                    if (pass == 0) *reinterpret_cast<float4*>(c.slot_prec + r.i * 4) = pr;  // visible through the barrier
                    mbar_arrive_expect_tx(bar, bytes * NS);
                    r.next(g.K);
                    ++t;
                }
            }
        }
    } else if (warp > kFlatProducerWarp && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish
        for (unsigned j = warp - kFlatPublishWarp0; j < nj; j += kFlatPublishWarps) {
            const Ring e = entry_of(j);
            mbar_wait(c.p1d0 + 8 * e.i, e.ph);
            float s1 = 0.f, s2 = 0.f;
            if (lane < kFlatConsumerWarps) {
                const float2 w = *reinterpret_cast<const float2*>(c.warp_part + (e.i * kFlatConsumerWarps + lane) * 4);
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
        for (unsigned j = warp - kFlatGatherWarp0; j < nj; j += kFlatGatherWarps) {
            const Ring e = entry_of(j);
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = pc.slab / C, ch = pc.slab - n * C;
            // no polling before this CTA's own piece is through P1 (the others are at the same point); the wait
            // also makes the piece's slab constants (copied by consumer thread 0) visible
            mbar_wait(c.p1d0 + 8 * e.i, e.ph);
            const volatile float* pr = c.prec + e.i * 4;
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
                float* cf = c.coefv + e.i * 8;
                *reinterpret_cast<float4*>(cf) = make_float4(a, B1, sizeof(T) == 4 ? B0c : fmaf(-B1, mean, B0c), mean);
                cf[4] = sizeof(T) == 4 ? beta : fmaf(-mean, a, beta);
                mbar_arrive(c.coef0 + 8 * e.i);
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
                                ll_wait1(g.ws_slab + (size_t)nn * C + ch, g.epoch, g.poll_backoff_ns, ra, rb);
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
        Ring r{0u, 0u};
        const uint32_t sb = c.stream_bytes;
        for (unsigned s = 0; s < nj + g.L; ++s) {
            if (s < nj) {  // P1(s)
                const Ring e = entry_of(s);
                const unsigned pv = piece_vecs(g, (s * G + cta) % g.P);
                mbar_wait(c.full0 + 8 * r.i, r.ph);
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                const float4 pr = *reinterpret_cast<const float4*>(c.slot_prec + r.i * 4);
                if (tid == 0) *reinterpret_cast<float4*>(c.prec + e.i * 4) = pr;  // for the gather, past the slot's life
                const float mean = pr.x, a = pr.y * pr.z;
                const float bq = sizeof(T) == 4 ? pr.w : fmaf(-mean, a, pr.w);
                float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 2
                for (unsigned v = tid; v < pv; v += kFlatConsumerThreads) {
                    float xf[VN], gf[VN], of[VN];
                    VecT<T>::unpack(lds128(base + v * 16), xf);
                    VecT<T>::unpack(lds128(base + sb + v * 16), gf);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(lds128(base + 2 * sb + v * 16), of);
#pragma unroll
                    for (int k = 0; k < VN; k += 2) {
                        const float g0 = bwd_masked<T, EPI>(xf[k], gf[k], EPI == MICN_EPI_ADD_LRELU ? of[k] : 0.f, mean, a, bq, p.slope);
                        const float g1 = bwd_masked<T, EPI>(xf[k + 1], gf[k + 1], EPI == MICN_EPI_ADD_LRELU ? of[k + 1] : 0.f, mean, a, bq, p.slope);
                        s1a += g0;
                        s1b += g1;
                        s2a = fmaf(g0, xf[k] - mean, s2a);
                        s2b = fmaf(g1, xf[k + 1] - mean, s2b);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
                const float s1 = warp_sum(s1a + s1b), s2 = warp_sum(s2a + s2b);
                if (lane == 0) {
                    *reinterpret_cast<float2*>(c.warp_part + (e.i * kFlatConsumerWarps + warp) * 4) = make_float2(s1, s2);
                    mbar_arrive(c.p1d0 + 8 * e.i);
                }
                r.next(g.K);
            }
            if (s >= g.L) {  // P2(s - L)
                const unsigned j = s - g.L;
                const Ring e = entry_of(j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* dxdst = reinterpret_cast<char*>(p.dx) + goff;
                char* drdst = EPI == MICN_EPI_ADD_LRELU ? reinterpret_cast<char*>(p.dres) + goff : nullptr;
                const uint32_t base = c.data0 + r.i * c.slot_bytes;
                mbar_wait(c.coef0 + 8 * e.i, e.ph);
                const float* cf = c.coefv + e.i * 8;
                const float4 cq = *reinterpret_cast<const float4*>(cf);
                const float A = cq.x, B1 = cq.y, B0 = cq.z, mean = cq.w, bq = cf[4];
                mbar_wait(c.full0 + 8 * r.i, r.ph);
#pragma unroll 2
                for (unsigned v = tid; v < pc.pv; v += kFlatConsumerThreads) {
                    float xf[VN], gf[VN], of[VN];
                    VecT<T>::unpack(lds128(base + v * 16), xf);
                    VecT<T>::unpack(lds128(base + sb + v * 16), gf);
                    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(lds128(base + 2 * sb + v * 16), of);
#pragma unroll
                    for (int k = 0; k < VN; ++k) {
                        const float gg = bwd_masked<T, EPI>(xf[k], gf[k], EPI == MICN_EPI_ADD_LRELU ? of[k] : 0.f, mean, A, bq, p.slope);
                        gf[k] = gg;
                        xf[k] = sizeof(T) == 4 ? fmaf(A, gg, fmaf(B1, xf[k] - mean, B0)) : fmaf(A, gg, fmaf(B1, xf[k], B0));
                    }
                    stg_stream(dxdst + (size_t)v * 16, VecT<T>::pack(xf));
                    if (EPI == MICN_EPI_ADD_LRELU) stg_stream(drdst + (size_t)v * 16, VecT<T>::pack(gf));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(c.empty0 + 8 * r.i);
                r.next(g.K);
            }
        }
    }
}

}  // namespace micn
