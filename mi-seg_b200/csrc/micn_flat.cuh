// micn_flat.cuh - flat-partition instance_cond forward / backward (sm_100a): the default large-slab path.
//
// Measured on B200 (tools/l2bw.cu, tools/streambw.cu, profiles/): HBM copy 6.5 TB/s, HBM read 7.1-7.3 TB/s, reads
// that hit the 126 MB L2 18-19 TB/s.  So the budget is HBM bytes, not L2<->SM bytes: a voxel should cross HBM
// once per tensor, while a second look at it is cheap as long as it is still in L2.  A ring of 1-D TMA bulk
// copies streams at the full HBM read rate from one persistent CTA per SM once 96 KB are in flight per SM
// (6 slots x 16 KB: 7.1 TB/s; 3 x 8 KB: 3.8 TB/s) and costs the consumer warps no issue slots.  What the
// earlier designs taught: cluster-per-slab cannot balance 48 slabs on 148 SMs;
// shared-memory-resident pieces cannot hide the cross-CTA exchange; a 220 KB-deep prefetch turns the memory
// system into a 5 us FIFO that every record store and poll queues in; helper warps that spin on mbarriers
// steal the issue slots the bf16 math needs; a kernel that outgrows the instruction cache loses everything.
//
// How: every (n, c) slab is cut into P pieces of <= PV 16-byte vectors; piece g = slab*P + k belongs to
// CTA g % G in its round g / G (G = one persistent CTA per SM, launched cooperatively so all are
// co-resident): every SM carries the same share whatever N*C is.  A CTA walks its pieces j = 0, 1, ... in
// steps; step s runs two tasks, each fed by its own ring of shared-memory slots and its own TMA producer lane:
//
//     ring A: P1(s)      statistics of piece s          (first touch: HBM -> L2 -> SM, L2 evict_last)
//     ring B: P2(s - L)  normalise / epilogue / backward formula of piece s - L, 128-bit streaming stores
//                                                       (second touch L steps later: served by L2, evict_first)
//
// Between the two the piece lives in L2 (L*G pieces, a few tens of MB), not in shared memory.  At launch the
// CTAs therefore read at full HBM speed for L steps while the first statistics are exchanged, and at the end
// the backlog of L P2 tasks hides the last exchange.  Roles inside a CTA:
//
//   producer A / B (1 lane each)  1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), KA / KB slots ahead;
//                           B issues P2(j)'s copy as soon as P1(j) is through, so it is in the slot long before use.
//   16 (8) consumer warps   each warp owns a fixed 1/16 (1/8) of every piece: fp32 shifted sums + warp shuffle in P1;
//                           FMA + pack + st.global.v4 in P2.  Every wait parks the warp (try_wait + suspend hint).
//   1 publish warp          folds the 16 warp partials of a piece and write the piece record to the workspace as
//                           soon as its P1 is done; never wait on another CTA.
//   4 (2) gather warps      poll the P records of the piece's slab (batches of loads in flight, re-polled in
//                           parallel), fold them in a fixed order (bit-identical in every CTA, no atomics)
//                           and publish the per-slab coefficients P2 needs; backward: also the per-slab
//                           sums and, for the last sample of a channel, d(gamma)/d(beta) per style.
//
// Statistics are folded without a chain of divisions: partials (n, mean, M2) are summed about a common
// reference `ref` (the first partial's mean - itself a mean, never an outlier) as
//     A = sum n_q (mean_q - ref),  B = sum [M2_q + n_q (mean_q - ref)^2]   ->   mean = ref + A/N,  M2 = B - A^2/N
// which is exact algebra, well conditioned because |mean_q - ref| is of the order of the spread, and uses one
// division per fold.
//
// Cross-CTA exchange is by 16-byte self-validating records {a, tag, b, tag}; the tag is a per-launch epoch kept in
// the workspace header and advanced by the kernel itself (flat_epoch_tag), so records of an earlier launch - or of
// an earlier replay of a captured CUDA graph - never look current and nothing has to be cleared.  Deadlock freedom: a slab of P pieces
// spans R <= ceil((P-1)/G)+1 rounds and the planner keeps L >= R, so every P1 a record depends on runs
// before anyone can block in a P2; publishes never block.  Every wait is bounded and traps instead of hanging.
//
// HBM traffic: forward reads x once and writes y once (2*E*s); backward reads x, dy [, act_out] once and
// writes dx [, dresidual] once (3*E*s / 5*E*s) - the algorithmic minimum; L2<->SM carries
// one extra read of the inputs.
//
// Reference semantics: networks/norms/conditional_instance_norm.py:59-60 (+ ATen instance_norm:
// biased variance, eps inside the sqrt), epilogues networks/blocks/dynunet_block.py:107-125.
// This header is included once per CTA SHAPE (micn_api.cu): MICN_FLAT_NS names the sub-namespace, MICN_FLAT_CW /
// MICN_FLAT_GW / MICN_FLAT_CPS the consumer warps, gather warps and persistent CTAs per SM of that shape
// (flat1: 16 / 4 / 1, fp32 forward and every backward; flat2: 8 / 2 / 2, 16-bit forward - the numbers in brackets above).
#include "micn_common.cuh"

#ifndef MICN_FLAT_NS
#define MICN_FLAT_NS flat1
#endif

namespace micn {
namespace MICN_FLAT_NS {

#ifndef MICN_FLAT_CW
#define MICN_FLAT_CW 16  // consumer warps per CTA
#endif
#ifndef MICN_FLAT_GW
#define MICN_FLAT_GW 4   // gather warps per CTA
#endif
#ifndef MICN_FLAT_CPS
#define MICN_FLAT_CPS 1  // persistent CTAs per SM
#endif
constexpr int kFlatConsumerWarps = MICN_FLAT_CW;
constexpr int kFlatCtasPerSm = MICN_FLAT_CPS;
constexpr int kFlatConsumerThreads = kFlatConsumerWarps * 32;  // 512 (flat1) / 256 (flat2)
constexpr int kFlatProducerWarpA = kFlatConsumerWarps;
constexpr int kFlatProducerWarpB = kFlatConsumerWarps + 1;
constexpr int kFlatPublishWarp0 = kFlatConsumerWarps + 2;
constexpr int kFlatPublishWarps = 1;
constexpr int kFlatGatherWarp0 = kFlatPublishWarp0 + kFlatPublishWarps;
constexpr int kFlatGatherWarps = MICN_FLAT_GW;  // a gather is a multi-microsecond latency chain: keep several in flight
constexpr int kFlatThreads = (kFlatConsumerWarps + 2 + kFlatPublishWarps + kFlatGatherWarps) * 32;  // 736 (flat1: 88 registers) / 416 (flat2)
constexpr int kFlatMaxSlots = 8;     // per ring
constexpr int kFlatNB = 32;          // per-piece control ring (partials, coefficients): piece j -> entry j % 32
constexpr int kFlatMaxLag = 30;      // L <= kFlatNB - 1: an entry is recycled only after its piece's P2 is done
constexpr int kFlatMaxPieces = 1024; // pieces per slab (workspace sizing); the planner enforces the round bound
constexpr int kFlatMinPieceVecs = 128;
constexpr uint32_t kFlatTmaChunk = 32768;

struct FlatGeom {
    unsigned long long V;  // 16-byte vectors per slab
    unsigned T;            // total pieces = num_slabs * P
    unsigned P;            // pieces per slab
    unsigned PV;           // vectors per piece ...
    unsigned PVlast;       // ... except the last piece of a slab
    unsigned KA, KB;       // shared-memory slots of ring A (P1) and ring B (P2)
    unsigned L;            // steps P2 trails P1
    unsigned slot_vecs;    // vectors reserved per stream per slot (>= PV, multiple of 8)
    unsigned* ws_ctl;      // workspace header: 64-bit word {launch epoch : CTAs arrived} (device-side, CUDA-graph safe)
    unsigned poll_delay_ns, poll_delay_tail_ns, poll_backoff_ns;
    FastDiv divP, divC;    // piece index -> slab, slab -> sample
    uint4* ws_piece;       // [T] piece records
    uint4* ws_slab;        // [num_slabs] per-slab records (backward parameter gradients)
    long long* trace;      // bring-up only: [grid][kFlatTraceSteps][16] %globaltimer stamps (ns) per piece, or null
};

constexpr int kFlatTraceSteps = 64;
enum { TR_LOAD = 0, TR_P1_BEGIN, TR_P1_END, TR_PUB_BEGIN, TR_PUB_END, TR_GA_BEGIN, TR_GA_POLLED, TR_GA_END,
       TR_P2_WAIT, TR_P2_BEGIN, TR_P2_END, TR_LOAD2 };
// compiled in only for the bring-up build (tools/Makefile target `trace`): the checks alone were 3 % of the
// executed instructions
#ifndef MICN_FLAT_TRACE
#define MICN_FLAT_TRACE 0
#endif
__device__ __forceinline__ void flat_trace(const FlatGeom& g, unsigned j, int ev) {
#if MICN_FLAT_TRACE
    if (g.trace && j < kFlatTraceSteps)
        g.trace[((size_t)blockIdx.x * kFlatTraceSteps + j) * 16 + ev] = (long long)globaltimer_ns();
#else
    (void)g;
    (void)j;
    (void)ev;
#endif
}

// The tag of this launch's records comes from the workspace itself, not from the host: a captured CUDA graph
// replays the same kernel parameters, and records left by the previous replay must not look current.  The header
// holds one 64-bit word {epoch : arrivals}; one lane per CTA adds 1 to it, which reads the epoch and registers the
// arrival in a single atomic (nothing to order).  The last CTA to arrive (all CTAs are co-resident) clears the
// arrivals and bumps the epoch for the next launch - by then every CTA has read it.  All of this happens off the
// critical path, during the first piece's load; nothing is added at kernel end.
__device__ __forceinline__ unsigned flat_epoch_tag(const FlatGeom& g) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(g.ws_ctl);  // low word: arrivals, high word: epoch
    const unsigned long long old = atomicAdd(w, 1ull);
    const unsigned e0 = (unsigned)(old >> 32);
    if ((unsigned)old == gridDim.x - 1) atomicAdd(w, (1ull << 32) - (unsigned long long)gridDim.x);
    return (e0 + 1u) | 0x80000000u;  // never 0: a zero-filled workspace holds no valid record
}

// control block: slot barriers of both rings + the per-piece ring
__host__ __device__ constexpr int flat_ctl_bytes() {
    return kFlatMaxSlots * (4 * 8 + 32) + kFlatNB * (2 * 8 + kFlatConsumerWarps * 16 + 64) + 16;
}
// the dual-norm kernels (MICN_EPI_NORM_ADD_LRELU) carry a second set of per-warp partials behind the control block
__host__ __device__ constexpr int flat_dual_extra_bytes() { return kFlatNB * kFlatConsumerWarps * 16; }
constexpr int kFlatCoefStride = 16;  // floats per entry of the coefficient ring

struct FlatCtx {
    uint32_t dataA, dataB, fullA, emptyA, fullB, emptyB, p1d0, coef0, tagbar;  // shared::cta addresses
    volatile unsigned* tagw;                     // this launch's record tag (written once by the publish warp)
    float* slot_prec;                            // [KA][8]  slab constants of the piece in the A slot (backward)
    float* warp_part;                            // [NB][16][4]
    float* warp_part2;                           // [NB][16][4]  second tensor of the dual-norm kernels (else unused)
    float* coefv;                                // [NB][kFlatCoefStride]
    uint32_t stream_bytes, slot_bytes, slot_bytes_b;  // one stream of a slot; a ring A slot; a ring B slot
};

template <int NSA, int NSB>
__device__ __forceinline__ FlatCtx flat_setup(unsigned char* smem, const FlatGeom& g) {
    FlatCtx c;
    c.stream_bytes = g.slot_vecs * 16u;
    c.slot_bytes = c.stream_bytes * NSA;
    c.slot_bytes_b = c.stream_bytes * NSB;
    c.dataA = smem_u32(smem);
    c.dataB = c.dataA + g.KA * c.slot_bytes;
    unsigned char* ctl = smem + (size_t)g.KA * c.slot_bytes + (size_t)g.KB * c.slot_bytes_b;
    c.fullA = smem_u32(ctl);
    c.emptyA = c.fullA + kFlatMaxSlots * 8;
    c.fullB = c.emptyA + kFlatMaxSlots * 8;
    c.emptyB = c.fullB + kFlatMaxSlots * 8;
    c.p1d0 = c.emptyB + kFlatMaxSlots * 8;
    c.coef0 = c.p1d0 + kFlatNB * 8;
    float* f = reinterpret_cast<float*>(ctl + kFlatMaxSlots * 32 + kFlatNB * 16);
    c.slot_prec = f;
    c.warp_part = c.slot_prec + kFlatMaxSlots * 8;
    c.coefv = c.warp_part + kFlatNB * kFlatConsumerWarps * 4;
    c.tagw = reinterpret_cast<volatile unsigned*>(c.coefv + kFlatNB * kFlatCoefStride);
    c.warp_part2 = reinterpret_cast<float*>(ctl + flat_ctl_bytes());  // (only the dual kernels reserve it)
    c.tagbar = smem_u32(const_cast<unsigned*>(c.tagw) + 2);
    // one barrier per thread (73 of them): a single thread doing all the inits costs ~0.4 us before the first TMA
    {
        const unsigned t = threadIdx.x;
        if (t < kFlatMaxSlots) {
            if (t < g.KA) {
                mbar_init(c.fullA + 8 * t, 1);
                mbar_init(c.emptyA + 8 * t, kFlatConsumerWarps);
            }
            if (t < g.KB) {
                mbar_init(c.fullB + 8 * t, 1);
                mbar_init(c.emptyB + 8 * t, kFlatConsumerWarps);
            }
        } else if (t < kFlatMaxSlots + kFlatNB) {
            mbar_init(c.p1d0 + 8 * (t - kFlatMaxSlots), kFlatConsumerWarps);
            mbar_init(c.coef0 + 8 * (t - kFlatMaxSlots), 1);
        } else if (t == kFlatMaxSlots + kFlatNB) {
            mbar_init(c.tagbar, 1);
        }
        if (t <= kFlatMaxSlots + kFlatNB) fence_mbar_init();
    }
    __syncthreads();
    // everything above touched shared memory only; from here on global memory (workspace epoch, records, tensors)
    pdl_wait();
    pdl_launch_dependents();
    return c;
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

// ring cursors with phase parity: slots advance once per task of their ring, piece-ring entries once per PIECE
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
// in ascending q per lane; `first(a0, b0)` sees record 0 (the common reference) before any fold.  A batch of up to four loads per lane is in flight at once and the missing ones
// are re-polled together, so a slab's record set costs one L2 round trip once everything is published.
template <typename First, typename Fold>
__device__ __forceinline__ void ll_gather(const uint4* recs, unsigned count, unsigned tag, unsigned backoff_ns, int lane,
                                          First first, Fold fold) {
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
        if (q0 == 0) first(__shfl_sync(0xffffffffu, a[0], 0), __shfl_sync(0xffffffffu, b[0], 0));  // record 0
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

// ---- the same records at SYSTEM scope: written into another GPU's memory over NVLink, polled by that GPU (micn_bwd_allreduce)
__device__ __forceinline__ void ll_store_sys(uint4* p, float a, float b, unsigned tag) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag),
                 "r"(__float_as_uint(b)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ll_try_sys(const uint4* p, unsigned tag, float& a, float& b) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    a = __uint_as_float(v.x);
    b = __uint_as_float(v.z);
    return v.y == tag && v.w == tag;
}
// Exchange buffer of one rank: 64-byte header (64-bit word {launch count : CTAs arrived}, as the workspace's), then
// [4 slots][world source ranks][S*C] records, slot = launch count mod 4.  The tag of a launch is its count on this buffer:
// every rank makes the same sequence of calls, so all ranks derive the same tag without talking to each other.  Four
// slots: launch k of this GPU folds slot k (or k-1, lagged mode) at its end; a peer's launch k+1 cannot finish before this
// GPU has emitted all of launch k (or k-1) - i.e. before launch k is running - and its launch k+2 not before launch k+1
// runs here, so the earliest launch of a peer that can overlap this GPU's launch k is k+2, writing slot k+2: never the
// slot (k or k-1) being folded, and k+3 (= k-1 mod 4) cannot have started.
constexpr unsigned kXchgHeader = 64;
constexpr unsigned kXchgSlots = 4;
__device__ __forceinline__ unsigned xchg_epoch_tag(void* local_buf) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(local_buf);
    const unsigned long long old = atomicAdd(w, 1ull);
    const unsigned e0 = (unsigned)(old >> 32);
    if ((unsigned)old == gridDim.x - 1) atomicAdd(w, (1ull << 32) - (unsigned long long)gridDim.x);
    return (e0 + 1u) | 0x80000000u;
}
__device__ __forceinline__ uint4* xchg_rec(void* buf, unsigned tag, unsigned src_rank, unsigned world, unsigned SC, unsigned idx) {
    return reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(buf) + kXchgHeader) +
           ((size_t)(tag & (kXchgSlots - 1u)) * world + src_rank) * SC + idx;
}
// one final (style, channel) entry of this rank -> every rank's buffer (own included)
__device__ __forceinline__ void xchg_emit(const BwdParams& p, unsigned xtag, unsigned idx, float dbeta, float dgamma) {
    // (into this rank's OWN buffer; xchg_forward sends it on to the peers at the kernel's end)
    const unsigned SC = (unsigned)p.num_styles * (unsigned)p.C, me = (unsigned)p.xchg_rank;
    ll_store_sys(xchg_rec(p.xchg_peers[me], xtag, me, (unsigned)p.xchg_world, SC, idx), dbeta, dgamma, xtag);
}
// the same for the S entries of one channel at once, called by a whole warp (one-sample case: the slab sums are final)
__device__ __forceinline__ void xchg_emit_channel(const BwdParams& p, unsigned xtag, unsigned ch, int style, float dbeta,
                                                  float dgamma, int lane) {
    const unsigned S = (unsigned)p.num_styles, C = (unsigned)p.C;
    for (unsigned s = lane; s < S; s += 32) {
        const bool hit = (int)s == style;
        xchg_emit(p, xtag, s * C + ch, hit ? dbeta : 0.f, hit ? dgamma : 0.f);
    }
}
// Forward this rank's records of launch `tag` from its OWN buffer to every peer's, one (entry, peer) pair per lane.  Called
// by the gather warps once their pieces are done: a store into NVLink peer memory holds the issuing warp for microseconds,
// which on the gather warps' critical path (mid-kernel, behind the record of a slab's first piece) cost the whole grid
// 3 us per launch; here it hides behind the consumers' last L normalise tasks.
__device__ __forceinline__ void xchg_forward(const BwdParams& p, unsigned tag, unsigned slot, unsigned nslots, int lane) {
    const unsigned SC = (unsigned)p.num_styles * (unsigned)p.C, world = (unsigned)p.xchg_world, me = (unsigned)p.xchg_rank;
    if ((p.xchg_mode >> 4) & 4) return;  // (bring-up: time the kernel without the peer stores)
    void* mine = p.xchg_peers[me];
    const unsigned total = SC * (world - 1u);
    for (unsigned q = slot * 32u + (unsigned)lane; q < total; q += nslots * 32u) {
        const unsigned idx = q / (world - 1u);
        unsigned r = q - idx * (world - 1u);
        r += r >= me ? 1u : 0u;
        const uint4* src = xchg_rec(mine, tag, me, world, SC, idx);
        float a, b;
        if (!ll_try_sys(src, tag, a, b)) {  // (the channel's own gather warp, in some other CTA, may still be on its way)
            const uint64_t t0 = globaltimer_ns();
            uint32_t spins = 0;
            do {
                __nanosleep(100);
                if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
            } while (!ll_try_sys(src, tag, a, b));
        }
        ll_store_sys(xchg_rec(p.xchg_peers[r], tag, me, world, SC, idx), a, b, tag);
    }
}

// Fold the exchange of launch `tag`: entry idx of every rank sits in THIS rank's buffer (or arrives within the skew between
// the GPUs); the (CTA, warp) pairs share the S*C entries, lane r polls rank r's record, the sum runs in rank order on every
// GPU alike (bit-identical results everywhere).  Called by whole warps; `slot` in [0, nslots) enumerates them.
__device__ __forceinline__ void xchg_fold(const BwdParams& p, unsigned tag, unsigned slot, unsigned nslots, int lane) {
    const unsigned SC = (unsigned)p.num_styles * (unsigned)p.C, world = (unsigned)p.xchg_world;
    void* mine = p.xchg_peers[p.xchg_rank];
    for (unsigned idx = slot; idx < SC; idx += nslots) {
        float a = 0.f, b = 0.f;
        if ((unsigned)lane < world) {
            const uint4* rec = xchg_rec(mine, tag, (unsigned)lane, world, SC, idx);
            if (!ll_try_sys(rec, tag, a, b)) {
                const uint64_t t0 = globaltimer_ns();
                uint32_t spins = 0;
                do {
                    __nanosleep(100);
                    if (((++spins) & 0xffu) == 0 && globaltimer_ns() - t0 > MICN_WAIT_TIMEOUT_NS) __trap();
                } while (!ll_try_sys(rec, tag, a, b));
                if ((p.xchg_mode >> 4) & 64) {  // (bring-up: how long, how often and for which rank do folds wait?)
                    unsigned* dbg = reinterpret_cast<unsigned*>(mine) + 4;
                    atomicMax(dbg, (unsigned)(globaltimer_ns() - t0));
                    atomicAdd(dbg + 1, 1u);
                    atomicAdd(dbg + 2, (unsigned)(globaltimer_ns() - t0));
                    atomicAdd(dbg + 3 + ((unsigned)lane == (unsigned)p.xchg_rank ? 0 : 1), 1u);
                }
            }
        }
        float sa = 0.f, sb = 0.f;
        for (unsigned r = 0; r < world; ++r) {
            sa += __shfl_sync(0xffffffffu, a, r);
            sb += __shfl_sync(0xffffffffu, b, r);
        }
        if (lane == 0) {
            p.dbeta[idx] = sa;
            p.dgamma[idx] = sb;
        }
    }
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

// A consumer thread's walk over the pv vectors of a piece, U vectors at a time (vector v = tid + 512*m).  The
// full batches run without a single bounds check or index recomputation; only the ragged tail is predicated.
template <int U, typename Load, typename Use>
__device__ __forceinline__ void piece_sweep(unsigned pv, unsigned tid, Load load, Use use) {
    constexpr unsigned kStep = U * kFlatConsumerThreads;
    const unsigned full = pv / kStep * kStep;
    unsigned v0 = tid;
    for (; v0 < full; v0 += kStep) {
#pragma unroll
        for (int i = 0; i < U; ++i) load(i, v0 + i * kFlatConsumerThreads);
#pragma unroll
        for (int i = 0; i < U; ++i) use(i, v0 + i * kFlatConsumerThreads);
    }
    if (v0 < pv) {
#pragma unroll
        for (int i = 0; i < U; ++i)
            if (v0 + i * kFlatConsumerThreads < pv) load(i, v0 + i * kFlatConsumerThreads);
#pragma unroll
        for (int i = 0; i < U; ++i)
            if (v0 + i * kFlatConsumerThreads < pv) use(i, v0 + i * kFlatConsumerThreads);
    }
}

// =================================================================================================
// forward
// =================================================================================================
template <typename T, int EPI>
__global__ void __launch_bounds__(kFlatThreads, kFlatCtasPerSm) micn_fwd_flat_kernel(const FwdParams p, const FlatGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VN = VecT<T>::N;
    constexpr bool DUAL = EPI == MICN_EPI_NORM_ADD_LRELU;   // y = lrelu(norm_a(x) + norm_b(x2)): statistics of TWO tensors
    constexpr int NSA = DUAL ? 2 : 1;                       // ring A: x [, x2]
    constexpr int NSB = (EPI == MICN_EPI_ADD_LRELU || DUAL) ? 2 : 1;  // ring B: x [, residual | x2]
    const FlatCtx c = flat_setup<NSA, NSB>(smem, g);
    const unsigned cta = blockIdx.x, G = gridDim.x;
    const unsigned nj = cta < g.T ? (g.T - cta + G - 1) / G : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned C = (unsigned)p.C;

    if (warp == kFlatProducerWarpA || warp == kFlatProducerWarpB) {
        // ------------------------------------------------------------------ producers (A: first touch, B: second touch)
        if (lane == 0) {
            const bool isA = warp == kFlatProducerWarpA;
            const uint64_t pol = isA ? l2_policy_evict_last() : l2_policy_evict_first();
            const unsigned K = isA ? g.KA : g.KB;
            const uint32_t full0 = isA ? c.fullA : c.fullB, empty0 = isA ? c.emptyA : c.emptyB;
            const uint32_t data0 = isA ? c.dataA : c.dataB;
            Ring r{0u, 0u};
            for (unsigned j = 0; j < nj; ++j) {
                const PieceId pc = piece_of(g, j * G + cta);
                const unsigned n = fastdiv(pc.slab, g.divC), ch = pc.slab - n * C;
                const char* src = reinterpret_cast<const char*>(p.x) +
                                  ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) +
                                  (size_t)pc.k * g.PV * 16;
                const uint32_t bytes = pc.pv * 16u, bar = full0 + 8 * r.i;
                if (!isA) {  // the second touch must find the piece in L2: not before its first touch is through
                    const Ring e = entry_of(j);
                    mbar_wait_park(c.p1d0 + 8 * e.i, e.ph);
                }
                if (j >= K) mbar_wait_park(empty0 + 8 * r.i, r.ph ^ 1u);
                flat_trace(g, j, isA ? TR_LOAD : TR_LOAD2);
                const uint32_t dst = data0 + r.i * (isA ? c.slot_bytes : c.slot_bytes_b);
                flat_issue(dst, src, bytes, bar, pol);
                const bool second = isA ? DUAL : NSB == 2;
                if (second) {  // the residual rides in the same slot (read once, from HBM); dual: the second tensor, both touches
                    const char* rsrc = reinterpret_cast<const char*>(DUAL ? p.x2 : p.res) +
                                       ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                    flat_issue(dst + c.stream_bytes, rsrc, bytes, bar, pol);
                }
                mbar_arrive_expect_tx(bar, second ? 2 * bytes : bytes);
                r.next(K);
            }
        }
    } else if (warp >= kFlatPublishWarp0 && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish: warp partials -> piece record
        if (lane == 0) {  // this launch's record tag, shared with the gather warps
            *c.tagw = flat_epoch_tag(g);
            mbar_arrive(c.tagbar);
        }
        __syncwarp();
        const unsigned tag = *c.tagw;
        for (unsigned j = warp - kFlatPublishWarp0; j < nj; j += kFlatPublishWarps) {
            const Ring e = entry_of(j);
            const unsigned gidx = j * G + cta;
            const unsigned k = gidx - fastdiv(gidx, g.divP) * g.P;
            const unsigned pv = piece_vecs(g, k);
            const float nw = lane < kFlatConsumerWarps ? (float)(warp_vecs(pv, lane) * VN) : 0.f;
            mbar_wait_park(c.p1d0 + 8 * e.i, e.ph);
            if (lane == 0) flat_trace(g, j, TR_PUB_BEGIN);
#pragma unroll
            for (int t = 0; t < NSA; ++t) {  // one record per normalised tensor (dual: the second set sits T records further)
                Stat st{0.f, 0.f, 0.f};
                if (lane < kFlatConsumerWarps) {
                    const float4 w = *reinterpret_cast<const float4*>((t ? c.warp_part2 : c.warp_part) +
                                                                     (e.i * kFlatConsumerWarps + lane) * 4);
                    st = stat_from_shifted(w.z, w.x, w.y, nw);
                }
                // fold about the first warp's mean (see the file header)
                const float ref = __shfl_sync(0xffffffffu, st.mean, 0);
                const float d = st.n > 0.f ? st.mean - ref : 0.f;
                const float A = warp_sum(st.n * d), B = warp_sum(fmaf(st.n * d, d, st.m2));
                const float N = (float)(pv * VN);
                const float m = A / N;
                if (lane == 0) ll_store(g.ws_piece + (size_t)t * g.T + gidx, ref + m, fmaxf(B - A * m, 0.f), tag);
            }
            if (lane == 0) flat_trace(g, j, TR_PUB_END);
        }
    } else if (warp >= kFlatGatherWarp0) {
        // ------------------------------------------------------------------ gather: slab records -> coefficients for P2
        mbar_wait_idle(c.tagbar, 0u);
        const unsigned tag = *c.tagw;
        for (unsigned j = warp - kFlatGatherWarp0; j < nj; j += kFlatGatherWarps) {
            const Ring e = entry_of(j);
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = fastdiv(pc.slab, g.divC), ch = pc.slab - n * C;
            // parameter loads first: their latency hides behind everything below
            const int style = load_style(p.styles, n, p.num_styles, p.status);
            float gamma, beta, gammaB = 1.f, betaB = 0.f;
            load_affine(p, style, ch, gamma, beta);
            if (DUAL && p.affine) {
                const float* gp = p.gamma2[0];
                const float* bp = p.beta2[0];
#pragma unroll
                for (int i = 1; i < kMaxStyles; ++i)
                    if (i == style) {
                        gp = p.gamma2[i];
                        bp = p.beta2[i];
                    }
                gammaB = __ldg(gp + ch);
                betaB = __ldg(bp + ch);
            }
            // no polling before this CTA's own piece is through P1: the other CTAs are at the same point
            mbar_wait_idle(c.p1d0 + 8 * e.i, e.ph);
            // ... and let their record stores land; the last L pieces of the CTA are the kernel's tail, where
            // nothing hides a late coefficient any more: poll eagerly there
            __nanosleep(j + g.L >= nj ? g.poll_delay_tail_ns : g.poll_delay_ns);
            if (lane == 0) flat_trace(g, j, TR_GA_BEGIN);
            float ref = 0.f, A = 0.f, B = 0.f;
            ll_gather(g.ws_piece + (size_t)pc.slab * g.P, g.P, tag, g.poll_backoff_ns, lane,
                      [&](float a0, float) { ref = a0; },
                      [&](unsigned q, float a, float b) {
                          const float nq = (float)(piece_vecs(g, q) * VN), d = a - ref;
                          A = fmaf(nq, d, A);
                          B += fmaf(nq * d, d, b);
                      });
            if (lane == 0) flat_trace(g, j, TR_GA_POLLED);
            A = warp_sum(A);
            B = warp_sum(B);
            float refB = 0.f, AB = 0.f, BB = 0.f;
            if (DUAL) {  // the second tensor's records of the same slab
                ll_gather(g.ws_piece + (size_t)g.T + (size_t)pc.slab * g.P, g.P, tag, g.poll_backoff_ns, lane,
                          [&](float a0, float) { refB = a0; },
                          [&](unsigned q, float a, float b) {
                              const float nq = (float)(piece_vecs(g, q) * VN), d = a - refB;
                              AB = fmaf(nq, d, AB);
                              BB += fmaf(nq * d, d, b);
                          });
                AB = warp_sum(AB);
                BB = warp_sum(BB);
            }
            if (lane == 0) {
                const float invM = 1.f / (float)p.M;
                const float m = A * invM;
                const float mean = ref + m;
                const float rstd = 1.f / sqrtf(fmaxf(B - A * m, 0.f) * invM + p.eps);  // biased variance, eps inside the sqrt
                const float a = rstd * gamma;
                // fp32: (x - mean) * a + beta.  16-bit: x * a + (beta - mean * a): one FMA per element.
                float* cf = c.coefv + e.i * kFlatCoefStride;
                *reinterpret_cast<float4*>(cf) =
                    make_float4(sizeof(T) == 4 ? mean : 0.f, a, sizeof(T) == 4 ? beta : fmaf(-mean, a, beta), 0.f);
                if (pc.k == 0 && p.save_mean) {
                    p.save_mean[pc.slab] = mean;
                    p.save_rstd[pc.slab] = rstd;
                }
                if (DUAL) {
                    const float mB = AB * invM;
                    const float meanB = refB + mB;
                    const float rstdB = 1.f / sqrtf(fmaxf(BB - AB * mB, 0.f) * invM + p.eps);
                    const float aB = rstdB * gammaB;
                    *reinterpret_cast<float4*>(cf + 4) =
                        make_float4(sizeof(T) == 4 ? meanB : 0.f, aB, sizeof(T) == 4 ? betaB : fmaf(-meanB, aB, betaB), 0.f);
                    if (pc.k == 0 && p.save_mean2) {
                        p.save_mean2[pc.slab] = meanB;
                        p.save_rstd2[pc.slab] = rstdB;
                    }
                }
                mbar_arrive(c.coef0 + 8 * e.i);
                flat_trace(g, j, TR_GA_END);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ consumers: P1(s), P2(s - L) in order
        Ring ra{0u, 0u}, rb{0u, 0u};
        const float slope = load_slope(p);
        for (unsigned s = 0; s < nj + g.L; ++s) {
            if (s < nj) {
                // ---- P1(s): statistics; the slot goes straight back, the piece stays in L2
                const Ring e = entry_of(s);
                const unsigned gidx = s * G + cta;
                const unsigned pv = piece_vecs(g, gidx - fastdiv(gidx, g.divP) * g.P);
                mbar_wait_park(c.fullA + 8 * ra.i, ra.ph);
                if (tid == 0) flat_trace(g, s, TR_P1_BEGIN);
                const uint32_t base0 = c.dataA + ra.i * c.slot_bytes;
#pragma unroll
                for (int t = 0; t < NSA; ++t) {  // the same sweep per normalised tensor of the slot
                    const uint32_t base = base0 + t * c.stream_bytes;
                    float Kw = 0.f;
                    f32x2 sacc = f2_splat(0.f), qacc = f2_splat(0.f), sacc2 = sacc, qacc2 = sacc;
                    if ((unsigned)warp * 32u < pv) {
                        Kw = first_elem<T>(base + warp * 512);  // shift = the warp's first element of the piece
                        const f32x2 K2 = f2_splat(Kw);
                        uint4 q[4];
                        piece_sweep<4>(
                            pv, tid, [&](int i, unsigned v) { q[i] = lds128(base + v * 16); },
                            [&](int i, unsigned) {
                                f32x2 f[VN / 2];
                                VecT<T>::unpack2(q[i], f);
#pragma unroll
                                for (int k = 0; k < VN / 2; k += 2) {  // two independent accumulator pairs
                                    const f32x2 d0 = f2_sub(f[k], K2), d1 = f2_sub(f[k + 1], K2);
                                    sacc = f2_add(sacc, d0);
                                    sacc2 = f2_add(sacc2, d1);
                                    qacc = f2_fma(d0, d0, qacc);
                                    qacc2 = f2_fma(d1, d1, qacc2);
                                }
                            });
                    }
                    if (t == NSA - 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(c.emptyA + 8 * ra.i);
                    }
                    const float s1 = warp_sum(f2_hsum(f2_add(sacc, sacc2))), s2 = warp_sum(f2_hsum(f2_add(qacc, qacc2)));
                    if (lane == 0)
                        *reinterpret_cast<float4*>((t ? c.warp_part2 : c.warp_part) + (e.i * kFlatConsumerWarps + warp) * 4) =
                            make_float4(s1, s2, Kw, 0.f);
                }
                if (lane == 0) mbar_arrive(c.p1d0 + 8 * e.i);
                if (tid == 0) flat_trace(g, s, TR_P1_END);
                ra.next(g.KA);
            }
            if (s >= g.L) {
                // ---- P2(s - L): normalise + epilogue from the piece's second (L2-served) copy
                const unsigned j = s - g.L;
                const Ring e = entry_of(j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* ydst = reinterpret_cast<char*>(p.y) + goff;
                const uint32_t base = c.dataB + rb.i * c.slot_bytes_b;
                if (tid == 0) flat_trace(g, j, TR_P2_WAIT);
                mbar_wait_park(c.coef0 + 8 * e.i, e.ph);
                const float4 cf = *reinterpret_cast<const float4*>(c.coefv + e.i * kFlatCoefStride);
                const f32x2 sub2 = f2_splat(cf.x), ca2 = f2_splat(cf.y), cb2 = f2_splat(cf.z);
                float4 cfB = make_float4(0.f, 0.f, 0.f, 0.f);
                if (DUAL) cfB = *reinterpret_cast<const float4*>(c.coefv + e.i * kFlatCoefStride + 4);
                const f32x2 subB = f2_splat(cfB.x), caB = f2_splat(cfB.y), cbB = f2_splat(cfB.z);
                mbar_wait_park(c.fullB + 8 * rb.i, rb.ph);
                if (tid == 0) flat_trace(g, j, TR_P2_BEGIN);
                uint4 q[2], rq[2];
                piece_sweep<2>(
                    pc.pv, tid,
                    [&](int i, unsigned v) {
                        q[i] = lds128(base + v * 16);
                        if (NSB == 2) rq[i] = lds128(base + c.stream_bytes + v * 16);
                    },
                    [&](int i, unsigned v) {
                        f32x2 f[VN / 2], rr[VN / 2];
                        VecT<T>::unpack2(q[i], f);
                        if (NSB == 2) VecT<T>::unpack2(rq[i], rr);
#pragma unroll
                        for (int k = 0; k < VN / 2; ++k) {
                            f32x2 o = sizeof(T) == 4 ? f2_fma(f2_sub(f[k], sub2), ca2, cb2) : f2_fma(f[k], ca2, cb2);
                            if (DUAL) o = f2_add(o, sizeof(T) == 4 ? f2_fma(f2_sub(rr[k], subB), caB, cbB) : f2_fma(rr[k], caB, cbB));
                            else if (NSB == 2) o = f2_add(o, rr[k]);
                            if (EPI != MICN_EPI_NONE) {
                                float lo, hi;
                                f2_split(o, lo, hi);
                                lo = lo > 0.f ? lo : lo * slope;
                                hi = hi > 0.f ? hi : hi * slope;
                                o = f2_make(lo, hi);
                            }
                            f[k] = o;
                        }
                        stg_stream(ydst + (size_t)v * 16, VecT<T>::pack2v(f));
                    });
                __syncwarp();
                if (lane == 0) mbar_arrive(c.emptyB + 8 * rb.i);
                if (tid == 0) flat_trace(g, j, TR_P2_END);
                rb.next(g.KB);
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

// the same on a packed pair (FFMA2 evaluates each half exactly like the scalar fmaf of the forward)
template <typename T, int EPI>
__device__ __forceinline__ f32x2 bwd_masked2(f32x2 x, f32x2 gy, f32x2 o, f32x2 mean2, f32x2 a2, f32x2 bq2, float slope) {
    if (EPI == MICN_EPI_NONE) return gy;
    const f32x2 pre = EPI == MICN_EPI_ADD_LRELU ? o : (sizeof(T) == 4 ? f2_fma(f2_sub(x, mean2), a2, bq2) : f2_fma(x, a2, bq2));
    // the reference masks on the ROUNDED output (in-place LeakyReLU on the norm's 16-bit result): an fp16 value in
    // (0, 2^-25] rounds to +0 and counts as "not positive"; bf16 and fp32 share fp32's exponent range
    const float zero = (EPI == MICN_EPI_LRELU && sizeof(T) == 2 && !VecT<T>::kWideExponent) ? 2.98023224e-8f : 0.f;
    float p0, p1, g0, g1;
    f2_split(pre, p0, p1);
    f2_split(gy, g0, g1);
    return f2_make(p0 > zero ? g0 : g0 * slope, p1 > zero ? g1 : g1 * slope);
}

// dual-norm epilogue: the mask is the sign of norm_a(x) + norm_b(x2), the SAME expression the forward evaluated
template <typename T>
__device__ __forceinline__ f32x2 bwd_masked_dual(f32x2 x, f32x2 x2, f32x2 gy, f32x2 mean2, f32x2 a2, f32x2 bq2, f32x2 meanB2,
                                                 f32x2 aB2, f32x2 bqB2, float slope) {
    const f32x2 pa = sizeof(T) == 4 ? f2_fma(f2_sub(x, mean2), a2, bq2) : f2_fma(x, a2, bq2);
    const f32x2 pb = sizeof(T) == 4 ? f2_fma(f2_sub(x2, meanB2), aB2, bqB2) : f2_fma(x2, aB2, bqB2);
    const f32x2 pre = f2_add(pa, pb);
    const float zero = (sizeof(T) == 2 && !VecT<T>::kWideExponent) ? 2.98023224e-8f : 0.f;  // (see bwd_masked2)
    float p0, p1, g0, g1;
    f2_split(pre, p0, p1);
    f2_split(gy, g0, g1);
    return f2_make(p0 > zero ? g0 : g0 * slope, p1 > zero ? g1 : g1 * slope);
}

// (mean, rstd, gamma, beta) of a slab's SECOND norm (dual-norm epilogue)
template <typename P>
__device__ __forceinline__ float4 slab_consts2(const P& p, unsigned slab, unsigned n, unsigned ch) {
    const int style = load_style(p.styles, n, p.num_styles, nullptr);
    float gamma = 1.f, beta = 0.f;
    if (p.affine) {
        const float* gp = p.gamma2[0];
        const float* bp = p.beta2[0];
#pragma unroll
        for (int i = 1; i < kMaxStyles; ++i)
            if (i == style) {
                gp = p.gamma2[i];
                bp = p.beta2[i];
            }
        gamma = __ldg(gp + ch);
        beta = __ldg(bp + ch);
    }
    return make_float4(__ldg(p.save_mean2 + slab), __ldg(p.save_rstd2 + slab), gamma, beta);
}

// (mean, rstd, gamma, beta) of a slab
template <typename P>
__device__ __forceinline__ float4 slab_consts(const P& p, unsigned slab, unsigned n, unsigned ch) {
    const int style = load_style(p.styles, n, p.num_styles, p.status);
    float gamma, beta;
    load_affine(p, style, ch, gamma, beta);
    return make_float4(__ldg(p.save_mean + slab), __ldg(p.save_rstd + slab), gamma, beta);
}

// DS (EPI_LRELU with a device slope only): also accumulate d(prelu)/d(slope) = sum over pre <= 0 of dy * pre; every
// consumer thread carries its share across ALL of the CTA's pieces and the CTA writes one partial at the end
template <typename T, int EPI, bool DS = false>
__global__ void __launch_bounds__(kFlatThreads, kFlatCtasPerSm) micn_bwd_flat_kernel(const BwdParams p, const FlatGeom g) {
    constexpr bool DUAL = EPI == MICN_EPI_NORM_ADD_LRELU;  // two normalised inputs: x, x2 (stream 2), one dy
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU || DUAL) ? 3 : 2;  // x, dy [, act_out | x2]
    constexpr int VN = VecT<T>::N;
    constexpr int U = NS == 3 ? 1 : 2;  // vectors per thread in flight per stream (register budget: 72)
    extern __shared__ __align__(128) unsigned char smem[];
    const FlatCtx c = flat_setup<NS, NS>(smem, g);
    const unsigned cta = blockIdx.x, G = gridDim.x;
    const unsigned nj = cta < g.T ? (g.T - cta + G - 1) / G : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned C = (unsigned)p.C;

    if (warp == kFlatProducerWarpA || warp == kFlatProducerWarpB) {
        // ------------------------------------------------------------------ producers (A: first touch, B: second touch)
        if (lane == 0) {
            const bool isA = warp == kFlatProducerWarpA;
            const uint64_t pol = isA ? l2_policy_evict_last() : l2_policy_evict_first();
            const unsigned K = isA ? g.KA : g.KB;
            const uint32_t full0 = isA ? c.fullA : c.fullB, empty0 = isA ? c.emptyA : c.emptyB;
            const uint32_t data0 = isA ? c.dataA : c.dataB;
            Ring r{0u, 0u};
            for (unsigned j = 0; j < nj; ++j) {
                const PieceId pc = piece_of(g, j * G + cta);
                const unsigned n = fastdiv(pc.slab, g.divC), ch = pc.slab - n * C;
                float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), prB = pr;
                if (isA) pr = slab_consts(p, pc.slab, n, ch);  // for P1: loads issued before the waits
                if (isA && DUAL) prB = slab_consts2(p, pc.slab, n, ch);
                const size_t poff = (size_t)pc.k * g.PV * 16;
                const size_t doff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + poff;
                const char* xsrc = reinterpret_cast<const char*>(p.x) +
                                   ((long long)n * p.x_sN + (long long)ch * p.x_sC) * (long long)sizeof(T) + poff;
                const uint32_t bytes = pc.pv * 16u, bar = full0 + 8 * r.i;
                const uint32_t dst = data0 + r.i * c.slot_bytes;
                if (!isA) {
                    const Ring e = entry_of(j);
                    mbar_wait_park(c.p1d0 + 8 * e.i, e.ph);
                }
                if (j >= K) mbar_wait_park(empty0 + 8 * r.i, r.ph ^ 1u);
                flat_trace(g, j, isA ? TR_LOAD : TR_LOAD2);
                flat_issue(dst, xsrc, bytes, bar, pol);
                flat_issue(dst + c.stream_bytes, reinterpret_cast<const char*>(p.dy) + doff, bytes, bar, pol);
                if (NS == 3)
                    flat_issue(dst + 2 * c.stream_bytes, reinterpret_cast<const char*>(DUAL ? p.x2 : p.act_out) + doff, bytes,
                               bar, pol);
                if (isA) *reinterpret_cast<float4*>(c.slot_prec + r.i * 8) = pr;  // visible through the barrier
                if (isA && DUAL) *reinterpret_cast<float4*>(c.slot_prec + r.i * 8 + 4) = prB;
                mbar_arrive_expect_tx(bar, bytes * NS);
                r.next(K);
            }
        }
    } else if (warp >= kFlatPublishWarp0 && warp < kFlatGatherWarp0) {
        // ------------------------------------------------------------------ publish
        if (lane == 0) {  // this launch's record tag, shared with the gather warps
            *c.tagw = flat_epoch_tag(g);
            mbar_arrive(c.tagbar);
        }
        __syncwarp();
        const unsigned tag = *c.tagw;
        for (unsigned j = warp - kFlatPublishWarp0; j < nj; j += kFlatPublishWarps) {
            const Ring e = entry_of(j);
            mbar_wait_park(c.p1d0 + 8 * e.i, e.ph);
            if (lane == 0) flat_trace(g, j, TR_PUB_BEGIN);
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (lane < kFlatConsumerWarps) {
                const float4 w = *reinterpret_cast<const float4*>(c.warp_part + (e.i * kFlatConsumerWarps + lane) * 4);
                s1 = w.x;
                s2 = w.y;
                s3 = w.z;
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (DUAL) s3 = warp_sum(s3);
            if (lane == 0) ll_store(g.ws_piece + (j * G + cta), s1, s2, tag);
            if (DUAL && lane == 0) ll_store(g.ws_piece + (size_t)g.T + (j * G + cta), s3, 0.f, tag);
            if (lane == 0) flat_trace(g, j, TR_PUB_END);
        }
    } else if (warp >= kFlatGatherWarp0) {
        // ------------------------------------------------------------------ gather
        const float invM = 1.f / (float)p.M;
        mbar_wait_idle(c.tagbar, 0u);
        const unsigned tag = *c.tagw;
        // peer exchange: its launch tag is fetched HERE, by the first gather warp (idle until the first piece is through P1),
        // not by the publish warp, whose first record is on the critical path of every CTA of the slab
        const bool xchg = p.xchg_world > 1;
        if (xchg) {
            if (warp == kFlatGatherWarp0 && lane == 0)
                c.tagw[1] = ((p.xchg_mode >> 4) & 2) ? (tag | 0x80000000u) : xchg_epoch_tag(p.xchg_peers[p.xchg_rank]);
            bar_sync(2, kFlatGatherWarps * 32);
        }
        const unsigned xtag = xchg ? c.tagw[1] : 0u;
        for (unsigned j = warp - kFlatGatherWarp0; j < nj; j += kFlatGatherWarps) {
            const Ring e = entry_of(j);
            const PieceId pc = piece_of(g, j * G + cta);
            const unsigned n = fastdiv(pc.slab, g.divC), ch = pc.slab - n * C;
            const float4 pr = slab_consts(p, pc.slab, n, ch);  // latency hides behind the wait below
            const float mean = pr.x, rstd = pr.y, gamma = pr.z, beta = pr.w;
            float4 prB = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DUAL) prB = slab_consts2(p, pc.slab, n, ch);
            // no polling before this CTA's own piece is through P1 (the others are at the same point)
            mbar_wait_idle(c.p1d0 + 8 * e.i, e.ph);
            __nanosleep(j + g.L >= nj ? g.poll_delay_tail_ns : g.poll_delay_ns);  // let the record stores land (tail: eager)
            if (lane == 0) flat_trace(g, j, TR_GA_BEGIN);
            float S1 = 0.f, S2 = 0.f;
            ll_gather(g.ws_piece + (size_t)pc.slab * g.P, g.P, tag, g.poll_backoff_ns, lane, [](float, float) {},
                      [&](unsigned, float a, float b) {
                          S1 += a;
                          S2 += b;
                      });
            if (lane == 0) flat_trace(g, j, TR_GA_POLLED);
            S1 = warp_sum(S1);
            S2 = warp_sum(S2);
            float S2B = 0.f;
            if (DUAL) {  // sum g * (x2 - mean2): the second record set of the same slab
                ll_gather(g.ws_piece + (size_t)g.T + (size_t)pc.slab * g.P, g.P, tag, g.poll_backoff_ns, lane, [](float, float) {},
                          [&](unsigned, float a, float) { S2B += a; });
                S2B = warp_sum(S2B);
            }
            const float a = rstd * gamma;
            const float S2r = S2 * rstd;  // sum g * xhat
            const float aB = prB.y * prB.z, S2Br = S2B * prB.y;
            if (lane == 0) {
                // dx = a*g - a*S1/M - a*rstd*(S2r/M)*(x - mean)
                const float B1 = -a * S2r * invM * rstd, B0c = -a * S1 * invM;
                float* cf = c.coefv + e.i * kFlatCoefStride;
                *reinterpret_cast<float4*>(cf) = make_float4(a, B1, sizeof(T) == 4 ? B0c : fmaf(-B1, mean, B0c), mean);
                cf[4] = sizeof(T) == 4 ? beta : fmaf(-mean, a, beta);
                if (DUAL) {
                    const float meanB = prB.x, rstdB = prB.y;
                    const float B1b = -aB * S2Br * invM * rstdB, B0cb = -aB * S1 * invM;
                    *reinterpret_cast<float4*>(cf + 8) =
                        make_float4(aB, B1b, sizeof(T) == 4 ? B0cb : fmaf(-B1b, meanB, B0cb), meanB);
                    cf[12] = sizeof(T) == 4 ? prB.w : fmaf(-meanB, aB, prB.w);
                }
                mbar_arrive(c.coef0 + 8 * e.i);
                flat_trace(g, j, TR_GA_END);
            }
            if (pc.k == 0 && p.dgamma) {
                const bool two = DUAL && p.dgamma2 != nullptr;
                if (p.N == 1) {
                    // one sample: this slab's sums ARE the gradients of its style's row
                    const int style = load_style(p.styles, 0, p.num_styles, nullptr);
                    if (xchg)  // final for this rank: straight into every rank's exchange buffer (folded later)
                        xchg_emit_channel(p, xtag, ch, style, S1, S2r, lane);
                    for (int s = lane; s < p.num_styles && !xchg; s += 32) {
                        p.dbeta[(size_t)s * C + ch] = s == style ? S1 : 0.f;
                        p.dgamma[(size_t)s * C + ch] = s == style ? S2r : 0.f;
                        if (two) {
                            p.dbeta2[(size_t)s * C + ch] = s == style ? S1 : 0.f;
                            p.dgamma2[(size_t)s * C + ch] = s == style ? S2Br : 0.f;
                        }
                    }
                } else {
                    const size_t slabs = (size_t)p.N * C;
                    if (lane == 0) ll_store(g.ws_slab + pc.slab, S1, S2r, tag);
                    if (two && lane == 0) ll_store(g.ws_slab + slabs + pc.slab, S2Br, 0.f, tag);
                    if (n == (unsigned)p.N - 1) {
                        // last sample of this channel: fold every sample's record per style, fixed order
                        for (int s = 0; s < p.num_styles; ++s) {
                            float ab = 0.f, ag = 0.f, ag2 = 0.f;
                            for (unsigned nn = lane; nn < (unsigned)p.N; nn += 32) {
                                float ra, rb, rc = 0.f, rd;
                                ll_wait1(g.ws_slab + (size_t)nn * C + ch, tag, g.poll_backoff_ns, ra, rb);
                                if (two) ll_wait1(g.ws_slab + slabs + (size_t)nn * C + ch, tag, g.poll_backoff_ns, rc, rd);
                                if (load_style(p.styles, nn, p.num_styles, nullptr) == s) {
                                    ab += ra;
                                    ag += rb;
                                    ag2 += rc;
                                }
                            }
                            ab = warp_sum(ab);
                            ag = warp_sum(ag);
                            if (two) ag2 = warp_sum(ag2);
                            if (lane == 0 && xchg) {
                                xchg_emit(p, xtag, (unsigned)s * C + ch, ab, ag);
                            } else if (lane == 0) {
                                p.dbeta[(size_t)s * C + ch] = ab;
                                p.dgamma[(size_t)s * C + ch] = ag;
                                if (two) {
                                    p.dbeta2[(size_t)s * C + ch] = ab;
                                    p.dgamma2[(size_t)s * C + ch] = ag2;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (xchg && p.dgamma)  // this launch's records: own buffer -> every peer's
            xchg_forward(p, xtag, cta * kFlatGatherWarps + (warp - kFlatGatherWarp0), G * kFlatGatherWarps, lane);
        if (xchg && p.dgamma && (p.xchg_mode & 15) == 1)  // synchronous mode: this call's own exchange, folded at the kernel's end
            xchg_fold(p, xtag, cta * kFlatGatherWarps + (warp - kFlatGatherWarp0), G * kFlatGatherWarps, lane);
        if (xchg && p.dgamma && (p.xchg_mode & 15) == 2 && (xtag & 0x7fffffffu) > 1u && !((p.xchg_mode >> 4) & 1))
            // lagged mode: dgamma / dbeta receive the all-reduced gradients of the PREVIOUS call - its records arrived a
            // whole kernel ago, nothing to wait for - while this call's records travel.  Done here, by gather warps that
            // have nothing left to do while the consumers drain the last L pieces: at the kernel's START the same fold
            // cost 2 us per step (measured at 2 GPUs: 80.7 against 78.6 us without it).
            xchg_fold(p, xtag - 1u, cta * kFlatGatherWarps + (warp - kFlatGatherWarp0), G * kFlatGatherWarps, lane);
    } else {
        // ------------------------------------------------------------------ consumers: P1(s), P2(s - L) in order
        Ring ra{0u, 0u}, rb{0u, 0u};
        const float slope = load_slope(p);
        f32x2 ds2 = f2_splat(0.f);
        const uint32_t sb = c.stream_bytes;
        for (unsigned s = 0; s < nj + g.L; ++s) {
            if (s < nj) {
                // ---- P1(s)
                const Ring e = entry_of(s);
                const unsigned gidx = s * G + cta;
                const unsigned pv = piece_vecs(g, gidx - fastdiv(gidx, g.divP) * g.P);
                mbar_wait_park(c.fullA + 8 * ra.i, ra.ph);
                if (tid == 0) flat_trace(g, s, TR_P1_BEGIN);
                const uint32_t base = c.dataA + ra.i * c.slot_bytes;
                const float4 pr = *reinterpret_cast<const float4*>(c.slot_prec + ra.i * 8);
                const float mean = pr.x, ca = pr.y * pr.z;
                const float bq = sizeof(T) == 4 ? pr.w : fmaf(-mean, ca, pr.w);
                const f32x2 mean2 = f2_splat(mean), ca2 = f2_splat(ca), bq2 = f2_splat(bq);
                f32x2 meanB2 = 0ull, caB2 = 0ull, bqB2 = 0ull, s3a = f2_splat(0.f), s3b = s3a;
                if (DUAL) {
                    const float4 prB = *reinterpret_cast<const float4*>(c.slot_prec + ra.i * 8 + 4);
                    const float caB = prB.y * prB.z;
                    meanB2 = f2_splat(prB.x);
                    caB2 = f2_splat(caB);
                    bqB2 = f2_splat(sizeof(T) == 4 ? prB.w : fmaf(-prB.x, caB, prB.w));
                }
                f32x2 s1a = f2_splat(0.f), s1b = s1a, s2a = s1a, s2b = s1a;
                uint4 qx[U], qg[U], qo[U];
                piece_sweep<U>(
                    pv, tid,
                    [&](int i, unsigned v) {
                        qx[i] = lds128(base + v * 16);
                        qg[i] = lds128(base + sb + v * 16);
                        if (NS == 3) qo[i] = lds128(base + 2 * sb + v * 16);
                    },
                    [&](int i, unsigned) {
                        f32x2 xf[VN / 2], gf[VN / 2], of[VN / 2];
                        VecT<T>::unpack2(qx[i], xf);
                        VecT<T>::unpack2(qg[i], gf);
                        if (NS == 3) VecT<T>::unpack2(qo[i], of);
#pragma unroll
                        for (int k = 0; k < VN / 2; k += 2) {
                            if (DS) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const f32x2 pre = sizeof(T) == 4 ? f2_fma(f2_sub(xf[k + h], mean2), ca2, bq2)
                                                                     : f2_fma(xf[k + h], ca2, bq2);
                                    float p0, p1;
                                    f2_split(pre, p0, p1);
                                    ds2 = f2_fma(gf[k + h], f2_make(p0 > 0.f ? 0.f : p0, p1 > 0.f ? 0.f : p1), ds2);
                                }
                            }
                            f32x2 g0, g1;
                            if (DUAL) {
                                g0 = bwd_masked_dual<T>(xf[k], of[k], gf[k], mean2, ca2, bq2, meanB2, caB2, bqB2, slope);
                                g1 = bwd_masked_dual<T>(xf[k + 1], of[k + 1], gf[k + 1], mean2, ca2, bq2, meanB2, caB2, bqB2, slope);
                                s3a = f2_fma(g0, f2_sub(of[k], meanB2), s3a);
                                s3b = f2_fma(g1, f2_sub(of[k + 1], meanB2), s3b);
                            } else {
                                g0 = bwd_masked2<T, EPI>(xf[k], gf[k], NS == 3 ? of[k] : 0ull, mean2, ca2, bq2, slope);
                                g1 = bwd_masked2<T, EPI>(xf[k + 1], gf[k + 1], NS == 3 ? of[k + 1] : 0ull, mean2, ca2, bq2, slope);
                            }
                            s1a = f2_add(s1a, g0);
                            s1b = f2_add(s1b, g1);
                            s2a = f2_fma(g0, f2_sub(xf[k], mean2), s2a);
                            s2b = f2_fma(g1, f2_sub(xf[k + 1], mean2), s2b);
                        }
                    });
                __syncwarp();
                if (lane == 0) mbar_arrive(c.emptyA + 8 * ra.i);
                const float s1 = warp_sum(f2_hsum(f2_add(s1a, s1b))), s2 = warp_sum(f2_hsum(f2_add(s2a, s2b)));
                const float s3 = DUAL ? warp_sum(f2_hsum(f2_add(s3a, s3b))) : 0.f;
                if (lane == 0) {
                    *reinterpret_cast<float4*>(c.warp_part + (e.i * kFlatConsumerWarps + warp) * 4) = make_float4(s1, s2, s3, 0.f);
                    mbar_arrive(c.p1d0 + 8 * e.i);
                }
                if (tid == 0) flat_trace(g, s, TR_P1_END);
                ra.next(g.KA);
            }
            if (s >= g.L) {
                // ---- P2(s - L)
                const unsigned j = s - g.L;
                const Ring e = entry_of(j);
                const PieceId pc = piece_of(g, j * G + cta);
                const size_t goff = ((size_t)pc.slab * (size_t)p.M) * sizeof(T) + (size_t)pc.k * g.PV * 16;
                char* dxdst = reinterpret_cast<char*>(p.dx) + goff;
                char* drdst = NS == 3 ? reinterpret_cast<char*>(DUAL ? p.dx2 : p.dres) + goff : nullptr;
                const uint32_t base = c.dataB + rb.i * c.slot_bytes;
                if (tid == 0) flat_trace(g, j, TR_P2_WAIT);
                mbar_wait_park(c.coef0 + 8 * e.i, e.ph);
                const float* cf = c.coefv + e.i * kFlatCoefStride;
                const float4 cq = *reinterpret_cast<const float4*>(cf);
                const f32x2 A2 = f2_splat(cq.x), B12 = f2_splat(cq.y), B02 = f2_splat(cq.z), mean2 = f2_splat(cq.w),
                            bq2 = f2_splat(cf[4]);
                f32x2 AB2 = 0ull, B1B2 = 0ull, B0B2 = 0ull, meanB2 = 0ull, bqB2 = 0ull;
                if (DUAL) {
                    const float4 cb = *reinterpret_cast<const float4*>(cf + 8);
                    AB2 = f2_splat(cb.x);
                    B1B2 = f2_splat(cb.y);
                    B0B2 = f2_splat(cb.z);
                    meanB2 = f2_splat(cb.w);
                    bqB2 = f2_splat(cf[12]);
                }
                mbar_wait_park(c.fullB + 8 * rb.i, rb.ph);
                if (tid == 0) flat_trace(g, j, TR_P2_BEGIN);
                uint4 qx[U], qg[U], qo[U];
                piece_sweep<U>(
                    pc.pv, tid,
                    [&](int i, unsigned v) {
                        qx[i] = lds128(base + v * 16);
                        qg[i] = lds128(base + sb + v * 16);
                        if (NS == 3) qo[i] = lds128(base + 2 * sb + v * 16);
                    },
                    [&](int i, unsigned v) {
                        f32x2 xf[VN / 2], gf[VN / 2], of[VN / 2];
                        VecT<T>::unpack2(qx[i], xf);
                        VecT<T>::unpack2(qg[i], gf);
                        if (NS == 3) VecT<T>::unpack2(qo[i], of);
#pragma unroll
                        for (int k = 0; k < VN / 2; ++k) {
                            const f32x2 gg = DUAL ? bwd_masked_dual<T>(xf[k], of[k], gf[k], mean2, A2, bq2, meanB2, AB2, bqB2, slope)
                                                  : bwd_masked2<T, EPI>(xf[k], gf[k], NS == 3 ? of[k] : 0ull, mean2, A2, bq2, slope);
                            // (dual: the second store carries d(x2) instead of d(residual))
                            gf[k] = DUAL ? f2_fma(AB2, gg, f2_fma(B1B2, sizeof(T) == 4 ? f2_sub(of[k], meanB2) : of[k], B0B2)) : gg;
                            xf[k] = f2_fma(A2, gg, f2_fma(B12, sizeof(T) == 4 ? f2_sub(xf[k], mean2) : xf[k], B02));
                        }
                        stg_stream(dxdst + (size_t)v * 16, VecT<T>::pack2v(xf));
                        if (NS == 3) stg_stream(drdst + (size_t)v * 16, VecT<T>::pack2v(gf));
                    });
                __syncwarp();
                if (lane == 0) mbar_arrive(c.emptyB + 8 * rb.i);
                if (tid == 0) flat_trace(g, j, TR_P2_END);
                rb.next(g.KB);
            }
        }
        if (DS) {  // one partial per CTA: warp sums in a fixed order (the control ring's partial area is free now)
            const float w = warp_sum(f2_hsum(ds2));
            if (lane == 0) c.warp_part[warp] = w;
            bar_sync(1, kFlatConsumerThreads);
            if (tid == 0) {
                float tot = 0.f;
                for (int i = 0; i < kFlatConsumerWarps; ++i) tot += c.warp_part[i];
                p.dslope[cta] = tot;
            }
        }
    }
}

// stand-alone fold of the LATEST call's exchange (micn_allreduce_fold): the tag is the buffer's launch count
__global__ void micn_xchg_fold_kernel(const BwdParams p) {
    const unsigned long long hdr = *reinterpret_cast<volatile unsigned long long*>(p.xchg_peers[p.xchg_rank]);
    const unsigned e = (unsigned)(hdr >> 32);
    if (e == 0u) return;  // no call has been made on this buffer yet
    xchg_fold(p, e | 0x80000000u, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5),
              (int)(threadIdx.x & 31));
}

// what the host-side planner needs to know about this shape
struct Traits {
    using Geom = FlatGeom;
    static constexpr int kThreads = kFlatThreads, kCtasPerSm = kFlatCtasPerSm, kConsumerThreads = kFlatConsumerThreads,
                         kMaxSlots = kFlatMaxSlots, kMaxLag = kFlatMaxLag, kMaxPieces = kFlatMaxPieces,
                         kMinPieceVecs = kFlatMinPieceVecs, kCtlBytes = flat_ctl_bytes(),
                         kDualExtraBytes = flat_dual_extra_bytes();
};

}  // namespace MICN_FLAT_NS
}  // namespace micn
