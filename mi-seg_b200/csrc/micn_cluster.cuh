// micn_cluster.cuh - cluster-resident instance_cond forward / backward for large slabs (sm_100a).
//
// One thread-block CLUSTER (1..16 CTAs, co-scheduled on one GPC) owns one (n, c) slab at a time and
// walks the slabs persistently.  Each CTA owns a contiguous 1/CS share of the slab and keeps as
// much of it as fits in its ~224 KB of shared memory:
//
//   producer warp (1 elected lane)   1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of
//                                    8 KB chunks into a ring of S slots, running up to S chunks
//                                    ahead - across the pass-1/pass-2 boundary and into the NEXT
//                                    slab - gated only by per-slot "empty" mbarriers.
//   16 consumer warps                work unit = a 2 KB quarter of a chunk (4 warp-wide 128-bit
//                                    rounds), dealt round-robin to the warps, so a warp pays one
//                                    barrier wait / release per 4 vectors.
//        pass 1  statistics: per-thread shifted sums in fp32 for every element type.
//                Partials: warp shuffle -> CTA -> pushed into every peer CTA's shared memory
//                (st.shared::cluster + remote mbarrier arrive), merged in rank order with Chan's
//                formula (bit-identical in every CTA, no atomics).
//        pass 2  normalise / epilogue straight out of shared memory, 128-bit streaming stores.
//
// Chunks that stayed in shared memory ("resident") are consumed twice without touching HBM or L2
// again; only when a CTA's share exceeds its S slots are the oldest chunks re-fetched (they were
// loaded with an L2 evict_last hint, the re-fetch hits L2).  With full residency each voxel crosses
// HBM exactly once per tensor: forward 2*E*s, backward 3*E*s bytes - the algorithmic minimum
// (SURVEY.md section 8d).
//
// Reference semantics: networks/norms/conditional_instance_norm.py:59-60 (+ ATen instance_norm),
// epilogues networks/blocks/dynunet_block.py:107-125.
#pragma once

#include "micn_common.cuh"

namespace micn {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;   // 512
constexpr int kClusterThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kChunkVecs = 512;                         // 16-byte vectors per TMA chunk
constexpr int kChunkBytes = kChunkVecs * 16;            // 8 KB
constexpr int kUnitVecs = 128;                          // work unit: 4 rounds of 32 lanes x 16 B
constexpr int kUnitBytes = kUnitVecs * 16;              // 2 KB
constexpr int kUnitsPerChunk = kChunkVecs / kUnitVecs;  // 4
constexpr int kMaxCluster = 16;
constexpr int kConsumerBarrier = 1;  // named barrier id for the 16 consumer warps

// bytes of shared memory besides the data slots
__host__ __device__ constexpr int cluster_smem_overhead(int S) {
    return S * 16 /*full+empty*/ + 16 /*2 partial-ready barriers*/ + 2 * kMaxCluster * 16 /*peer partials*/ +
           kConsumerWarps * 16 /*warp partials*/ + 16 /*flags*/;
}
__host__ __device__ constexpr int cluster_smem_bytes(int S, int NS) {
    return S * NS * kChunkBytes + cluster_smem_overhead(S);
}

struct Cursor {
    uint32_t slot, phase;
};
__device__ __forceinline__ Cursor cursor_add(Cursor c, uint32_t delta, uint32_t S) {
    c.slot += delta;
    while (c.slot >= S) {
        c.slot -= S;
        c.phase ^= 1u;
    }
    return c;
}

struct ClusterCtx {
    uint32_t data0, full0, empty0, pr0;  // shared::cta addresses
    float* peer_part;                    // [2][kMaxCluster][4]
    float* warp_part;                    // [kConsumerWarps][4]
    int* flags;
    uint32_t rank, CS, cid, G;
    long long v0, nv;  // this CTA's share of every slab, in 16-byte vectors
    int nchunks, r0;   // chunks per slab share; first resident chunk (= number of re-fetched chunks)
    int nunits;        // nchunks * kUnitsPerChunk (trailing units may be partial or empty)
};

template <int NS>
__device__ __forceinline__ ClusterCtx cluster_setup(unsigned char* smem, int S, long long vecs_per_slab) {
    ClusterCtx c;
    c.data0 = smem_u32(smem);
    unsigned char* ctl = smem + (size_t)S * NS * kChunkBytes;
    c.full0 = smem_u32(ctl);
    c.empty0 = c.full0 + S * 8;
    c.pr0 = c.empty0 + S * 8;
    c.peer_part = reinterpret_cast<float*>(ctl + S * 16 + 16);
    c.warp_part = c.peer_part + 2 * kMaxCluster * 4;
    c.flags = reinterpret_cast<int*>(c.warp_part + kConsumerWarps * 4);
    c.rank = cluster_ctarank();
    c.CS = cluster_nctarank();
    c.cid = cluster_id_x();
    c.G = nclusters_x();
    const long long base = vecs_per_slab / c.CS, rem = vecs_per_slab % c.CS;
    c.v0 = c.rank * base + (c.rank < rem ? c.rank : rem);
    c.nv = base + (c.rank < rem ? 1 : 0);
    c.nchunks = (int)((c.nv + kChunkVecs - 1) / kChunkVecs);
    c.r0 = c.nchunks > S ? c.nchunks - S : 0;
    c.nunits = c.nchunks * kUnitsPerChunk;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(c.full0 + 8 * i, 1);
            mbar_init(c.empty0 + 8 * i, kUnitsPerChunk);
        }
        mbar_init(c.pr0, c.CS);
        mbar_init(c.pr0 + 8, c.CS);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_sync_all();  // peers' barriers exist before anyone arrives on them remotely
    return c;
}

__device__ __forceinline__ int chunk_vecs(const ClusterCtx& c, int j) {
    const long long left = c.nv - (long long)j * kChunkVecs;
    return left < kChunkVecs ? (int)left : kChunkVecs;
}

// Producer: streams, per slab, chunks 0..nchunks-1 (pass 1) then 0..r0-1 again (pass-2 re-fetch of
// the chunks that were evicted from the ring), through slots q % S.
template <int NS, typename SrcFn>
__device__ __forceinline__ void cluster_producer(const ClusterCtx& c, int S, long long num_slabs, SrcFn src_of) {
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    Cursor cur{0u, 0u};
    long long q = 0;
    for (long long slab = c.cid; slab < num_slabs; slab += c.G) {
        const char* src[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) src[s] = src_of(slab, s) + c.v0 * 16;
        for (int pass = 0; pass < 2; ++pass) {
            const int jend = pass == 0 ? c.nchunks : c.r0;
            for (int j = 0; j < jend; ++j) {
                if (q >= S) mbar_wait(c.empty0 + 8 * cur.slot, cur.phase ^ 1u);
                const uint32_t bytes = (uint32_t)chunk_vecs(c, j) * 16u;
                const uint32_t bar = c.full0 + 8 * cur.slot;
                mbar_arrive_expect_tx(bar, bytes * NS);
                const uint64_t pol = (pass == 0 && j < c.r0) ? pol_keep : pol_stream;
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    tma_load_1d(c.data0 + (cur.slot * NS + s) * kChunkBytes, src[s] + (size_t)j * kChunkBytes, bytes,
                                bar, pol);
                ++q;
                cur = cursor_add(cur, 1, S);
            }
        }
    }
}

// Cluster-wide exchange of one small per-CTA record (K floats, K <= 3): every CTA pushes its record
// into slot [parity][rank] of every peer and arrives on the peer's "partials ready" barrier; after
// the wait each CTA holds all CS records locally.  Called by all consumer threads.
template <int K>
__device__ __forceinline__ void cluster_exchange(const ClusterCtx& c, int it, const float* rec /*warp0 lanes*/) {
    const int par = it & 1;
    if ((threadIdx.x >> 5) == 0) {
        const uint32_t lane = threadIdx.x & 31;
        if (lane < c.CS) {
            const uint32_t local = smem_u32(c.peer_part + (par * kMaxCluster + c.rank) * 4);
            const uint32_t dst = mapa(local, lane);
#pragma unroll
            for (int k = 0; k < K; ++k) st_cluster_f32(dst + 4 * k, rec[k]);
            mbar_arrive_remote(mapa(c.pr0 + 8 * par, lane));
        }
    }
    mbar_wait_cluster(c.pr0 + 8 * par, (uint32_t)(it >> 1) & 1u);
}

__device__ __forceinline__ void unit_release(uint32_t empty_bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar);
}

// ---------------------------------------------------------------------------------------------
// packed 16-bit helpers + tensor-core sums
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Half2Ops;

template <>
struct Half2Ops<__nv_bfloat16> {
    using T2 = __nv_bfloat162;
    static constexpr uint32_t kOnes = 0x3F803F80u;
    __device__ __forceinline__ static T2 as2(uint32_t u) { return *reinterpret_cast<T2*>(&u); }
    __device__ __forceinline__ static uint32_t bits(T2 h) { return *reinterpret_cast<uint32_t*>(&h); }
    __device__ __forceinline__ static uint32_t sub2(uint32_t a, uint32_t b) { return bits(__hsub2(as2(a), as2(b))); }
    __device__ __forceinline__ static uint32_t mul2(uint32_t a, uint32_t b) { return bits(__hmul2(as2(a), as2(b))); }
    __device__ __forceinline__ static uint32_t gt2_mask(uint32_t a, uint32_t b) { return __hgt2_mask(as2(a), as2(b)); }
    __device__ __forceinline__ static uint32_t bcast_rn(float f) { return bits(__float2bfloat162_rn(f)); }
    __device__ __forceinline__ static uint32_t bcast_rd(float f) {
        const __nv_bfloat16 h = __float2bfloat16_rd(f);
        return bits(__halves2bfloat162(h, h));
    }
    __device__ __forceinline__ static float low_to_float(uint32_t u) { return __uint_as_float(u << 16); }
    __device__ __forceinline__ static void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

template <>
struct Half2Ops<__half> {
    using T2 = __half2;
    static constexpr uint32_t kOnes = 0x3C003C00u;
    __device__ __forceinline__ static T2 as2(uint32_t u) { return *reinterpret_cast<T2*>(&u); }
    __device__ __forceinline__ static uint32_t bits(T2 h) { return *reinterpret_cast<uint32_t*>(&h); }
    __device__ __forceinline__ static uint32_t sub2(uint32_t a, uint32_t b) { return bits(__hsub2(as2(a), as2(b))); }
    __device__ __forceinline__ static uint32_t mul2(uint32_t a, uint32_t b) { return bits(__hmul2(as2(a), as2(b))); }
    __device__ __forceinline__ static uint32_t gt2_mask(uint32_t a, uint32_t b) { return __hgt2_mask(as2(a), as2(b)); }
    __device__ __forceinline__ static uint32_t bcast_rn(float f) { return bits(__float2half2_rn(f)); }
    __device__ __forceinline__ static uint32_t bcast_rd(float f) {
        const __half h = __float2half_rd(f);
        return bits(__halves2half2(h, h));
    }
    __device__ __forceinline__ static float low_to_float(uint32_t u) {
        return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu)));
    }
    __device__ __forceinline__ static void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

// With A = the warp's 32 vectors read as a 16x16 tile (thread (g = lane/4, t = lane%4) holds rows g and
// g+8), A x {a0,a2}-as-B puts sum_k A[m][k]*B'[m][k] for rows 0..7 on the diagonal of D (register
// c[g&1] of the lanes with t == g/2) and A x {a1,a3}-as-B does the same for rows 8..15 (register
// c[2 + (g&1)] of the same lanes).  A x ones replicates every row sum across the 8 columns: the
// t == 0 lanes hold them in c[0] (row g) and c[2] (row g+8).
__device__ __forceinline__ float mma_diag(const float (&d1)[4], const float (&d2)[4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float v = 0.f;
    if (t == (g >> 1)) v = (g & 1) ? d1[1] + d2[3] : d1[0] + d2[2];
    return v;
}
__device__ __forceinline__ float mma_rowsum(const float (&ds)[4]) {
    return ((threadIdx.x & 3) == 0) ? ds[0] + ds[2] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// forward statistics accumulators (one per warp; finish() returns the warp's Stat in every lane)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct FwdStats {  // per-thread shifted sums in fp32 (any element type), Chan-merged across the warp
    float K, s1a, s1b, s2a, s2b, n;
    bool have;
    __device__ __forceinline__ void init() {
        K = s1a = s1b = s2a = s2b = n = 0.f;
        have = false;
    }
    __device__ __forceinline__ void add(uint4 v, bool ok) {
        if (!ok) return;
        constexpr int VN = VecT<T>::N;
        float f[VN];
        VecT<T>::unpack(v, f);
        if (!have) {
            K = f[0];
            have = true;
        }
#pragma unroll
        for (int e = 0; e < VN; e += 2) {
            const float d0 = f[e] - K, d1 = f[e + 1] - K;
            s1a += d0;
            s1b += d1;
            s2a = fmaf(d0, d0, s2a);
            s2b = fmaf(d1, d1, s2b);
        }
        n += (float)VN;
    }
    __device__ __forceinline__ Stat finish() { return stat_warp_reduce(stat_from_shifted(K, s1a + s1b, s2a + s2b, n)); }
};

// pass-2 math on one vector
template <typename T, int EPI>
__device__ __forceinline__ uint4 fwd_apply(const uint4& xv, const uint4& rv, float sub, float a, float b, float slope) {
    constexpr int VN = VecT<T>::N;
    float f[VN], r[VN];
    VecT<T>::unpack(xv, f);
    if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(rv, r);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
        // fp32: (x - mean) * a + beta.  16-bit: x * a + (beta - mean * a) (sub == 0): one FMA, the fold
        // costs ~|mean|/std fp32 ulps, far below the output's own 8/11-bit rounding.
        float v = sizeof(T) == 4 ? fmaf(f[e] - sub, a, b) : fmaf(f[e], a, b);
        if (EPI == MICN_EPI_ADD_LRELU) v += r[e];
        if (EPI != MICN_EPI_NONE) v = v > 0.f ? v : v * slope;
        f[e] = v;
    }
    return VecT<T>::pack(f);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T, int EPI>
__global__ void __launch_bounds__(kClusterThreads, 1) micn_fwd_cluster_kernel(const FwdParams p, const int S) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    extern __shared__ __align__(128) unsigned char smem[];
    const long long num_slabs = p.N * p.C;
    const ClusterCtx c = cluster_setup<1>(smem, S, p.M * (long long)sizeof(T) / 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            cluster_producer<1>(c, S, num_slabs, [&](long long slab, int) {
                const long long n = slab / p.C, ch = slab - n * p.C;
                return reinterpret_cast<const char*>(p.x) + (n * p.x_sN + ch * p.x_sC) * (long long)sizeof(T);
            });
        }
    } else {
        Cursor base{0u, 0u};  // ring position of this slab's chunk 0
        int it = 0;
        const int nres = c.nchunks - c.r0;
        for (long long slab = c.cid; slab < num_slabs; slab += c.G, ++it) {
            const long long n = slab / p.C, ch = slab - n * p.C;
            const int style = load_style(p.styles, n, p.num_styles, p.status);
            float gamma, beta;
            load_affine(p, style, ch, gamma, beta);

            // ---- pass 1: statistics over this warp's units
            FwdStats<T> acc;
            acc.init();
            {
                Cursor cu = cursor_add(base, warp >> 2, S);
                for (int u = warp; u < c.nunits; u += kConsumerWarps) {
                    mbar_wait(c.full0 + 8 * cu.slot, cu.phase);
                    const uint32_t addr = c.data0 + cu.slot * kChunkBytes + (u & 3) * kUnitBytes + lane * 16;
                    const long long vb = (long long)u * kUnitVecs;
                    if (vb + kUnitVecs <= c.nv) {
                        uint4 v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = lds128(addr + i * 512);
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc.add(v[i], true);
                    } else if (vb < c.nv) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = vb + i * 32 + lane < c.nv;
                            uint4 v = make_uint4(0u, 0u, 0u, 0u);
                            if (ok) v = lds128(addr + i * 512);
                            acc.add(v, ok);
                        }
                    }
                    if ((u >> 2) < c.r0) unit_release(c.empty0 + 8 * cu.slot);  // not resident: hand it back
                    cu = cursor_add(cu, kConsumerWarps / kUnitsPerChunk, S);
                }
            }
            // warp -> CTA
            const Stat st = acc.finish();
            if (lane == 0) {
                c.warp_part[warp * 4 + 0] = st.n;
                c.warp_part[warp * 4 + 1] = st.mean;
                c.warp_part[warp * 4 + 2] = st.m2;
            }
            bar_sync(kConsumerBarrier, kConsumerThreads);
            float rec[3] = {0.f, 0.f, 0.f};
            if (warp == 0) {
                Stat w{0.f, 0.f, 0.f};
                if (lane < kConsumerWarps) {
                    w.n = c.warp_part[lane * 4 + 0];
                    w.mean = c.warp_part[lane * 4 + 1];
                    w.m2 = c.warp_part[lane * 4 + 2];
                }
                w = stat_warp_reduce(w);
                rec[0] = w.n;
                rec[1] = w.mean;
                rec[2] = w.m2;
            }
            // CTA -> cluster (DSMEM push + remote mbarrier arrive), merged in rank order
            cluster_exchange<3>(c, it, rec);
            const float* pp = c.peer_part + (it & 1) * kMaxCluster * 4;
            Stat tot{pp[0], pp[1], pp[2]};
            for (uint32_t r = 1; r < c.CS; ++r) tot = stat_merge(tot, Stat{pp[r * 4 + 0], pp[r * 4 + 1], pp[r * 4 + 2]});
            const float mean = tot.mean;
            const float rstd = 1.f / sqrtf(tot.m2 / tot.n + p.eps);  // biased variance, eps inside the sqrt
            if (c.rank == 0 && tid == 0 && p.save_mean) {
                p.save_mean[slab] = mean;
                p.save_rstd[slab] = rstd;
            }
            const float a = rstd * gamma;
            const float sub = sizeof(T) == 4 ? mean : 0.f;
            const float b = sizeof(T) == 4 ? beta : fmaf(-mean, a, beta);

            // ---- pass 2: resident chunks first (they free their slots for the next slab's
            //      prefetch), then the re-fetched ones.  Ring positions are contiguous from base+r0.
            char* ydst = reinterpret_cast<char*>(p.y) + (slab * p.M) * (long long)sizeof(T) + c.v0 * 16;
            const char* rsrc = nullptr;
            if (EPI == MICN_EPI_ADD_LRELU)
                rsrc = reinterpret_cast<const char*>(p.res) + (slab * p.M) * (long long)sizeof(T) + c.v0 * 16;
            {
                Cursor cu = cursor_add(base, c.r0 + (warp >> 2), S);
                for (int uk = warp; uk < c.nunits; uk += kConsumerWarps) {
                    const int k = uk >> 2;
                    const int j = k < nres ? c.r0 + k : k - nres;
                    const long long vb = ((long long)j * kUnitsPerChunk + (uk & 3)) * kUnitVecs;
                    const size_t goff = (size_t)(vb + lane) * 16;
                    uint4 rv[4];
                    if (EPI == MICN_EPI_ADD_LRELU) {  // residual: straight from HBM, 4 loads in flight per thread
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (vb + i * 32 + lane < c.nv) rv[i] = ldg_stream(rsrc + goff + i * 512);
                    }
                    mbar_wait(c.full0 + 8 * cu.slot, cu.phase);  // immediate for resident chunks
                    const uint32_t addr = c.data0 + cu.slot * kChunkBytes + (uk & 3) * kUnitBytes + lane * 16;
                    if (vb + kUnitVecs <= c.nv) {
                        uint4 v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = lds128(addr + i * 512);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            stg_stream(ydst + goff + i * 512, fwd_apply<T, EPI>(v[i], rv[i], sub, a, b, load_slope(p)));
                    } else if (vb < c.nv) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (vb + i * 32 + lane < c.nv)
                                stg_stream(ydst + goff + i * 512,
                                           fwd_apply<T, EPI>(lds128(addr + i * 512), rv[i], sub, a, b, load_slope(p)));
                    }
                    unit_release(c.empty0 + 8 * cu.slot);
                    cu = cursor_add(cu, kConsumerWarps / kUnitsPerChunk, S);
                }
            }
            base = cursor_add(base, (uint32_t)(c.nchunks + c.r0), S);  // next slab's chunk 0
        }
    }
    __syncwarp();
    __syncthreads();
    cluster_sync_all();  // nobody exits while a peer may still push into its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward:  g = dy * act'(.)   ;  S1 = sum g ; S2 = sum g*xhat ;
//            dx = gamma*rstd*(g - S1/M - xhat*S2/M) ; dresidual = g ; dgamma/dbeta from S2/S1
// ---------------------------------------------------------------------------------------------
struct BwdSlab {  // per-slab constants shared by both passes
    float mean, rstd, a, beta, slope;
    // 16-bit paths
    uint32_t K2;        // shift for the tensor-core sum(g*d): bf16(mean) when |mean| >> std, else 0
    float Kf;           // the same shift as a float
    uint32_t T2, flip;  // LeakyReLU mask: (x ^ flip) > T2  <=>  (x - mean)*a + beta > 0
    uint32_t slope2;
};

template <typename T, int EPI>
__device__ __forceinline__ BwdSlab bwd_slab_consts(float mean, float rstd, float gamma, float beta, float slope) {
    BwdSlab s;
    s.mean = mean;
    s.rstd = rstd;
    s.a = rstd * gamma;
    s.beta = beta;
    s.slope = slope;
    s.K2 = 0u;
    s.Kf = 0.f;
    s.T2 = 0u;
    s.flip = 0u;
    s.slope2 = 0u;
    if constexpr (sizeof(T) == 2) {
        using H = Half2Ops<T>;
        // sum(g*x) - mean*sum(g) cancels when |mean| >> std; there x - K with K = rn(mean) is exact
        // (same binade), so shift exactly in that regime and not at all otherwise.
        if (fabsf(mean) * rstd > 16.f) {
            s.K2 = H::bcast_rn(mean);
            s.Kf = H::low_to_float(s.K2);
        }
        s.slope2 = H::bcast_rn(slope);
        if (EPI == MICN_EPI_LRELU) {
            // pre > 0  <=>  x > T (a > 0) or x < T (a < 0) with T = mean - beta/a; for a 16-bit x,
            // x > T <=> x > round_down(T).  a < 0 is mapped onto the same compare by negating both.
            float Tt;
            if (s.a > 0.f)
                Tt = mean - beta / s.a;
            else if (s.a < 0.f) {
                Tt = -(mean - beta / s.a);
                s.flip = 0x80008000u;
            } else
                Tt = beta > 0.f ? -INFINITY : INFINITY;
            s.T2 = H::bcast_rd(Tt);
        }
    }
    return s;
}

// masked gradient of one packed pair (16-bit types)
template <typename T, int EPI>
__device__ __forceinline__ uint32_t bwd_mask2(const BwdSlab& s, uint32_t x2, uint32_t dy2, uint32_t o2) {
    using H = Half2Ops<T>;
    if (EPI == MICN_EPI_NONE) return dy2;
    const uint32_t m = EPI == MICN_EPI_LRELU ? H::gt2_mask(x2 ^ s.flip, s.T2) : H::gt2_mask(o2, 0u);
    const uint32_t dys = H::mul2(dy2, s.slope2);
    return (dy2 & m) | (dys & ~m);
}

template <typename T, int EPI>
struct BwdSums {  // 16-bit: the packed masked gradient (same mask as pass 2), sums in fp32
    float s1a, s1b, s2a, s2b;
    __device__ __forceinline__ void init() { s1a = s1b = s2a = s2b = 0.f; }
    __device__ __forceinline__ void add(const BwdSlab& s, const uint4& x, const uint4& dy, const uint4& o, bool ok) {
        if (!ok) return;
        const uint4 gq = make_uint4(bwd_mask2<T, EPI>(s, x.x, dy.x, o.x), bwd_mask2<T, EPI>(s, x.y, dy.y, o.y),
                                    bwd_mask2<T, EPI>(s, x.z, dy.z, o.z), bwd_mask2<T, EPI>(s, x.w, dy.w, o.w));
        float xf[8], gf[8];
        VecT<T>::unpack(x, xf);
        VecT<T>::unpack(gq, gf);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            s1a += gf[e];
            s1b += gf[e + 1];
            s2a = fmaf(gf[e], xf[e] - s.mean, s2a);
            s2b = fmaf(gf[e + 1], xf[e + 1] - s.mean, s2b);
        }
    }
    // warp totals (every lane): s1 = sum g, s2 = sum g * (x - mean)   [not yet scaled by rstd]
    __device__ __forceinline__ void finish(const BwdSlab&, float& s1, float& s2) {
        s1 = warp_sum(s1a + s1b);
        s2 = warp_sum(s2a + s2b);
    }
};

template <int EPI>
struct BwdSums<float, EPI> {
    float s1a, s1b, s2a, s2b;
    __device__ __forceinline__ void init() { s1a = s1b = s2a = s2b = 0.f; }
    __device__ __forceinline__ void add(const BwdSlab& s, const uint4& x, const uint4& dy, const uint4& o, bool ok) {
        if (!ok) return;
        float xf[4], gf[4], of[4];
        VecT<float>::unpack(x, xf);
        VecT<float>::unpack(dy, gf);
        if (EPI == MICN_EPI_ADD_LRELU) VecT<float>::unpack(o, of);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float d = xf[e] - s.mean;
            float g = gf[e];
            if (EPI == MICN_EPI_LRELU) g = fmaf(d, s.a, s.beta) > 0.f ? g : g * s.slope;
            if (EPI == MICN_EPI_ADD_LRELU) g = of[e] > 0.f ? g : g * s.slope;
            if (e & 1) {
                s1b += g;
                s2b = fmaf(g, d, s2b);
            } else {
                s1a += g;
                s2a = fmaf(g, d, s2a);
            }
        }
    }
    __device__ __forceinline__ void finish(const BwdSlab&, float& s1, float& s2) {
        s1 = warp_sum(s1a + s1b);
        s2 = warp_sum(s2a + s2b);
    }
};

// pass-2 math on one vector: dx = A*g + B1*(x - sub) + B0 ; gout = g (for dresidual)
template <typename T, int EPI>
__device__ __forceinline__ uint4 bwd_apply(const BwdSlab& s, const uint4& xv, const uint4& dyv, const uint4& ov, float A,
                                           float B1, float B0, float sub, uint4& gout) {
    constexpr int VN = VecT<T>::N;
    float xf[VN], gf[VN];
    VecT<T>::unpack(xv, xf);
    if constexpr (sizeof(T) == 2) {
        gout = make_uint4(bwd_mask2<T, EPI>(s, xv.x, dyv.x, ov.x), bwd_mask2<T, EPI>(s, xv.y, dyv.y, ov.y),
                          bwd_mask2<T, EPI>(s, xv.z, dyv.z, ov.z), bwd_mask2<T, EPI>(s, xv.w, dyv.w, ov.w));
        VecT<T>::unpack(gout, gf);
#pragma unroll
        for (int e = 0; e < VN; ++e) xf[e] = fmaf(A, gf[e], fmaf(B1, xf[e], B0));
    } else {
        float of[VN];
        VecT<T>::unpack(dyv, gf);
        if (EPI == MICN_EPI_ADD_LRELU) VecT<T>::unpack(ov, of);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const float d = xf[e] - sub;
            float g = gf[e];
            if (EPI == MICN_EPI_LRELU) g = fmaf(d, s.a, s.beta) > 0.f ? g : g * s.slope;
            if (EPI == MICN_EPI_ADD_LRELU) g = of[e] > 0.f ? g : g * s.slope;
            gf[e] = g;
            xf[e] = fmaf(A, g, fmaf(B1, d, B0));
        }
        if (EPI == MICN_EPI_ADD_LRELU) gout = VecT<T>::pack(gf);
    }
    return VecT<T>::pack(xf);
}

template <typename T, int EPI>
__global__ void __launch_bounds__(kClusterThreads, 1) micn_bwd_cluster_kernel(const BwdParams p, const int S) {
    pdl_wait();  // (programmatic dependent launch: nothing global is touched before the predecessor is through)
    pdl_launch_dependents();
    constexpr int NS = (EPI == MICN_EPI_ADD_LRELU) ? 3 : 2;  // x, dy [, act_out]
    extern __shared__ __align__(128) unsigned char smem[];
    const long long num_slabs = p.N * p.C;
    const ClusterCtx c = cluster_setup<NS>(smem, S, p.M * (long long)sizeof(T) / 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            cluster_producer<NS>(c, S, num_slabs, [&](long long slab, int s) {
                if (s == 0) {
                    const long long n = slab / p.C, ch = slab - n * p.C;
                    return reinterpret_cast<const char*>(p.x) + (n * p.x_sN + ch * p.x_sC) * (long long)sizeof(T);
                }
                const void* bs = s == 1 ? p.dy : p.act_out;
                return reinterpret_cast<const char*>(bs) + (slab * p.M) * (long long)sizeof(T);
            });
        }
    } else {
        Cursor base{0u, 0u};
        int it = 0;
        const float invM = 1.f / (float)p.M;
        const int nres = c.nchunks - c.r0;
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        for (long long slab = c.cid; slab < num_slabs; slab += c.G, ++it) {
            const long long n = slab / p.C, ch = slab - n * p.C;
            const int style = load_style(p.styles, n, p.num_styles, p.status);
            float gamma, beta;
            load_affine(p, style, ch, gamma, beta);
            const BwdSlab sc =
                bwd_slab_consts<T, EPI>(__ldg(p.save_mean + slab), __ldg(p.save_rstd + slab), gamma, beta, load_slope(p));

            // ---- pass 1
            BwdSums<T, EPI> acc;
            acc.init();
            {
                Cursor cu = cursor_add(base, warp >> 2, S);
                for (int u = warp; u < c.nunits; u += kConsumerWarps) {
                    mbar_wait(c.full0 + 8 * cu.slot, cu.phase);
                    const uint32_t addr = c.data0 + cu.slot * NS * kChunkBytes + (u & 3) * kUnitBytes + lane * 16;
                    const long long vb = (long long)u * kUnitVecs;
                    if (vb + kUnitVecs <= c.nv) {
                        uint4 xv[4], gv[4], ov[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            xv[i] = lds128(addr + i * 512);
                            gv[i] = lds128(addr + kChunkBytes + i * 512);
                            ov[i] = EPI == MICN_EPI_ADD_LRELU ? lds128(addr + 2 * kChunkBytes + i * 512) : zero4;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc.add(sc, xv[i], gv[i], ov[i], true);
                    } else if (vb < c.nv) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = vb + i * 32 + lane < c.nv;
                            uint4 xv = zero4, gv = zero4, ov = zero4;
                            if (ok) {
                                xv = lds128(addr + i * 512);
                                gv = lds128(addr + kChunkBytes + i * 512);
                                if (EPI == MICN_EPI_ADD_LRELU) ov = lds128(addr + 2 * kChunkBytes + i * 512);
                            }
                            acc.add(sc, xv, gv, ov, ok);
                        }
                    }
                    if ((u >> 2) < c.r0) unit_release(c.empty0 + 8 * cu.slot);
                    cu = cursor_add(cu, kConsumerWarps / kUnitsPerChunk, S);
                }
            }
            float s1, s2;
            acc.finish(sc, s1, s2);
            if (lane == 0) {
                c.warp_part[warp * 4 + 0] = s1;
                c.warp_part[warp * 4 + 1] = s2;
            }
            bar_sync(kConsumerBarrier, kConsumerThreads);
            float rec[2] = {0.f, 0.f};
            if (warp == 0) {
                rec[0] = warp_sum(lane < kConsumerWarps ? c.warp_part[lane * 4 + 0] : 0.f);
                rec[1] = warp_sum(lane < kConsumerWarps ? c.warp_part[lane * 4 + 1] : 0.f);
            }
            cluster_exchange<2>(c, it, rec);
            const float* pp = c.peer_part + (it & 1) * kMaxCluster * 4;
            float S1 = pp[0], S2 = pp[1];
            for (uint32_t r = 1; r < c.CS; ++r) {
                S1 += pp[r * 4 + 0];
                S2 += pp[r * 4 + 1];
            }
            S2 *= sc.rstd;  // sum g * xhat
            if (c.rank == 0 && tid == 0 && p.dgamma) {
                p.ws_sum_dy[slab] = S1;
                p.ws_sum_dyxh[slab] = S2;
            }
            // dx = a*g - a*S1/M - a*rstd*(S2/M)*(x - mean)
            const float A = sc.a;
            const float B1 = -sc.a * S2 * invM * sc.rstd;
            const float B0c = -sc.a * S1 * invM;
            const float sub = sizeof(T) == 4 ? sc.mean : 0.f;
            const float B0 = sizeof(T) == 4 ? B0c : fmaf(-B1, sc.mean, B0c);

            // ---- pass 2
            char* dxdst = reinterpret_cast<char*>(p.dx) + (slab * p.M) * (long long)sizeof(T) + c.v0 * 16;
            char* drdst = nullptr;
            if (EPI == MICN_EPI_ADD_LRELU)
                drdst = reinterpret_cast<char*>(p.dres) + (slab * p.M) * (long long)sizeof(T) + c.v0 * 16;
            {
                Cursor cu = cursor_add(base, c.r0 + (warp >> 2), S);
                for (int uk = warp; uk < c.nunits; uk += kConsumerWarps) {
                    const int k = uk >> 2;
                    const int j = k < nres ? c.r0 + k : k - nres;
                    const long long vb = ((long long)j * kUnitsPerChunk + (uk & 3)) * kUnitVecs;
                    const size_t goff = (size_t)(vb + lane) * 16;
                    mbar_wait(c.full0 + 8 * cu.slot, cu.phase);
                    const uint32_t addr = c.data0 + cu.slot * NS * kChunkBytes + (uk & 3) * kUnitBytes + lane * 16;
                    if (vb < c.nv) {
                        const bool full = vb + kUnitVecs <= c.nv;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (full || vb + i * 32 + lane < c.nv) {
                                const uint4 xv = lds128(addr + i * 512);
                                const uint4 gv = lds128(addr + kChunkBytes + i * 512);
                                const uint4 ov = EPI == MICN_EPI_ADD_LRELU ? lds128(addr + 2 * kChunkBytes + i * 512) : zero4;
                                uint4 gout = zero4;
                                const uint4 dxv = bwd_apply<T, EPI>(sc, xv, gv, ov, A, B1, B0, sub, gout);
                                stg_stream(dxdst + goff + i * 512, dxv);
                                if (EPI == MICN_EPI_ADD_LRELU) stg_stream(drdst + goff + i * 512, gout);
                            }
                        }
                    }
                    unit_release(c.empty0 + 8 * cu.slot);
                    cu = cursor_add(cu, kConsumerWarps / kUnitsPerChunk, S);
                }
            }
            base = cursor_add(base, (uint32_t)(c.nchunks + c.r0), S);
        }

        // ---- per-style parameter gradients: the LAST cluster to finish reduces the per-slab sums
        //      in a fixed order (deterministic; no float atomics) and re-arms the counter.
        if (p.dgamma && c.rank == 0) {
            if (tid == 0) {
                __threadfence();
                const unsigned int prev = atomicAdd(p.ws_counter, 1u);
                c.flags[0] = (prev == c.G - 1) ? 1 : 0;
            }
            bar_sync(kConsumerBarrier, kConsumerThreads);
            if (c.flags[0]) {
                __threadfence();
                const long long SC = (long long)p.num_styles * p.C;
                for (long long idx = tid; idx < SC; idx += kConsumerThreads) {
                    const int s = (int)(idx / p.C);
                    const long long ch = idx - (long long)s * p.C;
                    float acc_b = 0.f, acc_g = 0.f;
                    for (long long n = 0; n < p.N; ++n) {
                        if (load_style(p.styles, n, p.num_styles, nullptr) == s) {
                            acc_b += __ldcg(p.ws_sum_dy + n * p.C + ch);
                            acc_g += __ldcg(p.ws_sum_dyxh + n * p.C + ch);
                        }
                    }
                    p.dbeta[idx] = acc_b;
                    p.dgamma[idx] = acc_g;
                }
                if (tid == 0) *p.ws_counter = 0u;
            }
        }
    }
    __syncwarp();
    __syncthreads();
    cluster_sync_all();
}

}  // namespace micn
