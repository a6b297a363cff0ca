// micn_api.cu - the C ABI declared in include/micn.h: argument checks, path planning and launches.
//
// Three kernel families; the planner picks, per call,
//   * flat path    (micn_flat.cuh, default): 16-byte aligned slabs >= flat_min_bytes -> every slab cut
//                   into pieces dealt round-robin to one persistent CTA per SM, read-once/write-once
//   * small path   (micn_small.cuh): tiny slabs, or anything not 16-byte aligned -> warp / CTA per slab
//   * cluster path (micn_cluster.cuh): cluster-per-slab with DSMEM combine; kept selectable
//                   (force_path = 1) and as the fallback for slabs too large for the flat ring.
// Nothing here allocates or synchronises; the only state is per-process launch-attribute and
// occupancy caches.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "micn_cl.cuh"
#include "micn_cluster.cuh"
// the flat path in two CTA shapes: one 16-consumer-warp CTA per SM (fp32 forward, every backward) and two
// 8-consumer-warp CTAs per SM, out of phase with each other (16-bit forward: 31.5 us instead of 33.5 us at 1x48x96^3)
#define MICN_FLAT_NS flat1
#define MICN_FLAT_CW 16
#define MICN_FLAT_GW 4
#define MICN_FLAT_CPS 1
#include "micn_flat.cuh"
#undef MICN_FLAT_NS
#undef MICN_FLAT_CW
#undef MICN_FLAT_GW
#undef MICN_FLAT_CPS
#define MICN_FLAT_NS flat2
#define MICN_FLAT_CW 8
#define MICN_FLAT_GW 2
#define MICN_FLAT_CPS 2
#include "micn_flat.cuh"
#undef MICN_FLAT_NS
#undef MICN_FLAT_CW
#undef MICN_FLAT_GW
#undef MICN_FLAT_CPS
#include "micn_small.cuh"
#include "micn_res.cuh"

#include <map>
#include <tuple>

using namespace micn;

namespace {

// ------------------------------------------------------------------------------------------ options
struct Options {
    std::atomic<long long> cluster_size{-1};  // force CS (1..16)
    std::atomic<long long> force_path{-1};    // 0 small, 1 cluster, 2 flat, 4 resident (the codes `last_path` reports; 3 = channels-last)
    std::atomic<long long> flat_slots{-1};    // ring A slots (first touch, HBM) of the flat path
    std::atomic<long long> flat_slots_b{-1};  // ring B slots (second touch, L2)
    std::atomic<long long> flat_lag{-1};      // steps P2 trails P1
    std::atomic<long long> flat_l2_mb{-1};    // L2 budget (MB) the lag is sized for
    std::atomic<long long> flat_shape_fwd{-1}, flat_shape_bwd{-1};  // experiments: force CTA shape 1 / 2
    std::atomic<long long> flat_poll_delay_ns{-1}, flat_poll_backoff_ns{-1}, flat_poll_delay_tail_ns{-1};
    std::atomic<long long> flat_piece_vecs{-1};  // cap on vectors per piece
    std::atomic<long long> flat_min_bytes{-1};   // smallest slab the flat path takes
    std::atomic<long long> flat_grid{-1};     // cap on the persistent grid
    std::atomic<long long> flat_ovh_vecs{-1}; // planner: per-piece overhead in vector-equivalents
    std::atomic<long long> flat_coop{-1};     // 0: plain launch instead of a cooperative one (experiments)
    std::atomic<long long> res_cs{-1};        // resident path: force the cluster size (1..8)
    std::atomic<long long> res_min_bytes{-1}; // smallest slab the resident path takes
    std::atomic<long long> res_off{-1};       // 1: never the resident path (experiments)
    std::atomic<long long> xchg_dbg{-1};      // bring-up bits of the fused peer exchange: 1 no start fold, 2 no header atomic, 4 no peer stores
    std::atomic<long long> res_copies{-1};    // bulk copies per stream of a CTA's share (default 3)
    std::atomic<long long> res_trace{0};      // bring-up: device pointer of a [grid][8] int64 trace buffer
    std::atomic<long long> pdl{-1};           // 0: no programmatic dependent launch for the small / resident / channels-last kernels
    std::atomic<long long> flat_pdl{-1};      // 1: launch the flat kernels with programmatic stream serialization (PDL)
    std::atomic<long long> flat_refuse{-1};   // 1: behave as if the cooperative launch had been refused (tests the fallback chain)
    std::atomic<long long> flat_trace_which{0};  // 0 both, 1 forward only, 2 backward only
    std::atomic<long long> flat_trace{0};     // bring-up: device pointer of a [grid][64][16] int64 trace buffer
    std::atomic<long long> slots{-1};         // force ring slots S
    std::atomic<long long> max_clusters{-1};  // cap on co-resident clusters used
    std::atomic<long long> small_tps{-1};     // force 32 / 256 / 1024
    std::atomic<long long> cl_wide{-1};       // 0: never the 16-byte-load channels-last kernels (experiments)
    std::atomic<long long> small_reg{-1};     // 0: never the register-resident small kernels (experiments)
    std::atomic<long long> last_path{-1}, last_cs{-1}, last_slots{-1}, last_grid{-1}, last_lag{-1};  // read-back of the last plan
    std::atomic<long long> host_groups{-1};   // channel groups of the host-buffer path (default 10)
    std::atomic<long long> host_trace{0};     // 1: print a per-group timeline of the host-buffer path to stderr
    std::atomic<long long> host_taper{-1};    // 0: equal channel groups, else small groups at both ends (short fill and drain)
    std::atomic<long long> host_copy_2d{-1};  // 1: always 2-D copies (experiments)
    std::atomic<long long> launches{0};       // kernels launched by this library (bench "gpu_launches")
    std::atomic<long long> sm_bw_mbps{90000};   // per-SM bandwidth cap used by the planner (MB/s)
    std::atomic<long long> hbm_bw_mbps{6500000};
};
Options g_opt;

struct OptName {
    const char* name;
    std::atomic<long long>* v;
};
const OptName kOptNames[] = {
    {"cluster_size", &g_opt.cluster_size}, {"force_path", &g_opt.force_path}, {"slots", &g_opt.slots},
    {"max_clusters", &g_opt.max_clusters}, {"small_tps", &g_opt.small_tps}, {"small_reg", &g_opt.small_reg}, {"cl_wide", &g_opt.cl_wide},   {"last_path", &g_opt.last_path},
    {"last_cs", &g_opt.last_cs},           {"last_slots", &g_opt.last_slots}, {"last_grid", &g_opt.last_grid}, {"last_lag", &g_opt.last_lag},
    {"launches", &g_opt.launches}, {"host_groups", &g_opt.host_groups}, {"host_copy_2d", &g_opt.host_copy_2d}, {"host_taper", &g_opt.host_taper}, {"host_trace", &g_opt.host_trace},        {"sm_bw_mbps", &g_opt.sm_bw_mbps}, {"hbm_bw_mbps", &g_opt.hbm_bw_mbps},
    {"flat_slots", &g_opt.flat_slots},     {"flat_lag", &g_opt.flat_lag},     {"flat_piece_vecs", &g_opt.flat_piece_vecs},
    {"flat_min_bytes", &g_opt.flat_min_bytes}, {"flat_grid", &g_opt.flat_grid}, {"flat_ovh_vecs", &g_opt.flat_ovh_vecs},
    {"flat_coop", &g_opt.flat_coop},       {"flat_refuse", &g_opt.flat_refuse}, {"flat_pdl", &g_opt.flat_pdl}, {"pdl", &g_opt.pdl}, {"res_cs", &g_opt.res_cs}, {"res_min_bytes", &g_opt.res_min_bytes}, {"res_off", &g_opt.res_off}, {"xchg_dbg", &g_opt.xchg_dbg}, {"res_copies", &g_opt.res_copies}, {"res_trace", &g_opt.res_trace}, {"flat_trace", &g_opt.flat_trace},
    {"flat_trace_which", &g_opt.flat_trace_which}, {"flat_slots_b", &g_opt.flat_slots_b}, {"flat_l2_mb", &g_opt.flat_l2_mb},
    {"flat_shape_fwd", &g_opt.flat_shape_fwd}, {"flat_shape_bwd", &g_opt.flat_shape_bwd}, {"flat_poll_delay_ns", &g_opt.flat_poll_delay_ns}, {"flat_poll_delay_tail_ns", &g_opt.flat_poll_delay_tail_ns}, {"flat_poll_backoff_ns", &g_opt.flat_poll_backoff_ns},
};

// ------------------------------------------------------------------------------------------ device
struct DeviceInfo {
    int sm_count = 0, smem_optin = 0, cc_major = 0;
    bool valid = false;
};
std::mutex g_mu;
DeviceInfo g_dev[64];

int device_info(DeviceInfo** out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return MICN_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceInfo& d = g_dev[dev];
    if (!d.valid) {
        if ((e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
            return (int)e;
        if ((e = cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
        d.valid = true;
    }
    *out = &d;
    return 0;
}

// per (kernel, device): opt-in attributes set once; per (kernel, device, CS, smem): occupancy
struct KernelState {
    const void* fn;
    int dev;
    int smem_set;
    int occ[16];  // co-resident clusters for CS = 1..16 at smem_set bytes (-1 unknown)
};
std::deque<KernelState> g_kstate;  // a deque: entries never move, callers keep pointers to them

int cs_index(int cs) { return cs - 1; }

template <typename K>
int kernel_prepare(K kernel, int smem, int smem_optin, KernelState** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kernel);
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& k : g_kstate)
        if (k.fn == fn && k.dev == dev && k.smem_set == smem) {
            *out = &k;
            return 0;
        }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
    KernelState ks_new{fn, dev, smem, {}};
    for (int& o : ks_new.occ) o = -1;
    g_kstate.push_back(ks_new);
    *out = &g_kstate.back();
    return 0;
}

template <typename K>
int cluster_occupancy(K kernel, KernelState* ks, int cs, int smem) {
    const int ci = cs_index(cs);
    std::lock_guard<std::mutex> lk(g_mu);
    if (ks->occ[ci] >= 0) return ks->occ[ci];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    ks->occ[ci] = n;
    return n;
}

// Launch with programmatic stream serialization (PDL): the kernels start with griddepcontrol.wait, so their launch latency
// (and, for the resident kernels, their barrier set-up) overlaps the tail of whatever runs before them in the stream.
// Safe for every kernel family except the flat one, whose CTAs wait on each other across the grid (see launch_flat).
template <typename K, typename... Args>
int launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_opt.pdl.load() == 0 ? 0 : 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ------------------------------------------------------------------------------------------ planner
struct Plan {
    int path;  // 0 small, 1 cluster
    int tps;   // small: threads per slab
    int cs, slots, grid_clusters;
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename K>
int plan_cluster(K kernel, int NS_in, int NS_out, long long slabs, long long slab_bytes, const DeviceInfo& d, Plan* pl) {
    const int slots_max = (d.smem_optin - 2048) / (NS_in * kChunkBytes);
    int S = slots_max;
    const long long fs = g_opt.slots.load();
    if (fs > 0 && fs < S) S = (int)fs;
    // the 16 consumer warps span 4 consecutive chunks; parity waits are only sound when a slot cannot
    // be two phases away from any waiter, i.e. the ring holds more than those 4 chunks
    if (S < 5) S = 5;
    if (slots_max < 5) return MICN_ERR_BAD_ARG;
    const int smem = cluster_smem_bytes(S, NS_in);
    KernelState* ks = nullptr;
    int rc = kernel_prepare(kernel, smem, d.smem_optin, &ks);
    if (rc) return rc;

    const long long V = slab_bytes / 16;
    const double bw_sm = (double)g_opt.sm_bw_mbps.load() * 1e6, bw_hbm = (double)g_opt.hbm_bw_mbps.load() * 1e6;
    const long long forced = g_opt.cluster_size.load();
    double best = 1e30;
    int best_cs = 0, best_g = 0;
    for (int cs = 1; cs <= kMaxCluster; ++cs) {
        if (forced > 0 && cs != forced) continue;
        if (forced <= 0 && cs > 1 && V / cs < kChunkVecs / 2) continue;  // keep at least half a chunk per CTA
        int gmax = cluster_occupancy(kernel, ks, cs, smem);
        const long long cap = g_opt.max_clusters.load();
        if (cap > 0 && cap < gmax) gmax = (int)cap;
        if (gmax <= 0) continue;
        const long long g = slabs < gmax ? slabs : gmax;
        const long long share = (V + cs - 1) / cs;
        const long long nchunks = (share + kChunkVecs - 1) / kChunkVecs;
        const double reread = nchunks > S ? (double)(nchunks - S) / (double)nchunks : 0.0;
        const double traffic = (double)slab_bytes * (NS_in + NS_out) + 0.5 * reread * NS_in * (double)slab_bytes;
        auto t_wave = [&](long long active) {
            const double bw = std::min((double)cs * bw_sm, bw_hbm / (double)active);
            return traffic / bw + 1.0e-6;  // + per-slab reduce/exchange bubble
        };
        const long long full = slabs / g, rem = slabs % g;
        const double t = (double)full * t_wave(g) + (rem ? t_wave(rem) : 0.0);
        if (t < best * 0.999) {
            best = t;
            best_cs = cs;
            best_g = (int)g;
        }
    }
    if (!best_cs) return MICN_ERR_BAD_ARG;
    pl->path = 1;
    pl->cs = best_cs;
    pl->slots = S;
    pl->grid_clusters = best_g;
    return 0;
}

int small_tps(long long slab_bytes) {
    const long long f = g_opt.small_tps.load();
    if (f == 32 || f == 256 || f == 1024) return (int)f;
    // a warp per slab (8 slabs per CTA, no block barriers) up to 4 KB: [8,384,12^3] bf16 forward 10.3 us against
    // 23.1 us with a 256-thread CTA per slab; from 8 KB on the CTA per slab wins
    if (slab_bytes <= 4096) return 32;
    if (slab_bytes <= 128 * 1024) return 256;
    return 1024;
}

void record_plan(const Plan& pl) {
    g_opt.last_path.store(pl.path);
    g_opt.last_cs.store(pl.path ? pl.cs : pl.tps);
    g_opt.last_slots.store(pl.path ? pl.slots : 0);
    g_opt.last_grid.store(pl.grid_clusters);
    g_opt.launches.fetch_add(1);
}

template <typename K, typename P>
int launch_cluster(K kernel, const P& p, const Plan& pl, int NS, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pl.grid_clusters * pl.cs));
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = cluster_smem_bytes(pl.slots, NS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int S = pl.slots;
    return (int)cudaLaunchKernelEx(&cfg, kernel, p, S);
}


// ------------------------------------------------------------------------------------------ flat path
struct FlatWs {
    uint4* piece;
    uint4* slab;
    unsigned* ctl;  // header bytes 8..15: one 64-bit word, low half = CTAs arrived, high half = launch epoch
};
// Fixed prefix of every workspace, the same for all shapes (so a zero-filled workspace stays valid when calls of
// different shapes share it): the header words, then one arrival counter per channel (self-resetting).  The
// shape-dependent regions (per-slab sums, exchange records, channels-last partials) start at kWsData.
constexpr size_t kWsChanCounters = 16384;
constexpr size_t kWsData = 64 + kWsChanCounters * sizeof(unsigned);
constexpr size_t kWsHeader = 64;  // [0] counter (u32), [4] status (i32), [8] flat CTAs arrived (u32), [12] flat launch epoch (u32)

// upper bound of the pieces the flat planner can cut one slab into
long long flat_max_pieces(long long slab_bytes) {
    const long long V = slab_bytes / 16;
    long long p = (V + flat1::Traits::kMinPieceVecs - 1) / flat1::Traits::kMinPieceVecs;
    if (p < 1) p = 1;
    return p > flat1::Traits::kMaxPieces ? flat1::Traits::kMaxPieces : p;
}

struct WsLayout {
    size_t sums_off, slab_off, piece_off, total;
};
WsLayout ws_layout(long long N, long long C, long long M, int es) {
    WsLayout w;
    const size_t slabs = (size_t)N * (size_t)C;
    w.sums_off = kWsData;
    w.slab_off = (w.sums_off + slabs * 3 * sizeof(float) + 15) & ~(size_t)15;  // sum g, sum g*xhat, sum g*xhat2 (dual)
    w.piece_off = w.slab_off + slabs * 16 * 2;  // (two record sets: the dual-norm kernels exchange a second one)
    w.total = w.piece_off + slabs * (size_t)flat_max_pieces(M * es) * 16 * 2;
    w.total = (w.total + 255) & ~(size_t)255;
    return w;
}

template <typename TR>
struct FlatPlan {
    typename TR::Geom g;
    int grid, smem;
};

// occupancy of a flat kernel at its full shared-memory footprint (cached per kernel/device)
template <typename K>
int flat_blocks_per_sm(K kernel, int threads, int smem, int smem_optin) {
    KernelState* ks = nullptr;
    if (kernel_prepare(kernel, smem, smem_optin, &ks)) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    if (ks->occ[0] >= 0) return ks->occ[0];
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    ks->occ[0] = nb;
    return nb;
}

// returns 0 and fills *fp when the flat path can take the problem, 1 when it cannot, < 0 / > 0 codes on error
// NS = streams of a ring A slot (and, in L2, of a piece between its two touches); NSB = streams of a ring B slot
template <typename TR, typename KernelT>
int plan_flat(KernelT kernel, int NS, int NSB, long long slabs, long long C, long long slab_bytes, const DeviceInfo& d,
              FlatPlan<TR>* fp, int extra_smem = 0) {
    constexpr int kFlatCtasPerSm = TR::kCtasPerSm, kFlatMaxSlots = TR::kMaxSlots, kFlatMinPieceVecs = TR::kMinPieceVecs,
                  kFlatMaxPieces = TR::kMaxPieces, kFlatMaxLag = TR::kMaxLag, kFlatConsumerThreads = TR::kConsumerThreads;
    auto flat_ctl_bytes = [extra_smem] { return (long long)TR::kCtlBytes + extra_smem; };
    long long G = (long long)d.sm_count * kFlatCtasPerSm;
    const long long gcap = g_opt.flat_grid.load();
    if (gcap > 0 && gcap * kFlatCtasPerSm < G) G = gcap * kFlatCtasPerSm;
    const long long V = slab_bytes / 16;
    // shared memory of one SM (228 KB, 1 KB reserved per resident CTA) split between the co-resident CTAs
    const long long per_cta = kFlatCtasPerSm == 1 ? (long long)d.smem_optin : (233472LL / kFlatCtasPerSm - 1024);
    const long long ring = per_cta - flat_ctl_bytes() - 128;
    long long ovh = g_opt.flat_ovh_vecs.load();
    if (ovh < 0) ovh = 2200;

    // ring A holds the bytes in flight from HBM (~100 KB per SM streams at the full read rate, tools/streambw.cu;
    // anything much deeper only adds queueing delay to the record exchange); ring B re-reads from L2
    long long KA = g_opt.flat_slots.load(), KB = g_opt.flat_slots_b.load();
    if (KA <= 0) KA = 2;
    if (KB <= 0) KB = 2;
    KA = std::min<long long>(std::max<long long>(KA, 1), kFlatMaxSlots);
    KB = std::min<long long>(std::max<long long>(KB, 1), kFlatMaxSlots);
    // pieces as large as the rings allow (2 + 2 slots of ~56 KB): the 16 consumer warps of a CTA work on one
    // piece in lock step, so every per-piece latency chain (barrier wake-up, shuffles, hand-offs) is paid by
    // the whole SM and only amortises over large pieces (measured: 56 KB slots 0.69, 24 KB slots 0.58 of peak)
    long long pvmax = 1LL << 20;
    const long long cap = g_opt.flat_piece_vecs.load();
    if (cap >= kFlatMinPieceVecs) pvmax = cap;
    const long long fit = (ring / ((KA * NS + KB * NSB) * 16)) & ~7LL;
    if (pvmax > fit) pvmax = fit;
    if (pvmax > V) pvmax = V;
    if (pvmax < kFlatMinPieceVecs) return 1;
    const long long slot_vecs = (pvmax + 7) & ~7LL;
    const int smem = (int)((KA * NS + KB * NSB) * slot_vecs * 16 + flat_ctl_bytes());
    if (flat_blocks_per_sm(kernel, TR::kThreads, smem, d.smem_optin) < kFlatCtasPerSm) return 1;

    // a slab of P pieces spans R = ceil((P-1)/G)+1 rounds; P2 trails P1 by L >= R steps (L <= kFlatMaxLag)
    const long long pmax_hw = std::min<long long>(kFlatMaxPieces, (long long)(kFlatMaxLag - 1) * G + 1);
    const long long P0 = (V + pvmax - 1) / pvmax;
    if (P0 > pmax_hw) return 1;  // slab too large: not ours
    const long long Pend =
        std::min<long long>(pmax_hw, std::max<long long>(P0, (V + kFlatMinPieceVecs - 1) / kFlatMinPieceVecs));
    double best = 1e300;
    long long bestP = 0, bestPV = 0;
    for (long long P = P0; P <= Pend; ++P) {
        const long long PV = (V + P - 1) / P;
        const long long Pe = (V + PV - 1) / PV;
        if (Pe != P) continue;  // same split as a smaller P: already scored
        const long long T = slabs * Pe;
        if (T > 0x7fffffffLL) break;
        const long long rounds = (T + G - 1) / G;
        // a round costs its swept vectors (the consumer loops run whole 512-vector sweeps) plus a fixed latency
        // chain per piece (barrier wake-ups, shuffles, hand-offs: ~1 us, i.e. ~2200 vectors of HBM time per SM)
        const long long sweep = (PV + kFlatConsumerThreads - 1) / kFlatConsumerThreads * kFlatConsumerThreads;
        const double cost = (double)rounds * ((double)sweep + (double)ovh);
        if (cost < best * 0.9999) {
            best = cost;
            bestP = Pe;
            bestPV = PV;
        }
        if (PV * 3 < pvmax) break;  // far past the useful range
    }
    if (!bestP) return 1;
    // lag: long enough to cover the exchange at launch and at the tail (several microseconds of HBM time),
    // short enough that the L*G pieces waiting for their second touch stay well inside the 126 MB L2
    long long L_ = g_opt.flat_lag.load();
    if (L_ <= 0) {
        long long l2mb = g_opt.flat_l2_mb.load();
        if (l2mb <= 0) l2mb = 32;
        const double piece_bytes = (double)bestPV * 16.0 * NS;
        L_ = (long long)((double)l2mb * 1e6 / ((double)G * piece_bytes));
        if (L_ < 3) L_ = 3;
    }
    const long long need = (bestP - 1 + G - 1) / G + 1;  // R
    if (L_ < need) L_ = need;
    if (L_ > kFlatMaxLag) L_ = kFlatMaxLag;
    if (L_ < need) return 1;

    fp->g.V = (unsigned long long)V;
    fp->g.P = (unsigned)bestP;
    fp->g.PV = (unsigned)bestPV;
    fp->g.PVlast = (unsigned)(V - (bestP - 1) * bestPV);
    fp->g.T = (unsigned)(slabs * bestP);
    fp->g.KA = (unsigned)KA;
    fp->g.KB = (unsigned)KB;
    fp->g.L = (unsigned)L_;
    fp->g.slot_vecs = (unsigned)slot_vecs;
    fp->g.divP = fastdiv_make((unsigned)bestP);
    fp->g.divC = fastdiv_make((unsigned)C);
    const long long pd = g_opt.flat_poll_delay_ns.load(), pb = g_opt.flat_poll_backoff_ns.load();
    fp->g.poll_delay_ns = (unsigned)(pd >= 0 ? pd : 2500);
    fp->g.poll_backoff_ns = (unsigned)(pb >= 0 ? pb : 200);
    const long long pt = g_opt.flat_poll_delay_tail_ns.load();
    fp->g.poll_delay_tail_ns = (unsigned)(pt >= 0 ? pt : fp->g.poll_delay_ns / 4);
    fp->g.trace = reinterpret_cast<long long*>(g_opt.flat_trace.load());
    fp->grid = (int)std::min<long long>(G, (long long)fp->g.T);
    fp->smem = smem;
    return 0;
}

template <typename TR, typename K, typename P>
int launch_flat(K kernel, const P& p, const FlatPlan<TR>& fp, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)fp.grid);
    cfg.blockDim = dim3(TR::kThreads);
    cfg.dynamicSmemBytes = fp.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_opt.flat_coop.load() != 0) {
        attr[na].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait on each other's records
        attr[na].val.cooperative = 1;
        ++na;
    }
    if (g_opt.flat_pdl.load() == 1) {
        // programmatic dependent launch: this kernel's launch latency and barrier set-up overlap the tail of the kernel
        // before it in the stream (it blocks in griddepcontrol.wait before its first global-memory access)
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    typename TR::Geom g = fp.g;
    return (int)cudaLaunchKernelEx(&cfg, kernel, p, g);
}

template <typename TR>
void record_flat(const FlatPlan<TR>& fp) {
    g_opt.last_path.store(2);
    g_opt.last_cs.store(fp.g.P);
    g_opt.last_slots.store(fp.g.KA);
    g_opt.last_lag.store(fp.g.L);
    g_opt.last_grid.store(fp.grid);
    g_opt.launches.fetch_add(1);
}

// plan + launch of one flat kernel; kFlatNotTaken when the flat path declines (the caller tries the next path)
constexpr int kFlatNotTaken = -1000;
template <typename TR, typename KernelT, typename P>
int flat_run(KernelT kernel, int NS, int NSB, const P& p, long long slabs, long long slab_bytes, const FlatWs* ws_flat,
             const DeviceInfo& d, cudaStream_t st, int trace_mute, int extra_smem = 0) {
    FlatPlan<TR> fpl = {};
    const int rc = plan_flat<TR>(kernel, NS, NSB, slabs, p.C, slab_bytes, d, &fpl, extra_smem);
    if (rc == 1) return kFlatNotTaken;
    if (rc) return rc;
    if (g_opt.flat_trace_which.load() == trace_mute) fpl.g.trace = nullptr;
    fpl.g.ws_piece = ws_flat->piece;
    fpl.g.ws_slab = ws_flat->slab;
    fpl.g.ws_ctl = ws_flat->ctl;
    record_flat(fpl);
    const int lrc = g_opt.flat_refuse.load() == 1 ? (int)cudaErrorCooperativeLaunchTooLarge : launch_flat<TR>(kernel, p, fpl, st);
    // a device that cannot hold the whole persistent grid right now (MPS share, another resident kernel) refuses
    // the cooperative launch: take the cluster / small path instead of failing the call
    if (lrc != (int)cudaErrorCooperativeLaunchTooLarge) return lrc;
    cudaGetLastError();
    g_opt.launches.fetch_sub(1);  // (record_flat counted a launch that did not happen)
    return kFlatNotTaken;
}

long long flat_min_bytes() {
    const long long v = g_opt.flat_min_bytes.load();
    // measured with tools/calls_graph_probe.py (fwd+bwd pairs replayed from a CUDA graph, bf16): up to ~28 KB slabs the
    // small path ties or wins (96x22^3: 10.1 us against 11.7 us flat; 768x16^3: 18.0 against 21.6), from 44 KB on the
    // flat path does (96x28^3: 14.7 against 17.7; 48x40^3: 16.2 against 34.1)
    return v >= 0 ? v : 32 * 1024;
}

// ------------------------------------------------------------------------------------------ resident path
struct ResPlan {
    res::Geom g;
    int smem, grid;
    bool ok;
};
std::map<std::tuple<const void*, int, int, long long, long long, long long>, ResPlan> g_res_plans;

// Picks the cluster size for which slabs * CS CTAs fill the SMs in ONE wave with the smallest per-SM share, or declines
// (ok = false: more bytes than the shared memory of the whole chip holds, slabs too small to split, ...).  Cached.
template <typename K>
ResPlan plan_res(K kernel, int NS, long long slabs, long long slab_bytes, const DeviceInfo& d) {
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kernel);
    const long long forced = g_opt.res_cs.load();
    const auto key = std::make_tuple(fn, dev, NS, slabs, slab_bytes, forced * 64 + g_opt.res_copies.load());
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_res_plans.find(key);
        if (it != g_res_plans.end()) return it->second;
    }
    ResPlan best = {};
    best.ok = false;
    const long long V = slab_bytes / 16;
    double best_cost = 1e300;
    for (int cs = 1; cs <= res::kMaxCluster && V > 0 && V < 0x7fffffffLL; ++cs) {
        if (forced > 0 && cs != forced) continue;
        if (slabs * cs > 8LL * d.sm_count) break;
        const long long nvmax = (V + cs - 1) / cs;
        if (cs > 1 && nvmax < 256) break;  // 4 KB per CTA: below that the per-CTA fixed cost dominates
        const long long nch = (nvmax + res::kChunkVecs - 1) / res::kChunkVecs;
        if (nch * NS > res::kMaxChunks) continue;
        const int smem = res::smem_bytes(NS, (int)nch);
        if (smem > d.smem_optin) continue;
        KernelState* ks = nullptr;
        if (kernel_prepare(kernel, smem, d.smem_optin, &ks)) continue;
        int nclusters = 0;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            if (ks->occ[cs_index(cs)] < 0) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(cs * 64));
                cfg.blockDim = dim3(res::kThreads);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cs;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, kernel, &cfg) != cudaSuccess) {
                    cudaGetLastError();
                    nc = 0;
                }
                ks->occ[cs_index(cs)] = nc;
            }
            nclusters = ks->occ[cs_index(cs)];
        }
        if (nclusters < slabs) continue;  // would need a second wave: the flat path handles that regime better
        // per-SM bytes in the busiest SM (CTAs are dealt evenly over the SMs) + a small per-CTA charge
        const long long per_sm = (slabs * cs + d.sm_count - 1) / d.sm_count;
        const double cost = (double)per_sm * (double)nvmax * NS + 96.0 * cs;
        if (cost < best_cost * 0.999) {
            best_cost = cost;
            best.ok = true;
            best.smem = smem;
            best.grid = (int)(slabs * cs);
            best.g.V = (unsigned)V;
            best.g.CS = (unsigned)cs;
            best.g.nv_base = (unsigned)(V / cs);
            best.g.nv_rem = (unsigned)(V % cs);
            best.g.nch_max = (unsigned)nch;
            // a share travels as `copies` bulk copies of cv vectors (whole 8 KB granules): few enough that issuing them
            // (~0.15 us each) does not delay the data, enough that the statistics pass overlaps the arrival
            long long copies = g_opt.res_copies.load();
            if (copies <= 0) copies = 3;
            copies = std::max<long long>(copies, (nch + 3) / 4);  // at most 32 KB per copy (as the flat path's producers)
            copies = std::min<long long>(std::min<long long>(copies, res::kMaxCopies), nch);
            best.g.cv = (unsigned)(((nch + copies - 1) / copies) * res::kChunkVecs);
            best.g.trace = nullptr;
        }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_res_plans.size() > 4096) g_res_plans.clear();
    g_res_plans[key] = best;
    return best;
}

long long res_min_bytes() {
    const long long v = g_opt.res_min_bytes.load();
    return v >= 0 ? v : 32 * 1024;
}

// plan + launch; kFlatNotTaken when the resident path declines
constexpr int kNotTaken = -1000;
template <typename K, typename P>
int res_run(K kernel, int NS, const P& p, long long slabs, long long slab_bytes, const DeviceInfo& d, cudaStream_t st) {
    const ResPlan pl = plan_res(kernel, NS, slabs, slab_bytes, d);
    if (!pl.ok) return kNotTaken;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3(res::kThreads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.g.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // (see launch_pdl)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_opt.pdl.load() == 0 ? 1 : 2;
    g_opt.last_path.store(4);  // (3 is the channels-last path)
    g_opt.last_cs.store(pl.g.CS);
    g_opt.last_slots.store(pl.g.nch_max);
    g_opt.last_grid.store(pl.grid);
    g_opt.launches.fetch_add(1);
    res::Geom g = pl.g;
    g.trace = reinterpret_cast<long long*>(g_opt.res_trace.load());
    return (int)cudaLaunchKernelEx(&cfg, kernel, p, g);
}

inline bool res_wanted(long long fp, bool can_cluster, long long slab_bytes) {
    if (!can_cluster || g_opt.res_off.load() == 1) return false;
    return fp == 4 || (fp < 0 && slab_bytes >= res_min_bytes());
}

// ------------------------------------------------------------------------------------------ typed dispatch
template <typename T, int EPI>
int fwd_typed(const FwdParams& p, bool can_cluster, const FlatWs* ws_flat, const DeviceInfo& d, cudaStream_t st) {
    const long long slabs = p.N * p.C, slab_bytes = p.M * (long long)sizeof(T);
    Plan pl = {};
    const long long fp = g_opt.force_path.load();
    bool use_cluster = can_cluster && slab_bytes >= 32 * 1024;
    if (fp == 0) use_cluster = false;
    if (fp == 1 && can_cluster) use_cluster = true;
    if (res_wanted(fp, can_cluster, slab_bytes)) {  // everything fits the chip's shared memory in one wave: resident path
        const int rrc = res_run(res::micn_fwd_res_kernel<T, EPI>, EPI == MICN_EPI_ADD_LRELU ? 2 : 1, p, slabs, slab_bytes, d, st);
        if (rrc != kNotTaken) return rrc;
    }
    if (can_cluster && ws_flat && (fp == 2 || (fp < 0 && slab_bytes >= flat_min_bytes()))) {
        // 16-bit I/O: two half-size CTAs per SM hide each other's per-piece latency chains; fp32: one CTA per SM
        const long long shape = g_opt.flat_shape_fwd.load();
        int frc;
        if (shape == 2 || (shape != 1 && sizeof(T) == 2))
            frc = flat_run<flat2::Traits>(flat2::micn_fwd_flat_kernel<T, EPI>, 1, EPI == MICN_EPI_ADD_LRELU ? 2 : 1, p, slabs,
                                          slab_bytes, ws_flat, d, st, 2);
        else
            frc = flat_run<flat1::Traits>(flat1::micn_fwd_flat_kernel<T, EPI>, 1, EPI == MICN_EPI_ADD_LRELU ? 2 : 1, p, slabs,
                                          slab_bytes, ws_flat, d, st, 2);
        if (frc != kFlatNotTaken) return frc;
    }
    if (use_cluster) {
        auto kernel = micn_fwd_cluster_kernel<T, EPI>;
        int rc = plan_cluster(kernel, 1, EPI == MICN_EPI_ADD_LRELU ? 2 : 1, slabs, slab_bytes, d, &pl);
        if (rc) return rc;
        record_plan(pl);
        return launch_cluster(kernel, p, pl, 1, st);
    }
    pl.path = 0;
    pl.tps = small_tps(slab_bytes);
    // a CTA per slab with more than two 16-byte vectors per thread: the register-resident instantiation
    // (micn_small.cuh).  Measured neutral to slightly negative for the warp-per-slab shape, so not used there.
    const bool reg = pl.tps == 256 && slab_bytes / 16 > 2 * pl.tps && g_opt.small_reg.load() != 0;
    if (pl.tps == 32) {
        const unsigned grid = (unsigned)((slabs + 7) / 8);
        pl.grid_clusters = (int)grid;
        record_plan(pl);
        launch_pdl(micn_fwd_small_kernel<T, EPI, 32>, dim3(grid), dim3(256), 0, st, p);
    } else if (pl.tps == 256) {
        pl.grid_clusters = (int)slabs;
        record_plan(pl);
        if (reg)
            launch_pdl(micn_fwd_small_kernel<T, EPI, 256, true>, dim3((unsigned)slabs), dim3(256), 0, st, p);
        else
            launch_pdl(micn_fwd_small_kernel<T, EPI, 256>, dim3((unsigned)slabs), dim3(256), 0, st, p);
    } else {
        pl.grid_clusters = (int)slabs;
        record_plan(pl);
        launch_pdl(micn_fwd_small_kernel<T, EPI, 1024>, dim3((unsigned)slabs), dim3(1024), 0, st, p);
    }
    return (int)cudaGetLastError();
}

template <typename T, int EPI>
int bwd_typed(const BwdParams& p, bool can_cluster, const FlatWs* ws_flat, const DeviceInfo& d, cudaStream_t st) {
    const long long slabs = p.N * p.C, slab_bytes = p.M * (long long)sizeof(T);
    constexpr int NS = EPI == MICN_EPI_ADD_LRELU ? 3 : 2;
    Plan pl = {};
    const long long fp = g_opt.force_path.load();
    bool use_cluster = can_cluster && slab_bytes >= 32 * 1024;
    if (fp == 0) use_cluster = false;
    if (fp == 1 && can_cluster) use_cluster = true;
    const bool ds = EPI == MICN_EPI_LRELU && p.dslope != nullptr;
    if (ds) use_cluster = false;  // the slope-gradient partials are produced by the resident, flat and small kernels only
    const bool xchg = p.xchg_world > 1;  // fused peer exchange of the parameter gradients: flat kernels only
    if (!xchg && res_wanted(fp, can_cluster, slab_bytes) && !(p.dgamma && p.N > 1 && !p.ws_chan_cnt)) {
        int rrc;
        if (ds)
            rrc = res_run(res::micn_bwd_res_kernel<T, EPI, EPI == MICN_EPI_LRELU>, NS, p, slabs, slab_bytes, d, st);
        else
            rrc = res_run(res::micn_bwd_res_kernel<T, EPI, false>, NS, p, slabs, slab_bytes, d, st);
        if (rrc != kNotTaken) return rrc;
    }
    if (can_cluster && ws_flat && (fp == 2 || xchg || (fp < 0 && slab_bytes >= flat_min_bytes()))) {
        int frc;
        if (g_opt.flat_shape_bwd.load() == 2) {
            auto kernel = ds ? flat2::micn_bwd_flat_kernel<T, EPI, EPI == MICN_EPI_LRELU> : flat2::micn_bwd_flat_kernel<T, EPI, false>;
            frc = flat_run<flat2::Traits>(kernel, NS, NS, p, slabs, slab_bytes, ws_flat, d, st, 1);
        } else {
            auto kernel = ds ? flat1::micn_bwd_flat_kernel<T, EPI, EPI == MICN_EPI_LRELU> : flat1::micn_bwd_flat_kernel<T, EPI, false>;
            frc = flat_run<flat1::Traits>(kernel, NS, NS, p, slabs, slab_bytes, ws_flat, d, st, 1);
        }
        if (frc != kFlatNotTaken) return frc;
    }
    if (xchg) return MICN_ERR_UNSUPPORTED;
    if (use_cluster) {
        auto kernel = micn_bwd_cluster_kernel<T, EPI>;
        int rc = plan_cluster(kernel, NS, EPI == MICN_EPI_ADD_LRELU ? 2 : 1, slabs, slab_bytes, d, &pl);
        if (rc) return rc;
        record_plan(pl);
        return launch_cluster(kernel, p, pl, NS, st);
    }
    pl.path = 0;
    pl.tps = small_tps(slab_bytes);
    // a CTA per slab with more than two 16-byte vectors per thread: the register-resident instantiation
    // (micn_small.cuh).  Measured neutral to slightly negative for the warp-per-slab shape, so not used there.
    const bool reg = pl.tps == 256 && slab_bytes / 16 > 2 * pl.tps && g_opt.small_reg.load() != 0;
    if (pl.tps == 32) {
        const unsigned grid = (unsigned)((slabs + 7) / 8);
        pl.grid_clusters = (int)grid;
        record_plan(pl);
        launch_pdl(micn_bwd_small_kernel<T, EPI, 32>, dim3(grid), dim3(256), 0, st, p);
    } else if (pl.tps == 256) {
        pl.grid_clusters = (int)slabs;
        record_plan(pl);
        if (reg)
            launch_pdl(micn_bwd_small_kernel<T, EPI, 256, true>, dim3((unsigned)slabs), dim3(256), 0, st, p);
        else
            launch_pdl(micn_bwd_small_kernel<T, EPI, 256>, dim3((unsigned)slabs), dim3(256), 0, st, p);
    } else {
        pl.grid_clusters = (int)slabs;
        record_plan(pl);
        launch_pdl(micn_bwd_small_kernel<T, EPI, 1024>, dim3((unsigned)slabs), dim3(1024), 0, st, p);
    }
    return (int)cudaGetLastError();
}

template <typename T>
int fwd_by_epi(int epi, const FwdParams& p, bool cc, const FlatWs* wf, const DeviceInfo& d, cudaStream_t st) {
    switch (epi) {
        case MICN_EPI_NONE: return fwd_typed<T, MICN_EPI_NONE>(p, cc, wf, d, st);
        case MICN_EPI_LRELU: return fwd_typed<T, MICN_EPI_LRELU>(p, cc, wf, d, st);
        case MICN_EPI_ADD_LRELU: return fwd_typed<T, MICN_EPI_ADD_LRELU>(p, cc, wf, d, st);
    }
    return MICN_ERR_BAD_ARG;
}
template <typename T>
int bwd_by_epi(int epi, const BwdParams& p, bool cc, const FlatWs* wf, const DeviceInfo& d, cudaStream_t st) {
    switch (epi) {
        case MICN_EPI_NONE: return bwd_typed<T, MICN_EPI_NONE>(p, cc, wf, d, st);
        case MICN_EPI_LRELU: return bwd_typed<T, MICN_EPI_LRELU>(p, cc, wf, d, st);
        case MICN_EPI_ADD_LRELU: return bwd_typed<T, MICN_EPI_ADD_LRELU>(p, cc, wf, d, st);
    }
    return MICN_ERR_BAD_ARG;
}

int elem_size(int dtype) { return dtype == MICN_F32 ? 4 : (dtype == MICN_BF16 || dtype == MICN_F16) ? 2 : 0; }

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
// ------------------------------------------------------------------------------------------ channels-last path
namespace {
struct ClPlan {
    bool fused;  // short columns: statistics and apply in one launch (micn_cl_*_fused_kernel)
    int MS;
    long long rps;
    float4* part;
    float2* slab;
    int* status;
    unsigned* tile_cnt;
};
size_t cl_ws_bytes(int64_t N, int64_t C) {
    return kWsData + (size_t)N * kClMaxSplits * (size_t)C * sizeof(float4) + (size_t)N * (size_t)C * sizeof(float2);
}
int cl_plan(int64_t N, int64_t C, int64_t M, void* workspace, size_t workspace_bytes, const DeviceInfo& d, ClPlan* pl) {
    if (!workspace || workspace_bytes < cl_ws_bytes(N, C)) return MICN_ERR_WORKSPACE;
    const long long tiles = (C + kClTile - 1) / kClTile;
    pl->fused = M <= kClFusedMaxRows;
    long long ms = (2LL * d.sm_count + N * tiles - 1) / (N * tiles);  // at least two CTAs per SM
    ms = std::min<long long>(ms, kClMaxSplits);
    ms = std::min<long long>(ms, (M + 31) / 32);  // at least four rows per warp
    ms = std::max<long long>(ms, 1);
    if (pl->fused) ms = 1;
    pl->rps = (M + ms - 1) / ms;
    pl->MS = (int)((M + pl->rps - 1) / pl->rps);
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    pl->status = reinterpret_cast<int*>(w) + 1;
    pl->tile_cnt = (size_t)tiles <= kWsChanCounters ? reinterpret_cast<unsigned*>(w + kWsHeader) : nullptr;
    pl->part = reinterpret_cast<float4*>(w + kWsData);
    pl->slab = reinterpret_cast<float2*>(w + kWsData + (size_t)N * kClMaxSplits * (size_t)C * sizeof(float4));
    return 0;
}
// 16-byte row loads need whole vectors of channels per lane and aligned tensors (micn_cl.cuh, wide variant)
template <typename T>
bool cl_wide_ok(const ClParams& p) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.dy);
    return g_opt.cl_wide.load() != 0 && p.C % ClWide<T>::CPL == 0 && (bits & 15u) == 0;
}

// (the fused one-launch kernels are launched as programmatic dependents; the two-kernel route is not: measured on
// [1,384,24^3] bf16 it made the fwd+bwd pair slower, 36.0 -> 40.8 us)
template <typename T>
int cl_fwd_typed(const ClParams& p, dim3 grid, bool fused, cudaStream_t st) {
    if (fused) {
        launch_pdl(micn_cl_fwd_fused_kernel<T>, dim3(grid), dim3(kClFusedThreads), 0, st, p);
        return (int)cudaGetLastError();
    }
    if (cl_wide_ok<T>(p)) {
        micn_cl_fwd_stats_wide_kernel<T><<<grid, kClThreads, 0, st>>>(p);
        micn_cl_fwd_apply_wide_kernel<T><<<grid, kClThreads, 0, st>>>(p);
        return (int)cudaGetLastError();
    }
    micn_cl_fwd_stats_kernel<T><<<grid, kClThreads, 0, st>>>(p);
    micn_cl_fwd_apply_kernel<T><<<grid, kClThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}
template <typename T>
int cl_bwd_typed(const ClParams& p, dim3 grid, bool fused, cudaStream_t st) {
    if (fused) {
        launch_pdl(micn_cl_bwd_fused_kernel<T>, dim3(grid), dim3(kClFusedThreads), 0, st, p);
    } else {
        micn_cl_bwd_stats_kernel<T><<<grid, kClThreads, 0, st>>>(p);
        micn_cl_bwd_apply_kernel<T><<<grid, kClThreads, 0, st>>>(p);
    }
    if (p.dgamma && p.N > 1 && !p.tile_cnt) {  // (otherwise the backward kernels fold the parameter gradients themselves)
        const long long sc = (long long)p.num_styles * p.C;
        micn_cl_param_grads_kernel<<<(unsigned)((sc + 255) / 256), 256, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}
int cl_fill(ClParams& p, const float* const* gamma, const float* const* beta, int num_styles, const int64_t* styles,
            int64_t N, int64_t C, int64_t M, const ClPlan& pl) {
    p.affine = (gamma && beta) ? 1 : 0;
    for (int s = 0; s < kMaxStyles; ++s) {
        p.gamma[s] = (p.affine && s < num_styles) ? gamma[s] : nullptr;
        p.beta[s] = (p.affine && s < num_styles) ? beta[s] : nullptr;
        if (p.affine && s < num_styles && (!p.gamma[s] || !p.beta[s])) return MICN_ERR_BAD_ARG;
    }
    p.styles = reinterpret_cast<const long long*>(styles);
    p.status = pl.status;
    p.ws_part = pl.part;
    p.ws_slab = pl.slab;
    p.tile_cnt = pl.tile_cnt;
    p.N = N;
    p.C = C;
    p.M = M;
    p.MS = pl.MS;
    p.rows_per_split = pl.rps;
    p.num_styles = num_styles;
    return 0;
}
}  // namespace

// ------------------------------------------------------------------------------------------ dual-norm epilogue (typed)
namespace {
template <typename T>
int dual_supported_typed(int64_t N, int64_t C, int64_t M, int backward, const DeviceInfo& d) {
    const long long slabs = N * C, slab_bytes = M * (long long)sizeof(T);
    if ((M * (long long)sizeof(T)) % 16 != 0) return 0;
    const long long fp = g_opt.force_path.load();
    if (g_opt.res_off.load() != 1 && (fp < 0 || fp == 4)) {
        if (backward ? plan_res(res::micn_bwd_res_kernel<T, MICN_EPI_NORM_ADD_LRELU, false>, 3, slabs, slab_bytes, d).ok
                     : plan_res(res::micn_fwd_res_kernel<T, MICN_EPI_NORM_ADD_LRELU>, 2, slabs, slab_bytes, d).ok)
            return 1;
    }
    if ((fp < 0 && slab_bytes >= flat_min_bytes()) || fp == 2) {
        FlatPlan<flat1::Traits> fpl = {};
        const int extra = flat1::Traits::kDualExtraBytes;
        if (backward) return plan_flat<flat1::Traits>(flat1::micn_bwd_flat_kernel<T, MICN_EPI_NORM_ADD_LRELU, false>, 3, 3, slabs, C,
                                                      slab_bytes, d, &fpl, extra) == 0;
        return plan_flat<flat1::Traits>(flat1::micn_fwd_flat_kernel<T, MICN_EPI_NORM_ADD_LRELU>, 2, 2, slabs, C, slab_bytes, d,
                                        &fpl, extra) == 0;
    }
    return 0;
}
template <typename T>
int fwd_dual_typed(const FwdParams& p, const FlatWs* wf, const DeviceInfo& d, cudaStream_t st) {
    const long long slabs = p.N * p.C, slab_bytes = p.M * (long long)sizeof(T);
    const long long fp = g_opt.force_path.load();
    if (g_opt.res_off.load() != 1 && (fp < 0 || fp == 4)) {
        const int rc = res_run(res::micn_fwd_res_kernel<T, MICN_EPI_NORM_ADD_LRELU>, 2, p, slabs, slab_bytes, d, st);
        if (rc != kNotTaken) return rc;
    }
    if (wf && ((fp < 0 && slab_bytes >= flat_min_bytes()) || fp == 2)) {
        const int rc = flat_run<flat1::Traits>(flat1::micn_fwd_flat_kernel<T, MICN_EPI_NORM_ADD_LRELU>, 2, 2, p, slabs, slab_bytes,
                                               wf, d, st, 2, flat1::Traits::kDualExtraBytes);
        if (rc != kFlatNotTaken) return rc;
    }
    return MICN_ERR_UNSUPPORTED;
}
template <typename T>
int bwd_dual_typed(const BwdParams& p, const FlatWs* wf, const DeviceInfo& d, cudaStream_t st) {
    const long long slabs = p.N * p.C, slab_bytes = p.M * (long long)sizeof(T);
    const long long fp = g_opt.force_path.load();
    if (g_opt.res_off.load() != 1 && (fp < 0 || fp == 4) && !(p.dgamma && p.N > 1 && !p.ws_chan_cnt)) {
        const int rc = res_run(res::micn_bwd_res_kernel<T, MICN_EPI_NORM_ADD_LRELU, false>, 3, p, slabs, slab_bytes, d, st);
        if (rc != kNotTaken) return rc;
    }
    if (wf && ((fp < 0 && slab_bytes >= flat_min_bytes()) || fp == 2)) {
        const int rc = flat_run<flat1::Traits>(flat1::micn_bwd_flat_kernel<T, MICN_EPI_NORM_ADD_LRELU, false>, 3, 3, p, slabs,
                                               slab_bytes, wf, d, st, 1, flat1::Traits::kDualExtraBytes);
        if (rc != kFlatNotTaken) return rc;
    }
    return MICN_ERR_UNSUPPORTED;
}
}  // namespace

extern "C" {

int micn_version(void) { return MICN_VERSION; }

const char* micn_error_string(int code) {
    switch (code) {
        case MICN_OK: return "ok";
        case MICN_ERR_BAD_ARG: return "micn: bad argument";
        case MICN_ERR_BAD_DTYPE: return "micn: unsupported dtype (fp32, bf16, fp16 only)";
        case MICN_ERR_TOO_MANY_STYLES: return "micn: num_styles exceeds MICN_MAX_STYLES";
        case MICN_ERR_WORKSPACE: return "micn: workspace missing or too small";
        case MICN_ERR_UNALIGNED: return "micn: pointer not sufficiently aligned (element size; two elements for the channels-last calls; 16 bytes for the workspace)";
        case MICN_ERR_NO_DEVICE: return "micn: no usable CUDA device";
        case MICN_ERR_UNSUPPORTED: return "micn: no dual-norm kernel takes this shape (compose micn_fwd calls instead)";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "micn: unknown error";
}

int micn_set_option(const char* key, long long value) {
    if (!key) return MICN_ERR_BAD_ARG;
    for (const auto& o : kOptNames)
        if (!std::strcmp(o.name, key)) {
            o.v->store(value);
            return 0;
        }
    return MICN_ERR_BAD_ARG;
}

long long micn_get_option(const char* key) {
    if (!key) return -1;
    for (const auto& o : kOptNames)
        if (!std::strcmp(o.name, key)) return o.v->load();
    return -1;
}

size_t micn_workspace_bytes(int64_t N, int64_t C, int64_t M, int dtype, int num_styles) {
    (void)num_styles;
    const int es = elem_size(dtype);
    if (N < 0 || C < 0 || M < 0 || !es) return 0;
    return ws_layout(N, C, M, es).total;
}

int micn_read_status(void* workspace, void* stream, int* status_out) {
    if (!workspace || !status_out) return MICN_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int* w = reinterpret_cast<int*>(workspace) + 1;
    cudaError_t e = cudaMemcpyAsync(status_out, w, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(w, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(st);
}

static int fwd_impl(const void* x, void* y, const void* residual, const float* const* gamma, const float* const* beta,
                    int num_styles, const int64_t* styles, float* save_mean, float* save_rstd, int64_t N, int64_t C,
                    int64_t M, int64_t x_stride_n, int64_t x_stride_c, int dtype, int epilogue, float slope,
                    const float* slope_dev, float eps, void* workspace, size_t workspace_bytes, void* stream) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if (N == 0 || C == 0 || M == 0) return MICN_OK;
    if (!x || !y) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if (epilogue < MICN_EPI_NONE || epilogue > MICN_EPI_ADD_LRELU) return MICN_ERR_BAD_ARG;
    if (epilogue == MICN_EPI_ADD_LRELU && !residual) return MICN_ERR_BAD_ARG;
    if ((gamma == nullptr) != (beta == nullptr)) return MICN_ERR_BAD_ARG;
    if ((save_mean == nullptr) != (save_rstd == nullptr)) return MICN_ERR_BAD_ARG;
    if (N * C > 0x7fffffffLL) return MICN_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & (es - 1))
        return MICN_ERR_UNALIGNED;
    if (workspace && workspace_bytes < kWsHeader) return MICN_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MICN_ERR_UNALIGNED;  // 16-byte records, 64-bit header word

    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;

    FwdParams p = {};
    p.x = x;
    p.y = y;
    p.res = epilogue == MICN_EPI_ADD_LRELU ? residual : nullptr;
    p.affine = gamma != nullptr;
    for (int s = 0; s < num_styles && gamma; ++s) {
        if (!gamma[s] || !beta[s]) return MICN_ERR_BAD_ARG;
        p.gamma[s] = gamma[s];
        p.beta[s] = beta[s];
    }
    p.styles = reinterpret_cast<const long long*>(styles);
    p.save_mean = save_mean;
    p.save_rstd = save_rstd;
    p.status = workspace ? reinterpret_cast<int*>(workspace) + 1 : nullptr;
    p.N = N;
    p.C = C;
    p.M = M;
    p.x_sN = x_stride_n;
    p.x_sC = x_stride_c;
    p.num_styles = num_styles;
    p.eps = eps;
    p.slope = slope;
    p.slope_dev = slope_dev;

    const bool can_cluster = d->cc_major >= 9 && aligned16(x) && aligned16(y) && (!p.res || aligned16(p.res)) &&
                             ((M * es) % 16 == 0) && ((x_stride_n * es) % 16 == 0) && ((x_stride_c * es) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    FlatWs wf = {nullptr, nullptr, nullptr};
    const WsLayout wl = ws_layout(N, C, M, es);
    const bool have_flat_ws = workspace && workspace_bytes >= wl.total;
    if (have_flat_ws) {
        wf.slab = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.slab_off);
        wf.piece = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.piece_off);
        wf.ctl = reinterpret_cast<unsigned*>(workspace) + 2;
    }
    const FlatWs* wfp = have_flat_ws ? &wf : nullptr;
    switch (dtype) {
        case MICN_F32: return fwd_by_epi<float>(epilogue, p, can_cluster, wfp, *d, st);
        case MICN_BF16: return fwd_by_epi<__nv_bfloat16>(epilogue, p, can_cluster, wfp, *d, st);
        case MICN_F16: return fwd_by_epi<__half>(epilogue, p, can_cluster, wfp, *d, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

int micn_fwd(const void* x, void* y, const void* residual, const float* const* gamma, const float* const* beta,
             int num_styles, const int64_t* styles, float* save_mean, float* save_rstd, int64_t N, int64_t C, int64_t M,
             int64_t x_stride_n, int64_t x_stride_c, int dtype, int epilogue, float slope, float eps, void* workspace,
             size_t workspace_bytes, void* stream) {
    return fwd_impl(x, y, residual, gamma, beta, num_styles, styles, save_mean, save_rstd, N, C, M, x_stride_n, x_stride_c,
                    dtype, epilogue, slope, nullptr, eps, workspace, workspace_bytes, stream);
}

int micn_fwd_prelu(const void* x, void* y, const void* residual, const float* const* gamma, const float* const* beta,
                   int num_styles, const int64_t* styles, float* save_mean, float* save_rstd, int64_t N, int64_t C,
                   int64_t M, int64_t x_stride_n, int64_t x_stride_c, int dtype, int epilogue, const float* slope_dev,
                   float eps, void* workspace, size_t workspace_bytes, void* stream) {
    if (!slope_dev || epilogue == MICN_EPI_NONE) return MICN_ERR_BAD_ARG;
    return fwd_impl(x, y, residual, gamma, beta, num_styles, styles, save_mean, save_rstd, N, C, M, x_stride_n, x_stride_c,
                    dtype, epilogue, 0.f, slope_dev, eps, workspace, workspace_bytes, stream);
}

static int bwd_impl(const void* dy, const void* x, const void* act_out, const float* const* gamma,
                    const float* const* beta, int num_styles, const int64_t* styles, const float* save_mean,
                    const float* save_rstd, void* dx, void* dresidual, float* dgamma, float* dbeta, int64_t N, int64_t C,
                    int64_t M, int64_t x_stride_n, int64_t x_stride_c, int dtype, int epilogue, float slope,
                    const float* slope_dev, float* dslope_partial, void* workspace, size_t workspace_bytes, void* stream,
                    void* const* peer_bufs = nullptr, int peer_rank = 0, int peer_world = 1, int peer_mode = 1) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if ((dgamma == nullptr) != (dbeta == nullptr)) return MICN_ERR_BAD_ARG;
    if (peer_world > 1) {
        if (!peer_bufs || peer_world > MICN_MAX_PEERS || peer_rank < 0 || peer_rank >= peer_world || !dgamma) return MICN_ERR_BAD_ARG;
        if (N == 0 || C == 0 || M == 0) return MICN_ERR_UNSUPPORTED;  // (every rank must take part: no empty shards)
        for (int r = 0; r < peer_world; ++r)
            if (!peer_bufs[r] || (reinterpret_cast<uintptr_t>(peer_bufs[r]) & 15u)) return MICN_ERR_BAD_ARG;
    }
    if (N == 0 || C == 0 || M == 0) {
        if (dgamma && C > 0) {
            cudaStream_t st = (cudaStream_t)stream;
            cudaMemsetAsync(dgamma, 0, sizeof(float) * num_styles * C, st);
            cudaMemsetAsync(dbeta, 0, sizeof(float) * num_styles * C, st);
        }
        return MICN_OK;
    }
    if (!x || !dy || !dx || !save_mean || !save_rstd) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if (epilogue < MICN_EPI_NONE || epilogue > MICN_EPI_ADD_LRELU) return MICN_ERR_BAD_ARG;
    if (epilogue == MICN_EPI_ADD_LRELU && (!act_out || !dresidual)) return MICN_ERR_BAD_ARG;
    if ((gamma == nullptr) != (beta == nullptr)) return MICN_ERR_BAD_ARG;
    if ((dgamma == nullptr) != (dbeta == nullptr)) return MICN_ERR_BAD_ARG;
    if (N * C > 0x7fffffffLL) return MICN_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
         reinterpret_cast<uintptr_t>(act_out) | reinterpret_cast<uintptr_t>(dresidual)) &
        (es - 1))
        return MICN_ERR_UNALIGNED;
    if (dgamma && (!workspace || workspace_bytes < ws_layout(N, C, 0, es).slab_off)) return MICN_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MICN_ERR_UNALIGNED;  // 16-byte records, 64-bit header word

    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;

    BwdParams p = {};
    p.dy = dy;
    p.x = x;
    p.act_out = epilogue == MICN_EPI_ADD_LRELU ? act_out : nullptr;
    p.affine = gamma != nullptr;
    for (int s = 0; s < num_styles && gamma; ++s) {
        if (!gamma[s] || !beta[s]) return MICN_ERR_BAD_ARG;
        p.gamma[s] = gamma[s];
        p.beta[s] = beta[s];
    }
    p.styles = reinterpret_cast<const long long*>(styles);
    p.save_mean = save_mean;
    p.save_rstd = save_rstd;
    p.dx = dx;
    p.dres = epilogue == MICN_EPI_ADD_LRELU ? dresidual : nullptr;
    p.dgamma = dgamma;
    p.dbeta = dbeta;
    if (workspace) {
        unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
        p.ws_counter = reinterpret_cast<unsigned int*>(w);
        p.status = reinterpret_cast<int*>(w) + 1;
        if (dgamma) {
            p.ws_sum_dy = reinterpret_cast<float*>(w + kWsData);
            p.ws_sum_dyxh = p.ws_sum_dy + N * C;
            p.ws_chan_cnt = (size_t)C <= kWsChanCounters ? reinterpret_cast<unsigned int*>(w + kWsHeader) : nullptr;
        }
    }
    p.N = N;
    p.C = C;
    p.M = M;
    p.x_sN = x_stride_n;
    p.x_sC = x_stride_c;
    p.num_styles = num_styles;
    p.slope = slope;
    p.slope_dev = slope_dev;
    p.dslope = dslope_partial;
    p.xchg_rank = peer_rank;
    p.xchg_world = peer_world > 1 ? peer_world : 1;
    p.xchg_mode = peer_mode | ((int)std::max<long long>(0, g_opt.xchg_dbg.load()) << 4);  // (bring-up bits, see micn_flat.cuh)
    for (int r = 0; r < peer_world && peer_world > 1; ++r) p.xchg_peers[r] = peer_bufs[r];

    const bool can_cluster = d->cc_major >= 9 && aligned16(x) && aligned16(dy) && aligned16(dx) &&
                             (!p.act_out || aligned16(p.act_out)) && (!p.dres || aligned16(p.dres)) &&
                             ((M * es) % 16 == 0) && ((x_stride_n * es) % 16 == 0) && ((x_stride_c * es) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    FlatWs wf = {nullptr, nullptr, nullptr};
    const WsLayout wl = ws_layout(N, C, M, es);
    const bool have_flat_ws = workspace && workspace_bytes >= wl.total;
    if (have_flat_ws) {
        wf.slab = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.slab_off);
        wf.piece = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.piece_off);
        wf.ctl = reinterpret_cast<unsigned*>(workspace) + 2;
    }
    const FlatWs* wfp = have_flat_ws ? &wf : nullptr;
    switch (dtype) {
        case MICN_F32: return bwd_by_epi<float>(epilogue, p, can_cluster, wfp, *d, st);
        case MICN_BF16: return bwd_by_epi<__nv_bfloat16>(epilogue, p, can_cluster, wfp, *d, st);
        case MICN_F16: return bwd_by_epi<__half>(epilogue, p, can_cluster, wfp, *d, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

int micn_bwd(const void* dy, const void* x, const void* act_out, const float* const* gamma, const float* const* beta,
             int num_styles, const int64_t* styles, const float* save_mean, const float* save_rstd, void* dx,
             void* dresidual, float* dgamma, float* dbeta, int64_t N, int64_t C, int64_t M, int64_t x_stride_n,
             int64_t x_stride_c, int dtype, int epilogue, float slope, void* workspace, size_t workspace_bytes,
             void* stream) {
    return bwd_impl(dy, x, act_out, gamma, beta, num_styles, styles, save_mean, save_rstd, dx, dresidual, dgamma, dbeta, N, C,
                    M, x_stride_n, x_stride_c, dtype, epilogue, slope, nullptr, nullptr, workspace, workspace_bytes, stream);
}

size_t micn_peer_buffer_bytes(int64_t C, int num_styles, int world) {
    if (C <= 0 || num_styles < 1 || world < 1) return 0;
    // header, then [4 slots][world source ranks][S*C] 16-byte records
    return (size_t)64 + (size_t)4 * world * num_styles * (size_t)C * 16;
}

int micn_allreduce_fold(void* const* peer_bufs, int rank, int world, int64_t C, int num_styles, float* dgamma, float* dbeta,
                        void* stream) {
    if (!peer_bufs || world < 2 || world > MICN_MAX_PEERS || rank < 0 || rank >= world || C <= 0 || num_styles < 1 ||
        num_styles > MICN_MAX_STYLES || !dgamma || !dbeta)
        return MICN_ERR_BAD_ARG;
    BwdParams p = {};
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r]) return MICN_ERR_BAD_ARG;
        p.xchg_peers[r] = peer_bufs[r];
    }
    p.xchg_rank = rank;
    p.xchg_world = world;
    p.C = C;
    p.num_styles = num_styles;
    p.dgamma = dgamma;
    p.dbeta = dbeta;
    const long long entries = (long long)num_styles * C;
    const unsigned blocks = (unsigned)std::min<long long>(64, (entries + 7) / 8);
    g_opt.launches.fetch_add(1);
    flat1::micn_xchg_fold_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int micn_bwd_allreduce(const void* dy, const void* x, const void* act_out, const float* const* gamma, const float* const* beta,
                       int num_styles, const int64_t* styles, const float* save_mean, const float* save_rstd, void* dx,
                       void* dresidual, float* dgamma, float* dbeta, int64_t N, int64_t C, int64_t M, int64_t x_stride_n,
                       int64_t x_stride_c, int dtype, int epilogue, float slope, void* workspace, size_t workspace_bytes,
                       void* const* peer_bufs, int rank, int world, int fold_mode, void* stream) {
    if (world < 1 || fold_mode < MICN_FOLD_NONE || fold_mode > MICN_FOLD_PREVIOUS) return MICN_ERR_BAD_ARG;
    return bwd_impl(dy, x, act_out, gamma, beta, num_styles, styles, save_mean, save_rstd, dx, dresidual, dgamma, dbeta, N, C,
                    M, x_stride_n, x_stride_c, dtype, epilogue, slope, nullptr, nullptr, workspace, workspace_bytes, stream,
                    peer_bufs, rank, world, fold_mode);
}

int micn_bwd_prelu(const void* dy, const void* x, const void* act_out, const float* const* gamma,
                   const float* const* beta, int num_styles, const int64_t* styles, const float* save_mean,
                   const float* save_rstd, void* dx, void* dresidual, float* dgamma, float* dbeta, int64_t N, int64_t C,
                   int64_t M, int64_t x_stride_n, int64_t x_stride_c, int dtype, int epilogue, const float* slope_dev,
                   float* dslope_partial, void* workspace, size_t workspace_bytes, void* stream) {
    if (!slope_dev || epilogue == MICN_EPI_NONE) return MICN_ERR_BAD_ARG;
    if (dslope_partial && epilogue != MICN_EPI_LRELU) return MICN_ERR_BAD_ARG;
    return bwd_impl(dy, x, act_out, gamma, beta, num_styles, styles, save_mean, save_rstd, dx, dresidual, dgamma, dbeta, N, C,
                    M, x_stride_n, x_stride_c, dtype, epilogue, 0.f, slope_dev, dslope_partial, workspace, workspace_bytes,
                    stream);
}

// ------------------------------------------------------------------------------------------ dual-norm epilogue (C ABI)
int micn_dual_supported(int64_t N, int64_t C, int64_t M, int dtype, int backward) {
    if (N <= 0 || C <= 0 || M <= 0 || N * C > 0x7fffffffLL) return 0;
    DeviceInfo* d = nullptr;
    if (device_info(&d)) return 0;
    if (d->cc_major < 9) return 0;
    switch (dtype) {
        case MICN_F32: return dual_supported_typed<float>(N, C, M, backward, *d);
        case MICN_BF16: return dual_supported_typed<__nv_bfloat16>(N, C, M, backward, *d);
        case MICN_F16: return dual_supported_typed<__half>(N, C, M, backward, *d);
    }
    return 0;
}

int micn_fwd_dual(const void* a, const void* b, void* y, const float* const* gamma_a, const float* const* beta_a,
                  const float* const* gamma_b, const float* const* beta_b, int num_styles, const int64_t* styles,
                  float* save_mean_a, float* save_rstd_a, float* save_mean_b, float* save_rstd_b, int64_t N, int64_t C,
                  int64_t M, int dtype, float slope, float eps, void* workspace, size_t workspace_bytes, void* stream) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if (N == 0 || C == 0 || M == 0) return MICN_OK;
    if (!a || !b || !y) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if ((gamma_a == nullptr) != (beta_a == nullptr) || (gamma_b == nullptr) != (beta_b == nullptr) ||
        (gamma_a == nullptr) != (gamma_b == nullptr))
        return MICN_ERR_BAD_ARG;
    if ((save_mean_a == nullptr) != (save_rstd_a == nullptr) || (save_mean_b == nullptr) != (save_rstd_b == nullptr))
        return MICN_ERR_BAD_ARG;
    if (N * C > 0x7fffffffLL) return MICN_ERR_BAD_ARG;
    if (workspace && workspace_bytes < kWsHeader) return MICN_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MICN_ERR_UNALIGNED;
    if (!aligned16(a) || !aligned16(b) || !aligned16(y) || (M * es) % 16 != 0) return MICN_ERR_UNSUPPORTED;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;
    if (d->cc_major < 9) return MICN_ERR_UNSUPPORTED;
    FwdParams p = {};
    p.x = a;
    p.x2 = b;
    p.y = y;
    p.affine = gamma_a != nullptr;
    for (int s = 0; s < num_styles && gamma_a; ++s) {
        if (!gamma_a[s] || !beta_a[s] || !gamma_b[s] || !beta_b[s]) return MICN_ERR_BAD_ARG;
        p.gamma[s] = gamma_a[s];
        p.beta[s] = beta_a[s];
        p.gamma2[s] = gamma_b[s];
        p.beta2[s] = beta_b[s];
    }
    p.styles = reinterpret_cast<const long long*>(styles);
    p.save_mean = save_mean_a;
    p.save_rstd = save_rstd_a;
    p.save_mean2 = save_mean_b;
    p.save_rstd2 = save_rstd_b;
    p.status = workspace ? reinterpret_cast<int*>(workspace) + 1 : nullptr;
    p.N = N;
    p.C = C;
    p.M = M;
    p.x_sN = C * M;
    p.x_sC = M;
    p.num_styles = num_styles;
    p.eps = eps;
    p.slope = slope;
    cudaStream_t st = (cudaStream_t)stream;
    FlatWs wf = {nullptr, nullptr, nullptr};
    const WsLayout wl = ws_layout(N, C, M, es);
    const bool have_flat_ws = workspace && workspace_bytes >= wl.total;
    if (have_flat_ws) {
        wf.slab = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.slab_off);
        wf.piece = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.piece_off);
        wf.ctl = reinterpret_cast<unsigned*>(workspace) + 2;
    }
    const FlatWs* wfp = have_flat_ws ? &wf : nullptr;
    switch (dtype) {
        case MICN_F32: return fwd_dual_typed<float>(p, wfp, *d, st);
        case MICN_BF16: return fwd_dual_typed<__nv_bfloat16>(p, wfp, *d, st);
        case MICN_F16: return fwd_dual_typed<__half>(p, wfp, *d, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

int micn_bwd_dual(const void* dy, const void* a, const void* b, const float* const* gamma_a, const float* const* beta_a,
                  const float* const* gamma_b, const float* const* beta_b, int num_styles, const int64_t* styles,
                  const float* save_mean_a, const float* save_rstd_a, const float* save_mean_b, const float* save_rstd_b,
                  void* da, void* db, float* dgamma_a, float* dbeta_a, float* dgamma_b, float* dbeta_b, int64_t N,
                  int64_t C, int64_t M, int dtype, float slope, void* workspace, size_t workspace_bytes, void* stream) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if ((dgamma_a == nullptr) != (dbeta_a == nullptr) || (dgamma_b == nullptr) != (dbeta_b == nullptr)) return MICN_ERR_BAD_ARG;
    if (N == 0 || C == 0 || M == 0) {
        cudaStream_t st0 = (cudaStream_t)stream;
        for (float* g : {dgamma_a, dbeta_a, dgamma_b, dbeta_b})
            if (g && C > 0) cudaMemsetAsync(g, 0, sizeof(float) * num_styles * C, st0);
        return MICN_OK;
    }
    if (!dy || !a || !b || !da || !db || !save_mean_a || !save_rstd_a || !save_mean_b || !save_rstd_b) return MICN_ERR_BAD_ARG;
    if ((gamma_a == nullptr) != (beta_a == nullptr) || (gamma_b == nullptr) != (beta_b == nullptr) ||
        (gamma_a == nullptr) != (gamma_b == nullptr))
        return MICN_ERR_BAD_ARG;
    if (N * C > 0x7fffffffLL) return MICN_ERR_BAD_ARG;
    if (dgamma_b && !dgamma_a) return MICN_ERR_BAD_ARG;  // (the second norm's gradients ride on the first's fold)
    if (dgamma_a && (!workspace || workspace_bytes < ws_layout(N, C, 0, es).slab_off)) return MICN_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return MICN_ERR_UNALIGNED;
    if (!aligned16(a) || !aligned16(b) || !aligned16(dy) || !aligned16(da) || !aligned16(db) || (M * es) % 16 != 0)
        return MICN_ERR_UNSUPPORTED;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;
    if (d->cc_major < 9) return MICN_ERR_UNSUPPORTED;
    BwdParams p = {};
    p.dy = dy;
    p.x = a;
    p.x2 = b;
    p.affine = gamma_a != nullptr;
    for (int s = 0; s < num_styles && gamma_a; ++s) {
        if (!gamma_a[s] || !beta_a[s] || !gamma_b[s] || !beta_b[s]) return MICN_ERR_BAD_ARG;
        p.gamma[s] = gamma_a[s];
        p.beta[s] = beta_a[s];
        p.gamma2[s] = gamma_b[s];
        p.beta2[s] = beta_b[s];
    }
    p.styles = reinterpret_cast<const long long*>(styles);
    p.save_mean = save_mean_a;
    p.save_rstd = save_rstd_a;
    p.save_mean2 = save_mean_b;
    p.save_rstd2 = save_rstd_b;
    p.dx = da;
    p.dx2 = db;
    p.dgamma = dgamma_a;
    p.dbeta = dbeta_a;
    p.dgamma2 = dgamma_b;
    p.dbeta2 = dbeta_b;
    if (workspace) {
        unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
        p.ws_counter = reinterpret_cast<unsigned int*>(w);
        p.status = reinterpret_cast<int*>(w) + 1;
        if (dgamma_a) {
            p.ws_sum_dy = reinterpret_cast<float*>(w + kWsData);
            p.ws_sum_dyxh = p.ws_sum_dy + N * C;
            p.ws_sum_dyxh2 = p.ws_sum_dyxh + N * C;
            p.ws_chan_cnt = (size_t)C <= kWsChanCounters ? reinterpret_cast<unsigned int*>(w + kWsHeader) : nullptr;
        }
    }
    p.N = N;
    p.C = C;
    p.M = M;
    p.x_sN = C * M;
    p.x_sC = M;
    p.num_styles = num_styles;
    p.slope = slope;
    cudaStream_t st = (cudaStream_t)stream;
    FlatWs wf = {nullptr, nullptr, nullptr};
    const WsLayout wl = ws_layout(N, C, M, es);
    const bool have_flat_ws = workspace && workspace_bytes >= wl.total;
    if (have_flat_ws) {
        wf.slab = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.slab_off);
        wf.piece = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(workspace) + wl.piece_off);
        wf.ctl = reinterpret_cast<unsigned*>(workspace) + 2;
    }
    const FlatWs* wfp = have_flat_ws ? &wf : nullptr;
    switch (dtype) {
        case MICN_F32: return bwd_dual_typed<float>(p, wfp, *d, st);
        case MICN_BF16: return bwd_dual_typed<__nv_bfloat16>(p, wfp, *d, st);
        case MICN_F16: return bwd_dual_typed<__half>(p, wfp, *d, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

// ------------------------------------------------------------------------------------------ channels-last path (C ABI)
size_t micn_cl_workspace_bytes(int64_t N, int64_t C, int64_t M) {
    (void)M;
    if (N <= 0 || C <= 0) return 0;
    return (cl_ws_bytes(N, C) + 255) & ~(size_t)255;
}

int micn_fwd_cl(const void* x, void* y, const float* const* gamma, const float* const* beta, int num_styles,
                const int64_t* styles, float* save_mean, float* save_rstd, int64_t N, int64_t C, int64_t M, int dtype,
                float eps, void* workspace, size_t workspace_bytes, void* stream) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if (N == 0 || C == 0 || M == 0) return MICN_OK;
    if (!x || !y || (C & 1) || N > 65535) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    // the kernels load channel PAIRS (32-bit for 16-bit types, 64-bit for fp32)
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & (uintptr_t)(2 * es - 1)) return MICN_ERR_UNALIGNED;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;
    ClPlan pl;
    if ((rc = cl_plan(N, C, M, workspace, workspace_bytes, *d, &pl))) return rc;
    ClParams p = {};
    p.x = x;
    p.y = y;
    p.save_mean = save_mean;
    p.save_rstd = save_rstd;
    p.eps = eps;
    if ((rc = cl_fill(p, gamma, beta, num_styles, styles, N, C, M, pl))) return rc;
    const dim3 grid((unsigned)((C + kClTile - 1) / kClTile), (unsigned)pl.MS, (unsigned)N);
    g_opt.last_path.store(3);
    g_opt.last_cs.store(pl.MS);
    g_opt.last_grid.store((long long)grid.x * grid.y * grid.z);
    g_opt.launches.fetch_add(pl.fused ? 1 : 2);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MICN_F32: return cl_fwd_typed<float>(p, grid, pl.fused, st);
        case MICN_BF16: return cl_fwd_typed<__nv_bfloat16>(p, grid, pl.fused, st);
        case MICN_F16: return cl_fwd_typed<__half>(p, grid, pl.fused, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

int micn_bwd_cl(const void* dy, const void* x, const float* const* gamma, const float* const* beta, int num_styles,
                const int64_t* styles, const float* save_mean, const float* save_rstd, void* dx, float* dgamma,
                float* dbeta, int64_t N, int64_t C, int64_t M, int dtype, void* workspace, size_t workspace_bytes,
                void* stream) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N < 0 || C < 0 || M < 0) return MICN_ERR_BAD_ARG;
    if ((dgamma == nullptr) != (dbeta == nullptr)) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if (N == 0 || C == 0 || M == 0) {
        if (dgamma && C > 0) {
            cudaMemsetAsync(dgamma, 0, sizeof(float) * num_styles * C, (cudaStream_t)stream);
            cudaMemsetAsync(dbeta, 0, sizeof(float) * num_styles * C, (cudaStream_t)stream);
        }
        return MICN_OK;
    }
    if (!x || !dy || !dx || !save_mean || !save_rstd || (C & 1) || N > 65535) return MICN_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) &
        (uintptr_t)(2 * es - 1))
        return MICN_ERR_UNALIGNED;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc) return rc;
    ClPlan pl;
    if ((rc = cl_plan(N, C, M, workspace, workspace_bytes, *d, &pl))) return rc;
    ClParams p = {};
    p.x = x;
    p.dy = dy;
    p.y = dx;
    p.save_mean = const_cast<float*>(save_mean);
    p.save_rstd = const_cast<float*>(save_rstd);
    p.dgamma = dgamma;
    p.dbeta = dbeta;
    if ((rc = cl_fill(p, gamma, beta, num_styles, styles, N, C, M, pl))) return rc;
    const dim3 grid((unsigned)((C + kClTile - 1) / kClTile), (unsigned)pl.MS, (unsigned)N);
    g_opt.last_path.store(3);
    g_opt.last_cs.store(pl.MS);
    g_opt.last_grid.store((long long)grid.x * grid.y * grid.z);
    g_opt.launches.fetch_add((pl.fused ? 1 : 2) + ((dgamma && N > 1 && !pl.tile_cnt) ? 1 : 0));
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MICN_F32: return cl_bwd_typed<float>(p, grid, pl.fused, st);
        case MICN_BF16: return cl_bwd_typed<__nv_bfloat16>(p, grid, pl.fused, st);
        case MICN_F16: return cl_bwd_typed<__half>(p, grid, pl.fused, st);
    }
    return MICN_ERR_BAD_DTYPE;
}

// ------------------------------------------------------------------------------------------ host-buffer path
namespace {
constexpr int kHostStreams = 3;
struct HostState {  // per device: role streams and per-group hand-off events, created once and reused
    cudaStream_t hs[kHostStreams] = {};
    std::vector<cudaEvent_t> hev;
};
HostState g_host[64];
std::mutex g_host_mu;

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// Channel-group boundaries b[0..G] of the host-buffer path.  Tapered: groups weighted 1,2,3,4,...,4,3,2,1 so the
// first upload (nothing to overlap it with) and the last download are short.
int host_bounds(int64_t C, std::vector<int64_t>* b) {
    long long want = g_opt.host_groups.load();
    if (want <= 0) want = 10;  // measured at 1x48x96^3 bf16: 6..12 groups within 3 %, 24 and more clearly slower
    int G = (int)std::min<long long>(want, C);
    if (G < 1) G = 1;
    b->assign(1, 0);
    if (g_opt.host_taper.load() == 0 || G < 4) {
        const int64_t cg = (C + G - 1) / G;
        for (int64_t c = cg; c < C; c += cg) b->push_back(c);
    } else {
        auto wt = [&](int g) { return std::min(std::min(g + 1, G - g), 4); };
        long long wsum = 0, acc = 0;
        for (int g = 0; g < G; ++g) wsum += wt(g);
        for (int g = 0; g + 1 < G; ++g) {
            acc += wt(g);
            int64_t c = (C * acc + wsum / 2) / wsum;
            c = std::max<int64_t>(c, b->back() + 1);          // at least one channel per group
            c = std::min<int64_t>(c, C - (G - 1 - g));        // ... and for every group still to come
            b->push_back(c);
        }
    }
    b->push_back(C);
    return (int)b->size() - 1;
}

int64_t host_max_group(const std::vector<int64_t>& b) {
    int64_t m = 1;
    for (size_t g = 0; g + 1 < b.size(); ++g) m = std::max(m, b[g + 1] - b[g]);
    return m;
}

// [N rows of `width` bytes] between a pitched host tensor and a dense device block.  Few rows: one plain copy per
// row (a 1-D copy runs at the full PCIe rate; the pitched 2-D form is for many short rows).
cudaError_t host_copy(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int64_t rows,
                      cudaMemcpyKind kind, cudaStream_t st) {
    if (rows > 8 || g_opt.host_copy_2d.load() == 1) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
    for (int64_t r = 0; r < rows; ++r) {
        cudaError_t e = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + r * dpitch,
                                        static_cast<const unsigned char*>(src) + r * spitch, width, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
}  // namespace

size_t micn_host_scratch_bytes(int64_t N, int64_t C, int64_t M, int dtype, int num_styles, int with_backward) {
    const int es = elem_size(dtype);
    if (!es || N <= 0 || C <= 0 || M <= 0 || num_styles < 1) return 0;
    const size_t E = (size_t)N * C * M * es;
    std::vector<int64_t> cb;
    const int G = host_bounds(C, &cb);
    size_t b = 0;
    b += up256(E) * (with_backward ? 4 : 2);                        // x, y [, dy, dx]
    b += up256((size_t)N * C * 4) * 2;                              // mean, rstd
    b += up256((size_t)num_styles * C * 4) * 4;                     // gamma, beta, dgamma, dbeta
    b += up256((size_t)N * 8);                                      // styles
    b += (size_t)G * micn_workspace_bytes(N, host_max_group(cb), M, dtype, num_styles);
    return b + 4096;
}

int micn_fwd_bwd_host(const void* x_host, const void* dy_host, void* y_host, void* dx_host, const float* gamma_host,
                      const float* beta_host, int num_styles, const int64_t* styles_host, float* dgamma_host,
                      float* dbeta_host, int64_t N, int64_t C, int64_t M, int dtype, int epilogue, float slope, float eps,
                      void* dev_scratch, size_t dev_scratch_bytes) {
    const int es = elem_size(dtype);
    if (!es) return MICN_ERR_BAD_DTYPE;
    if (N <= 0 || C <= 0 || M <= 0 || !x_host || !y_host || !dev_scratch) return MICN_ERR_BAD_ARG;
    if (num_styles < 1) return MICN_ERR_BAD_ARG;
    if (num_styles > MICN_MAX_STYLES) return MICN_ERR_TOO_MANY_STYLES;
    if (epilogue != MICN_EPI_NONE && epilogue != MICN_EPI_LRELU) return MICN_ERR_BAD_ARG;
    const bool bwd = dy_host != nullptr;
    if (bwd && !dx_host) return MICN_ERR_BAD_ARG;
    if ((dgamma_host == nullptr) != (dbeta_host == nullptr)) return MICN_ERR_BAD_ARG;
    if (dev_scratch_bytes < micn_host_scratch_bytes(N, C, M, dtype, num_styles, bwd ? 1 : 0)) return MICN_ERR_WORKSPACE;

    std::lock_guard<std::mutex> lk(g_host_mu);
    cudaError_t e;
    int dev = -1;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return MICN_ERR_NO_DEVICE;
    cudaStream_t* hs = g_host[dev].hs;
    std::vector<cudaEvent_t>& hev = g_host[dev].hev;
    for (int i = 0; i < kHostStreams; ++i)
        if (!hs[i] && (e = cudaStreamCreateWithFlags(&hs[i], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;

    // carve the scratch
    unsigned char* w = reinterpret_cast<unsigned char*>(dev_scratch);
    auto take = [&](size_t bytes) {
        unsigned char* r = w;
        w += up256(bytes);
        return r;
    };
    const size_t E = (size_t)N * C * M * es;
    unsigned char* xd = take(E);
    unsigned char* yd = take(E);
    unsigned char* dyd = bwd ? take(E) : nullptr;
    unsigned char* dxd = bwd ? take(E) : nullptr;
    float* mean = reinterpret_cast<float*>(take((size_t)N * C * 4));
    float* rstd = reinterpret_cast<float*>(take((size_t)N * C * 4));
    const size_t SCb = (size_t)num_styles * C * 4;
    float* gd = reinterpret_cast<float*>(take(SCb));
    float* bd = reinterpret_cast<float*>(take(SCb));
    float* dgd = reinterpret_cast<float*>(take(SCb));
    float* dbd = reinterpret_cast<float*>(take(SCb));
    int64_t* sd = reinterpret_cast<int64_t*>(take((size_t)N * 8));
    std::vector<int64_t> cb;
    const int G = host_bounds(C, &cb);
    const size_t ws_each = micn_workspace_bytes(N, host_max_group(cb), M, dtype, num_styles);
    unsigned char* ws0 = take((size_t)G * ws_each);

    // Three role streams: uploads (H2D), kernels, downloads (D2H).  The upload engine never idles (x and dy of
    // group g, then group g+1, ...), the kernels of group g start as soon as its x has landed, and the download of
    // y / dx trails the kernels: PCIe runs full duplex for the whole call.
    cudaStream_t s_up = hs[0], s_comp = hs[1], s_down = hs[2];
    const bool affine = gamma_host != nullptr;
    if (affine) {
        if ((e = cudaMemcpyAsync(gd, gamma_host, SCb, cudaMemcpyHostToDevice, s_comp)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(bd, beta_host, SCb, cudaMemcpyHostToDevice, s_comp)) != cudaSuccess) return (int)e;
    }
    if (styles_host && (e = cudaMemcpyAsync(sd, styles_host, (size_t)N * 8, cudaMemcpyHostToDevice, s_comp)) != cudaSuccess)
        return (int)e;
    if ((e = cudaMemsetAsync(ws0, 0, (size_t)G * ws_each, s_comp)) != cudaSuccess) return (int)e;
    const size_t need_ev = (size_t)G * 4;
    while (hev.size() < need_ev) {
        cudaEvent_t ev;
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        hev.push_back(ev);
    }

    // bring-up aid: timestamps after every stage of every group (host_trace = 1)
    const bool trace = g_opt.host_trace.load() == 1;
    std::vector<cudaEvent_t> tev;
    auto stamp = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, s);
        tev.push_back(ev);
    };
    stamp(s_up);
    const auto host_t0 = std::chrono::steady_clock::now();
    double host_queued_us = 0.0;

    // channel groups: sub-tensors [N, cg, M] staged densely on the device, 2-D copies on the host side
    int rc = 0;
    const size_t row_pitch = (size_t)C * M * es;
    std::vector<const float*> gp(num_styles), bp(num_styles);
    size_t dev_off = 0;  // dense sub-tensors are packed one after the other
    for (int g = 0; g < G && !rc; ++g) {
        const int64_t c0 = cb[g], c1 = cb[g + 1];
        const int64_t cc = c1 - c0;
        cudaEvent_t ev_x = hev[4 * g], ev_dy = hev[4 * g + 1], ev_y = hev[4 * g + 2], ev_dx = hev[4 * g + 3];
        const size_t width = (size_t)cc * M * es;
        const size_t sub = (size_t)N * width;
        const unsigned char* xh = reinterpret_cast<const unsigned char*>(x_host) + (size_t)c0 * M * es;
        unsigned char* yh = reinterpret_cast<unsigned char*>(y_host) + (size_t)c0 * M * es;
        // ---- uploads
        if ((e = host_copy(xd + dev_off, width, xh, row_pitch, width, N, cudaMemcpyHostToDevice, s_up)) != cudaSuccess) {
            rc = (int)e;
            break;
        }
        cudaEventRecord(ev_x, s_up);
        stamp(s_up);
        if (bwd) {
            const unsigned char* dyh = reinterpret_cast<const unsigned char*>(dy_host) + (size_t)c0 * M * es;
            if ((e = host_copy(dyd + dev_off, width, dyh, row_pitch, width, N, cudaMemcpyHostToDevice, s_up)) !=
                cudaSuccess) {
                rc = (int)e;
                break;
            }
            cudaEventRecord(ev_dy, s_up);
            stamp(s_up);
        }
        // ---- kernels
        for (int s = 0; s < num_styles; ++s) {
            gp[s] = gd + (size_t)s * C + c0;
            bp[s] = bd + (size_t)s * C + c0;
        }
        float* gmean = mean + (size_t)N * c0;  // per-group stats live in the [N*C] arrays at a dense per-group offset
        float* grstd = rstd + (size_t)N * c0;
        unsigned char* ws = ws0 + (size_t)g * ws_each;
        cudaStreamWaitEvent(s_comp, ev_x, 0);
        rc = micn_fwd(xd + dev_off, yd + dev_off, nullptr, affine ? gp.data() : nullptr, affine ? bp.data() : nullptr,
                      num_styles, styles_host ? sd : nullptr, gmean, grstd, N, cc, M, cc * M, M, dtype, epilogue, slope, eps,
                      ws, ws_each, s_comp);
        if (rc) break;
        cudaEventRecord(ev_y, s_comp);
        stamp(s_comp);
        if (bwd) {
            // group-dense [S, cc] gradient blocks, scattered into [S, C] on the host afterwards
            float* gdg = dgamma_host ? dgd + (size_t)num_styles * c0 : nullptr;
            float* gdb = dgamma_host ? dbd + (size_t)num_styles * c0 : nullptr;
            cudaStreamWaitEvent(s_comp, ev_dy, 0);
            rc = micn_bwd(dyd + dev_off, xd + dev_off, nullptr, affine ? gp.data() : nullptr, affine ? bp.data() : nullptr,
                          num_styles, styles_host ? sd : nullptr, gmean, grstd, dxd + dev_off, nullptr, gdg, gdb, N, cc, M,
                          cc * M, M, dtype, epilogue, slope, ws, ws_each, s_comp);
            if (rc) break;
            cudaEventRecord(ev_dx, s_comp);
            stamp(s_comp);
        }
        // ---- downloads
        cudaStreamWaitEvent(s_down, ev_y, 0);
        if ((e = host_copy(yh, row_pitch, yd + dev_off, width, width, N, cudaMemcpyDeviceToHost, s_down)) != cudaSuccess) {
            rc = (int)e;
            break;
        }
        stamp(s_down);
        if (bwd) {
            unsigned char* dxh = reinterpret_cast<unsigned char*>(dx_host) + (size_t)c0 * M * es;
            cudaStreamWaitEvent(s_down, ev_dx, 0);
            if ((e = host_copy(dxh, row_pitch, dxd + dev_off, width, width, N, cudaMemcpyDeviceToHost, s_down)) !=
                cudaSuccess) {
                rc = (int)e;
                break;
            }
            stamp(s_down);
        }
        dev_off += sub;  // dense packing: the groups tile [0, E) exactly
    }
    if (trace) host_queued_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
    std::vector<float> tg, tb;
    if (!rc && bwd && dgamma_host) {
        tg.resize((size_t)num_styles * C);
        tb.resize((size_t)num_styles * C);
        for (int i = 0; i < kHostStreams; ++i) cudaStreamSynchronize(hs[i]);
        if ((e = cudaMemcpy(tg.data(), dgd, SCb, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = (int)e;
        if (!rc && (e = cudaMemcpy(tb.data(), dbd, SCb, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = (int)e;
        for (int g = 0; g < G && !rc; ++g) {
            const int64_t c0 = cb[g], cc = cb[g + 1] - cb[g];
            for (int s = 0; s < num_styles; ++s)
                for (int64_t c = 0; c < cc; ++c) {
                    dgamma_host[(size_t)s * C + c0 + c] = tg[(size_t)num_styles * c0 + (size_t)s * cc + c];
                    dbeta_host[(size_t)s * C + c0 + c] = tb[(size_t)num_styles * c0 + (size_t)s * cc + c];
                }
        }
    }
    for (int i = 0; i < kHostStreams; ++i) {
        e = cudaStreamSynchronize(hs[i]);
        if (e != cudaSuccess && !rc) rc = (int)e;
    }
    if (trace && !rc && bwd && tev.size() == 1 + 6 * (size_t)G) {
        std::fprintf(stderr, "micn host timeline (us after the first upload was queued; the host had queued everything after %.0f us): group x_up dy_up fwd bwd y_down dx_down\n", host_queued_us);
        for (int g = 0; g < G; ++g) {
            std::fprintf(stderr, "  %2d", g);
            for (int k = 0; k < 6; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, tev[0], tev[1 + 6 * g + k]);
                std::fprintf(stderr, " %8.1f", ms * 1e3f);
            }
            std::fprintf(stderr, "\n");
        }
    }
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    return rc;
}

}  // extern "C"
