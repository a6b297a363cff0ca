// micn_selftest.cu - standalone correctness + timing harness for libmicn.so (TEST INFRASTRUCTURE).
//
// Calls the kernels only through the C ABI of include/micn.h and checks them against a float64
// CPU computation written here (same closed form as oracle/micn_oracle.py).  Used on the GPU box
// for fast kernel bring-up and parameter sweeps: no Python / torch start-up cost.
//
//   micn_selftest --suite correctness            exit code != 0 on any mismatch
//   micn_selftest --suite perf [--out f.jsonl]   CUDA-event timings, rotating buffers > L2
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/micn.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(3);                                                                               \
        }                                                                                          \
    } while (0)

static int g_threads = 8;

template <typename F>
static void parallel_for(long long n, F f) {
    int nt = (int)std::min<long long>(g_threads, n);
    if (nt <= 1) {
        for (long long i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() {
            for (long long i = t; i < n; i += nt) f(i);
        });
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------------ dtype helpers
static size_t esize(int dt) { return dt == MICN_F32 ? 4 : 2; }
static const char* dname(int dt) { return dt == MICN_F32 ? "fp32" : dt == MICN_BF16 ? "bf16" : "fp16"; }

static void store_elem(void* base, size_t i, int dt, float v) {
    if (dt == MICN_F32)
        ((float*)base)[i] = v;
    else if (dt == MICN_BF16)
        ((__nv_bfloat16*)base)[i] = __float2bfloat16_rn(v);
    else
        ((__half*)base)[i] = __float2half_rn(v);
}
static float load_elem(const void* base, size_t i, int dt) {
    if (dt == MICN_F32) return ((const float*)base)[i];
    if (dt == MICN_BF16) return __bfloat162float(((const __nv_bfloat16*)base)[i]);
    return __half2float(((const __half*)base)[i]);
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t next() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (uint32_t)(s >> 32);
    }
    float uni() { return (next() >> 8) * (1.0f / 16777216.0f); }
    float normal() {  // sum of 4 uniforms, variance-normalised: cheap and good enough for tests
        float a = uni() + uni() + uni() + uni() - 2.0f;
        return a * 1.7320508f;
    }
};

// ------------------------------------------------------------------------------------------------ test case
struct Case {
    long long N, C, M;
    int dtype, epi;
    long long pad_c = 0;      // extra elements between channel slabs of x (strided input)
    long long misalign = 0;   // element offset applied to every buffer (unaligned path)
    int cs = -1, slots = -1, max_clusters = -1, force_path = -1, tps = -1;
    int fslots = -1, flag = -1, fpv = -1, fgrid = -1, fovh = -1, fcoop = -1, fpd = -1, fpb = -1;  // flat path knobs
    int num_styles = 2;
    bool affine = true;
    float mean = 1.0f, stdv = 2.0f;
    std::string tag;
};

struct Buffers {
    void *x = nullptr, *y = nullptr, *dy = nullptr, *dx = nullptr, *res = nullptr, *dres = nullptr;
    float *mean = nullptr, *rstd = nullptr, *gamma = nullptr, *beta = nullptr, *dgamma = nullptr, *dbeta = nullptr;
    int64_t* styles = nullptr;
    void* ws = nullptr;
    size_t ws_bytes = 0;
};

static double relerr(const std::vector<double>& ref, const void* got, int dt, size_t n, size_t off = 0) {
    double mx = 0, den = 0;
    for (size_t i = 0; i < n; ++i) {
        den = std::max(den, std::fabs(ref[i]));
        double dl = std::fabs(ref[i] - (double)load_elem(got, off + i, dt));
        if (!(dl == dl)) dl = 1e30;  // NaN (e.g. an element the kernel never wrote) must fail loudly
        mx = std::max(mx, dl);
    }
    return mx / (den > 0 ? den : 1.0);
}
static double relerr_f(const std::vector<double>& ref, const float* got, size_t n) {
    double mx = 0, den = 0;
    for (size_t i = 0; i < n; ++i) {
        den = std::max(den, std::fabs(ref[i]));
        double dl = std::fabs(ref[i] - (double)got[i]);
        if (!(dl == dl)) dl = 1e30;
        mx = std::max(mx, dl);
    }
    return mx / (den > 0 ? den : 1.0);
}

// generic library knobs from the command line (--opt name=value), applied after the per-case ones
static std::vector<std::pair<std::string, long long>> g_extra_opts;

static void set_opts(const Case& c) {
    micn_set_option("cluster_size", c.cs);
    micn_set_option("slots", c.slots);
    micn_set_option("max_clusters", c.max_clusters);
    micn_set_option("force_path", c.force_path);
    micn_set_option("small_tps", c.tps);
    micn_set_option("flat_slots", c.fslots);
    micn_set_option("flat_lag", c.flag);
    micn_set_option("flat_piece_vecs", c.fpv);
    micn_set_option("flat_grid", c.fgrid);
    micn_set_option("flat_ovh_vecs", c.fovh);
    micn_set_option("flat_coop", c.fcoop);
    micn_set_option("flat_poll_delay_ns", c.fpd);
    micn_set_option("flat_poll_backoff_ns", c.fpb);
    for (auto& o : g_extra_opts) micn_set_option(o.first.c_str(), o.second);
}

static int run_correctness(const Case& c, bool verbose) {
    const long long N = c.N, C = c.C, M = c.M, S = c.num_styles;
    const int dt = c.dtype;
    const size_t es = esize(dt);
    const long long sC = M + c.pad_c, sN = C * sC;
    const size_t xe = (size_t)N * sN + c.misalign, de = (size_t)N * C * M + c.misalign;
    const size_t off = c.misalign;

    // host data
    std::vector<unsigned char> hx(xe * es), hdy(de * es), hres(de * es);
    std::vector<float> hg(S * C, 1.f), hb(S * C, 0.f);
    std::vector<int64_t> hst(N);
    Rng r(1234 + N * 7 + C * 13 + M);
    for (size_t i = 0; i < xe; ++i) store_elem(hx.data(), i, dt, r.normal() * c.stdv + c.mean);
    for (size_t i = 0; i < de; ++i) store_elem(hdy.data(), i, dt, r.normal());
    for (size_t i = 0; i < de; ++i) store_elem(hres.data(), i, dt, r.normal() * 0.7f);
    if (c.affine)
        for (long long i = 0; i < S * C; ++i) {
            hg[i] = 1.f + 0.3f * r.normal();
            hb[i] = 0.3f * r.normal();
        }
    for (long long n = 0; n < N; ++n) hst[n] = (n % S) - ((n % 3 == 2) ? S : 0);  // some negative (python-style) ids

    // float64 reference, forward part
    std::vector<double> ry((size_t)N * C * M), rdx((size_t)N * C * M), rdres((size_t)N * C * M);
    std::vector<double> rmean(N * C), rrstd(N * C), rs1(N * C), rs2(N * C), rdg(S * C, 0.0), rdb(S * C, 0.0);
    const double slope = 0.01, eps = 1e-5;
    parallel_for(N * C, [&](long long slab) {
        const long long n = slab / C, ch = slab % C;
        long long st = hst[n];
        if (st < 0) st += S;
        const double g = hg[st * C + ch], b = hb[st * C + ch];
        const size_t xo = off + n * sN + ch * sC, yo = (size_t)slab * M;
        double sum = 0;
        for (long long m = 0; m < M; ++m) sum += load_elem(hx.data(), xo + m, dt);
        const double mean = sum / M;
        double var = 0;
        for (long long m = 0; m < M; ++m) {
            const double d = load_elem(hx.data(), xo + m, dt) - mean;
            var += d * d;
        }
        var /= M;
        const double rstd = 1.0 / std::sqrt(var + eps);
        rmean[slab] = mean;
        rrstd[slab] = rstd;
        for (long long m = 0; m < M; ++m) {
            const double xh = (load_elem(hx.data(), xo + m, dt) - mean) * rstd;
            double pre = xh * g + b;
            if (c.epi == MICN_EPI_ADD_LRELU) pre += load_elem(hres.data(), off + yo + m, dt);
            ry[yo + m] = (c.epi == MICN_EPI_NONE) ? pre : (pre > 0 ? pre : pre * slope);
        }
    });
    // backward part.  LeakyReLU's derivative is discontinuous at 0: an element whose pre-activation
    // rounds to the other side of 0 in fp32 than in float64 would flip its mask.  That is a
    // measure-zero ambiguity of the function, not a kernel property, so the reference takes the
    // mask exactly as the kernel defines it: sign of fmaf(x - mean, rstd*gamma, beta) in fp32 from
    // the saved statistics (LRELU), or sign of the stored forward output (ADD_LRELU).
    auto reference_backward = [&](const std::vector<float>& gmean, const std::vector<float>& grstd,
                                  const std::vector<unsigned char>& gy) {
        std::fill(rdg.begin(), rdg.end(), 0.0);
        std::fill(rdb.begin(), rdb.end(), 0.0);
        parallel_for(N * C, [&](long long slab) {
            const long long n = slab / C, ch = slab % C;
            long long st = hst[n];
            if (st < 0) st += S;
            const double g = hg[st * C + ch];
            const size_t xo = off + n * sN + ch * sC, yo = (size_t)slab * M;
            const double mean = rmean[slab], rstd = rrstd[slab];
            const float mean_f = gmean[slab], a_f = grstd[slab] * hg[st * C + ch], beta_f = hb[st * C + ch];
            double s1 = 0, s2 = 0;
            for (long long m = 0; m < M; ++m) {
                const float xf = load_elem(hx.data(), xo + m, dt);
                const double xh = ((double)xf - mean) * rstd;
                double gg = load_elem(hdy.data(), off + yo + m, dt);
                if (c.epi == MICN_EPI_LRELU) {
                    bool pos;
                    if (dt == MICN_F32) {
                        pos = fmaf(xf - mean_f, a_f, beta_f) > 0.f;
                    } else {
                        // 16-bit kernels compare x against the threshold T = mean - beta/a rounded DOWN to the
                        // element type (micn_cluster.cuh: bwd_slab_consts); a < 0 negates both sides
                        float T;
                        bool flip = false;
                        if (a_f > 0.f)
                            T = mean_f - beta_f / a_f;
                        else if (a_f < 0.f) {
                            T = -(mean_f - beta_f / a_f);
                            flip = true;
                        } else
                            T = beta_f > 0.f ? -INFINITY : INFINITY;
                        const float Td = dt == MICN_BF16 ? __bfloat162float(__float2bfloat16_rd(T)) : __half2float(__float2half_rd(T));
                        pos = (flip ? -xf : xf) > Td;
                    }
                    gg *= pos ? 1.0 : slope;
                }
                if (c.epi == MICN_EPI_ADD_LRELU) gg *= (load_elem(gy.data(), off + yo + m, dt) > 0.f ? 1.0 : slope);
                rdres[yo + m] = gg;
                s1 += gg;
                s2 += gg * xh;
            }
            rs1[slab] = s1;
            rs2[slab] = s2;
            for (long long m = 0; m < M; ++m) {
                const double xh = (load_elem(hx.data(), xo + m, dt) - mean) * rstd;
                rdx[yo + m] = g * rstd * (rdres[yo + m] - s1 / M - xh * s2 / M);
            }
        });
        for (long long n = 0; n < N; ++n) {
            long long st = hst[n];
            if (st < 0) st += S;
            for (long long ch = 0; ch < C; ++ch) {
                rdb[st * C + ch] += rs1[n * C + ch];
                rdg[st * C + ch] += rs2[n * C + ch];
            }
        }
    };

    // device
    Buffers d;
    CK(cudaMalloc(&d.x, xe * es));
    CK(cudaMalloc(&d.y, de * es));
    CK(cudaMalloc(&d.dy, de * es));
    CK(cudaMalloc(&d.dx, de * es));
    CK(cudaMalloc(&d.res, de * es));
    CK(cudaMalloc(&d.dres, de * es));
    CK(cudaMalloc(&d.mean, N * C * 4));
    CK(cudaMalloc(&d.rstd, N * C * 4));
    CK(cudaMalloc(&d.gamma, S * C * 4));
    CK(cudaMalloc(&d.beta, S * C * 4));
    CK(cudaMalloc(&d.dgamma, S * C * 4));
    CK(cudaMalloc(&d.dbeta, S * C * 4));
    CK(cudaMalloc(&d.styles, N * 8));
    d.ws_bytes = micn_workspace_bytes(N, C, M, dt, (int)S);
    CK(cudaMalloc(&d.ws, d.ws_bytes));
    CK(cudaMemset(d.ws, 0, d.ws_bytes));
    CK(cudaMemcpy(d.x, hx.data(), xe * es, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.dy, hdy.data(), de * es, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.res, hres.data(), de * es, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.gamma, hg.data(), S * C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.beta, hb.data(), S * C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.styles, hst.data(), N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(d.y, 0xff, de * es));
    CK(cudaMemset(d.dx, 0xff, de * es));
    CK(cudaMemset(d.dgamma, 0xff, S * C * 4));

    std::vector<const float*> gp(S), bp(S);
    for (long long s = 0; s < S; ++s) {
        gp[s] = d.gamma + s * C;
        bp[s] = d.beta + s * C;
    }
    auto P = [&](void* p) { return (void*)((unsigned char*)p + off * es); };
    set_opts(c);
    int fails = 0;
    // run twice: the second run checks that the self-resetting workspace really is reusable
    for (int rep = 0; rep < 2; ++rep) {
        int rc = micn_fwd(P(d.x), P(d.y), P(d.res), c.affine ? gp.data() : nullptr, c.affine ? bp.data() : nullptr, (int)S,
                          d.styles, d.mean, d.rstd, N, C, M, sN, sC, dt, c.epi, (float)slope, (float)eps, d.ws, d.ws_bytes,
                          nullptr);
        if (rc) {
            printf("FAIL %s: micn_fwd rc=%d (%s)\n", c.tag.c_str(), rc, micn_error_string(rc));
            return 1;
        }
        const long long fpath = micn_get_option("last_path"), fcs = micn_get_option("last_cs"),
                        fslots = micn_get_option("last_slots"), fgrid = micn_get_option("last_grid");
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("FAIL %s: forward kernel error %s\n", c.tag.c_str(), cudaGetErrorString(e));
            exit(4);  // context is gone
        }
        std::vector<unsigned char> gy(de * es), gdx(de * es), gdres(de * es);
        std::vector<float> gmean(N * C), grstd(N * C), gdg(S * C), gdb(S * C);
        CK(cudaMemcpy(gy.data(), d.y, de * es, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gmean.data(), d.mean, N * C * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(grstd.data(), d.rstd, N * C * 4, cudaMemcpyDeviceToHost));
        if (rep == 0) reference_backward(gmean, grstd, gy);
        rc = micn_bwd(P(d.dy), P(d.x), P(d.y), c.affine ? gp.data() : nullptr, c.affine ? bp.data() : nullptr, (int)S,
                      d.styles, d.mean, d.rstd, P(d.dx), P(d.dres), d.dgamma, d.dbeta, N, C, M, sN, sC, dt, c.epi,
                      (float)slope, d.ws, d.ws_bytes, nullptr);
        if (rc) {
            printf("FAIL %s: micn_bwd rc=%d (%s)\n", c.tag.c_str(), rc, micn_error_string(rc));
            return 1;
        }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("FAIL %s: backward kernel error %s\n", c.tag.c_str(), cudaGetErrorString(e));
            exit(4);
        }
        const long long bpath = micn_get_option("last_path"), bcs = micn_get_option("last_cs"),
                        bslots = micn_get_option("last_slots"), bgrid = micn_get_option("last_grid");
        CK(cudaMemcpy(gdx.data(), d.dx, de * es, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gdres.data(), d.dres, de * es, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gdg.data(), d.dgamma, S * C * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gdb.data(), d.dbeta, S * C * 4, cudaMemcpyDeviceToHost));
        int status = 0;
        micn_read_status(d.ws, nullptr, &status);
        const size_t tot = (size_t)N * C * M;
        const double ey = relerr(ry, gy.data(), dt, tot, off), edx = relerr(rdx, gdx.data(), dt, tot, off);
        const double edr = c.epi == MICN_EPI_ADD_LRELU ? relerr(rdres, gdres.data(), dt, tot, off) : 0.0;
        const double em = relerr_f(rmean, gmean.data(), N * C), er = relerr_f(rrstd, grstd.data(), N * C);
        const double edg = relerr_f(rdg, gdg.data(), S * C), edb = relerr_f(rdb, gdb.data(), S * C);
        // tolerances: BASELINE.json north_star (1e-5 fp32, 1e-2 half precision); conditioning factor for |mean|/std
        const double cond = std::max(1.0, std::fabs((double)c.mean) / c.stdv / 0.5);
        const double tol = (dt == MICN_F32 ? 1e-5 : 1e-2) * cond;
        const double tol_stat = 1e-5 * cond, tol_param = (dt == MICN_F32 ? 2e-5 : 2e-3) * cond;
        const bool ok = ey < tol && edx < tol && edr < tol && em < tol_stat && er < tol_stat && edg < tol_param &&
                        edb < tol_param && status == 0;
        if (!ok || verbose)
            printf("%s %-34s rep%d N=%lld C=%lld M=%lld %s epi=%d | fwd path=%lld cs/tps=%lld S=%lld g=%lld | bwd path=%lld "
                   "cs/tps=%lld S=%lld g=%lld | y %.2e dx %.2e dres %.2e mean %.2e rstd %.2e dgamma %.2e dbeta %.2e status %d\n",
                   ok ? "ok  " : "FAIL", c.tag.c_str(), rep, N, C, M, dname(dt), c.epi, fpath, fcs, fslots, fgrid, bpath, bcs,
                   bslots, bgrid, ey, edx, edr, em, er, edg, edb, status);
        fails += ok ? 0 : 1;
        CK(cudaMemset(d.y, 0xff, de * es));
        CK(cudaMemset(d.dx, 0xff, de * es));
        CK(cudaMemset(d.dgamma, 0xff, S * C * 4));
    }
    cudaFree(d.x); cudaFree(d.y); cudaFree(d.dy); cudaFree(d.dx); cudaFree(d.res); cudaFree(d.dres);
    cudaFree(d.mean); cudaFree(d.rstd); cudaFree(d.gamma); cudaFree(d.beta); cudaFree(d.dgamma); cudaFree(d.dbeta);
    cudaFree(d.styles); cudaFree(d.ws);
    return fails;
}

// ------------------------------------------------------------------------------------------------ perf
struct PerfResult {
    double fwd_us, bwd_us;
    long long f_cs, f_slots, f_grid, b_cs, b_slots, b_grid, f_path, b_path;
};

static PerfResult run_perf(const Case& c, int iters, int warm) {
    const long long N = c.N, C = c.C, M = c.M, S = c.num_styles;
    const int dt = c.dtype;
    const size_t es = esize(dt);
    const size_t E = (size_t)N * C * M, bytes = E * es;
    // rotate through enough buffer sets that the working set is > 2x the 126 MB L2
    int R = (int)std::max<size_t>(2, (size_t)(320e6 / (double)(bytes * 2)) + 1);
    if (R > 8) R = 8;
    std::vector<void*> x(R), y(R), dy(R), dx(R), res(R), dres(R);
    std::vector<unsigned char> h(bytes);
    Rng r(99);
    for (size_t i = 0; i < E; ++i) store_elem(h.data(), i, dt, r.normal() * 2.f + 1.f);
    for (int i = 0; i < R; ++i) {
        CK(cudaMalloc(&x[i], bytes));
        CK(cudaMalloc(&y[i], bytes));
        CK(cudaMalloc(&dy[i], bytes));
        CK(cudaMalloc(&dx[i], bytes));
        CK(cudaMemcpy(x[i], h.data(), bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dy[i], h.data(), bytes, cudaMemcpyHostToDevice));
        if (c.epi == MICN_EPI_ADD_LRELU) {
            CK(cudaMalloc(&res[i], bytes));
            CK(cudaMalloc(&dres[i], bytes));
            CK(cudaMemcpy(res[i], h.data(), bytes, cudaMemcpyHostToDevice));
        } else {
            res[i] = dres[i] = nullptr;
        }
    }
    float *mean, *rstd, *gamma, *beta, *dgamma, *dbeta;
    int64_t* styles;
    void* ws;
    const size_t wsb = micn_workspace_bytes(N, C, M, dt, (int)S);
    CK(cudaMalloc(&mean, N * C * 4));
    CK(cudaMalloc(&rstd, N * C * 4));
    CK(cudaMalloc(&gamma, S * C * 4));
    CK(cudaMalloc(&beta, S * C * 4));
    CK(cudaMalloc(&dgamma, S * C * 4));
    CK(cudaMalloc(&dbeta, S * C * 4));
    CK(cudaMalloc(&styles, N * 8));
    CK(cudaMalloc(&ws, wsb));
    CK(cudaMemset(ws, 0, wsb));
    std::vector<float> ones(S * C, 1.0f);
    std::vector<int64_t> hst(N);
    for (long long n = 0; n < N; ++n) hst[n] = n % S;
    CK(cudaMemcpy(gamma, ones.data(), S * C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(beta, 0, S * C * 4));
    CK(cudaMemcpy(styles, hst.data(), N * 8, cudaMemcpyHostToDevice));
    std::vector<const float*> gp(S), bp(S);
    for (long long s = 0; s < S; ++s) {
        gp[s] = gamma + s * C;
        bp[s] = beta + s * C;
    }
    set_opts(c);
    PerfResult pr = {};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto fwd = [&](int i) {
        int rc = micn_fwd(x[i], y[i], res[i], gp.data(), bp.data(), (int)S, styles, mean, rstd, N, C, M, C * M, M, dt, c.epi,
                          0.01f, 1e-5f, ws, wsb, nullptr);
        if (rc) {
            printf("perf fwd rc=%d %s\n", rc, micn_error_string(rc));
            exit(5);
        }
    };
    auto bwd = [&](int i) {
        int rc = micn_bwd(dy[i], x[i], y[i], gp.data(), bp.data(), (int)S, styles, mean, rstd, dx[i], dres[i], dgamma, dbeta,
                          N, C, M, C * M, M, dt, c.epi, 0.01f, ws, wsb, nullptr);
        if (rc) {
            printf("perf bwd rc=%d %s\n", rc, micn_error_string(rc));
            exit(5);
        }
    };
    for (int i = 0; i < warm; ++i) fwd(i % R);
    pr.f_path = micn_get_option("last_path");
    pr.f_cs = micn_get_option("last_cs");
    pr.f_slots = micn_get_option("last_slots");
    pr.f_grid = micn_get_option("last_grid");
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) fwd(i % R);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    pr.fwd_us = ms * 1e3 / iters;
    for (int i = 0; i < warm; ++i) bwd(i % R);
    pr.b_path = micn_get_option("last_path");
    pr.b_cs = micn_get_option("last_cs");
    pr.b_slots = micn_get_option("last_slots");
    pr.b_grid = micn_get_option("last_grid");
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) bwd(i % R);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    pr.bwd_us = ms * 1e3 / iters;
    for (int i = 0; i < R; ++i) {
        cudaFree(x[i]); cudaFree(y[i]); cudaFree(dy[i]); cudaFree(dx[i]);
        if (res[i]) cudaFree(res[i]);
        if (dres[i]) cudaFree(dres[i]);
    }
    cudaFree(mean); cudaFree(rstd); cudaFree(gamma); cudaFree(beta); cudaFree(dgamma); cudaFree(dbeta);
    cudaFree(styles); cudaFree(ws);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return pr;
}

static Case mk(const char* tag, long long N, long long C, long long M, int dt, int epi) {
    Case c;
    c.N = N; c.C = C; c.M = M; c.dtype = dt; c.epi = epi; c.tag = tag;
    return c;
}

int main(int argc, char** argv) {
    std::string suite = "correctness", out;
    bool verbose = false;
    double peak = 6542.1;
    long long oN = 1, oC = 48, oS = 96, oM = -1;
    int odt = MICN_BF16, oepi = MICN_EPI_NONE, ocs = -1, oslots = -1, oiters = 30, omaxcl = -1;
    int opath = -1, ofslots = -1, oflag = -1, ofpv = -1, ofgrid = -1, ofovh = -1, ofcoop = -1, ofpd = -1, ofpb = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--suite") && i + 1 < argc) suite = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--peak") && i + 1 < argc) peak = atof(argv[++i]);
        else if (!strcmp(argv[i], "-v")) verbose = true;
        else if (!strcmp(argv[i], "--N") && i + 1 < argc) oN = atoll(argv[++i]);
        else if (!strcmp(argv[i], "--C") && i + 1 < argc) oC = atoll(argv[++i]);
        else if (!strcmp(argv[i], "--S") && i + 1 < argc) oS = atoll(argv[++i]);
        else if (!strcmp(argv[i], "--M") && i + 1 < argc) oM = atoll(argv[++i]);
        else if (!strcmp(argv[i], "--dtype") && i + 1 < argc) { ++i; odt = !strcmp(argv[i], "fp32") ? MICN_F32 : !strcmp(argv[i], "fp16") ? MICN_F16 : MICN_BF16; }
        else if (!strcmp(argv[i], "--epi") && i + 1 < argc) oepi = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--cs") && i + 1 < argc) ocs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slots") && i + 1 < argc) oslots = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--maxcl") && i + 1 < argc) omaxcl = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--iters") && i + 1 < argc) oiters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--path") && i + 1 < argc) opath = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fslots") && i + 1 < argc) ofslots = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--flag") && i + 1 < argc) oflag = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fpv") && i + 1 < argc) ofpv = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fgrid") && i + 1 < argc) ofgrid = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fovh") && i + 1 < argc) ofovh = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fcoop") && i + 1 < argc) ofcoop = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fpd") && i + 1 < argc) ofpd = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fpb") && i + 1 < argc) ofpb = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--opt") && i + 1 < argc) {
            std::string kv = argv[++i];
            const size_t eq = kv.find('=');
            if (eq != std::string::npos) g_extra_opts.emplace_back(kv.substr(0, eq), atoll(kv.c_str() + eq + 1));
        }
    }
    g_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (g_threads > 32) g_threads = 32;
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s, %d SMs, smem opt-in %zu, L2 %d MB, micn %d, host threads %d\n", prop.name, prop.multiProcessorCount,
           prop.sharedMemPerBlockOptin, prop.l2CacheSize >> 20, micn_version(), g_threads);

    int fails = 0;
    if (suite == "correctness" || suite == "all") {
        std::vector<Case> cs;
        const int dts[3] = {MICN_F32, MICN_BF16, MICN_F16};
        const int epis[3] = {MICN_EPI_NONE, MICN_EPI_LRELU, MICN_EPI_ADD_LRELU};
        // --- small path: tiny / odd / unaligned slabs
        for (int dt : dts)
            for (int epi : epis) {
                cs.push_back(mk("small_warp_27", 3, 20, 27, dt, epi));
                cs.push_back(mk("small_warp_216", 2, 12, 216, dt, epi));
                cs.push_back(mk("small_cta_1728", 2, 8, 1728, dt, epi));
                cs.push_back(mk("small_cta_13824", 2, 6, 13824, dt, epi));
                Case u = mk("small_misaligned_1001", 2, 5, 1001, dt, epi);
                u.misalign = 3;
                cs.push_back(u);
                Case s = mk("small_strided_4096", 2, 4, 4096, dt, epi);
                s.pad_c = 24;
                cs.push_back(s);
            }
        {
            Case b = mk("small_big_odd_1024tps", 1, 3, 95 * 95 * 95, MICN_BF16, MICN_EPI_LRELU);
            cs.push_back(b);
            Case n1 = mk("small_nonaffine", 2, 6, 1728, MICN_F32, MICN_EPI_NONE);
            n1.affine = false;
            cs.push_back(n1);
            Case bm = mk("small_bigmean", 2, 4, 4096, MICN_F32, MICN_EPI_NONE);
            bm.mean = 50.f; bm.stdv = 0.1f;
            cs.push_back(bm);
            Case s3 = mk("small_3styles", 5, 4, 512, MICN_F32, MICN_EPI_LRELU);
            s3.num_styles = 3;
            cs.push_back(s3);
        }
        // --- cluster path
        // --- flat path (default for aligned slabs >= 32 KB)
        for (int dt : dts)
            for (int epi : epis) {
                cs.push_back(mk("flat_auto_48^3", 2, 6, 110592, dt, epi));
                cs.push_back(mk("flat_auto_96^3", 1, 3, 884736, dt, epi));
                Case a = mk("flat_smallgrid_manyrounds", 3, 5, 65536, dt, epi);  // 7 CTAs, dozens of rounds each
                a.fgrid = 7; a.fpv = 256;
                cs.push_back(a);
                Case b = mk("flat_ragged_pieces", 2, 3, 8 * (512 * 3 + 5) * 3 + 8, dt, epi);  // V not a multiple of anything
                b.fpv = 640;
                cs.push_back(b);
                Case c2 = mk("flat_tinypieces_emptywarps", 2, 4, 16384, dt, epi);  // pieces < 512 vectors: idle warps
                c2.fpv = 200; c2.force_path = 2;
                cs.push_back(c2);
                Case d2 = mk("flat_lag1_slots2", 2, 4, 110592, dt, epi);
                d2.flag = 1; d2.fslots = 2;
                cs.push_back(d2);
                Case e2 = mk("flat_128^3_multiround", 1, 1, 2097152, dt, epi);  // one slab spanning several rounds
                e2.fslots = 8; e2.fpv = 512;
                cs.push_back(e2);
                e2 = mk("flat_lag12_slots8", 2, 3, 110592, dt, epi);
                e2.flag = 12; e2.fslots = 8; e2.fpv = 300;
                cs.push_back(e2);
                e2 = mk("flat_lag3_slots3", 2, 3, 65536, dt, epi);
                e2.flag = 3; e2.fslots = 3;
                cs.push_back(e2);
                Case f2 = mk("flat_strided_x", 2, 3, 65536, dt, epi);
                f2.pad_c = 64;
                cs.push_back(f2);
                Case g2 = mk("flat_5samples_3styles", 5, 4, 32768, dt, epi);
                g2.num_styles = 3;
                cs.push_back(g2);
                Case h2 = mk("flat_40samples", 40, 2, 16384, dt, epi);  // > 32 samples: two records per lane in the style fold
                cs.push_back(h2);
            }
        {
            Case bm = mk("flat_bigmean", 1, 2, 110592, MICN_F32, MICN_EPI_NONE);
            bm.mean = 50.f; bm.stdv = 0.1f;
            cs.push_back(bm);
            Case na = mk("flat_nonaffine_1style", 2, 2, 110592, MICN_BF16, MICN_EPI_NONE);
            na.affine = false; na.num_styles = 1;
            cs.push_back(na);
            Case many = mk("flat_many_slabs", 2, 96, 32768, MICN_BF16, MICN_EPI_LRELU);
            cs.push_back(many);
            Case ns = mk("flat_north_star", 1, 48, 884736, MICN_BF16, MICN_EPI_NONE);
            cs.push_back(ns);
            Case f32 = mk("flat_fp32_128^3_auto", 1, 2, 2097152, MICN_F32, MICN_EPI_LRELU);
            cs.push_back(f32);
        }
        const size_t first_cluster_case = cs.size();
        for (int dt : dts)
            for (int epi : epis) {
                cs.push_back(mk("cluster_auto_48^3", 2, 6, 110592, dt, epi));
                Case a = mk("cluster_cs2_48^3", 2, 5, 110592, dt, epi);
                a.cs = 2;
                cs.push_back(a);
                Case b = mk("cluster_cs4_slots5_refetch", 2, 3, 110592, dt, epi);
                b.cs = 4; b.slots = 5;
                cs.push_back(b);
                Case e6 = mk("cluster_cs6_96^3", 1, 2, 884736, dt, epi);
                e6.cs = 6;
                cs.push_back(e6);
                Case e = mk("cluster_cs8_96^3", 1, 3, 884736, dt, epi);
                e.cs = 8;
                cs.push_back(e);
                Case f = mk("cluster_cs16_96^3", 1, 2, 884736, dt, epi);
                f.cs = 16;
                cs.push_back(f);
                Case g = mk("cluster_cs1_persist_maxcl3", 4, 10, 40960, dt, epi);
                g.cs = 1; g.max_clusters = 3;
                cs.push_back(g);
                Case h = mk("cluster_ragged_share", 2, 3, 8 * (512 * 3 + 5) * 3 + 8, dt, epi);
                h.cs = 4; h.force_path = 1;
                cs.push_back(h);
                Case st = mk("cluster_strided_x", 2, 3, 65536, dt, epi);
                st.pad_c = 64; st.cs = 2;
                cs.push_back(st);
            }
        {
            Case p = mk("cluster_cs2_partial_96^3", 1, 2, 884736, MICN_BF16, MICN_EPI_NONE);
            p.cs = 2;
            cs.push_back(p);
            Case q = mk("cluster_fp32_128^3_cs8", 1, 1, 2097152, MICN_F32, MICN_EPI_LRELU);
            q.cs = 8;
            cs.push_back(q);
            Case bm = mk("cluster_bigmean", 1, 2, 110592, MICN_F32, MICN_EPI_NONE);
            bm.mean = 50.f; bm.stdv = 0.1f;
            cs.push_back(bm);
            Case na = mk("cluster_nonaffine_1style", 2, 2, 110592, MICN_BF16, MICN_EPI_NONE);
            na.affine = false; na.num_styles = 1;
            cs.push_back(na);
            Case many = mk("cluster_auto_many_slabs", 2, 96, 32768, MICN_BF16, MICN_EPI_LRELU);
            cs.push_back(many);
        }
        for (size_t i = first_cluster_case; i < cs.size(); ++i)
            if (cs[i].force_path < 0) cs[i].force_path = 1;  // flat is the default now: pin these to the cluster kernels
        for (const auto& c : cs) fails += run_correctness(c, verbose);
        printf("correctness: %zu cases, %d failures\n", cs.size(), fails);
    }
    if (suite == "check") {
        Case c = mk("check", oN, oC, oM > 0 ? oM : oS * oS * oS, odt, oepi);
        c.cs = ocs; c.slots = oslots; c.max_clusters = omaxcl;
        c.force_path = opath; c.fslots = ofslots; c.flag = oflag; c.fpv = ofpv; c.fgrid = ofgrid; c.fovh = ofovh; c.fcoop = ofcoop; c.fpd = ofpd; c.fpb = ofpb;
        fails += run_correctness(c, true);
    }
    if (suite == "one") {
        const long long M = oS * oS * oS;
        Case c = mk("one", oN, oC, M, odt, oepi);
        c.cs = ocs; c.slots = oslots; c.max_clusters = omaxcl;
        c.force_path = opath; c.fslots = ofslots; c.flag = oflag; c.fpv = ofpv; c.fgrid = ofgrid; c.fovh = ofovh; c.fcoop = ofcoop; c.fpd = ofpd; c.fpb = ofpb;
        PerfResult r = run_perf(c, oiters, 3);
        const double E = (double)oN * oC * M * esize(odt);
        const double fb = (oepi == MICN_EPI_ADD_LRELU ? 3 : 2) * E, bb = (oepi == MICN_EPI_ADD_LRELU ? 4 : 3) * E;
        printf("{\"N\":%lld,\"C\":%lld,\"S\":%lld,\"dtype\":\"%s\",\"epi\":%d,\"fwd_us\":%.2f,\"bwd_us\":%.2f,\"fwd_gbps\":%.1f,"
               "\"bwd_gbps\":%.1f,\"tot_gbps\":%.1f,\"frac\":%.3f,\"f_cs\":%lld,\"f_slots\":%lld,\"f_grid\":%lld,\"b_cs\":%lld,"
               "\"b_slots\":%lld,\"b_grid\":%lld}\n",
               oN, oC, oS, dname(odt), oepi, r.fwd_us, r.bwd_us, fb / r.fwd_us * 1e-3, bb / r.bwd_us * 1e-3,
               (fb + bb) / (r.fwd_us + r.bwd_us) * 1e-3, (fb + bb) / (r.fwd_us + r.bwd_us) * 1e-3 / peak, r.f_cs, r.f_slots,
               r.f_grid, r.b_cs, r.b_slots, r.b_grid);
    }
    if (suite == "trace") {
        // per-piece SM-clock timeline of a few CTAs of the flat forward kernel (bring-up aid)
        const long long M = oS * oS * oS;
        Case c = mk("trace", oN, oC, M, odt, oepi);
        c.force_path = opath; c.fslots = ofslots; c.flag = oflag; c.fpv = ofpv; c.fgrid = ofgrid; c.fovh = ofovh; c.fcoop = ofcoop; c.fpd = ofpd; c.fpb = ofpb;
        const int trace_ctas = 2 * prop.multiProcessorCount;  // the two-CTA-per-SM shape launches 2 x SMs CTAs
        const size_t tb = (size_t)trace_ctas * 64 * 16 * sizeof(long long);
        long long* dtrace = nullptr;
        CK(cudaMalloc(&dtrace, tb));
        CK(cudaMemset(dtrace, 0, tb));
        micn_set_option("flat_trace", (long long)(uintptr_t)dtrace);
        micn_set_option("flat_trace_which", 1);  // forward unless --opt flat_trace_which=2 says otherwise
        PerfResult r = run_perf(c, 1, 2);
        micn_set_option("flat_trace", 0);
        std::vector<long long> ht(tb / sizeof(long long));
        CK(cudaMemcpy(ht.data(), dtrace, tb, cudaMemcpyDeviceToHost));
        printf("trace fwd_us %.2f bwd_us %.2f P=%lld KA=%lld L=%lld\n", r.fwd_us, r.bwd_us, r.f_cs, r.f_slots, micn_get_option("last_lag"));
        const char* names[12] = {"load", "p1b", "p1e", "pubb", "pube", "gab", "gapoll", "gae", "p2wait", "p2b", "p2e", "loadwait"};
        const int nsm = trace_ctas;
        long long t0 = 0;  // earliest stamp of the launch (%globaltimer is one clock for all SMs)
        for (size_t i = 0; i < ht.size(); ++i) if (ht[i] && (!t0 || ht[i] < t0)) t0 = ht[i];
        // per round: when was the LAST record of the round published, and when did the gathers finish
        printf("round:  last_p1e  last_pube  first_gae  last_gae  last_p2e   (ns since first stamp, over all CTAs)\n");
        for (int j = 0; j < 64; ++j) {
            long long lp1 = 0, lpub = 0, fga = 0, lga = 0, lp2 = 0;
            for (int cta = 0; cta < nsm; ++cta) {
                const long long* t = ht.data() + ((size_t)cta * 64 + j) * 16;
                if (!t[2]) continue;
                lp1 = std::max(lp1, t[2] - t0); lpub = std::max(lpub, t[4] - t0); lga = std::max(lga, t[7] - t0);
                lp2 = std::max(lp2, t[10] - t0);
                if (t[7] && (!fga || t[7] - t0 < fga)) fga = t[7] - t0;
            }
            if (!lp1) break;
            printf("%5d %9lld %9lld %9lld %9lld %9lld\n", j, lp1, lpub, fga, lga, lp2);
        }
        const int ctas[3] = {0, 73, (int)micn_get_option("last_grid") - 1};
        for (int ci = 0; ci < 3; ++ci) {
            const long long* t = ht.data() + (size_t)ctas[ci] * 64 * 16;
            printf("cta %d (ns since the launch's first stamp)\n  j ", ctas[ci]);
            for (int e = 0; e < 12; ++e) printf("%9s", names[e]);
            printf("\n");
            for (int j = 0; j < 64; ++j) {
                bool any = false;
                for (int e = 0; e < 12; ++e) any = any || t[j * 16 + e];
                if (!any) break;
                printf("%3d ", j);
                for (int e = 0; e < 12; ++e) printf("%9lld", t[j * 16 + e] ? t[j * 16 + e] - t0 : -1);
                printf("\n");
            }
        }
        cudaFree(dtrace);
    }
    if (suite == "perf" || suite == "all") {
        FILE* fo = out.empty() ? nullptr : fopen(out.c_str(), "w");
        struct Shape { long long N, C, S; int dt; };
        const Shape shapes[] = {{1, 48, 96, MICN_BF16}, {1, 48, 96, MICN_F32}, {4, 96, 48, MICN_BF16}, {4, 96, 48, MICN_F32},
                                {1, 24, 128, MICN_BF16}, {1, 24, 128, MICN_F32}, {8, 192, 24, MICN_BF16}, {8, 384, 12, MICN_BF16}};
        const int css[] = {-1, 1, 2, 3, 4, 5, 6, 7, 8, 12, 16};
        for (const auto& s : shapes)
            for (int cs : css) {
                const long long M = s.S * s.S * s.S;
                const long long slab_bytes = M * (long long)esize(s.dt);
                if (cs > 0 && slab_bytes < 32 * 1024) continue;
                if (cs > 1) continue;  // the cluster sweep was round-1 bring-up; -1 = auto (flat), 1 = cluster cs1
                if (cs > 0 && slab_bytes / cs < 4096) continue;
                Case c = mk("perf", s.N, s.C, M, s.dt, MICN_EPI_NONE);
                c.cs = cs;
                PerfResult r = run_perf(c, 30, 5);
                const double E = (double)s.N * s.C * M * esize(s.dt);
                const double fg = 2 * E / r.fwd_us * 1e-3, bg = 3 * E / r.bwd_us * 1e-3, tg = 5 * E / (r.fwd_us + r.bwd_us) * 1e-3;
                char line[1024];
                snprintf(line, sizeof line,
                         "{\"N\":%lld,\"C\":%lld,\"S\":%lld,\"dtype\":\"%s\",\"cs_req\":%d,\"fwd_us\":%.2f,\"bwd_us\":%.2f,"
                         "\"fwd_gbps\":%.1f,\"bwd_gbps\":%.1f,\"tot_gbps\":%.1f,\"frac\":%.3f,\"f_path\":%lld,\"f_cs\":%lld,\"f_slots\":%lld,"
                         "\"f_grid\":%lld,\"b_path\":%lld,\"b_cs\":%lld,\"b_slots\":%lld,\"b_grid\":%lld}",
                         s.N, s.C, s.S, dname(s.dt), cs, r.fwd_us, r.bwd_us, fg, bg, tg, tg / peak, r.f_path, r.f_cs, r.f_slots,
                         r.f_grid, r.b_path, r.b_cs, r.b_slots, r.b_grid);
                printf("%s\n", line);
                if (fo) {
                    fprintf(fo, "%s\n", line);
                    fflush(fo);
                }
            }
        if (fo) fclose(fo);
    }
    return fails ? 1 : 0;
}
