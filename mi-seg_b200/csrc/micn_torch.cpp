// micn_torch.cpp - thin C++ autograd binding of the C ABI (include/micn.h) for PyTorch: the same calls the Python/ctypes
// path in functional.py makes (`micn_fwd` / `micn_bwd` / `micn_fwd_prelu` / `micn_bwd_prelu`, `micn_fwd_cl` / `micn_bwd_cl`,
// `micn_fwd_dual` / `micn_bwd_dual`), issued from a torch::autograd::Function so that a forward + backward through the
// drop-in nn.Module costs tens of microseconds of host time instead of ~165 (ctypes marshalling of 21-24 arguments,
// torch.autograd.Function's Python trampolines, save_for_backward of 2S+5 tensors: ~7 ms per C-Swin-UNETR step, which is
// host-bound under DDP).  No kernel lives here: this file only allocates outputs with the caching allocator, picks the
// current stream and forwards raw pointers to libmicn.so.  Built in-tree by csrc/Makefile as mi-seg_b200/_micn_torch.so;
// when it is absent the Python/ctypes path (same kernels) is used.
//
// Reference semantics: networks/norms/conditional_instance_norm.py:59-60 and its autograd graph; epilogues
// networks/blocks/dynunet_block.py:107-125.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <vector>

#include "../../include/micn.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

int dtype_code(const Tensor& t) {
    switch (t.scalar_type()) {
        case at::kFloat: return MICN_F32;
        case at::kBFloat16: return MICN_BF16;
        case at::kHalf: return MICN_F16;
        default: TORCH_CHECK_TYPE(false, "instance_cond: unsupported dtype ", t.scalar_type(), " (float32, bfloat16, float16)");
    }
    return -1;
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, " failed: rc=", rc, " (", micn_error_string(rc), ")");
}

// dims 2.. form one dense block (a slab is M consecutive elements)
bool dense_spatial(const Tensor& x) {
    int64_t expect = 1;
    for (int64_t d = x.dim() - 1; d >= 2; --d) {
        if (x.size(d) != 1 && x.stride(d) != expect) return false;
        expect *= x.size(d);
    }
    return true;
}

struct Ncm {
    Tensor x;
    int64_t n, c, m, sn, sc;
};
Ncm as_ncm(const Tensor& x_in) {
    Ncm r;
    r.x = dense_spatial(x_in) ? x_in : x_in.contiguous();
    r.n = r.x.size(0);
    r.c = r.x.size(1);
    r.m = 1;
    for (int64_t d = 2; d < r.x.dim(); ++d) r.m *= r.x.size(d);
    r.sn = r.n > 1 ? r.x.stride(0) : r.c * r.m;
    r.sc = r.c > 1 ? r.x.stride(1) : r.m;
    if (r.sn < 0 || r.sc < 0 || (r.n > 1 && r.sn == 0) || (r.c > 1 && r.sc == 0)) {  // expanded / flipped views
        r.x = r.x.contiguous();
        r.sn = r.c * r.m;
        r.sc = r.m;
    }
    return r;
}

// fp32 contiguous view of every parameter (only the address is used; nothing is recorded inside forward)
std::vector<Tensor> f32_params(at::TensorList ps, const at::Device& dev) {
    std::vector<Tensor> out;
    out.reserve(ps.size());
    for (const Tensor& t : ps) {
        TORCH_CHECK(t.device() == dev, "instance_cond: parameter on ", t.device(), ", input on ", dev);
        out.push_back((t.scalar_type() == at::kFloat && t.is_contiguous()) ? t : t.detach().to(at::kFloat).contiguous());
    }
    return out;
}

void* styles_ptr(const c10::optional<Tensor>& s) { return (s.has_value() && s->defined()) ? s->data_ptr() : nullptr; }

// parameter gradients [rows, C] -> one entry per parameter; styles absent from the batch (mask bit clear) get an undefined
// tensor, i.e. `.grad` stays None as in the reference
void push_param_grads(variable_list& g, const Tensor& pg, int64_t rows, int64_t S, int64_t present_mask) {
    for (int64_t i = 0; i < rows; ++i) {
        if (!pg.defined() || (present_mask >= 0 && !((present_mask >> (i % S)) & 1)))
            g.emplace_back();
        else
            g.push_back(pg.select(0, i));
    }
}

// =================================================================================================
// NC* layout: micn_fwd / micn_bwd (+ the PReLU variants)
// =================================================================================================
struct InstanceCondFn : public torch::autograd::Function<InstanceCondFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& x, const c10::optional<Tensor>& styles,
                          const c10::optional<Tensor>& residual, const c10::optional<Tensor>& slope_t, const Tensor& ws,
                          double eps, int64_t epilogue, double slope, int64_t present_mask, int64_t S, at::TensorList params) {
        TORCH_CHECK(x.is_cuda(), "instance_cond (mi-seg_b200) runs on CUDA tensors only: there is no CPU fallback");
        const int code = dtype_code(x);
        const c10::cuda::CUDAGuard guard(x.device());
        const bool affine = params.size() > 0;
        TORCH_CHECK(!affine || (int64_t)params.size() == 2 * S, "instance_cond: expected ", 2 * S, " parameters");
        const std::vector<Tensor> ps = f32_params(params, x.device());
        const Ncm q = as_ncm(x);
        for (const Tensor& t : ps)
            TORCH_CHECK_VALUE(t.numel() == q.c, "instance_cond: parameter length does not match the channel count");
        Tensor y = at::empty(q.x.sizes(), q.x.options().memory_format(at::MemoryFormat::Contiguous));
        Tensor stats = at::empty({2, q.n * q.c}, q.x.options().dtype(at::kFloat));
        Tensor res;
        if (epilogue == MICN_EPI_ADD_LRELU) {
            TORCH_CHECK_VALUE(residual.has_value() && residual->defined() && residual->sizes() == q.x.sizes(),
                              "instance_cond: add_lrelu needs a residual of the input's shape");
            // (under autocast the block input can be fp32 while conv2's output is 16-bit: rounded like `out += residual`)
            res = residual->scalar_type() == q.x.scalar_type() ? residual->contiguous()
                                                               : residual->to(q.x.scalar_type()).contiguous();
        }
        const bool prelu = slope_t.has_value() && slope_t->defined();
        TORCH_CHECK_VALUE(!prelu || epilogue == MICN_EPI_LRELU, "instance_cond: a PReLU slope (tensor) needs the 'lrelu' epilogue");
        const float* gp[MICN_MAX_STYLES];
        const float* bp[MICN_MAX_STYLES];
        for (int64_t s = 0; s < S && affine; ++s) {
            gp[s] = ps[s].data_ptr<float>();
            bp[s] = ps[S + s].data_ptr<float>();
        }
        float* mean = stats.data_ptr<float>();
        void* stream = c10::cuda::getCurrentCUDAStream(x.device().index()).stream();
        int rc;
        if (!prelu)
            rc = micn_fwd(q.x.data_ptr(), y.data_ptr(), res.defined() ? res.data_ptr() : nullptr, affine ? gp : nullptr,
                          affine ? bp : nullptr, (int)S, (const int64_t*)styles_ptr(styles), mean, mean + q.n * q.c, q.n, q.c,
                          q.m, q.sn, q.sc, code, (int)epilogue, (float)slope, (float)eps, ws.data_ptr(), (size_t)ws.numel(), stream);
        else
            rc = micn_fwd_prelu(q.x.data_ptr(), y.data_ptr(), nullptr, affine ? gp : nullptr, affine ? bp : nullptr, (int)S,
                                (const int64_t*)styles_ptr(styles), mean, mean + q.n * q.c, q.n, q.c, q.m, q.sn, q.sc, code,
                                (int)epilogue, slope_t->data_ptr<float>(), (float)eps, ws.data_ptr(), (size_t)ws.numel(), stream);
        check_rc(rc, "micn_fwd");
        variable_list saved = {q.x, (styles.has_value() && styles->defined()) ? *styles : Tensor(), stats,
                               epilogue == MICN_EPI_ADD_LRELU ? y : Tensor(), prelu ? *slope_t : Tensor(), ws};
        for (const Tensor& t : ps) saved.push_back(t);
        ctx->save_for_backward(saved);
        // needs_input_grad() counts TENSOR inputs only (undefined optionals are not edges): remember where things sit
        const int64_t has_styles = (styles.has_value() && styles->defined()) ? 1 : 0;
        const int64_t has_res = (residual.has_value() && residual->defined()) ? 1 : 0;
        const int64_t slope_edge = 1 + has_styles + has_res, param_edge0 = slope_edge + (prelu ? 1 : 0) + 1;
        ctx->saved_data["meta"] = std::vector<int64_t>{q.n, q.c, q.m, q.sn, q.sc, epilogue, S, affine ? 1 : 0, present_mask,
                                                        res.defined() ? (int64_t)residual->scalar_type() : -1, prelu ? 1 : 0,
                                                        slope_edge, param_edge0};
        ctx->saved_data["slope"] = slope;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        const auto meta = ctx->saved_data["meta"].toIntVector();
        const int64_t n = meta[0], c = meta[1], m = meta[2], sn = meta[3], sc = meta[4], epilogue = meta[5], S = meta[6];
        const bool affine = meta[7] != 0, prelu = meta[10] != 0;
        const int64_t present_mask = meta[8], res_dtype = meta[9];
        const double slope = ctx->saved_data["slope"].toDouble();
        const variable_list saved = ctx->get_saved_variables();
        const Tensor &xs = saved[0], &styles = saved[1], &stats = saved[2], &act_out = saved[3], &slope_t = saved[4], &ws = saved[5];
        const c10::cuda::CUDAGuard guard(xs.device());
        Tensor dy = grad_outputs[0].contiguous();
        if (dy.scalar_type() != xs.scalar_type()) dy = dy.to(xs.scalar_type());
        Tensor dx = at::empty(dy.sizes(), dy.options());
        Tensor dres = res_dtype >= 0 ? at::empty_like(dx) : Tensor();
        bool need_pg = false;
        for (int64_t i = 0; affine && i < 2 * S; ++i) need_pg = need_pg || ctx->needs_input_grad(meta[12] + i);
        Tensor pg = need_pg ? at::empty({2 * S, c}, xs.options().dtype(at::kFloat)) : Tensor();
        const float* gp[MICN_MAX_STYLES];
        const float* bp[MICN_MAX_STYLES];
        for (int64_t s = 0; s < S && affine; ++s) {
            gp[s] = saved[6 + s].data_ptr<float>();
            bp[s] = saved[6 + S + s].data_ptr<float>();
        }
        const float* mean = stats.data_ptr<float>();
        float* dgamma = need_pg ? pg.data_ptr<float>() : nullptr;
        float* dbeta = need_pg ? dgamma + S * c : nullptr;
        void* stream = c10::cuda::getCurrentCUDAStream(xs.device().index()).stream();
        const void* st = styles.defined() ? styles.data_ptr() : nullptr;
        Tensor dslope;
        int rc;
        if (!prelu) {
            rc = micn_bwd(dy.data_ptr(), xs.data_ptr(), act_out.defined() ? act_out.data_ptr() : nullptr, affine ? gp : nullptr,
                          affine ? bp : nullptr, (int)S, (const int64_t*)st, mean, mean + n * c, dx.data_ptr(),
                          dres.defined() ? dres.data_ptr() : nullptr, dgamma, dbeta, n, c, m, sn, sc, dtype_code(xs),
                          (int)epilogue, (float)slope, ws.data_ptr(), (size_t)ws.numel(), stream);
        } else {
            Tensor part;
            if (ctx->needs_input_grad(meta[11])) part = at::zeros({std::max<int64_t>(n * c, 1024)}, xs.options().dtype(at::kFloat));
            rc = micn_bwd_prelu(dy.data_ptr(), xs.data_ptr(), nullptr, affine ? gp : nullptr, affine ? bp : nullptr, (int)S,
                                (const int64_t*)st, mean, mean + n * c, dx.data_ptr(), nullptr, dgamma, dbeta, n, c, m, sn, sc,
                                dtype_code(xs), (int)epilogue, slope_t.data_ptr<float>(),
                                part.defined() ? part.data_ptr<float>() : nullptr, ws.data_ptr(), (size_t)ws.numel(), stream);
            if (part.defined()) dslope = part.sum().reshape(slope_t.sizes()).to(slope_t.scalar_type());
        }
        check_rc(rc, "micn_bwd");
        if (dres.defined() && (int64_t)dres.scalar_type() != res_dtype) dres = dres.to((at::ScalarType)res_dtype);
        variable_list g = {dx, Tensor(), dres, dslope, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
        if (affine) push_param_grads(g, pg, 2 * S, S, present_mask);
        return g;
    }
};

// =================================================================================================
// channels-last (token-major) layout: micn_fwd_cl / micn_bwd_cl
// =================================================================================================
struct InstanceCondClFn : public torch::autograd::Function<InstanceCondClFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& x_cl, const c10::optional<Tensor>& styles, const Tensor& ws,
                          double eps, int64_t present_mask, int64_t S, at::TensorList params) {
        const int code = dtype_code(x_cl);
        const c10::cuda::CUDAGuard guard(x_cl.device());
        const bool affine = params.size() > 0;
        const std::vector<Tensor> ps = f32_params(params, x_cl.device());
        const int64_t n = x_cl.size(0), c = x_cl.size(-1);
        const int64_t m = x_cl.numel() / std::max<int64_t>(n * c, 1);
        for (const Tensor& t : ps)
            TORCH_CHECK_VALUE(t.numel() == c, "instance_cond: parameter length does not match the channel count");
        Tensor y = at::empty_like(x_cl);
        Tensor stats = at::empty({2, n * c}, x_cl.options().dtype(at::kFloat));
        const float* gp[MICN_MAX_STYLES];
        const float* bp[MICN_MAX_STYLES];
        for (int64_t s = 0; s < S && affine; ++s) {
            gp[s] = ps[s].data_ptr<float>();
            bp[s] = ps[S + s].data_ptr<float>();
        }
        float* mean = stats.data_ptr<float>();
        void* stream = c10::cuda::getCurrentCUDAStream(x_cl.device().index()).stream();
        check_rc(micn_fwd_cl(x_cl.data_ptr(), y.data_ptr(), affine ? gp : nullptr, affine ? bp : nullptr, (int)S,
                             (const int64_t*)styles_ptr(styles), mean, mean + n * c, n, c, m, code, (float)eps, ws.data_ptr(),
                             (size_t)ws.numel(), stream),
                 "micn_fwd_cl");
        variable_list saved = {x_cl, (styles.has_value() && styles->defined()) ? *styles : Tensor(), stats, ws};
        for (const Tensor& t : ps) saved.push_back(t);
        ctx->save_for_backward(saved);
        ctx->saved_data["meta"] = std::vector<int64_t>{n, c, m, S, affine ? 1 : 0, present_mask,
                                                        1 + ((styles.has_value() && styles->defined()) ? 1 : 0) + 1};
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        const auto meta = ctx->saved_data["meta"].toIntVector();
        const int64_t n = meta[0], c = meta[1], m = meta[2], S = meta[3], present_mask = meta[5];
        const bool affine = meta[4] != 0;
        const variable_list saved = ctx->get_saved_variables();
        const Tensor &x_cl = saved[0], &styles = saved[1], &stats = saved[2], &ws = saved[3];
        const c10::cuda::CUDAGuard guard(x_cl.device());
        Tensor dy = grad_outputs[0].contiguous();
        if (dy.scalar_type() != x_cl.scalar_type()) dy = dy.to(x_cl.scalar_type());
        Tensor dx = at::empty_like(x_cl);
        bool need_pg = false;
        for (int64_t i = 0; affine && i < 2 * S; ++i) need_pg = need_pg || ctx->needs_input_grad(meta[6] + i);
        Tensor pg = need_pg ? at::empty({2 * S, c}, x_cl.options().dtype(at::kFloat)) : Tensor();
        const float* gp[MICN_MAX_STYLES];
        const float* bp[MICN_MAX_STYLES];
        for (int64_t s = 0; s < S && affine; ++s) {
            gp[s] = saved[4 + s].data_ptr<float>();
            bp[s] = saved[4 + S + s].data_ptr<float>();
        }
        const float* mean = stats.data_ptr<float>();
        float* dgamma = need_pg ? pg.data_ptr<float>() : nullptr;
        void* stream = c10::cuda::getCurrentCUDAStream(x_cl.device().index()).stream();
        check_rc(micn_bwd_cl(dy.data_ptr(), x_cl.data_ptr(), affine ? gp : nullptr, affine ? bp : nullptr, (int)S,
                             (const int64_t*)(styles.defined() ? styles.data_ptr() : nullptr), mean, mean + n * c, dx.data_ptr(),
                             dgamma, need_pg ? dgamma + S * c : nullptr, n, c, m, dtype_code(x_cl), ws.data_ptr(),
                             (size_t)ws.numel(), stream),
                 "micn_bwd_cl");
        variable_list g = {dx, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
        if (affine) push_param_grads(g, pg, 2 * S, S, present_mask);
        return g;
    }
};

// =================================================================================================
// dual-norm epilogue: micn_fwd_dual / micn_bwd_dual   (params: wa[S], ba[S], wb[S], bb[S])
// =================================================================================================
struct DualNormFn : public torch::autograd::Function<DualNormFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& a, const Tensor& b, const c10::optional<Tensor>& styles,
                          const Tensor& ws, double eps, double slope, int64_t present_mask, int64_t S, at::TensorList params) {
        const int code = dtype_code(a);
        const c10::cuda::CUDAGuard guard(a.device());
        const bool affine = params.size() > 0;
        TORCH_CHECK(!affine || (int64_t)params.size() == 4 * S, "instance_cond_dual: expected ", 4 * S, " parameters");
        const std::vector<Tensor> ps = f32_params(params, a.device());
        const int64_t n = a.size(0), c = a.size(1);
        const int64_t m = a.numel() / std::max<int64_t>(n * c, 1);
        for (const Tensor& t : ps)
            TORCH_CHECK_VALUE(t.numel() == c, "instance_cond: parameter length does not match the channel count");
        Tensor y = at::empty_like(a);
        Tensor stats = at::empty({4, n * c}, a.options().dtype(at::kFloat));  // mean_a, rstd_a, mean_b, rstd_b
        const float* pp[4][MICN_MAX_STYLES];
        for (int k = 0; k < 4 && affine; ++k)
            for (int64_t s = 0; s < S; ++s) pp[k][s] = ps[k * S + s].data_ptr<float>();
        float* sp = stats.data_ptr<float>();
        const int64_t q = n * c;
        void* stream = c10::cuda::getCurrentCUDAStream(a.device().index()).stream();
        check_rc(micn_fwd_dual(a.data_ptr(), b.data_ptr(), y.data_ptr(), affine ? pp[0] : nullptr, affine ? pp[1] : nullptr,
                               affine ? pp[2] : nullptr, affine ? pp[3] : nullptr, (int)S, (const int64_t*)styles_ptr(styles), sp,
                               sp + q, sp + 2 * q, sp + 3 * q, n, c, m, code, (float)slope, (float)eps, ws.data_ptr(),
                               (size_t)ws.numel(), stream),
                 "micn_fwd_dual");
        variable_list saved = {a, b, (styles.has_value() && styles->defined()) ? *styles : Tensor(), stats, ws};
        for (const Tensor& t : ps) saved.push_back(t);
        ctx->save_for_backward(saved);
        ctx->saved_data["meta"] = std::vector<int64_t>{n, c, m, S, affine ? 1 : 0, present_mask,
                                                        2 + ((styles.has_value() && styles->defined()) ? 1 : 0) + 1};
        ctx->saved_data["slope"] = slope;
        return y;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        const auto meta = ctx->saved_data["meta"].toIntVector();
        const int64_t n = meta[0], c = meta[1], m = meta[2], S = meta[3], present_mask = meta[5];
        const bool affine = meta[4] != 0;
        const double slope = ctx->saved_data["slope"].toDouble();
        const variable_list saved = ctx->get_saved_variables();
        const Tensor &a = saved[0], &b = saved[1], &styles = saved[2], &stats = saved[3], &ws = saved[4];
        const c10::cuda::CUDAGuard guard(a.device());
        Tensor dy = grad_outputs[0].contiguous();
        if (dy.scalar_type() != a.scalar_type()) dy = dy.to(a.scalar_type());
        Tensor da = at::empty_like(a), db = at::empty_like(b);
        bool need_pg = false;
        for (int64_t i = 0; affine && i < 4 * S; ++i) need_pg = need_pg || ctx->needs_input_grad(meta[6] + i);
        // rows: dgamma_a[S], dbeta_a[S], dgamma_b[S], dbeta_b[S]
        Tensor pg = need_pg ? at::empty({4 * S, c}, a.options().dtype(at::kFloat)) : Tensor();
        const float* pp[4][MICN_MAX_STYLES];
        for (int k = 0; k < 4 && affine; ++k)
            for (int64_t s = 0; s < S; ++s) pp[k][s] = saved[5 + k * S + s].data_ptr<float>();
        const float* sp = stats.data_ptr<float>();
        const int64_t q = n * c, gq = S * c;
        float* gpg = need_pg ? pg.data_ptr<float>() : nullptr;
        void* stream = c10::cuda::getCurrentCUDAStream(a.device().index()).stream();
        check_rc(micn_bwd_dual(dy.data_ptr(), a.data_ptr(), b.data_ptr(), affine ? pp[0] : nullptr, affine ? pp[1] : nullptr,
                               affine ? pp[2] : nullptr, affine ? pp[3] : nullptr, (int)S,
                               (const int64_t*)(styles.defined() ? styles.data_ptr() : nullptr), sp, sp + q, sp + 2 * q, sp + 3 * q,
                               da.data_ptr(), db.data_ptr(), gpg, need_pg ? gpg + gq : nullptr, need_pg ? gpg + 2 * gq : nullptr,
                               need_pg ? gpg + 3 * gq : nullptr, n, c, m, dtype_code(a), (float)slope, ws.data_ptr(),
                               (size_t)ws.numel(), stream),
                 "micn_bwd_dual");
        variable_list g = {da, db, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
        if (affine) push_param_grads(g, pg, 4 * S, S, present_mask);
        return g;
    }
};

Tensor instance_cond(const Tensor& x, const c10::optional<Tensor>& styles, const c10::optional<Tensor>& residual,
                     const c10::optional<Tensor>& slope_t, const Tensor& ws, double eps, int64_t epilogue, double slope,
                     int64_t present_mask, int64_t S, std::vector<Tensor> params) {
    return InstanceCondFn::apply(x, styles, residual, slope_t, ws, eps, epilogue, slope, present_mask, S, at::TensorList(params));
}
Tensor instance_cond_cl(const Tensor& x_cl, const c10::optional<Tensor>& styles, const Tensor& ws, double eps,
                        int64_t present_mask, int64_t S, std::vector<Tensor> params) {
    return InstanceCondClFn::apply(x_cl, styles, ws, eps, present_mask, S, at::TensorList(params));
}
Tensor instance_cond_dual(const Tensor& a, const Tensor& b, const c10::optional<Tensor>& styles, const Tensor& ws, double eps,
                          double slope, int64_t present_mask, int64_t S, std::vector<Tensor> params) {
    return DualNormFn::apply(a, b, styles, ws, eps, slope, present_mask, S, at::TensorList(params));
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, mod) {
    mod.doc() = "C++ autograd binding of libmicn.so (include/micn.h) for the drop-in instance_cond modules";
    mod.def("instance_cond", &instance_cond);
    mod.def("instance_cond_cl", &instance_cond_cl);
    mod.def("instance_cond_dual", &instance_cond_dual);
    mod.def("micn_version", []() { return micn_version(); });
}
