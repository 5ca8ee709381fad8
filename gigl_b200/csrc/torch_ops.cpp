// torch_ops.cpp - the layers of gigl_b200.nn as torch.library custom ops over the C-ABI (include/gigl_b200.h).
//
// the reference's Trainer wraps `trainer.model` in DDP after `init_model`
// (python/gigl/src/training/v1/lib/training_process.py:298-303) and calls it like any nn.Module
// (base_trainer.py:16-37); ops registered with the dispatcher are what DDP / torch.compile / export see as graph
// nodes.  This file registers, in namespace `gigl_b200`:
//
//   csr_from_coo(src, dst, n) -> (rowptr, col)                         gigl_csr_from_coo_dev
//   sage_conv(x, rowptr, col, t_rowptr?, t_col?, Wl, bl?, Wr, relu, m)  forward gigl_sage_conv_train_fwd_dev,
//                                                                      backward gigl_sage_conv_bwd_dev (autograd)
//   gcn_conv(x, rowptr, col, t_rowptr, t_col, W, b?, relu)              gigl_gcn_conv_dev / gigl_gcn_conv_bwd_dev
//   sage_conv_fwd / sage_conv_bwd / gcn_conv_fwd / gcn_conv_bwd         the raw kernels (CUDA + Meta implementations)
//
// torch is plumbing here (tensors, streams, the autograd graph); every FLOP runs in libgigl_b200.so.  No CPU
// implementation is registered: a CPU tensor fails in the dispatcher ("no kernel for backend CPU").
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/autograd/custom_function.h>
#include <torch/library.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/gigl_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// one library context per (device, stream): kernels enqueue where torch's own work is ordered
gigl_ctx* ctx_for(const Tensor& t) {
    static std::mutex mu;
    static std::map<std::pair<int, void*>, gigl_ctx*> cache;
    const int dev = t.get_device();
    void* stream = (void*)c10::cuda::getCurrentCUDAStream(dev).stream();
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, stream});
    if (it != cache.end()) return it->second;
    gigl_ctx* ctx = nullptr;
    const int rc = gigl_ctx_create_on_stream(dev, stream, &ctx);
    TORCH_CHECK(rc == 0 && ctx, "gigl_b200: cannot create a context on cuda:", dev, " (needs a compute-capability 10.x device)");
    cache[{dev, stream}] = ctx;
    return ctx;
}

void check(int rc, gigl_ctx* ctx) { TORCH_CHECK(rc == 0, "gigl_b200: ", gigl_last_error(ctx), " (code ", rc, ")"); }

Tensor f32c(const Tensor& t, const char* what) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat, "gigl_b200: ", what, " must be a float32 CUDA tensor");
    return t.contiguous();
}
const float* fptr(const c10::optional<Tensor>& t) { return (t.has_value() && t->defined()) ? t->data_ptr<float>() : nullptr; }

// ---- raw kernels ------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor> csr_from_coo_cuda(const Tensor& src, const Tensor& dst, int64_t n) {
    TORCH_CHECK(src.is_cuda() && src.scalar_type() == at::kLong && dst.scalar_type() == at::kLong && src.numel() == dst.numel(),
                "gigl_b200: src / dst must be int64 CUDA tensors of one length");
    c10::cuda::CUDAGuard guard(src.device());
    const Tensor s = src.contiguous(), d = dst.contiguous();
    Tensor rowptr = at::empty({n + 1}, s.options());
    Tensor col = at::empty({std::max<int64_t>(s.numel(), 1)}, s.options().dtype(at::kInt));
    gigl_ctx* ctx = ctx_for(src);
    check(gigl_csr_from_coo_dev(ctx, n, s.numel(), s.data_ptr<int64_t>(), d.data_ptr<int64_t>(), rowptr.data_ptr<int64_t>(),
                                col.data_ptr<int32_t>()), ctx);
    return {rowptr, col};
}

std::tuple<Tensor, Tensor> sage_conv_fwd_cuda(const Tensor& x_, const Tensor& rowptr, const Tensor& col, const Tensor& Wl_,
                                              const c10::optional<Tensor>& bl_, const Tensor& Wr_, bool relu, int64_t m) {
    c10::cuda::CUDAGuard guard(x_.device());
    const Tensor x = f32c(x_, "x"), Wl = f32c(Wl_, "lin_l.weight"), Wr = f32c(Wr_, "lin_r.weight");
    c10::optional<Tensor> bl;
    if (bl_.has_value() && bl_->defined()) bl = f32c(*bl_, "lin_l.bias");
    const int64_t n = x.size(0), F = x.size(1), O = Wl.size(0), Fp = (F + 3) & ~(int64_t)3;
    TORCH_CHECK(m >= 0 && m <= n && Wl.size(1) == F && Wr.size(0) == O && Wr.size(1) == F, "gigl_b200: sage_conv shape mismatch");
    Tensor out = at::empty({m, O}, x.options()), saved = at::empty({m, 2 * Fp}, x.options());
    gigl_ctx* ctx = ctx_for(x);
    check(gigl_sage_conv_train_fwd_dev(ctx, n, m, (int32_t)F, (int32_t)O, rowptr.data_ptr<int64_t>(), col.data_ptr<int32_t>(),
                                       x.data_ptr<float>(), Wl.data_ptr<float>(), fptr(bl), Wr.data_ptr<float>(), out.data_ptr<float>(),
                                       saved.data_ptr<float>(), relu ? 1 : 0), ctx);
    return {out, saved};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> sage_conv_bwd_cuda(const Tensor& grad_out_, const Tensor& saved, const Tensor& out,
                                                              const Tensor& rowptr, const c10::optional<Tensor>& t_rowptr,
                                                              const c10::optional<Tensor>& t_col, const Tensor& Wl, const Tensor& Wr,
                                                              int64_t n, bool relu, bool need_x, bool need_w, bool need_b) {
    c10::cuda::CUDAGuard guard(grad_out_.device());
    const Tensor go = f32c(grad_out_, "grad_out");
    const int64_t m = go.size(0), O = Wl.size(0), F = Wl.size(1);
    TORCH_CHECK(!need_x || (t_rowptr.has_value() && t_col.has_value()), "gigl_b200: grad_x needs the CSR by source (t_rowptr / t_col)");
    Tensor gx = need_x ? at::empty({n, F}, go.options()) : Tensor();
    Tensor gWl = need_w ? at::empty_like(Wl) : Tensor(), gWr = need_w ? at::empty_like(Wr) : Tensor();
    Tensor gb = need_b ? at::empty({O}, go.options()) : Tensor();
    gigl_ctx* ctx = ctx_for(go);
    check(gigl_sage_conv_bwd_dev(ctx, n, m, (int32_t)F, (int32_t)O, rowptr.data_ptr<int64_t>(),
                                 need_x ? t_rowptr->data_ptr<int64_t>() : nullptr, need_x ? t_col->data_ptr<int32_t>() : nullptr,
                                 saved.data_ptr<float>(), Wl.data_ptr<float>(), Wr.data_ptr<float>(), out.data_ptr<float>(),
                                 go.data_ptr<float>(), need_x ? gx.data_ptr<float>() : nullptr, need_w ? gWl.data_ptr<float>() : nullptr,
                                 need_b ? gb.data_ptr<float>() : nullptr, need_w ? gWr.data_ptr<float>() : nullptr, relu ? 1 : 0), ctx);
    return {gx, gWl, gb, gWr};
}

Tensor gcn_conv_fwd_cuda(const Tensor& x_, const Tensor& rowptr, const Tensor& col, const Tensor& W_, const c10::optional<Tensor>& b_,
                         bool relu) {
    c10::cuda::CUDAGuard guard(x_.device());
    const Tensor x = f32c(x_, "x"), W = f32c(W_, "lin.weight");
    c10::optional<Tensor> b;
    if (b_.has_value() && b_->defined()) b = f32c(*b_, "bias");
    const int64_t n = x.size(0), F = x.size(1), O = W.size(0);
    TORCH_CHECK(W.size(1) == F, "gigl_b200: gcn_conv shape mismatch");
    Tensor out = at::empty({n, O}, x.options());
    gigl_ctx* ctx = ctx_for(x);
    check(gigl_gcn_conv_dev(ctx, n, (int32_t)F, (int32_t)O, rowptr.data_ptr<int64_t>(), col.data_ptr<int32_t>(), x.data_ptr<float>(),
                            W.data_ptr<float>(), fptr(b), out.data_ptr<float>(), relu ? 1 : 0), ctx);
    return out;
}

std::tuple<Tensor, Tensor, Tensor> gcn_conv_bwd_cuda(const Tensor& grad_out_, const Tensor& x, const Tensor& W, const Tensor& out,
                                                     const Tensor& rowptr, const Tensor& col, const Tensor& t_rowptr, const Tensor& t_col,
                                                     bool relu, bool need_x, bool need_w, bool need_b) {
    c10::cuda::CUDAGuard guard(grad_out_.device());
    const Tensor go = f32c(grad_out_, "grad_out");
    const int64_t n = x.size(0), F = x.size(1), O = W.size(0);
    Tensor gx = need_x ? at::empty_like(x) : Tensor(), gW = need_w ? at::empty_like(W) : Tensor();
    Tensor gb = need_b ? at::empty({O}, go.options()) : Tensor();
    gigl_ctx* ctx = ctx_for(go);
    check(gigl_gcn_conv_bwd_dev(ctx, n, (int32_t)F, (int32_t)O, rowptr.data_ptr<int64_t>(), col.data_ptr<int32_t>(),
                                t_rowptr.data_ptr<int64_t>(), t_col.data_ptr<int32_t>(), x.data_ptr<float>(), W.data_ptr<float>(),
                                out.data_ptr<float>(), go.data_ptr<float>(), need_x ? gx.data_ptr<float>() : nullptr,
                                need_w ? gW.data_ptr<float>() : nullptr, need_b ? gb.data_ptr<float>() : nullptr, relu ? 1 : 0), ctx);
    return {gx, gW, gb};
}

// ---- shape-only implementations (torch.compile / export trace through the ops without a device) ---------------------
std::tuple<Tensor, Tensor> csr_from_coo_meta(const Tensor& src, const Tensor&, int64_t n) {
    return {at::empty({n + 1}, src.options()), at::empty({std::max<int64_t>(src.numel(), 1)}, src.options().dtype(at::kInt))};
}
std::tuple<Tensor, Tensor> sage_conv_fwd_meta(const Tensor& x, const Tensor&, const Tensor&, const Tensor& Wl, const c10::optional<Tensor>&,
                                              const Tensor&, bool, int64_t m) {
    return {at::empty({m, Wl.size(0)}, x.options()), at::empty({m, 2 * ((x.size(1) + 3) & ~(int64_t)3)}, x.options())};
}
Tensor gcn_conv_fwd_meta(const Tensor& x, const Tensor&, const Tensor&, const Tensor& W, const c10::optional<Tensor>&, bool) {
    return at::empty({x.size(0), W.size(0)}, x.options());
}

// ---- autograd -------------------------------------------------------------------------------------------------------
class SageConvFn : public torch::autograd::Function<SageConvFn> {
public:
    static Tensor forward(AutogradContext* actx, const Tensor& x, const Tensor& rowptr, const Tensor& col, const c10::optional<Tensor>& t_rowptr,
                          const c10::optional<Tensor>& t_col, const Tensor& Wl, const c10::optional<Tensor>& bl, const Tensor& Wr, bool relu,
                          int64_t m) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("gigl_b200::sage_conv_fwd", "")
                             .typed<std::tuple<Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const Tensor&,
                                                               const c10::optional<Tensor>&, const Tensor&, bool, int64_t)>();
        auto [out, saved] = op.call(x, rowptr, col, Wl, bl, Wr, relu, m);
        actx->save_for_backward({saved, Wl, Wr, out, rowptr, t_rowptr.value_or(Tensor()), t_col.value_or(Tensor())});
        actx->saved_data["relu"] = relu;
        actx->saved_data["n"] = x.size(0);
        const bool has_bias = bl.has_value() && bl->defined();
        actx->saved_data["has_bias"] = has_bias;
        // needs_input_grad() counts the DEFINED tensor arguments only (an absent optional has no edge): x, rowptr, col,
        // [t_rowptr], [t_col], Wl, [bl], Wr
        const int64_t i_wl = 3 + ((t_rowptr.has_value() && t_rowptr->defined()) ? 1 : 0) + ((t_col.has_value() && t_col->defined()) ? 1 : 0);
        actx->saved_data["i_wl"] = i_wl;
        return out;
    }
    static variable_list backward(AutogradContext* actx, variable_list grads) {
        const auto s = actx->get_saved_variables();
        const bool relu = actx->saved_data["relu"].toBool(), has_bias = actx->saved_data["has_bias"].toBool();
        const int64_t n = actx->saved_data["n"].toInt(), i_wl = actx->saved_data["i_wl"].toInt();
        const int64_t i_wr = i_wl + 1 + (has_bias ? 1 : 0);
        const bool need_x = actx->needs_input_grad(0) && s[5].defined() && s[6].defined();
        const bool need_w = actx->needs_input_grad(i_wl) || actx->needs_input_grad(i_wr);
        const bool need_b = has_bias && actx->needs_input_grad(i_wl + 1);
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("gigl_b200::sage_conv_bwd", "")
                             .typed<std::tuple<Tensor, Tensor, Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const Tensor&,
                                                                               const c10::optional<Tensor>&, const c10::optional<Tensor>&,
                                                                               const Tensor&, const Tensor&, int64_t, bool, bool, bool, bool)>();
        c10::optional<Tensor> tr, tc;
        if (s[5].defined()) tr = s[5];
        if (s[6].defined()) tc = s[6];
        auto [gx, gWl, gb, gWr] = op.call(grads[0], s[0], s[3], s[4], tr, tc, s[1], s[2], n, relu, need_x, need_w, need_b);
        return {gx, Tensor(), Tensor(), Tensor(), Tensor(), gWl, gb, gWr, Tensor(), Tensor()};
    }
};

class GcnConvFn : public torch::autograd::Function<GcnConvFn> {
public:
    static Tensor forward(AutogradContext* actx, const Tensor& x, const Tensor& rowptr, const Tensor& col, const Tensor& t_rowptr,
                          const Tensor& t_col, const Tensor& W, const c10::optional<Tensor>& b, bool relu) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("gigl_b200::gcn_conv_fwd", "")
                             .typed<Tensor(const Tensor&, const Tensor&, const Tensor&, const Tensor&, const c10::optional<Tensor>&, bool)>();
        Tensor out = op.call(x, rowptr, col, W, b, relu);
        actx->save_for_backward({x, W, out, rowptr, col, t_rowptr, t_col});
        actx->saved_data["relu"] = relu;
        actx->saved_data["has_bias"] = b.has_value() && b->defined();
        return out;
    }
    static variable_list backward(AutogradContext* actx, variable_list grads) {
        const auto s = actx->get_saved_variables();
        const bool relu = actx->saved_data["relu"].toBool(), has_bias = actx->saved_data["has_bias"].toBool();
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("gigl_b200::gcn_conv_bwd", "")
                             .typed<std::tuple<Tensor, Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, const Tensor&, const Tensor&,
                                                                       const Tensor&, const Tensor&, const Tensor&, bool, bool, bool, bool)>();
        auto [gx, gW, gb] = op.call(grads[0], s[0], s[1], s[2], s[3], s[4], s[5], s[6], relu, actx->needs_input_grad(0),
                                    actx->needs_input_grad(5), has_bias && actx->needs_input_grad(6));
        return {gx, Tensor(), Tensor(), Tensor(), Tensor(), gW, gb, Tensor()};
    }
};

Tensor sage_conv_autograd(const Tensor& x, const Tensor& rowptr, const Tensor& col, const c10::optional<Tensor>& t_rowptr,
                          const c10::optional<Tensor>& t_col, const Tensor& Wl, const c10::optional<Tensor>& bl, const Tensor& Wr, bool relu,
                          int64_t m) {
    return SageConvFn::apply(x, rowptr, col, t_rowptr, t_col, Wl, bl, Wr, relu, m);
}
Tensor gcn_conv_autograd(const Tensor& x, const Tensor& rowptr, const Tensor& col, const Tensor& t_rowptr, const Tensor& t_col, const Tensor& W,
                         const c10::optional<Tensor>& b, bool relu) {
    return GcnConvFn::apply(x, rowptr, col, t_rowptr, t_col, W, b, relu);
}

}  // namespace

TORCH_LIBRARY(gigl_b200, m) {
    m.def("csr_from_coo(Tensor src, Tensor dst, int n) -> (Tensor, Tensor)");
    m.def("sage_conv_fwd(Tensor x, Tensor rowptr, Tensor col, Tensor Wl, Tensor? bl, Tensor Wr, bool relu, int m) -> (Tensor, Tensor)");
    m.def("sage_conv_bwd(Tensor grad_out, Tensor saved, Tensor out, Tensor rowptr, Tensor? t_rowptr, Tensor? t_col, Tensor Wl, Tensor Wr, "
          "int n, bool relu, bool need_x, bool need_w, bool need_b) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("gcn_conv_fwd(Tensor x, Tensor rowptr, Tensor col, Tensor W, Tensor? b, bool relu) -> Tensor");
    m.def("gcn_conv_bwd(Tensor grad_out, Tensor x, Tensor W, Tensor out, Tensor rowptr, Tensor col, Tensor t_rowptr, Tensor t_col, "
          "bool relu, bool need_x, bool need_w, bool need_b) -> (Tensor, Tensor, Tensor)");
    m.def("sage_conv(Tensor x, Tensor rowptr, Tensor col, Tensor? t_rowptr, Tensor? t_col, Tensor Wl, Tensor? bl, Tensor Wr, bool relu, int m) -> Tensor");
    m.def("gcn_conv(Tensor x, Tensor rowptr, Tensor col, Tensor t_rowptr, Tensor t_col, Tensor W, Tensor? b, bool relu) -> Tensor");
}

TORCH_LIBRARY_IMPL(gigl_b200, CUDA, m) {
    m.impl("csr_from_coo", csr_from_coo_cuda);
    m.impl("sage_conv_fwd", sage_conv_fwd_cuda);
    m.impl("sage_conv_bwd", sage_conv_bwd_cuda);
    m.impl("gcn_conv_fwd", gcn_conv_fwd_cuda);
    m.impl("gcn_conv_bwd", gcn_conv_bwd_cuda);
}

TORCH_LIBRARY_IMPL(gigl_b200, Meta, m) {
    m.impl("csr_from_coo", csr_from_coo_meta);
    m.impl("sage_conv_fwd", sage_conv_fwd_meta);
    m.impl("gcn_conv_fwd", gcn_conv_fwd_meta);
}

TORCH_LIBRARY_IMPL(gigl_b200, Autograd, m) {
    m.impl("sage_conv", sage_conv_autograd);
    m.impl("gcn_conv", gcn_conv_autograd);
}
