// Compiled PyTorch binding of the one-call loss path (dg_loss_forward / dg_loss_backward).
//
// The kernels live behind the C ABI of include/depthg_b200.h; this file is only the autograd plumbing that
// depthg_b200/modules.py used to do through ctypes (struct fill, arena / output allocation, the autograd node), moved
// to C++ because at ~0.2 ms of GPU work per step the Python version of it (~245 us per step) paced the whole step.
// PyTorch here supplies device memory, the current stream and the autograd graph - nothing is computed in this file.
// Mirrors what the reference does implicitly through autograd on src/modules.py:1280-1367: differentiable inputs are
// orig_code and orig_code_pos only; the four scalar losses are the differentiable outputs.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "../../include/depthg_b200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check(int rc, const char* what) {
  if (rc == DG_OK) return;
  std::string msg = std::string(what) + ": " + dg_last_error_string() + " (code " + std::to_string(rc) + ")";
  if (rc == DG_ERR_INVALID || rc == DG_ERR_UNSUPPORTED) throw py::value_error(msg);
  throw std::runtime_error(msg);
}

void fill_strides(int64_t* dst, const Tensor& t) {
  for (int i = 0; i < 4; ++i) dst[i] = t.stride(i);
}

const float* fptr(const c10::optional<Tensor>& t) { return t.has_value() && t->defined() ? t->data_ptr<float>() : nullptr; }

dg_loss_desc_t make_desc(const std::vector<int64_t>& ints, const std::vector<double>& shifts) {
  // ints: B, C, D, H, W, Hd, Wd, S, neg_samples, flags;  shifts: intra, inter, neg, depth
  dg_loss_desc_t d;
  memset(&d, 0, sizeof d);
  d.B = (int)ints[0]; d.C = (int)ints[1]; d.D = (int)ints[2]; d.H = (int)ints[3]; d.W = (int)ints[4];
  d.Hd = (int)ints[5]; d.Wd = (int)ints[6]; d.S = (int)ints[7]; d.neg_samples = (int)ints[8]; d.flags = (int)ints[9];
  if (shifts.size() >= 4) {
    d.pos_intra_shift = (float)shifts[0]; d.pos_inter_shift = (float)shifts[1];
    d.neg_inter_shift = (float)shifts[2]; d.depth_feat_shift = (float)shifts[3];
  }
  return d;
}

class CorrLossFn : public torch::autograd::Function<CorrLossFn> {
 public:
  // ints: B, C, D, H, W, Hd, Wd, S, neg_samples, flags [, perm_seed, perm_offset];  shifts: intra, inter, neg, depth
  static variable_list forward(AutogradContext* ctx, Tensor feats, Tensor feats_pos, Tensor code, Tensor code_pos,
                               c10::optional<Tensor> depth, c10::optional<Tensor> depth_pos,
                               c10::optional<Tensor> coords, c10::optional<Tensor> perms,
                               c10::optional<Tensor> aug_feats, std::vector<int64_t> ints, std::vector<double> shifts,
                               bool materialize, bool want_fd, bool prezero_grads, int64_t perms_event,
                               c10::optional<Tensor> dsign, c10::optional<Tensor> next_depth,
                               c10::optional<Tensor> next_depth_pos, c10::optional<Tensor> next_coords,
                               c10::optional<Tensor> next_dsign, c10::optional<Tensor> next_perms,
                               std::vector<int64_t> next_rng) {
    TORCH_CHECK((ints.size() == 10 || ints.size() == 12) && shifts.size() == 4,
                "corr_loss: ints[10 or 12] / shifts[4] expected");
    dg_loss_desc_t d = make_desc(ints, shifts);
    dg_loss_plan_t plan;
    check(dg_loss_plan(&d, &plan), "dg_loss_plan");
    // Undefined upstream gradients stay undefined: the default would hand backward() a ZERO tensor for every output
    // nobody differentiated — including the non-differentiable 151 MB arena, i.e. a 42 us fill kernel per step.
    ctx->set_materialize_grads(false);
    const c10::cuda::CUDAGuard guard(feats.device());
    auto stream = c10::cuda::getCurrentCUDAStream(feats.device().index());
    const auto u8 = feats.options().dtype(torch::kUInt8);
    const auto f32 = feats.options().dtype(torch::kFloat32);
    Tensor arena = torch::empty({(int64_t)plan.total}, u8);
    Tensor out8 = torch::empty({8}, f32);
    const int64_t P = (int64_t)d.S * d.S, np = plan.npairs;
    const bool has_depth = d.flags & DG_FLAG_DEPTH_TERM;
    Tensor cd_out, loss_out, dd_out, fd_dbg;
    if (materialize) {
      cd_out = torch::empty({np, d.B, P, P}, f32);
      loss_out = torch::empty({np, d.B, P, P}, f32);
      if (has_depth) dd_out = torch::empty({(int64_t)d.B, P, P}, f32);
    }
    if (want_fd && plan.kernel >= 1) fd_dbg = torch::zeros({np, d.B, plan.Prows, plan.Prows}, f32);
    dg_loss_io_t io;
    memset(&io, 0, sizeof io);
    io.feats = feats.data_ptr<float>(); io.feats_pos = feats_pos.data_ptr<float>();
    io.code = code.data_ptr<float>(); io.code_pos = code_pos.data_ptr<float>();
    fill_strides(io.feats_strides, feats); fill_strides(io.feats_pos_strides, feats_pos);
    fill_strides(io.code_strides, code); fill_strides(io.code_pos_strides, code_pos);
    io.depth = fptr(depth); io.depth_pos = fptr(depth_pos); io.coords = fptr(coords);
    io.perms = perms.has_value() && perms->defined() ? perms->data_ptr<int64_t>() : nullptr;
    io.arena = arena.data_ptr(); io.out8 = out8.data_ptr<float>();
    io.cd_out = cd_out.defined() ? cd_out.data_ptr<float>() : nullptr;
    io.loss_out = loss_out.defined() ? loss_out.data_ptr<float>() : nullptr;
    io.dd_out = dd_out.defined() ? dd_out.data_ptr<float>() : nullptr;
    io.fd_dbg = fd_dbg.defined() ? fd_dbg.data_ptr<float>() : nullptr;
    if (aug_feats.has_value() && aug_feats->defined()) {
      io.aug_feats = aug_feats->data_ptr<float>();
      fill_strides(io.aug_feats_strides, *aug_feats);
    }
    // the gradient buffers of the coming backward are allocated here and zeroed inside the forward's code gather
    // (dg_loss_io_t::clear): no fill launches between the correlation kernel and the backward kernel
    Tensor d_code, d_code_pos;
    if (prezero_grads) {   // the caller's grad mode (it is always off inside an autograd Function's forward)
      if (code.requires_grad() && code.is_non_overlapping_and_dense()) {
        d_code = torch::empty_strided(code.sizes(), code.strides(), f32);
        io.clear[0] = d_code.data_ptr(); io.clear_bytes[0] = (size_t)d_code.numel() * sizeof(float);
      }
      if (code_pos.requires_grad() && code_pos.is_non_overlapping_and_dense()) {
        d_code_pos = torch::empty_strided(code_pos.sizes(), code_pos.strides(), f32);
        io.clear[1] = d_code_pos.data_ptr(); io.clear_bytes[1] = (size_t)d_code_pos.numel() * sizeof(float);
      }
    }
    io.perms_ready = reinterpret_cast<void*>(perms_event);
    io.dsign = fptr(dsign);                       // sampled ahead of time (dg_loss_io_t::dsign)
    if (next_depth.has_value() && next_depth->defined()) {   // the next step's sampling rides in this forward
      TORCH_CHECK(next_depth_pos.has_value() && next_coords.has_value(), "corr_loss: incomplete next-sampling job");
      io.next_depth = next_depth->data_ptr<float>();
      io.next_depth_pos = next_depth_pos->data_ptr<float>();
      io.next_coords = next_coords->data_ptr<float>();
      io.next_dsign = next_dsign.has_value() && next_dsign->defined() ? next_dsign->data_ptr<float>() : nullptr;
      if (next_perms.has_value() && next_perms->defined()) {
        TORCH_CHECK(next_rng.size() == 2, "corr_loss: next_rng = (seed, offset) expected");
        io.next_perms = next_perms->data_ptr<int64_t>();
        io.next_n_perms = (int)next_perms->size(0);
        io.next_perm_seed = static_cast<unsigned long long>(next_rng[0]);
        io.next_perm_offset = static_cast<unsigned long long>(next_rng[1]);
      }
    }
    if (ints.size() >= 12) {   // the forward draws the permutations itself into `perms`
      io.gen_perms = 1;
      io.perm_seed = static_cast<unsigned long long>(ints[10]);    // two's-complement round trip of torch's uint64 seed
      io.perm_offset = static_cast<unsigned long long>(ints[11]);
    }
    check(dg_loss_forward(&d, &io, reinterpret_cast<dg_stream_t>(stream.stream())), "dg_loss_forward");

    ctx->saved_data["ints"] = ints;
    ctx->saved_data["shifts"] = shifts;
    ctx->saved_data["arena"] = arena;        // holds coords / panels / unit gradients for backward
    if (coords.has_value() && coords->defined()) ctx->saved_data["coords"] = *coords;
    if (perms.has_value() && perms->defined()) ctx->saved_data["perms"] = *perms;
    ctx->saved_data["code_sizes"] = code.sizes().vec();
    ctx->saved_data["code_strides"] = code.strides().vec();
    ctx->saved_data["code_pos_strides"] = code_pos.strides().vec();
    ctx->saved_data["need_code"] = code.requires_grad();
    ctx->saved_data["need_code_pos"] = code_pos.requires_grad();
    if (d_code.defined()) ctx->saved_data["d_code"] = d_code;            // consumed by the first backward
    if (d_code_pos.defined()) ctx->saved_data["d_code_pos"] = d_code_pos;

    auto u = out8.unbind(0);
    // outputs: 4 differentiable scalars, out8, the arena, then only the optional tensors that exist, in the order
    // [cd_out, loss_out][dd_out][fd_dbg] (an autograd Function may not return undefined tensors)
    variable_list outs = {u[0], u[2], u[4], u[6], out8.detach(), arena};
    for (const Tensor& t : {cd_out, loss_out, dd_out, fd_dbg})
      if (t.defined()) outs.push_back(t);
    ctx->mark_non_differentiable(variable_list(outs.begin() + 4, outs.end()));
    return outs;
  }

  static variable_list backward(AutogradContext* ctx, variable_list g) {
    const dg_loss_desc_t d = make_desc(ctx->saved_data["ints"].toIntVector(), ctx->saved_data["shifts"].toDoubleVector());
    Tensor arena = ctx->saved_data["arena"].toTensor();
    const c10::cuda::CUDAGuard guard(arena.device());
    auto stream = c10::cuda::getCurrentCUDAStream(arena.device().index());
    dg_loss_grads_t gr;
    memset(&gr, 0, sizeof gr);
    std::vector<Tensor> keep;
    for (int i = 0; i < 4; ++i) {
      if (!g[i].defined()) continue;
      Tensor t = g[i].scalar_type() == torch::kFloat32 ? g[i].contiguous() : g[i].to(torch::kFloat32);
      keep.push_back(t);
      gr.g[i] = t.data_ptr<float>();
    }
    const auto sizes = ctx->saved_data["code_sizes"].toIntVector();
    const auto f32 = arena.options().dtype(torch::kFloat32);
    Tensor d_code, d_code_pos;
    if (ctx->saved_data["need_code"].toBool()) {
      const auto st = ctx->saved_data["code_strides"].toIntVector();
      if (ctx->saved_data.count("d_code")) {   // pre-zeroed by the forward; a second backward (retain_graph) refills
        d_code = ctx->saved_data["d_code"].toTensor();
        ctx->saved_data.erase("d_code");
      } else {
        d_code = torch::empty_strided(sizes, st, f32).zero_();
      }
      gr.d_code = d_code.data_ptr<float>();
      for (int i = 0; i < 4; ++i) gr.d_code_strides[i] = st[i];
    }
    if (ctx->saved_data["need_code_pos"].toBool()) {
      const auto st = ctx->saved_data["code_pos_strides"].toIntVector();
      if (ctx->saved_data.count("d_code_pos")) {
        d_code_pos = ctx->saved_data["d_code_pos"].toTensor();
        ctx->saved_data.erase("d_code_pos");
      } else {
        d_code_pos = torch::empty_strided(sizes, st, f32).zero_();
      }
      gr.d_code_pos = d_code_pos.data_ptr<float>();
      for (int i = 0; i < 4; ++i) gr.d_code_pos_strides[i] = st[i];
    }
    dg_loss_io_t io;
    memset(&io, 0, sizeof io);
    io.arena = arena.data_ptr();
    if (ctx->saved_data.count("coords")) io.coords = ctx->saved_data["coords"].toTensor().data_ptr<float>();
    if (ctx->saved_data.count("perms")) io.perms = ctx->saved_data["perms"].toTensor().data_ptr<int64_t>();
    check(dg_loss_backward(&d, &io, &gr, reinterpret_cast<dg_stream_t>(stream.stream())), "dg_loss_backward");
    variable_list out(22);
    out[2] = d_code;
    out[3] = d_code_pos;
    return out;
  }
};

std::vector<Tensor> corr_loss(Tensor feats, Tensor feats_pos, Tensor code, Tensor code_pos, c10::optional<Tensor> depth,
                              c10::optional<Tensor> depth_pos, c10::optional<Tensor> coords, c10::optional<Tensor> perms,
                              c10::optional<Tensor> aug_feats, std::vector<int64_t> ints, std::vector<double> shifts,
                              bool materialize, bool want_fd, bool prezero_grads, int64_t perms_event,
                              c10::optional<Tensor> dsign, c10::optional<Tensor> next_depth,
                              c10::optional<Tensor> next_depth_pos, c10::optional<Tensor> next_coords,
                              c10::optional<Tensor> next_dsign, c10::optional<Tensor> next_perms,
                              std::vector<int64_t> next_rng) {
  return CorrLossFn::apply(feats, feats_pos, code, code_pos, depth, depth_pos, coords, perms, aug_feats, ints, shifts,
                           materialize, want_fd, prezero_grads, perms_event, dsign, next_depth, next_depth_pos,
                           next_coords, next_dsign, next_perms, next_rng);
}

std::vector<int64_t> loss_plan(std::vector<int64_t> ints) {
  TORCH_CHECK(ints.size() >= 10, "loss_plan: ints[10] expected");
  dg_loss_desc_t d = make_desc(ints, {});
  dg_loss_plan_t p;
  check(dg_loss_plan(&d, &p), "dg_loss_plan");
  return {(int64_t)p.total, (int64_t)p.coords, (int64_t)p.dC1, (int64_t)p.dC2, p.kernel, p.Prows, p.ldf, p.ldc, p.npairs};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "depthg_b200: compiled autograd binding of dg_loss_forward / dg_loss_backward";
  m.def("corr_loss", &corr_loss, "fused ContrastiveCorrelationLoss forward (autograd-aware)");
  m.def("loss_plan", &loss_plan, "dg_loss_plan as a list: total, coords, dC1, dC2, kernel, Prows, ldf, ldc, npairs");
}
