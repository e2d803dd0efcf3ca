// One-call forward / backward of the loss: plans the arena, builds the gather tables for
// all four source tensors and enqueues the same kernels the per-stage entry points launch.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "fps_body.cuh"

namespace dg {

static float fov_factor() {  // 2*tan(90/2 rad) in fp32, bits 0x404f54cb (src/modules.py:988-989 called with fov=90)
  const uint32_t bits = 0x404F54CBu;
  float f;
  memcpy(&f, &bits, sizeof f);
  return f;
}
static const float kFarPlane = 5.0f;
static const float kNormEps = 1e-10f;

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static int check_desc(const dg_loss_desc_t* d) {
  DG_REQUIRE(d, DG_ERR_INVALID, "dg_loss: null descriptor");
  DG_REQUIRE(d->B > 0 && d->C > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->S > 0, DG_ERR_INVALID, "dg_loss: bad sizes");
  DG_REQUIRE(d->S * d->S <= d->H * d->W, DG_ERR_INVALID, "dg_loss: S*S=%d exceeds the %dx%d grid", d->S * d->S, d->H, d->W);
  DG_REQUIRE(d->neg_samples >= 0 && d->neg_samples + 2 <= DG_MAX_PAIRS, DG_ERR_INVALID, "dg_loss: neg_samples=%d", d->neg_samples);
  DG_REQUIRE(d->neg_samples == 0 || d->B >= 2, DG_ERR_INVALID, "dg_loss: negatives need a batch of at least 2");
  DG_REQUIRE(round_up(d->D, 32) <= 128, DG_ERR_UNSUPPORTED, "dg_loss: code dim %d > 128 not supported", d->D);
  if (d->flags & (DG_FLAG_FPS | DG_FLAG_DEPTH_TERM))
    DG_REQUIRE(d->Hd > 0 && d->Wd > 0, DG_ERR_INVALID, "dg_loss: depth size missing");
  return DG_OK;
}

static void make_plan(const dg_loss_desc_t* d, dg_loss_plan_t* p) {
  const int P = d->S * d->S;
  p->npairs = 2 + d->neg_samples;
  p->ldf = round_up(d->C, 32);
  p->ldc = round_up(d->D, 32);
  // tcgen05 kernel: up to 256 points a CTA walks all column tiles; up to 1024 ("dense") the columns are cut into
  // groups of two tiles (panels padded to a multiple of 256 rows so that every group is whole)
  // kernel 2 = the persistent double-buffered tcgen05 kernel (corr_pipe.cu, default); 1 = the round-1 kernel
  // (corr_umma.cu, DEPTHG_B200_CORR=umma1, kept for comparison); 0 = the generic CUDA-core kernel
  // Measured on B200 (profiles/r02_*): up to 128 points (one tile per pair and image) the persistent kernel wins
  // (69 vs 73 us at cfg2); above, a CTA of the round-1 kernel that walks a 2 x 2 tile block (S = 12: no row-mean
  // pre-pass, 0.50 vs 0.60 ms a step) or a group of two column tiles (dense: first-operand chunks loaded once per two
  // tiles, 3.13 vs 3.91 ms) moves fewer operand bytes per tile, which is what bounds both.
  p->kernel = (P <= 1024 && p->ldc <= 128 && !(d->flags & DG_FLAG_FORCE_SIMT)) ? (P <= 128 ? 2 : 1) : 0;
  if (p->kernel) {
    const char* e = getenv("DEPTHG_B200_CORR");
    if (e && strcmp(e, "umma1") == 0) p->kernel = 1;
    if (e && strcmp(e, "pipe") == 0) p->kernel = 2;
  }
  p->Prows = p->kernel == 2 ? round_up(P, 128)
                            : (p->kernel == 1 ? (P <= 256 ? round_up(P, 128) : round_up(P, 256)) : round_up(P, 64));
  const size_t ni = p->kernel ? p->Prows / 128 : 1;  // dC2 partial buffers (one per 128-row tile of the first operand)
  // dC1 partial buffers: one per 128-column tile (kernel 2) / per 256-column group above 256 points (kernel 1)
  const size_t nj = p->kernel == 2 ? p->Prows / 128 : ((p->kernel == 1 && P > 256) ? p->Prows / 256 : 1);
  const size_t np = p->npairs, B = d->B, Pr = p->Prows;
  const size_t nf = np + ((d->flags & DG_FLAG_AUG_INTRA) ? 1 : 0);   // feature panel slots (+1: depth-augmented features)
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  p->coords = take((size_t)2 * B * P * 2 * 4);
  p->frn = take(nf * B * Pr * 4);
  p->fmean = take(nf * B * 16 * p->ldf * 4);  // up to 16 partial means per (slot, image)
  p->crn = take(np * B * Pr * 4);
  p->dsign = take(B * Pr * 4);
  p->ws_bytes = corr_workspace_bytes(p->npairs, d->B, P);
  p->ws = take(p->ws_bytes);
  p->dC1 = take((np + 1) * nj * B * Pr * p->ldc * 4);
  p->dC2 = take((np + 1) * ni * B * Pr * p->ldc * 4);
  if (p->kernel == 2) {   // 16-bit panels interleaved per 32-channel chunk [32 hi | 32 lo]: one buffer of pitch 2 * ld
    p->c_hi = take(np * B * Pr * p->ldc * 4);
    p->c_lo = take(np * B * Pr * p->ldc * 4);
    p->cb_hi = take(np * B * Pr * p->ldc * 4);
    p->f_hi = take(nf * B * Pr * p->ldf * 4);
    p->cb_lo = p->f_lo = 0;
  } else if (p->kernel) {
    p->c_hi = take(np * B * Pr * p->ldc * 4);
    p->c_lo = take(np * B * Pr * p->ldc * 4);
    p->cb_hi = take(np * B * Pr * p->ldc * 2);
    p->cb_lo = take(np * B * Pr * p->ldc * 2);
    p->f_hi = take(nf * B * Pr * p->ldf * 2);
    p->f_lo = take(nf * B * Pr * p->ldf * 2);
  } else {
    p->c_hi = take(np * B * Pr * p->ldc * 4);
    p->f_hi = take(nf * B * Pr * p->ldf * 4);
    p->c_lo = p->cb_hi = p->cb_lo = p->f_lo = 0;
  }
  p->stage = 0;
  if (p->kernel && (d->flags & DG_FLAG_STAGE_NHWC))
    p->stage = take((size_t)((d->flags & DG_FLAG_AUG_INTRA) ? 3 : 2) * B * d->C * d->H * d->W * 4);
  p->total = off;
}

static void set_desc(SetDesc& s, const float* src, const int64_t* st, int coord, int slot, int perm_row) {
  s.src = src;
  s.sb = st[0]; s.sc = st[1]; s.sh = st[2]; s.sw = st[3];
  s.coord = coord; s.slot = slot; s.perm_row = perm_row;
}

// slot 0 = own tensor at coords1, slot 1 = positive tensor at coords2, slots 2.. = own tensor[perm_k] at coords2
static int build_sets(SetTable& tab, const float* own, const int64_t* own_st, const float* pos, const int64_t* pos_st,
                      int nneg) {
  set_desc(tab.s[0], own, own_st, 0, 0, -1);
  set_desc(tab.s[1], pos, pos_st, 1, 1, -1);
  for (int k = 0; k < nneg; ++k) set_desc(tab.s[2 + k], own, own_st, 1, 2 + k, k);
  return 2 + nneg;
}

}  // namespace dg

extern "C" int dg_loss_plan(const dg_loss_desc_t* desc, dg_loss_plan_t* plan) {
  using namespace dg;
  int rc = check_desc(desc);
  if (rc != DG_OK) return rc;
  DG_REQUIRE(plan, DG_ERR_INVALID, "dg_loss_plan: null plan");
  make_plan(desc, plan);
  return DG_OK;
}

extern "C" int dg_loss_forward(const dg_loss_desc_t* d, const dg_loss_io_t* io, dg_stream_t stream) {
  using namespace dg;
  int rc = check_desc(d);
  if (rc != DG_OK) return rc;
  DG_REQUIRE(io && io->feats && io->feats_pos && io->code && io->code_pos && io->arena && io->out8, DG_ERR_INVALID,
             "dg_loss_forward: null pointer");
  DG_REQUIRE(d->neg_samples == 0 || io->perms, DG_ERR_INVALID, "dg_loss_forward: perms missing");
  const bool fps = d->flags & DG_FLAG_FPS, depth_term = d->flags & DG_FLAG_DEPTH_TERM;
  DG_REQUIRE(!fps || (io->depth && io->depth_pos), DG_ERR_INVALID, "dg_loss_forward: fps sampling needs depth and depth_pos");
  DG_REQUIRE(!depth_term || io->depth || (!fps && io->dsign), DG_ERR_INVALID, "dg_loss_forward: the depth term needs depth");
  DG_REQUIRE(fps || io->coords, DG_ERR_INVALID, "dg_loss_forward: coords missing");
  const bool aug = d->flags & DG_FLAG_AUG_INTRA;
  DG_REQUIRE(!aug || io->aug_feats, DG_ERR_INVALID, "dg_loss_forward: DG_FLAG_AUG_INTRA needs aug_feats");
  DG_REQUIRE(!aug || !depth_term, DG_ERR_INVALID, "dg_loss_forward: the depth-augmented variant has no depth term");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dg_loss_plan_t pl;
  make_plan(d, &pl);
  uint8_t* A = static_cast<uint8_t*>(io->arena);
  const int B = d->B, S = d->S, P = S * S, np = pl.npairs;
  const float* coords = io->coords;
  const bool pointwise = d->flags & DG_FLAG_POINTWISE;
  float* fmean = pointwise ? reinterpret_cast<float*>(A + pl.fmean) : nullptr;
  float* dsign = depth_term ? reinterpret_cast<float*>(A + pl.dsign) : nullptr;
  PermJob pj;
  memset(&pj, 0, sizeof pj);
  bool draw = io->gen_perms && d->neg_samples > 0;
  if (draw) {
    pj.seed = io->perm_seed; pj.offset = io->perm_offset; pj.n = d->neg_samples; pj.B = B;
    pj.out = const_cast<int64_t*>(io->perms);
  }
  if (fps) {  // FPS of both depth tensors; the same CTAs also emit the depth signs the depth term needs
    float* c = reinterpret_cast<float*>(A + pl.coords);
    const bool fuse = draw && pj.n <= 256 && (size_t)pj.n * B * sizeof(int) <= 64 * 1024;
    rc = launch_fps(io->depth, io->depth_pos, B, d->Hd, d->Wd, d->H, d->W, S, fov_factor(), kFarPlane, 1, c, nullptr, st,
                    dsign, pl.Prows, kNormEps, fuse ? &pj : nullptr);
    if (rc != DG_OK) return rc;
    coords = c;
    if (fuse) draw = false;
  }
  if (draw) {   // no FPS launch to ride on (or too many permutations for one CTA): the sampler's own launch
    rc = dg_super_perms(pj.seed, pj.offset, pj.n, pj.B, pj.out, stream);
    if (rc != DG_OK) return rc;
  }
  if (!fps && depth_term) {
    if (io->dsign) {   // sampled ahead of time (dg_loss_presample / the previous forward's next_* job)
      dsign = const_cast<float*>(io->dsign);
    } else {
      rc = launch_depth_sign(io->depth, B, d->Hd, d->Wd, S, kNormEps, pl.Prows, dsign, st);
      if (rc != DG_OK) return rc;
    }
  }
  if (io->perms_ready) DG_CUDA_OK(cudaStreamWaitEvent(st, reinterpret_cast<cudaEvent_t>(io->perms_ready), 0));
  SetTable ftab, ctab;
  GatherOut fo, co;
  // backbone features (NCHW sources are first staged as channels-last copies so the gather reads whole lines)
  const float* feats = io->feats;
  const float* feats_pos = io->feats_pos;
  int64_t fst[4], fpst[4];
  for (int i = 0; i < 4; ++i) { fst[i] = io->feats_strides[i]; fpst[i] = io->feats_pos_strides[i]; }
  const int64_t HW = (int64_t)d->H * d->W;
  if (pl.stage && fst[3] == 1 && fst[2] == d->W && fst[1] == HW && fpst[3] == 1 && fpst[2] == d->W && fpst[1] == HW) {
    float* sa = reinterpret_cast<float*>(A + pl.stage);
    float* sb2 = sa + (size_t)B * d->C * HW;
    rc = launch_nchw_to_nhwc(feats, feats_pos, B, d->C, (int)HW, fst[0], fpst[0], sa, sb2, st);
    if (rc != DG_OK) return rc;
    feats = sa; feats_pos = sb2;
    fst[0] = fpst[0] = (int64_t)d->C * HW; fst[1] = fpst[1] = 1; fst[2] = fpst[2] = (int64_t)d->W * d->C; fst[3] = fpst[3] = d->C;
  }
  int nsets = build_sets(ftab, feats, fst, feats_pos, fpst, d->neg_samples);
  const int ncsets = nsets;
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];
  for (int k = 0; k < np; ++k) { fs1[k] = 0; fs2[k] = k; }
  if (aug) {   // one more feature panel slot: the depth-augmented features at coords1; the intra pair correlates it with itself
    const float* af = io->aug_feats;
    int64_t ast[4];
    for (int i = 0; i < 4; ++i) ast[i] = io->aug_feats_strides[i];
    if (pl.stage && feats != io->feats && ast[3] == 1 && ast[2] == d->W && ast[1] == HW) {
      float* sc3 = reinterpret_cast<float*>(A + pl.stage) + (size_t)2 * B * d->C * HW;
      rc = launch_nchw_to_nhwc(af, nullptr, B, d->C, (int)HW, ast[0], 0, sc3, nullptr, st);
      if (rc != DG_OK) return rc;
      af = sc3;
      ast[0] = (int64_t)d->C * HW; ast[1] = 1; ast[2] = (int64_t)d->W * d->C; ast[3] = d->C;
    }
    set_desc(ftab.s[nsets], af, ast, 0, np, -1);
    ++nsets;
    fs1[0] = fs2[0] = np;
  }
  fo.out = reinterpret_cast<float*>(A + pl.f_hi);
  fo.out_lo = nullptr;
  fo.hi16 = reinterpret_cast<__half*>(A + pl.f_hi);
  fo.lo16 = reinterpret_cast<__half*>(A + pl.f_lo);
  fo.interleave = pl.kernel == 2;
  fo.rnorm = reinterpret_cast<float*>(A + pl.frn);
  fo.meanvec = fmean;
  const int ffmt = pl.kernel ? FMT_FEATS_SPLIT : FMT_F32;
  const int nsplit = gather_nsplit(ffmt, ftab, nsets, d->C, pl.ldf);
  // code
  build_sets(ctab, io->code, io->code_strides, io->code_pos, io->code_pos_strides, d->neg_samples);  // ncsets sets
  co.out = reinterpret_cast<float*>(A + pl.c_hi);
  co.out_lo = pl.kernel ? reinterpret_cast<float*>(A + pl.c_lo) : nullptr;
  co.hi16 = pl.kernel ? reinterpret_cast<__half*>(A + pl.cb_hi) : nullptr;
  co.lo16 = pl.kernel ? reinterpret_cast<__half*>(A + pl.cb_lo) : nullptr;
  co.interleave = pl.kernel == 2;
  co.rnorm = reinterpret_cast<float*>(A + pl.crn);
  co.meanvec = nullptr;
  // (a single merged launch — launch_gather_all — was measured slower: the code CTAs inherit the feature
  //  kernel's register footprint and lose the occupancy that hides their latency; kept for experiments)
  bool dots_done = false;
  rc = DG_ERR_UNSUPPORTED;
  if (pl.kernel && !aug && getenv("DEPTHG_B200_GATHER_ALL")) {
    rc = launch_gather_all(ftab, ctab, nsets, B, d->C, d->D, d->H, d->W, coords, S, io->perms, kNormEps, pl.Prows, pl.ldf,
                           pl.ldc, nsplit, fo, co, st);
    if (rc == DG_OK)   // this experimental launch has no job slot: the caller's buffers are cleared by plain memsets
      for (int q = 0; q < 2; ++q)
        if (io->clear[q] && io->clear_bytes[q]) DG_CUDA_OK(cudaMemsetAsync(io->clear[q], 0, io->clear_bytes[q], st));
  }
  if (rc == DG_ERR_UNSUPPORTED) {
    rc = launch_gather(ffmt, ftab, nsets, B, d->C, d->H, d->W, coords, S, io->perms, kNormEps, pl.Prows, pl.ldf, nsplit, fo,
                       st);
    if (rc != DG_OK) return rc;
    DotsJob job;
    memset(&job, 0, sizeof job);
    for (int q = 0; q < 2; ++q) {   // the caller's buffers to clear: inside the code gather when it has a job slot
      if (!io->clear[q] || !io->clear_bytes[q]) continue;
      if (pl.kernel && reinterpret_cast<uintptr_t>(io->clear[q]) % 16 == 0 && io->clear_bytes[q] % 16 == 0) {
        job.clr[q] = static_cast<float4*>(io->clear[q]);
        job.clr_n16[q] = io->clear_bytes[q] / 16;
      } else {
        DG_CUDA_OK(cudaMemsetAsync(io->clear[q], 0, io->clear_bytes[q], st));
      }
    }
    if (pl.kernel) {   // the pair dots / flag reset of the tcgen05 kernel ride along as extra CTAs of the code gather
      job.fmean = fmean;
      umma_ws_layout(A + pl.ws, &job.err, &job.dots);
      job.nsplit = nsplit; job.npairs = np; job.B = B; job.ldf = pl.ldf;
      for (int k = 0; k < DG_MAX_PAIRS; ++k) { job.fs1[k] = k < np ? fs1[k] : 0; job.fs2[k] = k < np ? fs2[k] : 0; }
      dots_done = true;
    }
    rc = launch_gather(pl.kernel ? FMT_CODE_SPLIT : FMT_F32, ctab, ncsets, B, d->D, d->H, d->W, coords, S, io->perms,
                       kNormEps, pl.Prows, pl.ldc, 1, co, st, pl.kernel ? &job : nullptr);
  }
  if (rc != DG_OK) return rc;

  float shifts[DG_MAX_PAIRS];
  int32_t groups[DG_MAX_PAIRS];
  shifts[0] = d->pos_intra_shift; groups[0] = DG_GROUP_INTRA;
  shifts[1] = d->pos_inter_shift; groups[1] = DG_GROUP_INTER;
  for (int k = 2; k < np; ++k) { shifts[k] = d->neg_inter_shift; groups[k] = DG_GROUP_NEG; }
  const int kflags = d->flags & (DG_FLAG_POINTWISE | DG_FLAG_ZERO_CLAMP | DG_FLAG_STABALIZE);
  float* dC1 = reinterpret_cast<float*>(A + pl.dC1);
  float* dC2 = reinterpret_cast<float*>(A + pl.dC2);
  // the next step's sampling (io->next_*): extra CTAs of the persistent correlation kernel when it is the one that
  // runs, else a launch of its own behind this forward
  FpsArgs next;
  size_t next_smem = 0;
  const bool has_next = io->next_depth != nullptr;
  if (has_next) {
    DG_REQUIRE(io->next_depth_pos && io->next_coords && d->Hd > 0 && d->Wd > 0, DG_ERR_INVALID,
               "dg_loss_forward: next_depth needs next_depth_pos, next_coords and the depth size");
    DG_REQUIRE(io->next_n_perms == 0 || io->next_perms, DG_ERR_INVALID, "dg_loss_forward: next_perms missing");
    PermJob npj;
    memset(&npj, 0, sizeof npj);
    npj.seed = io->next_perm_seed; npj.offset = io->next_perm_offset; npj.n = io->next_n_perms; npj.B = B;
    npj.out = io->next_perms;
    rc = make_fps_args(&next, &next_smem, io->next_depth, io->next_depth_pos, B, d->Hd, d->Wd, d->H, d->W, S,
                       fov_factor(), kFarPlane, 1, io->next_coords, nullptr, io->next_dsign, pl.Prows, kNormEps,
                       npj.n > 0 ? &npj : nullptr);
    if (rc != DG_OK) return rc;
  }
  const bool ride = has_next && pl.kernel == 2 && corr_pipe_can_ride(next, next_smem) && !getenv("DEPTHG_B200_NO_RIDE");
  if (pl.kernel) {
    dg_panels_t pan;
    pan.format = DG_PANEL_CODE_SPLIT;
    pan.f_hi = A + pl.f_hi; pan.f_lo = A + pl.f_lo; pan.c_hi = A + pl.c_hi; pan.c_lo = A + pl.c_lo;
    pan.cb_hi = A + pl.cb_hi; pan.cb_lo = A + pl.cb_lo;
    if (pl.kernel == 2)
      rc = corr_loss_pipe(&pan, fmean, nsplit, dsign, np, B, P, pl.Prows, pl.ldf, pl.ldc, shifts, groups,
                          d->depth_feat_shift, kflags, io->out8, dC1, dC2, io->cd_out, io->loss_out, io->dd_out,
                          io->fd_dbg, A + pl.ws, st, fs1, fs2, aug ? np + 1 : np, dots_done, ride ? &next : nullptr,
                          next_smem);
    else
      rc = corr_loss_umma(&pan, fmean, nsplit, dsign, np, B, P, pl.Prows, pl.ldf, pl.ldc, shifts, groups,
                          d->depth_feat_shift, kflags, io->out8, dC1, dC2, io->cd_out, io->loss_out, io->dd_out,
                          io->fd_dbg, A + pl.ws, st, fs1, fs2, aug ? np + 1 : np, dots_done);
  } else {
    rc = corr_loss_simt(reinterpret_cast<const float*>(A + pl.f_hi), reinterpret_cast<const float*>(A + pl.c_hi), fmean,
                        nsplit, dsign, np, B, P, pl.Prows, pl.ldf, pl.ldc, shifts, groups, d->depth_feat_shift, kflags,
                        io->out8, dC1, dC2, io->cd_out, io->loss_out, io->dd_out, A + pl.ws, st, fs1, fs2);
  }
  if (rc != DG_OK) return rc;
  if (has_next && !ride)
    return launch_fps(io->next_depth, io->next_depth_pos, B, d->Hd, d->Wd, d->H, d->W, S, fov_factor(), kFarPlane, 1,
                      io->next_coords, nullptr, st, io->next_dsign, pl.Prows, kNormEps, next.pj.n > 0 ? &next.pj : nullptr);
  return DG_OK;
}

extern "C" int dg_loss_presample(const dg_loss_desc_t* d, const float* depth, const float* depth_pos, float* coords,
                                 float* dsign, int64_t* perms, int n_perms, unsigned long long perm_seed,
                                 unsigned long long perm_offset, dg_stream_t stream) {
  using namespace dg;
  int rc = check_desc(d);
  if (rc != DG_OK) return rc;
  DG_REQUIRE(depth && depth_pos && coords && d->Hd > 0 && d->Wd > 0, DG_ERR_INVALID, "dg_loss_presample: null pointer / depth size");
  DG_REQUIRE(n_perms == 0 || perms, DG_ERR_INVALID, "dg_loss_presample: perms missing");
  dg_loss_plan_t pl;
  make_plan(d, &pl);
  PermJob pj;
  memset(&pj, 0, sizeof pj);
  pj.seed = perm_seed; pj.offset = perm_offset; pj.n = n_perms; pj.B = d->B; pj.out = perms;
  const bool fuse = n_perms > 0 && n_perms <= 256 && (size_t)n_perms * d->B * sizeof(int) <= 64 * 1024;
  rc = launch_fps(depth, depth_pos, d->B, d->Hd, d->Wd, d->H, d->W, d->S, fov_factor(), kFarPlane, 1, coords, nullptr,
                  reinterpret_cast<cudaStream_t>(stream), dsign, pl.Prows, kNormEps, fuse ? &pj : nullptr);
  if (rc != DG_OK) return rc;
  if (n_perms > 0 && !fuse) return dg_super_perms(perm_seed, perm_offset, n_perms, d->B, perms, stream);
  return DG_OK;
}

extern "C" int dg_loss_backward(const dg_loss_desc_t* d, const dg_loss_io_t* io, const dg_loss_grads_t* gr,
                                dg_stream_t stream) {
  using namespace dg;
  int rc = check_desc(d);
  if (rc != DG_OK) return rc;
  DG_REQUIRE(io && io->arena && gr, DG_ERR_INVALID, "dg_loss_backward: null pointer");
  DG_REQUIRE(d->neg_samples == 0 || io->perms, DG_ERR_INVALID, "dg_loss_backward: perms missing");
  dg_loss_plan_t pl;
  make_plan(d, &pl);
  uint8_t* A = static_cast<uint8_t*>(io->arena);
  const float* coords = (d->flags & DG_FLAG_FPS) ? reinterpret_cast<const float*>(A + pl.coords) : io->coords;
  DG_REQUIRE(coords, DG_ERR_INVALID, "dg_loss_backward: coords missing");
  // destination sets: slot 0 + negatives scatter into d_code, slot 1 into d_code_pos
  SetTable tab;
  int n = 0;
  if (gr->d_code) {
    set_desc(tab.s[n++], gr->d_code, gr->d_code_strides, 0, 0, -1);
    for (int k = 0; k < d->neg_samples; ++k) set_desc(tab.s[n++], gr->d_code, gr->d_code_strides, 1, 2 + k, k);
  }
  if (gr->d_code_pos) set_desc(tab.s[n++], gr->d_code_pos, gr->d_code_pos_strides, 1, 1, -1);
  if (n == 0) return DG_OK;
  PairTable pt;
  pt.group[0] = DG_GROUP_INTRA; pt.scale[0] = 1.f;
  pt.group[1] = DG_GROUP_INTER; pt.scale[1] = 1.f;
  for (int k = 2; k < pl.npairs; ++k) { pt.group[k] = DG_GROUP_NEG; pt.scale[k] = 1.f / (float)d->neg_samples; }
  GroupW gw;
  gw.arr = nullptr;
  for (int g = 0; g < DG_NUM_GROUPS; ++g) gw.ptr[g] = gr->g[g];
  return launch_gather_bwd(tab, n, d->B, d->D, d->H, d->W, coords, d->S, io->perms, kNormEps, pl.Prows, pl.ldc,
                           reinterpret_cast<const float*>(A + pl.c_hi),
                           pl.kernel ? reinterpret_cast<const float*>(A + pl.c_lo) : nullptr,
                           reinterpret_cast<const float*>(A + pl.crn), reinterpret_cast<const float*>(A + pl.dC1),
                           reinterpret_cast<const float*>(A + pl.dC2), pl.npairs, pt,
                           (d->flags & DG_FLAG_DEPTH_TERM) ? 1 : 0, gw, reinterpret_cast<cudaStream_t>(stream),
                           pl.kernel ? pl.Prows / 128 : 1,
                           pl.kernel == 2 ? pl.Prows / 128 : ((pl.kernel == 1 && d->S * d->S > 256) ? pl.Prows / 256 : 1),
                           pl.kernel == 2 ? 128 : 256);
}
