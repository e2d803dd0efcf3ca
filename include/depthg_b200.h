/* depthg_b200 — C ABI of the B200-native DepthG hot path.
 *
 * This is the drop-in boundary.  Every entry point
 *   - is extern "C", takes plain pointers and sizes, no torch / C++ types;
 *   - takes DEVICE pointers unless the parameter is documented "host";
 *   - enqueues its kernels on `stream` and returns without synchronising;
 *   - never allocates: the caller owns every buffer including workspaces
 *     (sizes come from the dg_*_workspace_bytes helpers);
 *   - returns DG_OK (0) or a negative DG_ERR_* code; the text of the last
 *     failure on the calling thread is available from dg_last_error_string().
 * Kernels are compiled for sm_100a only; there is no CPU path.
 *
 * The reference (leonsick/depthg) is pure Python: what each entry point
 * replaces is a function of /root/reference/src/modules.py or
 * src/precompute_knns.py, cited per declaration.
 */
#ifndef DEPTHG_B200_H_
#define DEPTHG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dg_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define DG_API __attribute__((visibility("default")))
#else
#define DG_API
#endif

enum {
  DG_OK = 0,
  DG_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, bad pitch) */
  DG_ERR_UNSUPPORTED = -2, /* shape outside what the kernels are built for            */
  DG_ERR_WORKSPACE = -3,   /* workspace too small                                     */
  DG_ERR_CUDA = -4         /* a CUDA runtime call / launch failed                     */
};

#define DG_MAX_PAIRS 32 /* helper() calls per forward: 1 intra + 1 inter + neg_samples */
#define DG_MAX_SETS 32  /* coordinate sets gathered from one source tensor per launch  */

/* Pair groups: which scalar of the reference's output tuple a pair feeds. */
enum { DG_GROUP_INTRA = 0, DG_GROUP_INTER = 1, DG_GROUP_NEG = 2, DG_GROUP_DEPTH = 3, DG_NUM_GROUPS = 4 };

/* Panel formats written by dg_gather_norm and consumed by dg_corr_loss.
 *   F32         : fp32 rows (generic CUDA-core correlation kernel).
 *   FEATS_SPLIT : two bf16 panels hi/lo with x ~= hi + lo (tcgen05 kind::f16, 3-term product).
 *   CODE_SPLIT  : fp32 hi (tf32-rounded) and lo = x - hi (tcgen05 kind::tf32, 3-term product)
 *                 plus bf16 hi/lo panels of the same rows for the gradient GEMMs. */
enum { DG_PANEL_F32 = 0, DG_PANEL_FEATS_SPLIT = 1, DG_PANEL_CODE_SPLIT = 2 };

/* Flags of ContrastiveCorrelationLoss.helper (src/modules.py:1231-1254). */
enum { DG_FLAG_POINTWISE = 1, DG_FLAG_ZERO_CLAMP = 2, DG_FLAG_STABALIZE = 4 };

DG_API int dg_version(void);
DG_API const char* dg_last_error_string(void);
/* Number of kernels this library has launched in this process (for launch accounting). */
DG_API unsigned long long dg_kernel_launches(void);
/* Diagnostics: when enabled, every kernel this library launches is bracketed by CUDA events on
 * its own stream; dg_profile_collect synchronises and writes "name\tlaunches\ttotal_us\n" lines
 * (returns the buffer size needed).  Enabling/disabling clears the records. */
DG_API int dg_profile_enable(int on);
DG_API size_t dg_profile_collect(char* buf, size_t buf_bytes);
/* Debug: device buffer [grid][16] int64 receiving %globaltimer stamps of the tcgen05 kernel phases (NULL = off). */
DG_API int dg_debug_set_clock_buffer(long long* dev_ptr);

/* Pitch (in floats) of a panel row holding `channels` values: rounded up to a
 * multiple of 32 so a row is a whole number of 128-byte lines. */
DG_API int dg_panel_ld(int channels);
/* Rows of a panel holding P sample points: rounded up to a multiple of 64. */
DG_API int dg_panel_rows(int P);

/* ---------------------------------------------------------------------------
 * Depth-guided farthest point sampling.
 * Replaces farthest_point_sampling_depth + depth2points + fps
 * (src/modules.py:999-1037, :988-996, :939-985) for up to two depth tensors in
 * one launch (one CTA per image, nothing leaves the device).
 *
 *   depth_a, depth_b : [B,1,Hd,Wd] fp32 contiguous; depth_b may be NULL.
 *   H, W             : feature-grid size the depth is average-pooled to
 *                      (adaptive_avg_pool2d window rule); H*W <= 4096.
 *   S                : feature_samples; S*S <= H*W points are selected.
 *   factor, far_plane: 2*tan(fov/2) as an fp32 value computed by the caller
 *                      (the reference feeds 90 to tan in radians) and 5.0.
 *   affine           : 0 -> coords in [0,1) as the reference function returns
 *                      them; 1 -> coords*2-1 as the loss consumes them.
 *   coords           : out [nimg,S,S,2] (row/H, col/W), raster order.
 *   idx              : out [nimg,S*S] int32 flat indices row*W+col, ascending
 *                      (may be NULL).   nimg = B or 2B (depth_a then depth_b).
 * Arithmetic is fp32 with one rounding per operation and no FMA contraction, so
 * index sets are bit-identical to the reference's NumPy path. */
DG_API int dg_fps_coords(const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H, int W, int S,
                  float factor, float far_plane, int affine, float* coords, int32_t* idx, dg_stream_t stream);

/* All neg_samples permutations of super_perm (src/modules.py:1184-1188, called at :1341) in one
 * launch: Fisher-Yates on Philox4x32-10 keyed by (seed, offset), fixed points bumped by one, mod B.
 *   out : [n,B] int64.  Same distribution as torch.randperm-based super_perm, different stream. */
DG_API int dg_super_perms(unsigned long long seed, unsigned long long offset, int n, int B, int64_t* out,
                          dg_stream_t stream);

/* norm(interpolate(depth,(S,S),bilinear,align_corners=True)) of
 * depth_feature_correlation (src/modules.py:1261-1265): s = d / max(|d|, eps).
 *   depth [B,1,Hd,Wd] -> out [B, out_pitch] (first S*S entries per image, rest zeroed). */
DG_API int dg_depth_sign(const float* depth, int B, int Hd, int Wd, int S, float eps, int out_pitch, float* out,
                  dg_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused bilinear gather + L2 normalise.  Replaces sample() + norm()
 * (src/modules.py:822-825, :789-790) and the orig_feats[perm] copies (:1342).
 *
 *   t        : source tensor, logical shape [B,C,H,W], element strides
 *              strides[4] = (b,c,h,w) (host array) — NCHW and channels-last
 *              both work without a copy.
 *   coords   : [ncoord,B,S*S,2] in [-1,1]; (.,0) addresses x/width, (.,1) y/height
 *              (grid_sample convention, padding 'border', align_corners=True).
 *   nsets    : number of output panels; for set s (host arrays of length nsets)
 *                set_coord[s] : which coordinate block to use,
 *                set_slot[s]  : which panel slot of `out` to write,
 *              and row b of the panel samples image perm[s*B+b] (perm NULL -> b).
 *   out      : [nslots,B,Prows,ld]; row p = h*S+w holds the sample at coordinate
 *              index w*S+h (the reference's axis swap), L2-normalised over C
 *              with eps; rows >= S*S and columns >= C are zero-filled.
 *   format   : DG_PANEL_*; element type / meaning of out, out_lo (and, for CODE_SPLIT, the bf16
 *              panels out16_hi, out16_lo) as described at the enum (unused ones NULL).
 *   rnorm    : [nslots,B,Prows]  1/max(||x||,eps) per row (needed by backward).
 *   meanvec  : [nslots,B,ld]     mean over the S*S rows of the normalised panel
 *              (NULL to skip; needed for pointwise centring). */
DG_API int dg_gather_norm(const float* t, const int64_t* strides, int B, int C, int H, int W, const float* coords, int S,
                   int nsets, const int32_t* set_coord, const int32_t* set_slot, const int64_t* perm, float eps,
                   int Prows, int ld, int format, void* out, void* out_lo, void* out16_hi, void* out16_lo,
                   float* rnorm, float* meanvec, dg_stream_t stream);

/* norm() of the reference (src/modules.py:789-790, F.normalize(t, dim=1, eps)): the tensor is viewed as
 * [N, C, inner] with element strides (sN, sC, sP) — NCHW-contiguous (sC = inner, sP = 1) and channels-last
 * (sC = 1, sP = C) both work without a copy; `out` uses the same strides.  Forward only. */
DG_API int dg_norm_dim1(const float* t, int N, int C, long long inner, long long sN, long long sC, long long sP,
                        float eps, float* out, dg_stream_t stream);

/* Backward of dg_gather_norm for the code tensors: combines the unit
 * gradients of dg_corr_loss with the upstream scalars, goes back through the
 * normalisation and scatter-adds through the bilinear weights (atomicAdd; the
 * caller zero-initialises `grad`).
 *
 *   grad     : gradient w.r.t. the source tensor, same logical shape/strides as t.
 *   cn,cn_lo,rnorm : the code panels (cn_lo NULL for DG_PANEL_F32; hi + lo for CODE_SPLIT) and
 *              reciprocal norms written by the forward gather.
 *   dC1,dC2  : [npairs+1,B,Prows,ld] unit gradients from dg_corr_loss.
 *   group_w  : device [DG_NUM_GROUPS] upstream gradients of the four scalar losses.
 *   pair_group (host [npairs]) / pair_scale (host [npairs]): group of pair k and
 *              the factor that turns its mean into the group's mean (1/neg_samples
 *              for negatives, 1 otherwise).  Pair k reads code slot k as its
 *              second operand and slot 0 as its first; index npairs is the depth term.
 * Slot 0 rows receive sum_k w_k dC1[k] + w_0 dC2[0] (+ depth), slot s>0 rows w_s dC2[s]. */
DG_API int dg_gather_norm_bwd(float* grad, const int64_t* strides, int B, int C, int H, int W, const float* coords, int S,
                       int nsets, const int32_t* set_coord, const int32_t* set_slot, const int64_t* perm, float eps,
                       int Prows, int ld, const float* cn, const float* cn_lo, const float* rnorm, const float* dC1,
                       const float* dC2,
                       int npairs, const int32_t* pair_group, const float* pair_scale, int has_depth,
                       const float* group_w, dg_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused correlation loss, forward + unit gradients.  Replaces helper() for
 * every pair and depth_feature_correlation() (src/modules.py:1231-1278) plus the
 * einsum of tensor_correlation (:797-809) without materialising fd/cd.
 *
 *   pan : the normalised panels written by dg_gather_norm, slot k = second operand of
 *         pair k, slot 0 = first operand of every pair (k = 0 is the intra pair):
 *           format DG_PANEL_F32   : f_hi = fp32 [npairs,B,Prows,ldf], c_hi = fp32 [npairs,B,Prows,ldc]
 *                                   (generic CUDA-core kernel, any S);
 *           split formats         : f_hi/f_lo bf16 [npairs,B,Prows,ldf]; c_hi/c_lo fp32 [npairs,B,Prows,ldc];
 *                                   cb_hi/cb_lo bf16 [npairs,B,Prows,ldc] (tcgen05 kernel, S*S <= 1024) with
 *                                   Prows = round_up(S*S,128) up to 256 points, round_up(S*S,256) above.
 *   fmean : [npairs,B,ldf] panel row means (only read with DG_FLAG_POINTWISE).
 *   dsign : [B,Prows] depth signs from dg_depth_sign, or NULL for no depth term.
 *   pair_shift / pair_group : host [npairs].
 *   out8  : device [8] = (intra loss, intra cd mean, inter loss, inter cd mean,
 *           neg loss mean, neg cd mean, depth loss, depth dd mean); NaN if the tcgen05
 *           pipeline reported a timeout.
 *   dC1,dC2 : out [npairs+1,B,Prows,ldc] unit gradients of each pair's MEAN loss
 *           w.r.t. its first / second normalised code operand (index npairs = depth).
 *           Split formats write partial buffers that dg_gather_norm_bwd sums:
 *           dC2 [npairs+1, Prows/128, B,Prows,ldc] (one per 128-row tile of the first operand) and,
 *           above 256 points, dC1 [npairs+1, Prows/256, B,Prows,ldc] (one per 256-column group).
 *   cd_out, loss_out : optional dense [npairs,B,P,P] (NULL to skip) — the
 *           reference's 5-D tensors, [b,h,w,i,j] flattened; dd_out optional [B,P,P].
 *   fd_dbg : optional [npairs,B,Prows,Prows] raw feature correlations (tcgen05 path; tests).
 *   ws : workspace of dg_corr_loss_workspace_bytes(). */
typedef struct dg_panels {
  int format; /* DG_PANEL_F32, or DG_PANEL_FEATS_SPLIT / DG_PANEL_CODE_SPLIT for the split set */
  const void* f_hi;
  const void* f_lo;
  const void* c_hi;
  const void* c_lo;
  const void* cb_hi;
  const void* cb_lo;
} dg_panels_t;

DG_API size_t dg_corr_loss_workspace_bytes(int npairs, int B, int P);
DG_API int dg_corr_loss(const dg_panels_t* pan, const float* fmean, const float* dsign, int npairs, int B, int P,
                        int Prows, int C, int ldf, int D, int ldc, const float* pair_shift, const int32_t* pair_group,
                        float depth_shift, int flags, float* out8, float* dC1, float* dC2, float* cd_out,
                        float* loss_out, float* dd_out, float* fd_dbg, void* ws, size_t ws_bytes, dg_stream_t stream);

/* ---------------------------------------------------------------------------
 * The whole loss in one call.  dg_loss_forward enqueues FPS (or takes coordinates),
 * the gathers, the depth signs and the fused correlation loss; dg_loss_backward the
 * scatter backward.  Same kernels as the per-stage entry points above — this form
 * only removes per-stage host overhead (one FFI crossing and one allocation per step).
 * Replaces ContrastiveCorrelationLoss.forward (src/modules.py:1280-1367) and its autograd
 * backward.  The caller owns one `arena` of dg_loss_plan().total bytes that must stay
 * alive until dg_loss_backward has run. */
enum { DG_FLAG_DEPTH_TERM = 8,  /* cfg.depth_feat_correlation_loss                          */
       DG_FLAG_FPS = 16,        /* cfg.depth_sampling == "fps": coordinates from depth maps */
       DG_FLAG_FORCE_SIMT = 32, /* use the generic CUDA-core correlation kernel             */
       DG_FLAG_STAGE_NHWC = 64, /* feats / feats_pos are NCHW-contiguous: stage a channels-last copy in the arena */
       DG_FLAG_AUG_INTRA = 128  /* DepthContrastiveCorrelationLoss (src/modules.py:1370-1463): the intra pair correlates
                                   io->aug_feats (depth-augmented features) instead of io->feats */ };

typedef struct dg_loss_desc {
  int B, C, D, H, W, Hd, Wd, S, neg_samples;
  int flags; /* DG_FLAG_POINTWISE | ZERO_CLAMP | STABALIZE | DEPTH_TERM | FPS | FORCE_SIMT */
  float pos_intra_shift, pos_inter_shift, neg_inter_shift, depth_feat_shift;
} dg_loss_desc_t;

typedef struct dg_loss_plan { /* byte offsets into the arena, all 256-byte aligned */
  size_t total, coords, frn, fmean, crn, f_hi, f_lo, c_hi, c_lo, cb_hi, cb_lo, dsign, dC1, dC2, ws, ws_bytes, stage;
  int kernel; /* 0 = generic CUDA-core kernel, 1 = tcgen05 kernel */
  int Prows, ldf, ldc, npairs;
} dg_loss_plan_t;

typedef struct dg_loss_io {
  const float* feats;     /* [B,C,H,W], element strides below */
  const float* feats_pos;
  const float* code;      /* [B,D,H,W] */
  const float* code_pos;
  int64_t feats_strides[4], feats_pos_strides[4], code_strides[4], code_pos_strides[4];
  const float* depth;     /* [B,1,Hd,Wd] contiguous; needed with DG_FLAG_FPS or DG_FLAG_DEPTH_TERM */
  const float* depth_pos; /* needed with DG_FLAG_FPS */
  const float* coords;    /* [2,B,S*S,2] in [-1,1]; ignored with DG_FLAG_FPS (written to arena+coords) */
  const int64_t* perms;   /* [neg_samples,B] source image of each negative */
  void* arena;
  float* out8;            /* device [8], see dg_corr_loss */
  float* cd_out;          /* optional dense outputs, see dg_corr_loss */
  float* loss_out;
  float* dd_out;
  float* fd_dbg;
  const float* aug_feats; /* [B,C,H,W] with DG_FLAG_AUG_INTRA (strides below), else NULL */
  int64_t aug_feats_strides[4];
  void* perms_ready;      /* optional cudaEvent_t: `perms` is produced on another stream; the gathers wait for it
                             (FPS and the depth signs, which do not need perms, are enqueued before the wait) */
  /* gen_perms != 0: `perms` is an OUTPUT - the forward draws the neg_samples permutations itself (the dg_super_perms
     stream for (perm_seed, perm_offset)), in one extra CTA of the FPS launch when there is one: no launch of its own,
     nothing on the step's critical path.  dg_loss_backward reads the same buffer. */
  unsigned long long perm_seed, perm_offset;
  int gen_perms;
  /* Optional: up to two caller buffers the forward sets to zero on its stream — meant for the d_code / d_code_pos
     buffers of the coming dg_loss_backward (which accumulates into them with atomics), so that the caller needs no
     fill launches between forward and backward.  The work rides in the code-gather launch.  NULL / 0 = nothing. */
  void* clear[2];
  size_t clear_bytes[2];
  /* Sampling done ahead of time (dg_loss_presample, or the next_* job of the previous forward): without DG_FLAG_FPS the
     forward takes `coords` and `perms` from the caller anyway; `dsign` (optional, [B, Prows] as dg_loss_presample
     writes it) replaces the depth-sign launch of the depth term. */
  const float* dsign;
  /* The NEXT step's sampling, computed inside THIS forward: when next_depth is set, FPS of (next_depth, next_depth_pos)
     - coordinates, depth signs and (next_n_perms > 0) that step's permutations from (next_perm_seed, next_perm_offset)
     - is written to next_coords [2,B,S*S,2] / next_dsign [B,Prows] (may be NULL) / next_perms [n,B], exactly as
     dg_loss_presample would.  The work rides as extra CTAs of the correlation kernel, on SMs that its item list
     leaves idle, so the next forward (called with these buffers as coords / dsign / perms and without DG_FLAG_FPS)
     has no FPS on its critical path.  Shapes are this step's (B, Hd, Wd, H, W, S). */
  const float* next_depth;
  const float* next_depth_pos;
  float* next_coords;
  float* next_dsign;
  int64_t* next_perms;
  unsigned long long next_perm_seed, next_perm_offset;
  int next_n_perms;
} dg_loss_io_t;

typedef struct dg_loss_grads {
  const float* g[4];      /* device scalars: upstream gradients of (intra, inter, neg mean, depth); NULL = 0 */
  float* d_code;          /* zero-initialised by the caller; NULL to skip */
  float* d_code_pos;
  int64_t d_code_strides[4], d_code_pos_strides[4];
} dg_loss_grads_t;

DG_API int dg_loss_plan(const dg_loss_desc_t* desc, dg_loss_plan_t* plan);
/* The sampling of one forward as its own launch (any stream - e.g. a side stream under the backbone's forward pass):
 * farthest_point_sampling_depth of depth and depth_pos (src/modules.py:999-1037) -> coords [2,B,S*S,2] in [-1,1],
 * the depth signs of the depth term -> dsign [B,Prows] (NULL to skip; Prows from dg_loss_plan), and n_perms > 0
 * permutations of super_perm (src/modules.py:1184-1188) from (perm_seed, perm_offset) -> perms [n_perms,B].
 * Uses desc->B, Hd, Wd, H, W, S. */
DG_API int dg_loss_presample(const dg_loss_desc_t* desc, const float* depth, const float* depth_pos, float* coords,
                             float* dsign, int64_t* perms, int n_perms, unsigned long long perm_seed,
                             unsigned long long perm_offset, dg_stream_t stream);
DG_API int dg_loss_forward(const dg_loss_desc_t* desc, const dg_loss_io_t* io, dg_stream_t stream);
DG_API int dg_loss_backward(const dg_loss_desc_t* desc, const dg_loss_io_t* io, const dg_loss_grads_t* grads,
                            dg_stream_t stream);

/* ---------------------------------------------------------------------------
 * Cosine-similarity k-nearest-neighbour build.  Replaces the einsum + topk
 * loop of src/precompute_knns.py:99-113 for a block of query rows.
 *
 *   q  : [Nq,F] query rows, db : [N,F] database rows (fp32, row pitch F; unit-norm for cosine similarity — the
 *        tensor-core pass measures the row norms and scales its error bound by |q| * max|d|, so other norms stay exact).
 *   idx : out [Nq,k] int64, sorted by descending fp32 similarity (k <= 32).
 *   sims: optional out [Nq,k] fp32 similarities.
 *   ws  : dg_knn_workspace_bytes() bytes.  Once the stream has drained, its first three int32 hold diagnostics of the
 *        tensor-core path: [0] pipeline error flag (0 = ok), [1] number of query rows whose candidate list could not be
 *        certified and were recomputed by the exact fp32 kernel, [2] bit pattern of the largest squared database row norm. */
DG_API size_t dg_knn_workspace_bytes(int Nq, int N, int F, int k);
DG_API int dg_knn_topk(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                size_t ws_bytes, dg_stream_t stream);

/* The same build for a query-row SHARD of a multi-GPU job (query rows shard, the database is
 * all-gathered), in phases so that the exchange overlaps useful work.  The shard's query rows are database rows
 * [row_lo, row_lo + Nq).
 *   dg_knn_shard_begin : needs only the local rows `local` [Nq,F].  DG_KNN_SPLIT_LOCAL resets the diagnostics header and
 *                        writes the rows' bf16 hi/lo panels into the workspace (panel rows [row_lo, row_lo+Nq));
 *                        DG_KNN_PASS_LOCAL ranks the rows against themselves - the part of the database this rank
 *                        already holds, i.e. what it can do while the exchange is in flight (it leaves a quarter of the
 *                        SMs to the communication kernels unless DG_KNN_ALL_SMS says the exchange needs none).
 *   dg_knn_shard_finish: `db` is the complete [N,F] fp32 database (rows [row_lo, row_lo+Nq) equal to `local`).
 *                        DG_KNN_SPLIT_REMOTE writes the panels of all other rows (skip it when the host copied the
 *                        peers' panel rows into the workspace itself: dg_knn_panel_layout), DG_KNN_PASS_REMOTE continues
 *                        the candidate lists of phase 1 over the remote rows, DG_KNN_RERANK finishes exactly like
 *                        dg_knn_topk (`db` is only read by SPLIT_REMOTE and RERANK).  npeer_max: number of header slots
 *                        ws int32[32 .. 32+npeer_max) that hold the squared-norm maxima (bit patterns, ws int32[2] of the
 *                        other ranks) when their panels were copied in; 0 otherwise.
 * Same workspace (dg_knn_workspace_bytes(Nq,N,F,k), Nq = the LARGEST shard if workspaces are to be laid out alike),
 * untouched between the calls; same stream order; same diagnostics header.  Result identical to
 * dg_knn_topk(db + row_lo*F, db, Nq, N, ...).  Replaces the per-shard form of src/precompute_knns.py:99-113. */
#define DG_KNN_SPLIT_LOCAL 1
#define DG_KNN_PASS_LOCAL 2
#define DG_KNN_BEGIN_ALL 3
#define DG_KNN_ALL_SMS 8      /* with DG_KNN_PASS_LOCAL: the exchange needs no SMs (copy engines) - use all of them */
#define DG_KNN_SPLIT_REMOTE 1
#define DG_KNN_PASS_REMOTE 2
#define DG_KNN_RERANK 4
#define DG_KNN_FINISH_ALL 7
DG_API int dg_knn_shard_begin(const float* local, int Nq, int row_lo, int N, int F, int k, void* ws, size_t ws_bytes,
                       int phases, dg_stream_t stream);
DG_API int dg_knn_shard_finish(const float* db, int Nq, int row_lo, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                        size_t ws_bytes, int phases, int npeer_max, dg_stream_t stream);
/* Byte offsets of the bf16 hi / lo panels ([N, row_bytes / 2] each) inside a KNN workspace and the pitch of a panel row. */
DG_API int dg_knn_panel_layout(int N, int F, size_t* hi_offset, size_t* lo_offset, size_t* row_bytes);
/* How many of those panels the tensor pass reads, i.e. what a host-side exchange has to move: 1 = the hi panel only
 * (the default fast pass: one fp16 panel), 2 = hi and lo (DEPTHG_B200_KNN_PASS=split3: the 3-term bf16 pass). */
DG_API int dg_knn_panel_count(void);

/* n device-to-device copies, copy i enqueued on streams[i] (copy engines; peer-mapped pointers allowed): the exchange
 * step of the sharded KNN build when the host moves panel rows itself. */
DG_API int dg_memcpy_batch(int n, void* const* dst, const void* const* src, const size_t* bytes, const dg_stream_t* streams);

/* Mean-pool + L2-normalise of get_feats (src/precompute_knns.py:19):
 * t [N,C,H,W] with element strides (host) -> out [N,C], eps = 1e-12. */
DG_API int dg_pool_normalize(const float* t, const int64_t* strides, int N, int C, int H, int W, float eps, float* out,
                      dg_stream_t stream);

/* ---------------------------------------------------------------------------
 * Probe losses on the detached code map.  Both train only
 * their own parameters, so the entry points return the loss AND the unit
 * gradients of that loss in one call (forward and backward are one pass).
 * code : [B,D,h,w] fp32 with element strides (host array of 4), D <= 128.
 * ws   : dg_probe_workspace_bytes(B,h,w,D,K) bytes of device scratch.       */
DG_API size_t dg_probe_workspace_bytes(int B, int h, int w, int D, int K);

/* Linear probe: replaces  linear_probe(code) -> F.interpolate(bilinear,
 * align_corners=False, size=label.shape[-2:]) -> permute/reshape -> [mask] ->
 * CrossEntropyLoss  (src/train_segmentation.py:419-437).
 *   weight [K,D], bias [K] (may be NULL): the 1x1 conv, K <= 32.
 *   labels [B,Hl,Wl] int64 with element strides; entries outside [0,K) are masked out.
 *   loss_out: device scalar = mean CE over the unmasked label pixels.
 *   dweight [K,D], dbias [K]: d loss / d weight, d loss / d bias (NULL = forward only). */
DG_API int dg_linear_probe_ce(const float* code, const int64_t* strides, int B, int D, int h, int w, const float* weight,
                       const float* bias, int K, const int64_t* labels, const int64_t* label_strides, int Hl, int Wl,
                       float* loss_out, float* dweight, float* dbias, void* ws, size_t ws_bytes, dg_stream_t stream);

/* Cluster probe: replaces ClusterLookup.forward (src/modules.py:659-675).
 *   clusters [N,D] (un-normalised parameter), N <= 32.
 *   mode 0: alpha=None  -> probs = one_hot(argmax), loss, dclusters (NULL = skip)
 *   mode 1: alpha given -> probs = softmax(inner*alpha), loss (forward only)
 *   mode 2: log_probs   -> probs_out = log_softmax(inner*alpha), no loss
 *   probs_out: [B,h,w,N] (the reference returns the same memory permuted to [B,N,h,w]); may be NULL in modes 0/1. */
DG_API int dg_cluster_probe(const float* code, const int64_t* strides, int B, int D, int h, int w, const float* clusters,
                     int N, int mode, float alpha, float* loss_out, float* probs_out, float* dclusters, void* ws,
                     size_t ws_bytes, dg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEPTHG_B200_H_ */
