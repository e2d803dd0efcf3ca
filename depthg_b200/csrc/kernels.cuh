// Internal launcher interface shared by the per-kernel C-ABI wrappers and the one-call
// loss entry points (loss_api.cu).  Everything here enqueues on `st` and returns DG_* codes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dg {

// One gathered coordinate set: which tensor to read (pointer + element strides), which
// coordinate block, which panel slot to write, and which row of the perms table remaps the
// batch index (-1 = identity).  Lets ONE launch serve several source tensors.
struct SetDesc {
  const float* src;
  int64_t sb, sc, sh, sw;
  int32_t coord, slot, perm_row;
};
struct SetTable {
  SetDesc s[DG_MAX_SETS];
};

enum { FMT_F32 = 0, FMT_FEATS_SPLIT = 1, FMT_CODE_SPLIT = 2 };

// The 16-bit split panels hold fp16 hi/lo of the normalised rows times a power of two: x * s = hi + lo + r with
// |r| <= 2^-22 |x * s| (fp16 keeps 11 significand bits; the scale keeps lo out of the subnormal range for every
// element down to ~1e-6 of a unit-norm row), so a 3-term product hi.hi + hi.lo + lo.hi is fp32-grade (~2^-22) where
// the bf16 split of round 1 stopped at 2^-16.  Normalised rows have |x| <= 1, so nothing can overflow.
constexpr float F16_FEAT_SCALE = 64.f;    // backbone feature panels
constexpr float F16_CODE_SCALE = 256.f;   // code panels (gradient-GEMM operands)

struct GatherOut {
  // FMT_F32        : out (fp32 [slot,b,Prows,ld])
  // FMT_FEATS_SPLIT: hi16/lo16 (fp16 [slot,b,Prows,ld]) : F16_FEAT_SCALE * x ~= hi + lo
  // FMT_CODE_SPLIT : out = tf32-rounded hi, out_lo = x - hi (fp32 [slot,b,Prows,ld]) for the cd product,
  //                  hi16/lo16 (fp16 [slot,b,Prows,ld]) : F16_CODE_SCALE * x, the same rows for the gradient GEMMs
  float* out;
  float* out_lo;
  __half* hi16;
  __half* lo16;
  // interleave != 0: ONE fp16 panel of pitch 2*ld at hi16, every 32-channel chunk stored as [32 hi | 32 lo] (128 bytes),
  // so a TMA box row of the persistent tcgen05 kernel carries a chunk's hi and lo halves in one 128-byte line
  int interleave;
  float* rnorm;
  float* meanvec;
};

// element offsets of channel c (hi part; the lo part sits 32 elements further) in an interleaved 16-bit panel row
__host__ __device__ __forceinline__ size_t il_col(int c) { return ((size_t)(c >> 5) << 6) + (size_t)(c & 31); }

struct PairTable {
  int32_t group[DG_MAX_PAIRS + 1];
  float scale[DG_MAX_PAIRS + 1];
};

// upstream gradients of the four scalar losses: either one device array [4] or four separate
// device scalars (null = 0), so the caller needs no stack/cat kernel
struct GroupW {
  const float* arr;
  const float* ptr[DG_NUM_GROUPS];
};

// pair_dots work that the one-call loss path appends to the code gather's grid (the tcgen05 correlation kernel
// needs dots[k,b] = <mean row of F1[b], mean row of F2[k,b]> and a cleared error flag / completion counter)
struct FpsArgs;   // fps_body.cuh

struct DotsJob {
  const float* fmean;  // [slot,b,nsplit,ldf] partial means, or null (not pointwise): only the flags are cleared
  float* dots;         // [npairs,B]
  int* err;            // [4] error flag + completion counter
  int nsplit, npairs, B, ldf;
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];
  float4* clr[2];           // caller buffers the code-gather CTAs set to zero (dg_loss_io_t::clear), or null
  unsigned long long clr_n16[2];  // their sizes in 16-byte units
};

// One 256-thread block per (pair k, image b); block 0 also clears the flags.
__device__ __forceinline__ void pair_dots_body(const DotsJob& j, int w) {
  __shared__ float pd_red[8];
  if (w == 0 && threadIdx.x < 4 && j.err) j.err[threadIdx.x] = 0;
  if (j.fmean == nullptr) return;
  const int k = w / j.B, b = w - k * j.B;
  const float* m1 = j.fmean + ((size_t)j.fs1[k] * j.B + b) * j.nsplit * j.ldf;   // nsplit partial means each
  const float* m2 = j.fmean + ((size_t)j.fs2[k] * j.B + b) * j.nsplit * j.ldf;
  float s = 0.f;
  for (int c = threadIdx.x; c < j.ldf; c += 256) {
    float a = 0.f, bb = 0.f;
    for (int i = 0; i < j.nsplit; ++i) {
      a += __ldg(m1 + (size_t)i * j.ldf + c);
      bb += __ldg(m2 + (size_t)i * j.ldf + c);
    }
    s += a * bb;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) pd_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += pd_red[i];
    j.dots[w] = t;
  }
}

// dsign != null: also writes the depth signs of depth_a (what launch_depth_sign computes) from the same CTA
// Negative-pair permutations drawn by one extra CTA of the FPS launch (n = 0: none)
struct PermJob {
  unsigned long long seed, offset;
  int n, B;
  int64_t* out;   // [n,B]
};
int launch_fps(const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H, int W, int S, float factor,
               float far_plane, int affine, float* coords, int32_t* idx, cudaStream_t st, float* dsign = nullptr,
               int sign_pitch = 0, float sign_eps = 0.f, const PermJob* perm_job = nullptr);
int make_fps_args(FpsArgs* a, size_t* smem_out, const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H,
                  int W, int S, float factor, float far_plane, int affine, float* coords, int32_t* idx, float* dsign,
                  int sign_pitch, float sign_eps, const PermJob* perm_job);
int launch_depth_sign(const float* depth, int B, int Hd, int Wd, int S, float eps, int out_pitch, float* out,
                      cudaStream_t st);
// meanvec is written as `nsplit` partial means per (slot, image): [slot,b,nsplit,ld], each already divided by P
int gather_nsplit(int fmt, const SetTable& tab, int nsets, int C, int ld);
int launch_gather(int fmt, const SetTable& tab, int nsets, int B, int C, int H, int W, const float* coords, int S,
                  const int64_t* perms, float eps, int Prows, int ld, int nsplit, const GatherOut& o, cudaStream_t st,
                  const DotsJob* tail = nullptr);  // tail: extra CTAs of the code gather that do the pair_dots work
int launch_nchw_to_nhwc(const float* a, const float* b, int B, int C, int HW, int64_t sb_a, int64_t sb_b, float* out_a,
                        float* out_b, cudaStream_t st);
int launch_gather_all(const SetTable& fsets, const SetTable& csets, int nsets, int B, int C, int D, int H, int W,
                      const float* coords, int S, const int64_t* perms, float eps, int Prows, int ldf, int ldc, int fsplit,
                      const GatherOut& fo, const GatherOut& co, cudaStream_t st);
int launch_gather_bwd(const SetTable& tab, int nsets, int B, int C, int H, int W, const float* coords, int S,
                      const int64_t* perms, float eps, int Prows, int ld, const float* cn, const float* cn_lo,
                      const float* rnorm, const float* dC1, const float* dC2, int npairs, const PairTable& pt,
                      int has_depth, const GroupW& gw, cudaStream_t st, int ni = 1, int nj = 1, int njw = 256);
int launch_corr_finalize(const float* partials, int npairs, int B, int P, const int32_t* group, int has_depth,
                         const int* err, float* out8, int n_pt, cudaStream_t st);
int corr_loss_simt(const float* fn, const float* cn, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P,
                   int Prows, int ldf, int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift,
                   int flags, float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out,
                   void* ws, cudaStream_t st, const int32_t* fslot1 = nullptr, const int32_t* fslot2 = nullptr);
int corr_loss_umma(const dg_panels_t* pan, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P, int Prows, int ldf,
                   int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift, int flags,
                   float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out, float* fd_dbg,
                   void* ws, cudaStream_t st, const int32_t* fslot1 = nullptr, const int32_t* fslot2 = nullptr,
                   int nfslots = 0, bool dots_done = false);
// the persistent, double-buffered tcgen05 kernel (corr_pipe.cu): same arguments; Prows a multiple of 128, partial
// gradient buffers dC1 [npairs+1, Prows/128 (column tile), ...] and dC2 [npairs+1, Prows/128 (row tile), ...]
// `ride` (optional): an FPS launch (make_fps_args) executed by extra CTAs of this kernel - the NEXT step's sampling, see
// PipeParams::ride; only when corr_pipe_can_ride(*ride, ride_smem).
int corr_loss_pipe(const dg_panels_t* pan, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P, int Prows, int ldf,
                   int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift, int flags,
                   float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out, float* fd_dbg,
                   void* ws, cudaStream_t st, const int32_t* fslot1 = nullptr, const int32_t* fslot2 = nullptr,
                   int nfslots = 0, bool dots_done = false, const FpsArgs* ride = nullptr, size_t ride_smem = 0);
bool corr_pipe_can_ride(const FpsArgs& a, size_t smem);
size_t corr_pipe_workspace_floats(int npairs, int B, int P);
// where corr_loss_umma keeps its error flag / completion counter and the pair dots inside its workspace
void umma_ws_layout(void* ws, int** err, float** dots);
size_t corr_workspace_bytes(int npairs, int B, int P);

}  // namespace dg
