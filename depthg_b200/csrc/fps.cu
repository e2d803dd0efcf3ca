// Depth-guided farthest point sampling, one CTA per image.
//
// Replaces farthest_point_sampling_depth / depth2points / fps of the reference
// (/root/reference/src/modules.py:999-1037, :988-996, :939-985), which pools on
// the device, copies every image to the host and runs a 120-round NumPy loop.
// Here the whole chain stays in one kernel:
//   1. adaptive average pooling of the [Hd,Wd] depth to [H,W] straight from global
//      memory: a warp per row of the feature grid with coalesced 128-bit loads when
//      the windows are the aligned 8 x 8 blocks (a thread per window otherwise), every
//      window summed row-major with sequential fp32 adds, then ATen's sum / kh / kw;
//   2. lifting to 3-D with one explicit rounding per operation (no FMA);
//   3. S*S-1 rounds of (distance to last pick, running min, block argmax).  The
//      candidate key is the int view of the non-negative fp32 distance, folded with a
//      SIGNED integer minimum (a picked point drops to key 0 by itself: no marking), so
//      a warp argmax is two redux.sync instructions (max of the key, then min of the
//      index among the maxima = NumPy's first-argmax); eight warp results meet in
//      shared memory behind ONE barrier per round;
//   4. the selection ORDER is discarded like the reference does: a block scan of
//      the taken flags emits the picks in raster order as indices and as
//      normalised (row/H, col/W) coordinates.
#include <string.h>

#include "fps_body.cuh"

namespace dg {

template <int PPT, int RT, int PR>
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const __grid_constant__ FpsArgs a) {
  extern __shared__ __align__(16) float fps_dyn_smem[];
  pdl_trigger();
  pdl_wait();   // (the depth maps may be the previous kernel's output; nothing is written before this point)
  if (a.pj.n > 0 && (int)blockIdx.x == a.nimg) {   // one extra CTA: the step's negative-pair permutations
    super_perms_block(a.pj.seed, a.pj.offset, a.pj.n, a.pj.B, a.pj.out, reinterpret_cast<int*>(fps_dyn_smem));
    return;
  }
  if (a.stage)
    fps_cta<PPT, RT, PR, true>(a, (int)blockIdx.x, fps_dyn_smem);
  else
    fps_cta<PPT, RT, PR, false>(a, (int)blockIdx.x, fps_dyn_smem);
}

// s = d / max(|d|, eps) of the align_corners=True bilinear resample of depth to SxS.
__global__ void depth_sign_kernel(const float* __restrict__ depth, int B, int Hd, int Wd, int S, float eps,
                                  int out_pitch, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * out_pitch) return;
  const int b = t / out_pitch, p = t - b * out_pitch;
  out[t] = depth_sign_value(depth + (size_t)b * Hd * Wd, Hd, Wd, S, p, eps);
}

// Validates the arguments of an FPS launch and fills `a`; `smem` = dynamic shared memory one CTA needs.
int make_fps_args(FpsArgs* a, size_t* smem_out, const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H,
                  int W, int S, float factor, float far_plane, int affine, float* coords, int32_t* idx, float* dsign,
                  int sign_pitch, float sign_eps, const PermJob* perm_job) {
  PermJob pj;
  memset(&pj, 0, sizeof pj);
  if (perm_job) pj = *perm_job;
  DG_REQUIRE(depth_a && coords, DG_ERR_INVALID, "dg_fps_coords: null pointer");
  DG_REQUIRE(B > 0 && Hd > 0 && Wd > 0 && H > 0 && W > 0 && S > 0, DG_ERR_INVALID, "dg_fps_coords: bad sizes");
  DG_REQUIRE(H <= Hd && W <= Wd, DG_ERR_UNSUPPORTED, "dg_fps_coords: pooling must not upsample (%dx%d -> %dx%d)", Hd,
             Wd, H, W);
  const int npts = H * W;
  DG_REQUIRE(S * S <= npts, DG_ERR_INVALID, "dg_fps_coords: S*S=%d exceeds H*W=%d points", S * S, npts);
  DG_REQUIRE(npts <= 4096, DG_ERR_UNSUPPORTED, "dg_fps_coords: H*W=%d > 4096 not supported", npts);
  // fast pooling path: aligned 8 x 8 windows and 16-byte aligned image rows (the reference's 224 x 224 -> 28 x 28)
  const bool stage = Hd == 8 * H && Wd == 8 * W && (Wd % 4) == 0 && ((Hd * Wd) % 4) == 0 &&
                     ((reinterpret_cast<uintptr_t>(depth_a) | reinterpret_cast<uintptr_t>(depth_b)) % 16 == 0);
  size_t smem = fps_smem_bytes(npts);
  if (pj.n > 0) {   // the extra CTA shuffles in shared memory and needs one thread per permutation
    DG_REQUIRE(pj.n <= FPS_THREADS && (size_t)pj.n * pj.B * sizeof(int) <= 64 * 1024, DG_ERR_UNSUPPORTED,
               "fps: %d permutations of %d do not fit the fused draw", pj.n, pj.B);
    if (smem < (size_t)pj.n * pj.B * sizeof(int)) smem = (size_t)pj.n * pj.B * sizeof(int);
  }
  a->depth_a = depth_a; a->depth_b = depth_b;
  a->B = B; a->Hd = Hd; a->Wd = Wd; a->H = H; a->W = W; a->nsel = S * S;
  a->factor = factor; a->far_plane = far_plane; a->affine = affine;
  a->coords = coords; a->idx = idx; a->dsign = dsign;
  a->sign_S = S; a->sign_pitch = sign_pitch; a->sign_eps = sign_eps;
  a->stage = stage ? 1 : 0;
  a->nimg = depth_b ? 2 * B : B;
  a->pj = pj;
  a->clk = nullptr;
  *smem_out = smem;
  return DG_OK;
}

int launch_fps(const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H, int W, int S, float factor,
               float far_plane, int affine, float* coords, int32_t* idx, cudaStream_t st, float* dsign, int sign_pitch,
               float sign_eps, const PermJob* perm_job) {
  FpsArgs a;
  size_t smem = 0;
  int rc = make_fps_args(&a, &smem, depth_a, depth_b, B, Hd, Wd, H, W, S, factor, far_plane, affine, coords, idx, dsign,
                         sign_pitch, sign_eps, perm_job);
  if (rc != DG_OK) return rc;
  const int npts = H * W;
#define DG_FPS_LAUNCH(PPT, RTV, PRV)                                                                              \
  do {                                                                                                            \
    static PerDevice configured_pd = {};                                                                          \
    size_t& configured = per_device(configured_pd);                                                               \
    if (smem > 48 * 1024 && smem > configured) {                                                                  \
      DG_CUDA_OK(cudaFuncSetAttribute(fps_kernel<PPT, RTV, PRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      configured = smem;                                                                                          \
    }                                                                                                             \
    DG_PRE(st);                                                                                                   \
    launch_pdl(fps_kernel<PPT, RTV, PRV>, dim3(a.nimg + (a.pj.n > 0 ? 1 : 0)), dim3(FPS_THREADS), smem, st, a);   \
  } while (0)
  static int rt_env = -1;  // DEPTHG_B200_FPS_RT = 128 | 256 round threads (experiments); default FPS_DEFAULT_RT
  if (rt_env < 0) {
    const char* e = getenv("DEPTHG_B200_FPS_RT");
    rt_env = e ? atoi(e) : 0;
  }
  if (npts <= 896) {  // the 28x28 grid of the reference (784 points) lives here
    if (rt_env == 128) DG_FPS_LAUNCH(4, 128, 8); else DG_FPS_LAUNCH(4, 256, 4);
  } else if (npts <= 8 * FPS_THREADS) {
    DG_FPS_LAUNCH(8, 256, 8);
  } else {
    DG_FPS_LAUNCH(16, 256, 16);
  }
#undef DG_FPS_LAUNCH
  DG_LAUNCH_OK("fps_kernel");
  return DG_OK;
}

int launch_depth_sign(const float* depth, int B, int Hd, int Wd, int S, float eps, int out_pitch, float* out,
                      cudaStream_t st) {
  DG_REQUIRE(depth && out, DG_ERR_INVALID, "dg_depth_sign: null pointer");
  DG_REQUIRE(B > 0 && Hd > 0 && Wd > 0 && S > 0 && out_pitch >= S * S, DG_ERR_INVALID, "dg_depth_sign: bad sizes");
  const int n = B * out_pitch;
  DG_PRE(st);
  depth_sign_kernel<<<ceil_div(n, 256), 256, 0, st>>>(depth, B, Hd, Wd, S, eps, out_pitch, out);
  DG_LAUNCH_OK("depth_sign_kernel");
  return DG_OK;
}

}  // namespace dg

extern "C" int dg_fps_coords(const float* depth_a, const float* depth_b, int B, int Hd, int Wd, int H, int W, int S,
                             float factor, float far_plane, int affine, float* coords, int32_t* idx,
                             dg_stream_t stream) {
  return dg::launch_fps(depth_a, depth_b, B, Hd, Wd, H, W, S, factor, far_plane, affine, coords, idx,
                        reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dg_depth_sign(const float* depth, int B, int Hd, int Wd, int S, float eps, int out_pitch, float* out,
                             dg_stream_t stream) {
  return dg::launch_depth_sign(depth, B, Hd, Wd, S, eps, out_pitch, out, reinterpret_cast<cudaStream_t>(stream));
}
