// Shared helpers for the depthg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/depthg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "depthg_b200 kernels are written for sm_100a (B200) only"
#endif

namespace dg {

void set_error(const char* fmt, ...);
void count_launch();  // bumps the process-wide kernel-launch counter (dg_kernel_launches)
// optional per-kernel CUDA-event timing (dg_profile_enable / dg_profile_collect): no-ops unless enabled
void profile_pre(cudaStream_t st);
void profile_post(const char* name);

#define DG_PRE(st) ::dg::profile_pre(st)

#define DG_REQUIRE(cond, code, ...)     \
  do {                                  \
    if (!(cond)) {                      \
      ::dg::set_error(__VA_ARGS__);     \
      return (code);                    \
    }                                   \
  } while (0)

#define DG_CUDA_OK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      ::dg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return DG_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

#define DG_LAUNCH_OK(name)                                                                   \
  do {                                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess) {                                                                 \
      ::dg::set_error("launch of %s failed: %s", name, cudaGetErrorString(e_));              \
      return DG_ERR_CUDA;                                                                    \
    }                                                                                        \
    ::dg::count_launch();                                                                    \
    ::dg::profile_post(name);                                                                \
  } while (0)

// cudaFuncSetAttribute applies to the CURRENT device only, so "already configured" is remembered per device
// (one process per GPU is the normal deployment, but tests may touch several devices from one process).
struct PerDevice {
  size_t v[64];
};
inline size_t& per_device(PerDevice& s) {
  int d = 0;
  cudaGetDevice(&d);
  return s.v[(d >= 0 && d < 64) ? d : 0];
}

// ---- programmatic dependent launch (PDL).  The kernels of a step form a chain in one stream; launched with the
// "programmatic stream serialization" attribute, kernel n+1 is set up on the SMs while kernel n drains, instead of
// after it has drained: each of them calls pdl_wait() before it touches global memory (it blocks until the previous
// kernel has completed and flushed) and pdl_trigger() as soon as it runs (the next kernel may be scheduled whenever
// SM resources allow).  Takes the ~3-5 us launch gap per boundary off a 0.3 ms step.  DEPTHG_B200_PDL=0 disables it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Bilinear corner set of torch.grid_sample(padding_mode='border', align_corners=True)
// for one normalised coordinate pair; weights in ATen's nw/ne/sw/se order.
struct Corners {
  int x0, y0, x1, y1;      // x1/y1 may equal W/H (out of range) -> weight is dropped
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
  bool x1_ok, y1_ok;
};

__device__ __forceinline__ Corners bilinear_corners(float gx, float gy, int H, int W) {
  float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  float fx = floorf(ix), fy = floorf(iy);
  Corners c;
  c.x0 = (int)fx;
  c.y0 = (int)fy;
  c.x1 = c.x0 + 1;
  c.y1 = c.y0 + 1;
  float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy;  // distance to the se corner
  float dx = ix - fx, dy = iy - fy;
  c.w00 = ex * ey;
  c.w01 = dx * ey;
  c.w10 = ex * dy;
  c.w11 = dx * dy;
  c.x1_ok = c.x1 < W;
  c.y1_ok = c.y1 < H;
  return c;
}

}  // namespace dg
