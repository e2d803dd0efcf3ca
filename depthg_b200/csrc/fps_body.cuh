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
#pragma once
#include "kernels.cuh"
#include "sampler.cuh"

namespace dg {

constexpr int FPS_SCRATCH_FLOATS = 64;   // head of the dynamic shared memory: s_kv[2][8] (int2) + s_scan[8], padded to 256 B

// Everything one FPS launch needs (kernel parameter of fps_kernel; a member of the correlation kernel's parameters when
// the next step's sampling rides along).
struct FpsArgs {
  const float* depth_a;
  const float* depth_b;        // may be null
  int B, Hd, Wd, H, W, nsel;
  float factor, far_plane;
  int affine;
  float* coords;               // [nimg, nsel, 2]
  int32_t* idx;                // [nimg, nsel] or null
  float* dsign;                // [B, sign_pitch] or null
  int sign_S, sign_pitch;
  float sign_eps;
  int stage;                   // 1: the fast pooling path (aligned 8 x 8 windows, see fps_cta<.., FAST8>)
  int nimg;                    // CTAs 0 .. nimg-1 run images; CTA nimg (when pj.n > 0) draws the permutations
  PermJob pj;
  long long* clk;              // debug: [gridDim.x][16] %globaltimer stamps (slots 3-4: pooled+lifted / rounds done)
};


constexpr int FPS_THREADS = 256;  // setup (pooling, lifting) and emission threads; the rounds use the first RT of them
constexpr int FPS_WARPS = FPS_THREADS / 32;
// measured on B200 (28x28 grid, S=11, one image per CTA; scripts/micro/fps_micro.cu): cycles per round 350 at RT=256,
// 468 at 128, 367 at 512, 576 at 64 — a warp issues one instruction every other cycle, so fewer round threads pay in
// issue slots what they save on the barrier; the floor of the chain is LDS 29 + redux pair 49 + STS/BAR/LDS 59 +
// redux pair 42 cycles.
constexpr int FPS_DEFAULT_RT = 256;  // round threads for the <= 896-point case (see fps_rounds)

// s = d / max(|d|, eps) of the align_corners=True bilinear resample of one [Hd,Wd] image at point p of the SxS grid
// (0 for p >= S*S).  `d` may point to global or shared memory.
__device__ __forceinline__ float depth_sign_value(const float* d, int Hd, int Wd, int S, int p, float eps) {
  if (p >= S * S) return 0.f;
  const int h = p / S, w = p - h * S;
  const float sy = S > 1 ? __fdiv_rn((float)(Hd - 1), (float)(S - 1)) : 0.f;
  const float sx = S > 1 ? __fdiv_rn((float)(Wd - 1), (float)(S - 1)) : 0.f;
  const float fy = __fmul_rn(sy, (float)h), fx = __fmul_rn(sx, (float)w);
  const int y0 = min((int)fy, Hd - 1), x0 = min((int)fx, Wd - 1);
  const int y1 = y0 + (y0 < Hd - 1 ? 1 : 0), x1 = x0 + (x0 < Wd - 1 ? 1 : 0);
  const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
  const float v00 = d[(size_t)y0 * Wd + x0], v01 = d[(size_t)y0 * Wd + x1];
  const float v10 = d[(size_t)y1 * Wd + x0], v11 = d[(size_t)y1 * Wd + x1];
  const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  return v / fmaxf(fabsf(v), eps);
}

// Packed fp32 pairs (FADD2 / FMUL2 of sm_100): two IEEE-rounded operations per instruction, bit-identical to the scalar
// ones.  Only subtraction and multiplication are packed: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// (ONE rounding) whatever -fmad says, which would break the bit-exact distances, so the two additions stay scalar
// (scripts/sass_evidence.py checks that no FFMA is left in this kernel's round loop).
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_sqr(unsigned long long a) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(a));
  return r;
}

// The S*S-1 selection rounds, run by the first RT threads of the CTA: thread t owns points t, t+RT, ... (PR per thread,
// in registers as packed pairs).  Per round (measured on B200, scripts/micro/fps_micro.cu: 350 cycles at RT = 256 /
// PR = 4 against 588 for the round-1 loop):
//   * key = min(key, int view of the distance) as a SIGNED integer minimum: a point that can no longer be picked needs no
//     test and no marking, because the distance of a picked point to itself is exactly +0, so its key drops to 0 in the
//     round after its pick and can only win again when every remaining key is 0 too (coincident points) — that case
//     (block maximum == 0) takes a slow path that consults the taken flags, like np.delete does in the reference;
//   * warp argmax = redux.max of the key, then redux.min of the lowest owned index holding that key; the RT/32 warp
//     results meet in shared memory as one 8-byte word each behind ONE named barrier and every warp folds them with the
//     same two redux ops.
template <int PR, int RT>
__device__ __forceinline__ void fps_rounds(const float* sX, const float* sY, const float* sZ, unsigned char* sTaken,
                                           int npts, int nsel, int2 (*s_kv)[FPS_WARPS]) {
  constexpr int RW = RT / 32;
  static_assert(PR % 2 == 0 && RW <= FPS_WARPS && (RW & (RW - 1)) == 0, "fps_rounds: packed pairs, power-of-two warps");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long X2[PR / 2], Y2[PR / 2], Z2[PR / 2];
  int key[PR];  // int view of the running min distance; -1 = not a point
#pragma unroll
  for (int j = 0; j < PR; j += 2) {
    const int i0 = j * RT + tid, i1 = i0 + RT;
    const bool ok0 = i0 < npts, ok1 = i1 < npts;
    X2[j / 2] = f2_pack(ok0 ? sX[i0] : 0.f, ok1 ? sX[i1] : 0.f);
    Y2[j / 2] = f2_pack(ok0 ? sY[i0] : 0.f, ok1 ? sY[i1] : 0.f);
    Z2[j / 2] = f2_pack(ok0 ? sZ[i0] : 0.f, ok1 ? sZ[i1] : 0.f);
    key[j] = ok0 ? 0x7f800000 : -1;  // +inf; point 0 (the first pick) drops to 0 in round 1
    key[j + 1] = ok1 ? 0x7f800000 : -1;
  }
  int last = 0;
  for (int r = 1; r < nsel; ++r) {
    const float lx = sX[last], ly = sY[last], lz = sZ[last];
    const unsigned long long lx2 = f2_pack(lx, lx), ly2 = f2_pack(ly, ly), lz2 = f2_pack(lz, lz);
    int bk = -1;
#pragma unroll
    for (int j = 0; j < PR; j += 2) {
      float xx0, xx1, yy0, yy1, zz0, zz1;
      f2_unpack(f2_sqr(f2_sub(lx2, X2[j / 2])), xx0, xx1);
      f2_unpack(f2_sqr(f2_sub(ly2, Y2[j / 2])), yy0, yy1);
      f2_unpack(f2_sqr(f2_sub(lz2, Z2[j / 2])), zz0, zz1);
      const float d0 = __fadd_rn(__fadd_rn(xx0, yy0), zz0), d1 = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
      key[j] = min(key[j], __float_as_int(d0));  // non-negative floats: int order == float order; -1 stays -1
      key[j + 1] = min(key[j + 1], __float_as_int(d1));
      bk = max(bk, max(key[j], key[j + 1]));
    }
    const int wk = __reduce_max_sync(0xffffffffu, bk);
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = PR - 1; j >= 0; --j) bi = key[j] == wk ? j * RT + tid : bi;  // lowest owned index holding the maximum
    const int wi = __reduce_min_sync(0xffffffffu, bi);
    if (RW == 1) {
      last = wi;
      if (wk == 0) {  // coincident points only: first index that is not taken yet (np.delete semantics)
        int fi = 0x7fffffff;
#pragma unroll
        for (int j = PR - 1; j >= 0; --j)
          if (key[j] == 0 && !sTaken[j * RT + tid]) fi = j * RT + tid;
        last = __reduce_min_sync(0xffffffffu, fi);
      }
    } else {
      const int buf = r & 1;
      if (lane == 0) s_kv[buf][warp] = make_int2(wk, wi);
      asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
      const int2 c = s_kv[buf][lane & (RW - 1)];
      const int gk = __reduce_max_sync(0xffffffffu, c.x);
      last = __reduce_min_sync(0xffffffffu, c.x == gk ? c.y : 0x7fffffff);
      if (gk == 0) {  // block-uniform: coincident points only
        int fi = 0x7fffffff;
#pragma unroll
        for (int j = PR - 1; j >= 0; --j)
          if (key[j] == 0 && !sTaken[j * RT + tid]) fi = j * RT + tid;
        fi = __reduce_min_sync(0xffffffffu, fi);
        asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");  // everybody has read slot `buf`
        if (lane == 0) s_kv[buf][warp].y = fi;
        asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
        last = __reduce_min_sync(0xffffffffu, s_kv[buf][lane & (RW - 1)].y);
      }
    }
    if (tid == 0) sTaken[last] = 1;  // read by the slow path (two barriers later at the earliest) and by the emission
  }
}

// One image of the FPS launch, run by FPS_THREADS threads of a CTA (the whole CTA of fps_kernel; the first 256 threads of
// a correlation-kernel CTA when the NEXT step's sampling rides along there, see corr_pipe.cu).  `img` = image index
// (0 .. 2B-1: depth_a then depth_b); `fps_smem` = the CTA's dynamic shared memory (fps_smem_bytes()).  CTA-wide
// synchronisation is a named barrier over exactly these threads.
__device__ __forceinline__ void fps_sync() { asm volatile("bar.sync 2, %0;" ::"n"(FPS_THREADS) : "memory"); }
__device__ __forceinline__ void fps_stamp(const FpsArgs& a, int slot) {
  if (a.clk && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.clk[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

// FAST8 (template parameter; host-checked: Hd == 8 H, Wd == 8 W, 16-byte aligned rows): the adaptive-average-pooling
// windows are the aligned 8 x 8 blocks, and a WARP pools one row of the feature grid at a time straight from global
// memory - lane l loads the two float4 of pixel l in each of the 8 image rows (a row of the image = 28 lanes x 32
// contiguous bytes: fully coalesced), all 16 loads of a lane and all rows of a warp in flight at once, then sums its 64
// values in the reference's row-major order.  This replaced "stage the whole image in shared memory with cp.async,
// then one thread per window": 6 us of staging + 4-10 us of pooling out of bank-conflicted shared memory per launch
// (measured with %globaltimer stamps), and 200 KB of shared memory per CTA.
template <int PPT, int RT, int PR, bool FAST8>
__device__ __forceinline__ void fps_cta(const FpsArgs& a, int img, float* fps_smem) {
  const float* __restrict__ depth_a = a.depth_a;
  const float* __restrict__ depth_b = a.depth_b;
  const int B = a.B, Hd = a.Hd, Wd = a.Wd, H = a.H, W = a.W, nsel = a.nsel, affine = a.affine;
  const float factor = a.factor, far_plane = a.far_plane;
  float* __restrict__ coords = a.coords;
  int32_t* __restrict__ idx_out = a.idx;
  float* __restrict__ dsign = a.dsign;
  const int sign_S = a.sign_S, sign_pitch = a.sign_pitch;
  const float sign_eps = a.sign_eps;
  const int npts = H * W;
  const int npad = (npts + 3) & ~3;
  int2 (*s_kv)[FPS_WARPS] = reinterpret_cast<int2 (*)[FPS_WARPS]>(fps_smem);       // [2][FPS_WARPS]
  int* s_scan = reinterpret_cast<int*>(fps_smem) + 4 * FPS_WARPS;                   // [FPS_WARPS]
  float* sX = fps_smem + FPS_SCRATCH_FLOATS;
  float* sY = sX + npad;
  float* sZ = sY + npad;
  unsigned char* sTaken = reinterpret_cast<unsigned char*>(sZ + npad);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gdepth = (img < B ? depth_a + (size_t)img * Hd * Wd : depth_b + (size_t)(img - B) * Hd * Wd);
  const float* depth = gdepth;
  const float halfH = (float)H / 2.0f, halfW = (float)W / 2.0f;
  // pooled value -> lifted point (depth2points, one rounding per operation) -> shared memory
  auto put_point = [&](int py, int px, float acc, int kh, int kw) {
    const int i = py * W + px;
    const float pooled = __fdiv_rn(__fdiv_rn(acc, (float)kh), (float)kw);  // ATen: sum / kh / kw
    const float fd = __fmul_rn(factor, pooled);
    sX[i] = __fdiv_rn(__fmul_rn(fd, __fsub_rn((float)px, halfW)), (float)W);
    sY[i] = __fdiv_rn(__fmul_rn(fd, __fsub_rn((float)py, halfH)), (float)H);
    sZ[i] = __fmul_rn(-pooled, far_plane);
    sTaken[i] = 0;
  };
  if (FAST8) {
    // The image comes from HBM exactly once and a warp's rows are consumed one after the other (16 loads, then a chain
    // of 64 additions): without help every row iteration pays a full DRAM round trip (4 x ~2 us per warp, measured).
    // One bulk L2 prefetch per warp for its share of the image starts all of it moving at once; the loads below then
    // find their lines in L2 or already on their way.
    if (lane == 0) {
      const size_t total = (size_t)Hd * Wd * sizeof(float);                  // a multiple of 16 (host-checked)
      const size_t chunk = ((total / FPS_WARPS) + 15) & ~(size_t)15;
      const size_t off = (size_t)warp * chunk;
      if (off < total) {
        const unsigned bytes = (unsigned)(total - off < chunk ? total - off : chunk);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(gdepth) + off), "r"(bytes) : "memory");
      }
    }
    for (int py = warp; py < H; py += FPS_WARPS) {
      for (int px = lane; px < W; px += 32) {
        float4 v[16];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4* row = reinterpret_cast<const float4*>(gdepth + (size_t)(8 * py + r) * Wd + 8 * px);
          v[2 * r] = __ldg(row);
          v[2 * r + 1] = __ldg(row + 1);
        }
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          acc = __fadd_rn(acc, v[r].x); acc = __fadd_rn(acc, v[r].y);
          acc = __fadd_rn(acc, v[r].z); acc = __fadd_rn(acc, v[r].w);
        }
        put_point(py, px, acc, 8, 8);
      }
    }
  } else {
    for (int i = tid; i < npts; i += FPS_THREADS) {
      const int py = i / W, px = i - py * W;
      const int ys = (py * Hd) / H, ye = ((py + 1) * Hd + H - 1) / H;
      const int xs = (px * Wd) / W, xe = ((px + 1) * Wd + W - 1) / W;
      float acc = 0.f;
      for (int y = ys; y < ye; ++y)
        for (int x = xs; x < xe; ++x) acc = __fadd_rn(acc, __ldg(depth + (size_t)y * Wd + x));
      put_point(py, px, acc, ye - ys, xe - xs);
    }
  }
  fps_sync();
  fps_stamp(a, 3);
  if (tid == 0) sTaken[0] = 1;  // point 0 is the first pick

  if (tid < RT) fps_rounds<PR, RT>(sX, sY, sZ, sTaken, npts, nsel, s_kv);
  fps_sync();
  fps_stamp(a, 4);

  // Raster-order emission: each thread scans a contiguous chunk of point indices.
  const int chunk = (npts + FPS_THREADS - 1) / FPS_THREADS;
  const int beg = tid * chunk, end = min(beg + chunk, npts);
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += sTaken[i];
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_scan[warp] = incl;
  fps_sync();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_scan[w];
  int rank = base + incl - cnt;
  float* cimg = coords + (size_t)img * nsel * 2;
  for (int i = beg; i < end; ++i) {
    if (sTaken[i]) {
      const int py = i / W, px = i - py * W;
      float cy = __fdiv_rn((float)py, (float)H), cx = __fdiv_rn((float)px, (float)W);
      if (affine) {
        cy = __fsub_rn(__fmul_rn(cy, 2.0f), 1.0f);
        cx = __fsub_rn(__fmul_rn(cx, 2.0f), 1.0f);
      }
      cimg[2 * rank + 0] = cy;
      cimg[2 * rank + 1] = cx;
      if (idx_out) idx_out[(size_t)img * nsel + rank] = i;
      ++rank;
    }
  }
  // fused depth_sign_kernel for the first depth tensor (the depth term of the loss only uses `depth`, not depth_pos)
  if (dsign != nullptr && img < B) {
    for (int p = tid; p < sign_pitch; p += FPS_THREADS)
      dsign[(size_t)img * sign_pitch + p] = depth_sign_value(depth, Hd, Wd, sign_S, p, sign_eps);
  }
}


// dynamic shared memory of one FPS CTA
__host__ __device__ inline size_t fps_smem_bytes(int npts) {
  return (size_t)FPS_SCRATCH_FLOATS * 4 + (size_t)((npts + 3) & ~3) * 3 * sizeof(float) + (size_t)((npts + 15) & ~15);
}

}  // namespace dg
