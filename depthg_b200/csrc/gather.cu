// Fused bilinear gather + L2 normalise (forward) and its scatter-add backward.
//
// Forward replaces sample() + norm() of the reference
// (/root/reference/src/modules.py:822-825, :789-790) and the orig_feats[perm]
// copies it makes for every negative (:1342-1343): one launch gathers ALL the
// coordinate sets that read the same source tensor (own coords + one set per
// negative), straight from the tensor's own strides (NCHW or channels-last),
// and writes K-major row panels [slot][b][p][ld] ready for the correlation
// kernel, plus 1/||x|| per row and the per-image mean row (pointwise centring).
//
// One CTA per (set, image); a warp owns a sample point at a time, lanes run over
// channels (128-bit loads when the channel stride is 1), so a channels-last
// source is read in whole 128-byte lines and every panel row is written once,
// coalesced.  HBM-bound: no data reuse beyond the four bilinear corners.
#include <string.h>

#include <stdlib.h>

#include "kernels.cuh"
#include "umma.cuh"

namespace dg {

__device__ __forceinline__ float umma_tf32(float x) {  // round-to-nearest tf32 (10-bit mantissa) in an fp32 container
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

constexpr int GATHER_THREADS = 512;
constexpr int GATHER_WARPS = GATHER_THREADS / 32;

__device__ __forceinline__ void split_f16(float v, __half& h, __half& l) {  // v already carries the panel scale
  h = __float2half_rn(v);
  l = __float2half_rn(v - __half2float(h));
}

// dynamic smem: row staging [GATHER_WARPS][ld] + mean accumulators [GATHER_WARPS][ld]
template <int FMT>
__global__ void __launch_bounds__(GATHER_THREADS)
    gather_norm_kernel(const __grid_constant__ SetTable sets, int B, int C, int H, int W,
                       const float* __restrict__ coords, int S, const int64_t* __restrict__ perms, float eps, int Prows,
                       int ld, GatherOut o) {
  extern __shared__ float gsm[];
  const int set = blockIdx.x / B, b = blockIdx.x - set * B;
  const int P = S * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* row = gsm + (size_t)warp * ld;
  float* macc = gsm + (size_t)(GATHER_WARPS + warp) * ld;
  const SetDesc& sd = sets.s[set];
  const float* t = sd.src;
  const int64_t sb = sd.sb, sc = sd.sc, sh = sd.sh, sw = sd.sw;
  const int slot = sd.slot;
  const int64_t src = sd.perm_row >= 0 ? perms[(size_t)sd.perm_row * B + b] : (int64_t)b;
  const float* timg = t + src * sb;
  const float* cset = coords + ((size_t)sd.coord * B + b) * P * 2;
  const size_t pbase = ((size_t)slot * B + b) * Prows;
  float* rpanel = o.rnorm + pbase;

  for (int c = lane; c < ld; c += 32) macc[c] = 0.f;
  const bool vec = (sc == 1) && ((C & 3) == 0) && ((sb & 3) == 0) && ((sh & 3) == 0) && ((sw & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(t) & 15) == 0);

  for (int p = warp; p < Prows; p += GATHER_WARPS) {
    const size_t ro = (pbase + p) * ld;
    float r = 0.f;
    if (p < P) {
      const int h = p / S, w = p - h * S;
      const float* cc = cset + 2 * (w * S + h);  // the reference's S-axis swap (coords.permute(0,2,1,3))
      const Corners k = bilinear_corners(__ldg(cc), __ldg(cc + 1), H, W);
      const float* p00 = timg + k.y0 * sh + k.x0 * sw;
      const float* p01 = k.x1_ok ? p00 + sw : p00;
      const float* p10 = k.y1_ok ? p00 + sh : p00;
      const float* p11 = p10 + (k.x1_ok ? sw : 0);
      const float w00 = k.w00, w01 = k.x1_ok ? k.w01 : 0.f, w10 = k.y1_ok ? k.w10 : 0.f,
                  w11 = (k.x1_ok && k.y1_ok) ? k.w11 : 0.f;
      float ss = 0.f;
      if (vec) {
#pragma unroll 2
        for (int c = lane * 4; c < C; c += 128) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p00 + c));
          const float4 bq = __ldg(reinterpret_cast<const float4*>(p01 + c));
          const float4 cq = __ldg(reinterpret_cast<const float4*>(p10 + c));
          const float4 d = __ldg(reinterpret_cast<const float4*>(p11 + c));
          float4 v;
          v.x = a.x * w00 + bq.x * w01 + cq.x * w10 + d.x * w11;
          v.y = a.y * w00 + bq.y * w01 + cq.y * w10 + d.y * w11;
          v.z = a.z * w00 + bq.z * w01 + cq.z * w10 + d.z * w11;
          v.w = a.w * w00 + bq.w * w01 + cq.w * w10 + d.w * w11;
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          *reinterpret_cast<float4*>(row + c) = v;
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const int64_t off = (int64_t)c * sc;
          const float v =
              __ldg(p00 + off) * w00 + __ldg(p01 + off) * w01 + __ldg(p10 + off) * w10 + __ldg(p11 + off) * w11;
          ss += v * v;
          row[c] = v;
        }
      }
      ss = warp_sum(ss);
      r = 1.f / fmaxf(sqrtf(ss), eps);
    }
    __syncwarp();
    // normalise, accumulate the panel mean, write the row in the requested format (zeros for padding)
    if (vec || p >= P) {
      for (int c = lane * 4; c < ld; c += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P && c < C) {
          v = *reinterpret_cast<const float4*>(row + c);
          v.x *= r; v.y *= r; v.z *= r; v.w *= r;
          float4 m = *reinterpret_cast<float4*>(macc + c);
          m.x += v.x; m.y += v.y; m.z += v.z; m.w += v.w;
          *reinterpret_cast<float4*>(macc + c) = m;
        }
        if (FMT == FMT_F32) {
          *reinterpret_cast<float4*>(o.out + ro + c) = v;
        } else if (FMT == FMT_FEATS_SPLIT) {
          __half h[4], l[4];
          constexpr float sc = F16_FEAT_SCALE;
          split_f16(v.x * sc, h[0], l[0]); split_f16(v.y * sc, h[1], l[1]);
          split_f16(v.z * sc, h[2], l[2]); split_f16(v.w * sc, h[3], l[3]);
          __half* ph = o.interleave ? o.hi16 + 2 * ro + il_col(c) : o.hi16 + ro + c;
          __half* pl = o.interleave ? ph + 32 : o.lo16 + ro + c;
          *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(h);
          *reinterpret_cast<uint2*>(pl) = *reinterpret_cast<uint2*>(l);
        } else {
          float4 hi = make_float4(umma_tf32(v.x), umma_tf32(v.y), umma_tf32(v.z), umma_tf32(v.w));
          *reinterpret_cast<float4*>(o.out + ro + c) = hi;
          *reinterpret_cast<float4*>(o.out_lo + ro + c) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
          __half h[4], l[4];
          constexpr float sc = F16_CODE_SCALE;
          split_f16(v.x * sc, h[0], l[0]); split_f16(v.y * sc, h[1], l[1]);
          split_f16(v.z * sc, h[2], l[2]); split_f16(v.w * sc, h[3], l[3]);
          __half* ph = o.interleave ? o.hi16 + 2 * ro + il_col(c) : o.hi16 + ro + c;
          __half* pl = o.interleave ? ph + 32 : o.lo16 + ro + c;
          *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(h);
          *reinterpret_cast<uint2*>(pl) = *reinterpret_cast<uint2*>(l);
        }
      }
    } else {
      for (int c = lane; c < ld; c += 32) {
        float v = 0.f;
        if (c < C) {
          v = row[c] * r;
          macc[c] += v;
        }
        if (FMT == FMT_F32) {
          o.out[ro + c] = v;
        } else if (FMT == FMT_FEATS_SPLIT) {
          __half h, l;
          split_f16(v * F16_FEAT_SCALE, h, l);
          if (o.interleave) {
            o.hi16[2 * ro + il_col(c)] = h;
            o.hi16[2 * ro + il_col(c) + 32] = l;
          } else {
            o.hi16[ro + c] = h;
            o.lo16[ro + c] = l;
          }
        } else {
          const float hi = umma_tf32(v);
          o.out[ro + c] = hi;
          o.out_lo[ro + c] = v - hi;
          __half h, l;
          split_f16(v * F16_CODE_SCALE, h, l);
          if (o.interleave) {
            o.hi16[2 * ro + il_col(c)] = h;
            o.hi16[2 * ro + il_col(c) + 32] = l;
          } else {
            o.hi16[ro + c] = h;
            o.lo16[ro + c] = l;
          }
        }
      }
    }
    if (lane == 0) rpanel[p] = r;
    __syncwarp();
  }
  if (o.meanvec == nullptr) return;
  __syncthreads();
  if (o.meanvec != nullptr) {
    float* mv = o.meanvec + ((size_t)slot * B + b) * ld;
    const float invP = 1.f / (float)P;
    for (int c = threadIdx.x; c < ld; c += GATHER_THREADS) {
      float s = 0.f;
#pragma unroll
      for (int wdx = 0; wdx < GATHER_WARPS; ++wdx) s += gsm[(size_t)(GATHER_WARPS + wdx) * ld + c];
      mv[c] = s * invP;
    }
  }
}

// Fast path for channels-last sources with C a multiple of 128 (the backbone features on the live
// path): the sampled row, its running panel mean and the bilinear corner data all stay in
// registers (no smem staging), every lane issues its 4 x NV 128-bit corner loads back to back, and a
// (set, image) is split over `nsplit` CTAs so ~900 CTAs cover the GPU evenly.  Each CTA writes its
// share of the panel mean (already divided by P); consumers add the nsplit partials.
#ifndef GF_MINBLOCKS
#define GF_MINBLOCKS 2
#endif
constexpr int GF_THREADS = 256;
constexpr int GF_WARPS = GF_THREADS / 32;

template <int NV, int FMT>
__device__ __forceinline__ void gather_feats_body(const SetTable& sets, int B, int C, int H, int W,
                                                  const float* __restrict__ coords, int S,
                                                  const int64_t* __restrict__ perms, float eps, int Prows, int nsplit,
                                                  const GatherOut& o, int vblock, float* gfs) {
  const int split = vblock % nsplit, sbi = vblock / nsplit;
  const int set = sbi / B, b = sbi - set * B;
  const int P = S * S, ld = C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SetDesc& sd = sets.s[set];
  const int64_t sh = sd.sh, sw = sd.sw;
  const int slot = sd.slot;
  const int64_t src = sd.perm_row >= 0 ? perms[(size_t)sd.perm_row * B + b] : (int64_t)b;
  const float* timg = sd.src + src * sd.sb;
  const float* cset = coords + ((size_t)sd.coord * B + b) * P * 2;
  const size_t pbase = ((size_t)slot * B + b) * Prows;
  const int chunk = (Prows + nsplit - 1) / nsplit;
  const int p_end = min((split + 1) * chunk, Prows);

  float4 macc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) macc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int p = split * chunk + warp; p < p_end; p += GF_WARPS) {
    const size_t ro = (pbase + p) * ld;
    float4 v[NV];
    float r = 0.f;
    if (p < P) {
      const int h = p / S, w = p - h * S;
      const float* cc = cset + 2 * (w * S + h);  // the reference's S-axis swap
      const Corners k = bilinear_corners(__ldg(cc), __ldg(cc + 1), H, W);
      const float* p00 = timg + k.y0 * sh + k.x0 * sw + lane * 4;
      const float* p01 = k.x1_ok ? p00 + sw : p00;
      const float* p10 = k.y1_ok ? p00 + sh : p00;
      const float* p11 = p10 + (k.x1_ok ? sw : 0);
      const float w00 = k.w00, w01 = k.x1_ok ? k.w01 : 0.f, w10 = k.y1_ok ? k.w10 : 0.f,
                  w11 = (k.x1_ok && k.y1_ok) ? k.w11 : 0.f;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p00 + 128 * i));
        const float4 bq = __ldg(reinterpret_cast<const float4*>(p01 + 128 * i));
        const float4 cq = __ldg(reinterpret_cast<const float4*>(p10 + 128 * i));
        const float4 d = __ldg(reinterpret_cast<const float4*>(p11 + 128 * i));
        v[i].x = a.x * w00 + bq.x * w01 + cq.x * w10 + d.x * w11;
        v[i].y = a.y * w00 + bq.y * w01 + cq.y * w10 + d.y * w11;
        v[i].z = a.z * w00 + bq.z * w01 + cq.z * w10 + d.z * w11;
        v[i].w = a.w * w00 + bq.w * w01 + cq.w * w10 + d.w * w11;
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
      ss = warp_sum(ss);
      r = 1.f / fmaxf(sqrtf(ss), eps);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P) {
        x = make_float4(v[i].x * r, v[i].y * r, v[i].z * r, v[i].w * r);
        macc[i].x += x.x; macc[i].y += x.y; macc[i].z += x.z; macc[i].w += x.w;
      }
      const int c = lane * 4 + 128 * i;
      if (FMT == FMT_F32) {
        *reinterpret_cast<float4*>(o.out + ro + c) = x;
      } else {
        __half hh[4], ll[4];
        constexpr float sc = F16_FEAT_SCALE;
        split_f16(x.x * sc, hh[0], ll[0]); split_f16(x.y * sc, hh[1], ll[1]);
        split_f16(x.z * sc, hh[2], ll[2]); split_f16(x.w * sc, hh[3], ll[3]);
        if (o.interleave) {
          // lanes 2m / 2m+1 hold channels 8m .. 8m+7: the even lane collects both hi quads, the odd lane both lo quads,
          // so every lane writes 16 contiguous bytes and one store instruction covers four whole 128-byte lines
          const uint2 mh = *reinterpret_cast<uint2*>(hh), ml = *reinterpret_cast<uint2*>(ll);
          const bool odd = lane & 1;
          const uint2 give = odd ? mh : ml;       // what the partner needs from me
          uint2 got;
          got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
          got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
          const uint4 outv = odd ? make_uint4(got.x, got.y, ml.x, ml.y) : make_uint4(mh.x, mh.y, got.x, got.y);
          const int c8 = c & ~7;                  // first channel of the lane pair's 8
          *reinterpret_cast<uint4*>(o.hi16 + 2 * ro + il_col(c8) + (odd ? 32 : 0)) = outv;
        } else {
          *reinterpret_cast<uint2*>(o.hi16 + ro + c) = *reinterpret_cast<uint2*>(hh);
          *reinterpret_cast<uint2*>(o.lo16 + ro + c) = *reinterpret_cast<uint2*>(ll);
        }
      }
    }
    if (lane == 0) o.rnorm[pbase + p] = r;
  }
  if (o.meanvec == nullptr) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(gfs + (size_t)warp * C + lane * 4 + 128 * i) = macc[i];
  __syncthreads();
  float* mv = o.meanvec + (((size_t)slot * B + b) * nsplit + split) * ld;
  const float invP = 1.f / (float)P;
  for (int c = threadIdx.x; c < C; c += GF_THREADS) {
    float sum = 0.f;
#pragma unroll
    for (int wdx = 0; wdx < GF_WARPS; ++wdx) sum += gfs[(size_t)wdx * C + c];
    mv[c] = sum * invP;
  }
}

template <int NV, int FMT>
__global__ void __launch_bounds__(GF_THREADS, GF_MINBLOCKS)
    gather_feats_kernel(const __grid_constant__ SetTable sets, int B, int C, int H, int W,
                        const float* __restrict__ coords, int S, const int64_t* __restrict__ perms, float eps,
                        int Prows, int nsplit, GatherOut o) {
  extern __shared__ float gfs[];  // [GF_WARPS][C] for the final cross-warp mean
  gather_feats_body<NV, FMT>(sets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, blockIdx.x, gfs);
}

// ---------------------------------------------------------------------------------------------------------------
// The same gather as a PERSISTENT kernel of self-feeding warps (channels-last sources, C a multiple of 128, split fp16
// panels).  The register-resident kernel above keeps a point's four corner rows in flight in registers (128 registers
// per thread, 16 warps per SM; ncu: waiting on memory with 24 % of the warp slots active at 45 % of the DRAM peak), and
// every warp pays the address chain of a point - permutation row, coordinates, corners - in front of its loads.
// Here each of the 8 warps of a CTA owns two 4 x C-float slots of shared memory and runs its own double-buffered
// pipeline: it turns its next-but-one point into four 1-D bulk copies (cp.async.bulk, one per corner row, completion
// on the slot's mbarrier) and meanwhile blends / normalises / splits / stores the point whose bytes have landed.  No
// warp ever waits for another one (the first version of this kernel had a producer warp and a consumer barrier per 8
// points: the barriers and the producer's serial address chain cost more than the loads).  The address work is done
// for 32 points at a time, one per lane, half a batch ahead of its first use.
//
// A CTA owns a contiguous range of panel rows, dealt round-robin to its warps.  Panel means: a warp sums the rows it
// normalised per (set, image) panel and writes its own partial - 16 partials per panel: 8 warps x (the CTA that holds
// the panel's first row | the CTA that holds the rest); a range is never shorter than a panel, so two CTAs at most
// share one.  Fixed order everywhere: the means are deterministic.
constexpr int GB_WARPS = 8;
constexpr int GB_THREADS = 32 * GB_WARPS;
constexpr int GB_PARTS = 2 * GB_WARPS;      // mean partials per (set, image) panel

struct GatherBulkArgs {
  SetTable sets;
  GatherOut o;
  const float* coords;
  const int64_t* perms;
  int B, C, H, W, S, Prows;
  long long total_rows;                     // nsets * B * Prows
  int dbg;                                  // timing experiments (DEPTHG_B200_GATHER_DBG) bits: 1 no panel stores, 2 no loads
  int l2_hints;                             // evict-first source loads / evict-last panel stores (DEPTHG_B200_L2HINTS=1; measured: slower, off by default)
  float eps;
};

template <int NV>
__global__ void __launch_bounds__(GB_THREADS, 1) gather_bulk_kernel(const __grid_constant__ GatherBulkArgs a) {
  using namespace umma;
  extern __shared__ uint8_t gb_raw[];
  constexpr int C = 128 * NV;
  uint8_t* base = gb_raw + ((128u - (smem_u32(gb_raw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(base) + (size_t)warp * 2 * 4 * C;        // this warp's two slots [2][4][C]
  float* meta = reinterpret_cast<float*>(base) + (size_t)GB_WARPS * 2 * 4 * C + warp * 8;   // [2][4] corner weights
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(base) + (size_t)GB_WARPS * 2 * 4 * C +
                                               GB_WARPS * 8) + warp * 2;                    // [2]
  const int P = a.S * a.S, Prows = a.Prows;
  if (lane == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncwarp();
  pdl_trigger();
  pdl_wait();   // coordinates / permutations come from the kernels before; nothing global is touched above

  // this CTA's rows [r0, r1): 16-row units, contiguous, at least one whole panel unless there are fewer panels than CTAs
  const long long units = a.total_rows / 16;
  const long long r0 = (units * blockIdx.x / gridDim.x) * 16, r1 = (units * (blockIdx.x + 1) / gridDim.x) * 16;
  const int nw = (int)((r1 - r0 - warp + GB_WARPS - 1) / GB_WARPS);   // rows of this warp: r0 + warp + 8 n
  if (nw <= 0) return;

  // address work for 32 of the warp's rows at a time (lane i <-> row n0 + i): raw global loads first, the arithmetic later
  struct Raw { int64_t perm; float cx, cy; int set, b, p; bool in_range, valid; };
  struct Desc { const float *p00, *p01, *p10, *p11; float w0, w1, w2, w3; };
  auto fetch = [&](int n0) {
    Raw q;
    const int n = n0 + lane;
    q.in_range = n < nw;
    q.valid = false;
    q.perm = 0; q.cx = q.cy = 0.f; q.set = q.b = q.p = 0;
    if (!q.in_range) return q;
    const unsigned gr = (unsigned)(r0 + warp) + (unsigned)GB_WARPS * (unsigned)n;   // total_rows < 2^31 (checked at launch)
    const unsigned g = gr / (unsigned)Prows;
    q.p = (int)(gr - g * (unsigned)Prows);
    q.set = (int)(g / (unsigned)a.B);
    q.b = (int)(g - (unsigned)q.set * (unsigned)a.B);
    q.valid = q.p < P;
    const SetDesc& sd = a.sets.s[q.set];
    q.perm = sd.perm_row >= 0 ? a.perms[(size_t)sd.perm_row * a.B + q.b] : (int64_t)q.b;
    if (q.valid) {
      const int h = q.p / a.S, w = q.p - h * a.S;
      const float* cc = a.coords + (((size_t)sd.coord * a.B + q.b) * P + (w * a.S + h)) * 2;   // the reference's S-axis swap
      q.cx = __ldg(cc);
      q.cy = __ldg(cc + 1);
    }
    return q;
  };
  auto resolve = [&](const Raw& q) {
    Desc d;
    const SetDesc& sd = a.sets.s[q.set];
    const float* timg = sd.src + q.perm * sd.sb;
    d.p00 = d.p01 = d.p10 = d.p11 = timg;   // padding rows (p >= P) read pixel 0 with zero weights: they come out as zeros
    d.w0 = d.w1 = d.w2 = d.w3 = 0.f;
    if (q.valid) {
      const Corners k = bilinear_corners(q.cx, q.cy, a.H, a.W);
      d.p00 = timg + k.y0 * sd.sh + k.x0 * sd.sw;
      d.p01 = k.x1_ok ? d.p00 + sd.sw : d.p00;
      d.p10 = k.y1_ok ? d.p00 + sd.sh : d.p00;
      d.p11 = d.p10 + (k.x1_ok ? sd.sw : 0);
      d.w0 = k.w00;
      d.w1 = k.x1_ok ? k.w01 : 0.f;
      d.w2 = k.y1_ok ? k.w10 : 0.f;
      d.w3 = (k.x1_ok && k.y1_ok) ? k.w11 : 0.f;
    }
    return d;
  };
  // lane (n % 32) issues row n's copies into slot n & 1 from the batch `d` that holds it
  auto issue = [&](int n, const Desc& d) {
    if (n < nw && lane == (n & 31)) {
      const int s = n & 1;
      float* m = meta + 4 * s;
      m[0] = d.w0; m[1] = d.w1; m[2] = d.w2; m[3] = d.w3;
      float* dst = ring + (size_t)s * 4 * C;
      if (a.dbg & 2) {
        mbar_arrive(&full[s]);
      } else {
        fence_proxy_async_smem();     // the slot's previous contents were read through the generic proxy
        mbar_arrive_expect_tx(&full[s], (uint32_t)(16 * C));
        // the sources stream through (each pixel is wanted by a handful of points at about the same time); the
        // panels this kernel writes are what the correlation kernel reads next: they should be what stays in L2
        const uint64_t pol = a.l2_hints ? L2_EVICT_FIRST : 0x1000000000000000ull;
        bulk_load_hint(dst, d.p00, (uint32_t)(4 * C), &full[s], pol);
        bulk_load_hint(dst + C, d.p01, (uint32_t)(4 * C), &full[s], pol);
        bulk_load_hint(dst + 2 * C, d.p10, (uint32_t)(4 * C), &full[s], pol);
        bulk_load_hint(dst + 3 * C, d.p11, (uint32_t)(4 * C), &full[s], pol);
      }
    }
  };

  Desc cur = resolve(fetch(0));
  Raw nxt_raw = fetch(32);
  Desc nxt = cur;
  issue(0, cur);
  issue(1, cur);

  float4 macc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) macc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float invP = 1.f / (float)P;

  int g = (int)((r0 + warp) / Prows);
  int p = (int)((r0 + warp) - (long long)g * Prows);
  int set = g / a.B, b = g - set * a.B;
  for (int n = 0; n < nw; ++n) {
    const int nb = n & 31;
    if (nb == 16) nxt = resolve(nxt_raw);                       // (its loads were issued 16 rows ago)
    // row identity (uniform over the warp), advanced incrementally at the bottom of the loop
    const int slot = a.sets.s[set].slot;
    const size_t pbase = ((size_t)slot * a.B + b) * Prows;
    const size_t ro = (pbase + p) * C;
    const int s = n & 1;
    mbar_wait(&full[s], (n >> 1) & 1);
    const float4 wq = *reinterpret_cast<const float4*>(meta + 4 * s);
    const float* src = ring + (size_t)s * 4 * C + lane * 4;
    float4 x[NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 q0 = *reinterpret_cast<const float4*>(src + 128 * i);
      const float4 q1 = *reinterpret_cast<const float4*>(src + C + 128 * i);
      const float4 q2 = *reinterpret_cast<const float4*>(src + 2 * C + 128 * i);
      const float4 q3 = *reinterpret_cast<const float4*>(src + 3 * C + 128 * i);
      // (same operation order as the register-resident kernel: the two are bit-identical)
      x[i].x = q0.x * wq.x + q1.x * wq.y + q2.x * wq.z + q3.x * wq.w;
      x[i].y = q0.y * wq.x + q1.y * wq.y + q2.y * wq.z + q3.y * wq.w;
      x[i].z = q0.z * wq.x + q1.z * wq.y + q2.z * wq.z + q3.z * wq.w;
      x[i].w = q0.w * wq.x + q1.w * wq.y + q2.w * wq.z + q3.w * wq.w;
      ss += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
    }
    __syncwarp();                      // the slot is in registers: refill it with the warp's next-but-one row
    if (nb == 30) { cur = nxt; nxt_raw = fetch(n + 2 + 32); }   // rows n + 2 .. n + 33 are the next batch
    issue(n + 2, cur);
    ss = warp_sum(ss);
    const float r = p < P ? 1.f / fmaxf(sqrtf(ss), a.eps) : 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (p >= P) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // padding row: exact zeros whatever pixel 0 holds
      x[i] = make_float4(x[i].x * r, x[i].y * r, x[i].z * r, x[i].w * r);
      macc[i].x += x[i].x; macc[i].y += x[i].y; macc[i].z += x[i].z; macc[i].w += x[i].w;
    }
    if (!(a.dbg & 1)) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane * 4 + 128 * i;
        __half hh[4], ll[4];
        constexpr float sc = F16_FEAT_SCALE;
        split_f16(x[i].x * sc, hh[0], ll[0]); split_f16(x[i].y * sc, hh[1], ll[1]);
        split_f16(x[i].z * sc, hh[2], ll[2]); split_f16(x[i].w * sc, hh[3], ll[3]);
        if (a.o.interleave) {   // see gather_feats_body: even lanes store both hi quads, odd lanes both lo quads
          const uint2 mh = *reinterpret_cast<uint2*>(hh), ml = *reinterpret_cast<uint2*>(ll);
          const bool odd = lane & 1;
          const uint2 give = odd ? mh : ml;
          uint2 got;
          got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
          got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
          const uint4 outv = odd ? make_uint4(got.x, got.y, ml.x, ml.y) : make_uint4(mh.x, mh.y, got.x, got.y);
          const int c8 = c & ~7;
          if (a.l2_hints)
            st_global_v4_hint(a.o.hi16 + 2 * ro + il_col(c8) + (odd ? 32 : 0), outv, L2_EVICT_LAST);
          else
            *reinterpret_cast<uint4*>(a.o.hi16 + 2 * ro + il_col(c8) + (odd ? 32 : 0)) = outv;
        } else {
          *reinterpret_cast<uint2*>(a.o.hi16 + ro + c) = *reinterpret_cast<uint2*>(hh);
          *reinterpret_cast<uint2*>(a.o.lo16 + ro + c) = *reinterpret_cast<uint2*>(ll);
        }
      }
    }
    if (lane == 0) a.o.rnorm[pbase + p] = r;
    // last row of this warp in panel g (its rows are 8 apart): write the warp's partial of the panel mean
    const bool last_in_panel = (n + 1 == nw) || (p + GB_WARPS >= Prows);
    if (last_in_panel && a.o.meanvec != nullptr) {
      const long long g0 = (long long)g * Prows;                 // the panel's first row
      const int part = (g0 >= r0 ? 0 : GB_WARPS) + warp;         // second half of the partials: the panel began in another CTA
      float* mv = a.o.meanvec + (((size_t)slot * a.B + b) * GB_PARTS + part) * C + lane * 4;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        *reinterpret_cast<float4*>(mv + 128 * i) =
            make_float4(macc[i].x * invP, macc[i].y * invP, macc[i].z * invP, macc[i].w * invP);
        macc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (g0 >= r0 && g0 + Prows <= r1) {                        // the whole panel is this CTA's: the other half is zero
        float* mz = mv + (size_t)GB_WARPS * C;
#pragma unroll
        for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(mz + 128 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    p += GB_WARPS;                     // the warp's next row
    if (p >= Prows) {
      p -= Prows;
      ++g;
      if (++b == a.B) { b = 0; ++set; }
    }
  }
}

// Code tensors (D <= 128, any strides): lanes over channels with scalar loads, everything in registers, a
// (set, image) split over `nsplit` CTAs.  Writes the tf32 hi / fp32 remainder panels for the cd product and the
// bf16 hi/lo panels for the gradient GEMMs, plus 1/||x||.
template <int R>
__device__ __forceinline__ void gather_code_body(const SetTable& sets, int B, int C, int H, int W,
                                                 const float* __restrict__ coords, int S,
                                                 const int64_t* __restrict__ perms, float eps, int Prows, int nsplit,
                                                 const GatherOut& o, int vblock) {
  const int split = vblock % nsplit, sbi = vblock / nsplit;
  const int set = sbi / B, b = sbi - set * B;
  const int P = S * S, ld = 32 * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SetDesc& sd = sets.s[set];
  const int64_t sc = sd.sc, sh = sd.sh, sw = sd.sw;
  const int64_t src = sd.perm_row >= 0 ? perms[(size_t)sd.perm_row * B + b] : (int64_t)b;
  const float* timg = sd.src + src * sd.sb;
  const float* cset = coords + ((size_t)sd.coord * B + b) * P * 2;
  const size_t pbase = ((size_t)sd.slot * B + b) * Prows;
  const int chunk = (Prows + nsplit - 1) / nsplit;
  const int p_end = min((split + 1) * chunk, Prows);
  for (int p = split * chunk + warp; p < p_end; p += GF_WARPS) {
    float v[R];
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = 0.f;
    if (p < P) {
      const int h = p / S, w = p - h * S;
      const float* cc = cset + 2 * (w * S + h);
      const Corners k = bilinear_corners(__ldg(cc), __ldg(cc + 1), H, W);
      const float* p00 = timg + k.y0 * sh + k.x0 * sw;
      const float* p01 = k.x1_ok ? p00 + sw : p00;
      const float* p10 = k.y1_ok ? p00 + sh : p00;
      const float* p11 = p10 + (k.x1_ok ? sw : 0);
      const float w00 = k.w00, w01 = k.x1_ok ? k.w01 : 0.f, w10 = k.y1_ok ? k.w10 : 0.f,
                  w11 = (k.x1_ok && k.y1_ok) ? k.w11 : 0.f;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          const int64_t off = (int64_t)c * sc;
          v[j] = __ldg(p00 + off) * w00 + __ldg(p01 + off) * w01 + __ldg(p10 + off) * w10 + __ldg(p11 + off) * w11;
          ss += v[j] * v[j];
        }
      }
      ss = warp_sum(ss);
      r = 1.f / fmaxf(sqrtf(ss), eps);
    }
    const size_t ro = (pbase + p) * ld;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float x = v[j] * r;
      const float hi = umma_tf32(x);
      __half h, l;
      split_f16(x * F16_CODE_SCALE, h, l);
      const int c = lane + 32 * j;
      o.out[ro + c] = hi;
      o.out_lo[ro + c] = x - hi;
      if (o.interleave) {   // c = lane + 32 j: chunk j, hi at 64 j + lane, lo 32 further
        o.hi16[2 * ro + 64 * j + lane] = h;
        o.hi16[2 * ro + 64 * j + 32 + lane] = l;
      } else {
        o.hi16[ro + c] = h;
        o.lo16[ro + c] = l;
      }
    }
    if (lane == 0) o.rnorm[pbase + p] = r;
  }
}

template <int R>
__global__ void __launch_bounds__(GF_THREADS)
    gather_code_kernel(const __grid_constant__ SetTable sets, int B, int C, int H, int W,
                       const float* __restrict__ coords, int S, const int64_t* __restrict__ perms, float eps,
                       int Prows, int nsplit, GatherOut o, const __grid_constant__ DotsJob job, int ncode_blocks) {
  pdl_trigger();
  pdl_wait();
  // the pair_dots CTAs (independent of the code gather, same stream slot) come FIRST: as the last blocks of the grid
  // they were a serial tail of ~9 us (each is a latency chain over 2 x 16 partial means)
  const int ndots = (int)gridDim.x - ncode_blocks;
  if ((int)blockIdx.x < ndots) {
    pair_dots_body(job, (int)blockIdx.x);
    return;
  }
  // the zero fill of the caller's gradient buffers (dg_loss_io_t::clear) rides here: a few 16-byte stores per thread
  // in a latency-bound kernel instead of two fill launches between the forward and the backward
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (job.clr[q] == nullptr) continue;
    const unsigned long long step = (unsigned long long)ncode_blocks * GF_THREADS;
    for (unsigned long long i = (unsigned long long)((int)blockIdx.x - ndots) * GF_THREADS + threadIdx.x; i < job.clr_n16[q];
         i += step)
      job.clr[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  gather_code_body<R>(sets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, (int)blockIdx.x - ndots);
}

// Both gathers of a forward pass in ONE launch: the first `nfeat_blocks` CTAs gather the backbone features (bf16
// hi/lo panels), the remaining ones the code tensors.  Each kind alone is latency-bound; together they fill the GPU.
struct GatherAllArgs {
  SetTable fsets, csets;
  GatherOut fo, co;
  int B, C, D, H, W, S, Prows, fsplit, csplit, nfeat_blocks;
  float eps;
  const float* coords;
  const int64_t* perms;
};

template <int NV, int R>
__global__ void __launch_bounds__(GF_THREADS, 2) gather_all_kernel(const __grid_constant__ GatherAllArgs a) {
  extern __shared__ float gfs[];
  // code CTAs outnumber feature CTAs 2:1 (csplit = 2 * fsplit): interleave them f,c,c,f,c,c,... so both kinds are
  // resident together from the first wave on
  const int q = blockIdx.x / 3, r = blockIdx.x - 3 * q;
  if (r == 0)
    gather_feats_body<NV, FMT_FEATS_SPLIT>(a.fsets, a.B, a.C, a.H, a.W, a.coords, a.S, a.perms, a.eps, a.Prows,
                                           a.fsplit, a.fo, q, gfs);
  else
    gather_code_body<R>(a.csets, a.B, a.D, a.H, a.W, a.coords, a.S, a.perms, a.eps, a.Prows, a.csplit, a.co,
                        2 * q + (r - 1));
}

// NCHW -> channels-last staging copy for the feature tensors: [B][C][HW] -> [B][HW][C] through a 32x33 smem tile
// (both reads and writes coalesced).  NCHW sources would otherwise make every gathered value its own 32-byte sector.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ a, const float* __restrict__ b2,
                                                           int B, int C, int HW, int64_t sb_a, int64_t sb_b,
                                                           float* __restrict__ out_a, float* __restrict__ out_b) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const float* src = img < B ? a + (int64_t)img * sb_a : b2 + (int64_t)(img - B) * sb_b;
  float* dst = (img < B ? out_a + (size_t)img * C * HW : out_b + (size_t)(img - B) * C * HW);
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, p = p0 + tx;
    tile[ty + 8 * r][tx] = (c < C && p < HW) ? __ldg(src + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r, c = c0 + tx;
    if (p < HW && c < C) dst[(size_t)p * C + c] = tile[tx][ty + 8 * r];
  }
}

int launch_nchw_to_nhwc(const float* a, const float* b, int B, int C, int HW, int64_t sb_a, int64_t sb_b, float* out_a,
                        float* out_b, cudaStream_t st) {
  DG_PRE(st);
  nchw_to_nhwc_kernel<<<dim3(ceil_div(HW, 32), ceil_div(C, 32), b ? 2 * B : B), 256, 0, st>>>(a, b, B, C, HW, sb_a, sb_b,
                                                                                               out_a, out_b);
  DG_LAUNCH_OK("nchw_to_nhwc_kernel");
  return DG_OK;
}

// One warp per panel row: compose the row's gradient from the unit gradients,
// back through x/max(||x||,eps), then atomically scatter through the 4 corners.
template <int RT>
__global__ void __launch_bounds__(256)
    gather_norm_bwd_kernel(const __grid_constant__ SetTable sets, int B, int C, int H, int W,
                           const float* __restrict__ coords, int S, const int64_t* __restrict__ perms, float eps,
                           int Prows, int ld, const float* __restrict__ cn, const float* __restrict__ cn_lo,
                           const float* __restrict__ rnorm, const float* __restrict__ dC1,
                           const float* __restrict__ dC2, int npairs, const __grid_constant__ PairTable pairs,
                           int has_depth, const __grid_constant__ GroupW gws, int nsets, int ni, int nj, int njw) {
  pdl_trigger();
  pdl_wait();
  const int P = S * S;
  // ni / nj: pitch (in panels) of the dC2 per-row-tile and dC1 per-column-tile (njw = 128) or per-column-group
  // (njw = 256) partial buffers; only the tiles / groups that contain real points were written
  const int ni_used = min(ni, (P + 127) / 128), nj_used = min(nj, (P + njw - 1) / njw);
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp_global >= nsets * B * P) return;
  const int set = warp_global / (B * P);
  const int rem = warp_global - set * B * P;
  const int b = rem / P, p = rem - b * P;
  const SetDesc& sd = sets.s[set];
  const int slot = sd.slot;
  const size_t panel = (size_t)B * Prows * ld;
  const size_t rowoff = ((size_t)b * Prows + p) * ld;
  // ld is a multiple of 32; RT = ld / 32 chunks handled in registers (RT = 0: runtime count, up to 8)
  constexpr int RJ = RT > 0 ? RT : 8;
  const int R = RT > 0 ? RT : ld / 32;
  float gw[DG_NUM_GROUPS];
#pragma unroll
  for (int g = 0; g < DG_NUM_GROUPS; ++g)
    gw[g] = gws.arr ? __ldg(gws.arr + g) : (gws.ptr[g] ? __ldg(gws.ptr[g]) : 0.f);
  float g[RJ];
#pragma unroll
  for (int j = 0; j < RJ; ++j) g[j] = 0.f;
  if (slot == 0) {
    for (int k = 0; k < npairs; ++k) {
      const float wk = gw[pairs.group[k]] * pairs.scale[k];
      for (int t = 0; t < nj_used; ++t) {   // dC1 comes as `nj` partial buffers (one per column group; 1 unless dense)
        const float* a = dC1 + ((size_t)k * nj + t) * panel + rowoff;
#pragma unroll
        for (int j = 0; j < RJ; ++j)
          if (j < R) g[j] += wk * __ldg(a + lane + 32 * j);
      }
    }
    for (int t = 0; t < ni_used; ++t) {  // dC2 comes as `ni` partial buffers (one per 128-row tile of the first operand)
      const float w0 = gw[pairs.group[0]] * pairs.scale[0];
      const float* a = dC2 + (size_t)t * panel + rowoff;
#pragma unroll
      for (int j = 0; j < RJ; ++j)
        if (j < R) g[j] += w0 * __ldg(a + lane + 32 * j);
    }
    if (has_depth) {
      const float wd = gw[DG_GROUP_DEPTH];
      for (int t = 0; t < nj_used; ++t) {
        const float* a1 = dC1 + ((size_t)npairs * nj + t) * panel + rowoff;
#pragma unroll
        for (int j = 0; j < RJ; ++j)
          if (j < R) g[j] += wd * __ldg(a1 + lane + 32 * j);
      }
      for (int t = 0; t < ni_used; ++t) {
        const float* a2 = dC2 + ((size_t)npairs * ni + t) * panel + rowoff;
#pragma unroll
        for (int j = 0; j < RJ; ++j)
          if (j < R) g[j] += wd * __ldg(a2 + lane + 32 * j);
      }
    }
  } else {
    const float ws = gw[pairs.group[slot]] * pairs.scale[slot];
    for (int t = 0; t < ni_used; ++t) {
      const float* a = dC2 + ((size_t)slot * ni + t) * panel + rowoff;
#pragma unroll
      for (int j = 0; j < RJ; ++j)
        if (j < R) g[j] += ws * __ldg(a + lane + 32 * j);
    }
  }
  const float* xh = cn + (size_t)slot * panel + rowoff;
  const float* xl = cn_lo ? cn_lo + (size_t)slot * panel + rowoff : nullptr;  // split panels: x = hi + lo exactly
  const float r = __ldg(rnorm + ((size_t)slot * B + b) * Prows + p);
  float x[RJ];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < RJ; ++j) {
    x[j] = (j < R) ? __ldg(xh + lane + 32 * j) : 0.f;
    if (xl && j < R) x[j] += __ldg(xl + lane + 32 * j);
    dot += g[j] * x[j];
  }
  dot = warp_sum(dot);
  const bool clamped = r * eps >= 0.9999f;  // ||x|| <= eps: the denominator was the constant eps
  if (clamped) dot = 0.f;

  const int h = p / S, w = p - h * S;
  const float* cc = coords + (((size_t)sd.coord * B + b) * P + (w * S + h)) * 2;
  const Corners k = bilinear_corners(__ldg(cc), __ldg(cc + 1), H, W);
  const int64_t src = sd.perm_row >= 0 ? perms[(size_t)sd.perm_row * B + b] : (int64_t)b;
  const int64_t sb = sd.sb, sc = sd.sc, sh = sd.sh, sw = sd.sw;
  float* g00 = const_cast<float*>(sd.src) + src * sb + k.y0 * sh + k.x0 * sw;  // sd.src is the gradient tensor here
#pragma unroll
  for (int j = 0; j < RJ; ++j) {
    const int c = lane + 32 * j;
    if (j < R && c < C) {
      const float dx = (g[j] - x[j] * dot) * r;
      float* q = g00 + (int64_t)c * sc;
      atomicAdd(q, dx * k.w00);
      if (k.x1_ok) atomicAdd(q + sw, dx * k.w01);
      if (k.y1_ok) atomicAdd(q + sh, dx * k.w10);
      if (k.x1_ok && k.y1_ok) atomicAdd(q + sh + sw, dx * k.w11);
    }
  }
}

// get_feats pooling: mean over H,W then L2 normalise (src/precompute_knns.py:19).
__global__ void __launch_bounds__(256) pool_normalize_kernel(const float* __restrict__ t, int64_t sb, int64_t sc,
                                                             int64_t sh, int64_t sw, int C, int H, int W, float eps,
                                                             float* __restrict__ out) {
  extern __shared__ float psm[];  // [C]
  __shared__ float red[8];
  const float* timg = t + (int64_t)blockIdx.x * sb;
  const int HW = H * W;
  const float inv = 1.f / (float)HW;
  float ss = 0.f;
  if (sc == 1) {  // channels-last: threads over channels, loop over pixels (coalesced)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) s += __ldg(timg + y * sh + x * sw + c);
      s *= inv;
      psm[c] = s;
      ss += s * s;
    }
  } else {  // NCHW: a warp per channel plane, lanes over pixels
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < C; c += nw) {
      float s = 0.f;
      for (int i = lane; i < HW; i += 32) {
        const int y = i / W, x = i - y * W;
        s += __ldg(timg + (int64_t)c * sc + y * sh + x * sw);
      }
      s = warp_sum(s) * inv;
      if (lane == 0) {
        psm[c] = s;
        ss += s * s;
      }
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float r = 1.f / fmaxf(sqrtf(tot), eps);
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)blockIdx.x * C + c] = psm[c] * r;
}

// norm() of the reference (src/modules.py:789-790): x / max(||x||_2 over dim 1, eps) on a tensor viewed as
// [N, C, inner] with element strides (sN, sC, sP), same strides for the output.
//   sC == 1 (channels-last): a warp per position, lanes over channels;  otherwise: a thread per position (coalesced
//   along the position axis), two passes over the channels.
__global__ void __launch_bounds__(256) norm_dim1_kernel(const float* __restrict__ t, int N, int C, long long inner,
                                                        long long sN, long long sC, long long sP, float eps,
                                                        float* __restrict__ out) {
  const long long total = (long long)N * inner;
  if (sC == 1) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= total) return;
    const long long n = w / inner, p = w - n * inner;
    const float* x = t + n * sN + p * sP;
    float* y = out + n * sN + p * sP;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = __ldg(x + c);
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float d = fmaxf(sqrtf(ss), eps);
    for (int c = lane; c < C; c += 32) y[c] = __ldg(x + c) / d;
  } else {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / inner, p = i - n * inner;
    const float* x = t + n * sN + p * sP;
    float* y = out + n * sN + p * sP;
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(x + (long long)c * sC);
      ss = fmaf(v, v, ss);
    }
    const float d = fmaxf(sqrtf(ss), eps);
    for (int c = 0; c < C; ++c) y[(long long)c * sC] = __ldg(x + (long long)c * sC) / d;
  }
}

static int fill_sets(const char* fn, const float* src, const int64_t* strides, int nsets, const int32_t* set_coord,
                     const int32_t* set_slot, bool has_perm, SetTable* tab) {
  DG_REQUIRE(nsets > 0 && nsets <= DG_MAX_SETS, DG_ERR_INVALID, "%s: nsets=%d out of range", fn, nsets);
  DG_REQUIRE(set_coord && set_slot, DG_ERR_INVALID, "%s: null set tables", fn);
  for (int s = 0; s < nsets; ++s) {
    DG_REQUIRE(set_coord[s] >= 0 && set_slot[s] >= 0, DG_ERR_INVALID, "%s: negative set entry", fn);
    SetDesc& d = tab->s[s];
    d.src = src;
    d.sb = strides[0]; d.sc = strides[1]; d.sh = strides[2]; d.sw = strides[3];
    d.coord = set_coord[s];
    d.slot = set_slot[s];
    d.perm_row = has_perm ? s : -1;
  }
  return DG_OK;
}

// DEPTHG_B200_GATHER=regs keeps the register-resident kernel (comparison / fallback)
static bool gather_bulk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DEPTHG_B200_GATHER");
    v = (e && e[0] == 'r') ? 0 : 1;
  }
  return v != 0;
}

// How many CTAs a (set, image) is split over: 8 on the register-resident fast path, else 1.
int gather_nsplit(int fmt, const SetTable& tab, int nsets, int C, int ld) {
  if (fmt == FMT_CODE_SPLIT || C != ld || (C % 128) != 0 || C / 128 > 8) return 1;
  const int nv = C / 128;
  if (nv != 1 && nv != 2 && nv != 3 && nv != 4 && nv != 6 && nv != 8) return 1;
  for (int s = 0; s < nsets; ++s) {
    const SetDesc& d = tab.s[s];
    if (d.sc != 1 || (d.sb & 3) || (d.sh & 3) || (d.sw & 3) || (reinterpret_cast<uintptr_t>(d.src) & 15)) return 1;
  }
  // split fp16 panels whose 2 x 8 point slots (16 C floats each) fit in shared memory: the bulk-copy kernel, which
  // writes GB_PARTS mean partials per panel; otherwise the register-resident kernel with 8 CTAs per panel
  if (fmt == FMT_FEATS_SPLIT && gather_bulk_enabled() && (size_t)GB_WARPS * 2 * 16 * C <= 200 * 1024) return GB_PARTS;
  return 8;
}

template <int NV>
static int launch_gather_fast(int fmt, const SetTable& tab, int nsets, int B, int C, int H, int W, const float* coords,
                              int S, const int64_t* perms, float eps, int Prows, int nsplit, const GatherOut& o,
                              cudaStream_t st) {
  if (fmt == FMT_FEATS_SPLIT && nsplit == GB_PARTS) {   // (gather_nsplit chose the self-feeding bulk-copy kernel)
    const size_t smem = (size_t)GB_WARPS * 2 * 16 * C + GB_WARPS * (32 + 16) + 128;
    static PerDevice attr_pd = {}, sm_pd = {};
    size_t& attr = per_device(attr_pd);
    size_t& sms = per_device(sm_pd);
    if (attr < smem) {
      DG_CUDA_OK(cudaFuncSetAttribute(gather_bulk_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    if (!sms) {
      int dev = 0, n = 0;
      DG_CUDA_OK(cudaGetDevice(&dev));
      DG_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
      sms = (size_t)(n > 0 ? n : 148);
    }
    GatherBulkArgs ga;
    ga.sets = tab; ga.o = o; ga.coords = coords; ga.perms = perms;
    ga.B = B; ga.C = C; ga.H = H; ga.W = W; ga.S = S; ga.Prows = Prows;
    ga.total_rows = (long long)nsets * B * Prows; ga.eps = eps;
    DG_REQUIRE(ga.total_rows < (1LL << 31), DG_ERR_UNSUPPORTED, "gather: too many panel rows");
    ga.dbg = getenv("DEPTHG_B200_GATHER_DBG") ? atoi(getenv("DEPTHG_B200_GATHER_DBG")) : 0;
    ga.l2_hints = getenv("DEPTHG_B200_L2HINTS") && getenv("DEPTHG_B200_L2HINTS")[0] == '1';
    const int npanels = nsets * B;             // a CTA's row range must hold at least one whole panel
    const int grid = npanels < (int)sms ? npanels : (int)sms;
    DG_PRE(st);
    launch_pdl(gather_bulk_kernel<NV>, dim3(grid), dim3(GB_THREADS), smem, st, ga);
    DG_LAUNCH_OK("gather_bulk_kernel");
    return DG_OK;
  }
  const size_t smem = (size_t)GF_WARPS * C * sizeof(float);
  DG_PRE(st);
  if (fmt == FMT_F32)
    gather_feats_kernel<NV, FMT_F32><<<nsets * B * nsplit, GF_THREADS, smem, st>>>(tab, B, C, H, W, coords, S, perms, eps,
                                                                                  Prows, nsplit, o);
  else
    gather_feats_kernel<NV, FMT_FEATS_SPLIT><<<nsets * B * nsplit, GF_THREADS, smem, st>>>(tab, B, C, H, W, coords, S,
                                                                                          perms, eps, Prows, nsplit, o);
  DG_LAUNCH_OK("gather_feats_kernel");
  return DG_OK;
}

int launch_gather(int fmt, const SetTable& tab, int nsets, int B, int C, int H, int W, const float* coords, int S,
                  const int64_t* perms, float eps, int Prows, int ld, int nsplit, const GatherOut& o, cudaStream_t st,
                  const DotsJob* tail) {
  DG_REQUIRE(!tail || (fmt == FMT_CODE_SPLIT && ld <= 128), DG_ERR_INVALID, "gather: a dots tail needs the code-split path");
  if (fmt == FMT_CODE_SPLIT && ld <= 128) {  // register-resident code gather
    const int ns = 16;  // one or two points per warp: the two dependent global round trips per point are the cost
    DotsJob job;
    memset(&job, 0, sizeof job);
    if (tail) job = *tail;
    const int ncode = nsets * B * ns, grid = ncode + (tail ? tail->npairs * tail->B : 0);
    DG_PRE(st);
    switch (ld / 32) {
      case 1: launch_pdl(gather_code_kernel<1>, dim3(grid), dim3(GF_THREADS), 0, st, tab, B, C, H, W, coords, S, perms, eps, Prows, ns, o, job, ncode); break;
      case 2: launch_pdl(gather_code_kernel<2>, dim3(grid), dim3(GF_THREADS), 0, st, tab, B, C, H, W, coords, S, perms, eps, Prows, ns, o, job, ncode); break;
      case 3: launch_pdl(gather_code_kernel<3>, dim3(grid), dim3(GF_THREADS), 0, st, tab, B, C, H, W, coords, S, perms, eps, Prows, ns, o, job, ncode); break;
      default: launch_pdl(gather_code_kernel<4>, dim3(grid), dim3(GF_THREADS), 0, st, tab, B, C, H, W, coords, S, perms, eps, Prows, ns, o, job, ncode); break;
    }
    DG_LAUNCH_OK("gather_code_kernel");
    return DG_OK;
  }
  if (nsplit > 1) {
    DG_REQUIRE(nsplit == gather_nsplit(fmt, tab, nsets, C, ld), DG_ERR_INVALID, "gather: inconsistent nsplit");
    switch (C / 128) {
      case 1: return launch_gather_fast<1>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
      case 2: return launch_gather_fast<2>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
      case 3: return launch_gather_fast<3>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
      case 4: return launch_gather_fast<4>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
      case 6: return launch_gather_fast<6>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
      default: return launch_gather_fast<8>(fmt, tab, nsets, B, C, H, W, coords, S, perms, eps, Prows, nsplit, o, st);
    }
  }
  size_t smem = (size_t)2 * GATHER_WARPS * ld * sizeof(float);
  DG_REQUIRE(smem <= 200 * 1024, DG_ERR_UNSUPPORTED, "gather: C=%d too large for the row staging buffer", C);
#define DG_GATHER_LAUNCH(F)                                                                                        \
  do {                                                                                                             \
    static PerDevice configured_pd = {};                                                                           \
    size_t& configured = per_device(configured_pd);                                                                \
    if (smem > configured) {                                                                                       \
      DG_CUDA_OK(cudaFuncSetAttribute(gather_norm_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      configured = smem;                                                                                           \
    }                                                                                                              \
    DG_PRE(st);                                                                                                    \
    gather_norm_kernel<F><<<nsets * B, GATHER_THREADS, smem, st>>>(tab, B, C, H, W, coords, S, perms, eps, Prows, ld, o); \
  } while (0)
  if (fmt == FMT_F32) DG_GATHER_LAUNCH(FMT_F32);
  else if (fmt == FMT_FEATS_SPLIT) DG_GATHER_LAUNCH(FMT_FEATS_SPLIT);
  else DG_GATHER_LAUNCH(FMT_CODE_SPLIT);
#undef DG_GATHER_LAUNCH
  DG_LAUNCH_OK("gather_norm_kernel");
  return DG_OK;
}

template <int NV>
static int launch_gather_all_nv(const GatherAllArgs& a, int total_blocks, int R, cudaStream_t st) {
  const size_t smem = (size_t)GF_WARPS * a.C * sizeof(float);
  DG_PRE(st);
  switch (R) {
    case 1: gather_all_kernel<NV, 1><<<total_blocks, GF_THREADS, smem, st>>>(a); break;
    case 2: gather_all_kernel<NV, 2><<<total_blocks, GF_THREADS, smem, st>>>(a); break;
    case 3: gather_all_kernel<NV, 3><<<total_blocks, GF_THREADS, smem, st>>>(a); break;
    default: gather_all_kernel<NV, 4><<<total_blocks, GF_THREADS, smem, st>>>(a); break;
  }
  DG_LAUNCH_OK("gather_all_kernel");
  return DG_OK;
}

// Features (split bf16 panels, register fast path) and code (split panels) in one launch; returns DG_ERR_UNSUPPORTED if the
// shapes/strides do not qualify so the caller can fall back to the two separate launches.
int launch_gather_all(const SetTable& fsets, const SetTable& csets, int nsets, int B, int C, int D, int H, int W,
                      const float* coords, int S, const int64_t* perms, float eps, int Prows, int ldf, int ldc, int fsplit,
                      const GatherOut& fo, const GatherOut& co, cudaStream_t st) {
  if (fsplit <= 1 || ldc > 128 || C != ldf) return DG_ERR_UNSUPPORTED;
  GatherAllArgs a;
  a.fsets = fsets; a.csets = csets; a.fo = fo; a.co = co;
  a.B = B; a.C = C; a.D = D; a.H = H; a.W = W; a.S = S; a.Prows = Prows; a.fsplit = fsplit; a.csplit = 2 * fsplit;  // the 2:1 interleave of gather_all_kernel relies on this
  a.nfeat_blocks = nsets * B * fsplit;
  a.eps = eps; a.coords = coords; a.perms = perms;
  const int total = a.nfeat_blocks + nsets * B * a.csplit;
  const int R = ldc / 32;
  switch (C / 128) {
    case 1: return launch_gather_all_nv<1>(a, total, R, st);
    case 2: return launch_gather_all_nv<2>(a, total, R, st);
    case 3: return launch_gather_all_nv<3>(a, total, R, st);
    case 4: return launch_gather_all_nv<4>(a, total, R, st);
    case 6: return launch_gather_all_nv<6>(a, total, R, st);
    case 8: return launch_gather_all_nv<8>(a, total, R, st);
    default: return DG_ERR_UNSUPPORTED;
  }
}

int launch_gather_bwd(const SetTable& tab, int nsets, int B, int C, int H, int W, const float* coords, int S,
                      const int64_t* perms, float eps, int Prows, int ld, const float* cn, const float* cn_lo,
                      const float* rnorm, const float* dC1, const float* dC2, int npairs, const PairTable& pt,
                      int has_depth, const GroupW& gw, cudaStream_t st, int ni, int nj, int njw) {
  const long long rows = (long long)nsets * B * S * S;
  const int blocks = (int)((rows * 32 + 255) / 256);
  DG_PRE(st);
#define DG_BWD(RT)                                                                                                    \
  launch_pdl(gather_norm_bwd_kernel<RT>, dim3(blocks), dim3(256), 0, st, tab, B, C, H, W, coords, S, perms, eps, Prows, \
             ld, cn, cn_lo, rnorm, dC1, dC2, npairs, pt, has_depth, gw, nsets, ni, nj, njw)
  switch (ld / 32) {   // the common code widths get their register loops sized at compile time
    case 1: DG_BWD(1); break;
    case 2: DG_BWD(2); break;
    case 3: DG_BWD(3); break;
    case 4: DG_BWD(4); break;
    default: DG_BWD(0); break;
  }
#undef DG_BWD
  DG_LAUNCH_OK("gather_norm_bwd_kernel");
  return DG_OK;
}

}  // namespace dg

extern "C" int dg_panel_ld(int channels) { return dg::round_up(channels, 32); }
extern "C" int dg_panel_rows(int P) { return dg::round_up(P, 64); }

extern "C" int dg_gather_norm(const float* t, const int64_t* strides, int B, int C, int H, int W, const float* coords,
                              int S, int nsets, const int32_t* set_coord, const int32_t* set_slot, const int64_t* perm,
                              float eps, int Prows, int ld, int format, void* out, void* out_lo, void* out16_hi,
                              void* out16_lo, float* rnorm, float* meanvec, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(t && strides && coords && out && rnorm, DG_ERR_INVALID, "dg_gather_norm: null pointer");
  DG_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && S > 0, DG_ERR_INVALID, "dg_gather_norm: bad sizes");
  DG_REQUIRE(ld >= C && (ld % 32) == 0, DG_ERR_INVALID, "dg_gather_norm: ld=%d must be a multiple of 32 >= C=%d", ld, C);
  DG_REQUIRE(Prows >= S * S, DG_ERR_INVALID, "dg_gather_norm: Prows=%d < S*S=%d", Prows, S * S);
  DG_REQUIRE(format >= DG_PANEL_F32 && format <= DG_PANEL_CODE_SPLIT, DG_ERR_INVALID, "dg_gather_norm: bad format");
  DG_REQUIRE(format != DG_PANEL_CODE_SPLIT || (out_lo && out16_hi), DG_ERR_INVALID,
             "dg_gather_norm: code-split format needs out_lo (fp32 remainder) and the 16-bit panel");
  SetTable tab;
  int rc = fill_sets("dg_gather_norm", t, strides, nsets, set_coord, set_slot, perm != nullptr, &tab);
  if (rc != DG_OK) return rc;
  GatherOut o;
  o.out = static_cast<float*>(out);
  o.out_lo = static_cast<float*>(out_lo);
  const bool code = format == DG_PANEL_CODE_SPLIT;
  o.hi16 = static_cast<__half*>(code ? out16_hi : out);
  o.lo16 = static_cast<__half*>(code ? out16_lo : out_lo);
  o.interleave = o.lo16 == nullptr ? 1 : 0;   // no separate lo panel given: the interleaved layout
  o.rnorm = rnorm;
  o.meanvec = meanvec;
  if (format == DG_PANEL_F32) o.interleave = 0;
  return launch_gather(format, tab, nsets, B, C, H, W, coords, S, perm, eps, Prows, ld, 1, o,
                       reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dg_gather_norm_bwd(float* grad, const int64_t* strides, int B, int C, int H, int W, const float* coords,
                                  int S, int nsets, const int32_t* set_coord, const int32_t* set_slot,
                                  const int64_t* perm, float eps, int Prows, int ld, const float* cn,
                                  const float* cn_lo, const float* rnorm, const float* dC1, const float* dC2,
                                  int npairs, const int32_t* pair_group, const float* pair_scale, int has_depth,
                                  const float* group_w, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(grad && strides && coords && cn && rnorm && dC1 && dC2 && group_w && pair_group && pair_scale,
             DG_ERR_INVALID, "dg_gather_norm_bwd: null pointer");
  DG_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && S > 0, DG_ERR_INVALID, "dg_gather_norm_bwd: bad sizes");
  DG_REQUIRE(ld >= C && (ld % 32) == 0 && ld <= 256, DG_ERR_UNSUPPORTED,
             "dg_gather_norm_bwd: ld=%d must be a multiple of 32 in [C,256]", ld);
  DG_REQUIRE(npairs > 0 && npairs <= DG_MAX_PAIRS, DG_ERR_INVALID, "dg_gather_norm_bwd: npairs=%d", npairs);
  SetTable tab;
  int rc = fill_sets("dg_gather_norm_bwd", grad, strides, nsets, set_coord, set_slot, perm != nullptr, &tab);
  if (rc != DG_OK) return rc;
  PairTable pt;
  for (int k = 0; k < npairs; ++k) {
    DG_REQUIRE(pair_group[k] >= 0 && pair_group[k] < DG_NUM_GROUPS, DG_ERR_INVALID, "dg_gather_norm_bwd: bad group");
    pt.group[k] = pair_group[k];
    pt.scale[k] = pair_scale[k];
  }
  for (int s = 0; s < nsets; ++s)
    DG_REQUIRE(set_slot[s] < npairs, DG_ERR_INVALID, "dg_gather_norm_bwd: slot %d >= npairs", set_slot[s]);
  GroupW gw;
  gw.arr = group_w;
  for (int g = 0; g < DG_NUM_GROUPS; ++g) gw.ptr[g] = nullptr;
  // split (tcgen05) panels carry one dC2 buffer per 128-row tile and one dC1 buffer per 128-column tile
  const int ni = (cn_lo != nullptr) ? Prows / 128 : 1;
  const int nj = ni;
  return launch_gather_bwd(tab, nsets, B, C, H, W, coords, S, perm, eps, Prows, ld, cn, cn_lo, rnorm, dC1, dC2, npairs,
                           pt, has_depth, gw, reinterpret_cast<cudaStream_t>(stream), ni, nj, 128);
}

extern "C" int dg_pool_normalize(const float* t, const int64_t* strides, int N, int C, int H, int W, float eps,
                                 float* out, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(t && strides && out, DG_ERR_INVALID, "dg_pool_normalize: null pointer");
  DG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && C <= 12288, DG_ERR_INVALID, "dg_pool_normalize: bad sizes");
  DG_PRE(reinterpret_cast<cudaStream_t>(stream));
  pool_normalize_kernel<<<N, 256, (size_t)C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      t, strides[0], strides[1], strides[2], strides[3], C, H, W, eps, out);
  DG_LAUNCH_OK("pool_normalize_kernel");
  return DG_OK;
}

extern "C" int dg_norm_dim1(const float* t, int N, int C, long long inner, long long sN, long long sC, long long sP,
                            float eps, float* out, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(t && out, DG_ERR_INVALID, "dg_norm_dim1: null pointer");
  DG_REQUIRE(N > 0 && C > 0 && inner > 0, DG_ERR_INVALID, "dg_norm_dim1: bad sizes");
  const long long total = (long long)N * inner;
  const long long threads = sC == 1 ? total * 32 : total;
  DG_REQUIRE((threads + 255) / 256 < (1LL << 31), DG_ERR_UNSUPPORTED, "dg_norm_dim1: tensor too large for one launch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DG_PRE(st);
  norm_dim1_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(t, N, C, inner, sN, sC, sP, eps, out);
  DG_LAUNCH_OK("norm_dim1_kernel");
  return DG_OK;
}
