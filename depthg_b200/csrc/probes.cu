// The trainer's two probe losses on the DETACHED code map:
//   * linear probe: 1x1 conv -> bilinear upsample to the label size -> masked cross-entropy
//     (src/train_segmentation.py:419-437).  The reference materialises the upsampled logits
//     [B,K,Hl,Wl] four times over (interpolate, permute+reshape, boolean gather, log-softmax);
//     here the upsample, the softmax, the loss and the whole backward down to the low-resolution
//     logit gradient are one pass over the label map, which is read exactly once.
//   * cluster probe: ClusterLookup.forward (src/modules.py:659-675): cosine similarity to the
//     cluster centres, hard (alpha=None) or soft assignment, loss and centre gradients.
// Both only train their own small parameters (the code map is detached), so the "backward" is a
// by-product of the forward pass: the entry points return unit gradients of the loss.
#include <math.h>

#include "common.cuh"

namespace dg {

constexpr int KP = 32;  // padded class pitch of the low-resolution logits / their gradient

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Linear probe, step 1: low-resolution logits  L[pix][k] = bias[k] + sum_d W[k][d] * code[pix][d]
// A warp takes 4 pixels at a time (lane = class, 4 accumulators per lane, the pixels' channels staged in
// shared memory as float4 so one broadcast load feeds 4 FMAs).  Also clears what the later steps accumulate into.
__global__ void __launch_bounds__(256) probe_logits_kernel(const float* __restrict__ code, int64_t sb, int64_t sc,
                                                           int64_t sh, int64_t sw, int D, int h, int w, int npix,
                                                           const float* __restrict__ weight,
                                                           const float* __restrict__ bias, int K,
                                                           float* __restrict__ logits, float* __restrict__ dlogits,
                                                           double* __restrict__ loss_sum,
                                                           unsigned long long* __restrict__ count,
                                                           float* __restrict__ dweight, float* __restrict__ dbias) {
  extern __shared__ float4 lsm4[];
  const int DP = D | 1;
  float4* xs = lsm4;                                        // [8 warps][128] channel d of 4 pixels
  float* wsm = reinterpret_cast<float*>(lsm4 + 8 * 128);    // [K][D|1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += 8)
    for (int d = lane; d < D; d += 32) wsm[k * DP + d] = weight[k * D + d];
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      *loss_sum = 0.0;
      *count = 0ull;
    }
    if (dweight)
      for (int i = threadIdx.x; i < K * D; i += blockDim.x) dweight[i] = 0.f;
    if (dbias && threadIdx.x < K) dbias[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const float* wr = wsm + (lane < K ? lane : 0) * DP;
  const float bk = (bias && lane < K) ? bias[lane] : 0.f;
  float4* x4 = xs + warp * 128;
  float* x1 = reinterpret_cast<float*>(x4);
  for (int pix0 = (blockIdx.x * 8 + warp) * 4; pix0 < npix; pix0 += gridDim.x * 32) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pix = min(pix0 + j, npix - 1);
      const int b = pix / (h * w), r = pix - b * h * w, y = r / w, x = r - y * w;
      const float* cp = code + b * sb + y * sh + x * sw;
      for (int d = lane; d < D; d += 32) x1[d * 4 + j] = __ldg(cp + d * sc);
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 6
    for (int d = 0; d < D; ++d) {
      const float wv = wr[d];
      const float4 xv = x4[d];
      a0 = fmaf(wv, xv.x, a0);
      a1 = fmaf(wv, xv.y, a1);
      a2 = fmaf(wv, xv.z, a2);
      a3 = fmaf(wv, xv.w, a3);
    }
    const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (pix0 + j < npix) {
        logits[(size_t)(pix0 + j) * KP + lane] = lane < K ? acc[j] + bk : 0.f;
        dlogits[(size_t)(pix0 + j) * KP + lane] = 0.f;
      }
    }
  }
}

// Source index of ATen's upsample_bilinear2d with align_corners=False (area_pixel_compute_source_index).
__device__ __forceinline__ float up_src(int dst, float scale) {
  const float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}
// first destination index whose source cell (int)src is >= cell
__device__ __forceinline__ int up_first(int cell, float scale, int n_out) {
  if (cell <= 0) return 0;
  int i = (int)ceilf(((float)cell + 0.5f) / scale - 0.5f);
  i = max(0, min(i, n_out));
  while (i > 0 && (int)up_src(i - 1, scale) >= cell) --i;
  while (i < n_out && (int)up_src(i, scale) < cell) ++i;
  return i;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Linear probe, step 2: one warp per low-resolution CELL (y0,x0) = the label pixels whose bilinear
// footprint is the corner set {y0,y0+1}x{x0,x0+1}.  Phase A (lane = label pixel) evaluates the
// upsampled logits, softmax and loss; phase B (lane = class) folds the per-pixel softmax gradient
// back onto the four corners, so the only global atomics are 4 per class per cell.
// KS = class slots (K rounded up); slots >= K hold -1e30 logits and fall out of the softmax.
// The softmax offset is the bilinear blend of the four corner maxima: an upper bound of every
// upsampled logit (so exp never overflows) known before the class loop starts; the rare pixel
// whose classes all sit > 80 below that bound is redone with its exact maximum.
template <int KS>
__global__ void __launch_bounds__(256, 3) probe_ce_kernel(const float* __restrict__ logits,
                                                          const int64_t* __restrict__ labels, int64_t lsb,
                                                          int64_t lsh, int64_t lsw, int ncell, int h, int w, int Hl,
                                                          int Wl, int K, float scale_y, float scale_x,
                                                          float* __restrict__ dlogits, double* __restrict__ loss_sum,
                                                          unsigned long long* __restrict__ count) {
  constexpr int CH = 32, GS = KS | 1;
  constexpr int PER_WARP = KS * 4 + CH * GS + CH * 4 + (4 - (KS * 4 + CH * GS + CH * 4) % 4) % 4;  // float4-aligned
  constexpr float LOG2E = 1.4426950408889634f;
  extern __shared__ float4 csm4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = reinterpret_cast<float*>(csm4) + warp * PER_WARP;
  float4* Cc = reinterpret_cast<float4*>(my);             // [KS] corner logits per class
  float4* Wt = reinterpret_cast<float4*>(my + KS * 4);    // [CH] corner weights per pixel
  float* G = my + KS * 4 + CH * 4;                        // [CH][GS] softmax gradient per pixel
  const int cell = blockIdx.x * 8 + warp;
  if (cell >= ncell) return;
  const int b = cell / (h * w), r = cell - b * h * w, y0 = r / w, x0 = r - y0 * w;
  const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
  const size_t p00 = ((size_t)(b * h + y0) * w + x0) * KP, p01 = p00 + (size_t)xp * KP;
  const size_t p10 = p00 + (size_t)yp * w * KP, p11 = p10 + (size_t)xp * KP;
  const float4 cmine = lane < K ? make_float4(logits[p00 + lane], logits[p01 + lane], logits[p10 + lane],
                                              logits[p11 + lane])
                                : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
  if (lane < KS) Cc[lane] = cmine;
  const float4 M = make_float4(warp_max(cmine.x), warp_max(cmine.y), warp_max(cmine.z), warp_max(cmine.w));
  const int i0 = up_first(y0, scale_y, Hl), i1 = y0 == h - 1 ? Hl : up_first(y0 + 1, scale_y, Hl);
  const int j0 = up_first(x0, scale_x, Wl), j1 = x0 == w - 1 ? Wl : up_first(x0 + 1, scale_x, Wl);
  const int ncol = j1 - j0, npx = (i1 - i0) * ncol;
  const float inv_ncol = 1.f / (float)max(ncol, 1);
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, loss = 0.f;
  int cnt = 0;
  for (int base = 0; base < npx; base += CH) {
    __syncwarp();
    {  // ---- phase A: lane = pixel
      const int p = min(base + lane, npx - 1);
      const bool act = base + lane < npx;
      const int pr = (int)(((float)p + 0.5f) * inv_ncol);  // p / ncol (exact: p + 0.5 is never a multiple of ncol)
      const int i = i0 + pr, j = j0 + (p - pr * ncol);
      const float ly = up_src(i, scale_y) - (float)y0, lx = up_src(j, scale_x) - (float)x0;
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
      const long long lab = labels[b * lsb + i * lsh + j * lsw];
      const bool valid = act && lab >= 0 && lab < K;
      const int lab32 = valid ? (int)lab : 0;
      float mb = fmaf(w11, M.w, fmaf(w10, M.z, fmaf(w01, M.y, w00 * M.x)));
      float e[KS];
      float s = 0.f;
      {
        const float m2 = mb * LOG2E;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const float4 c = Cc[k];
          const float l = fmaf(w11, c.w, fmaf(w10, c.z, fmaf(w01, c.y, w00 * c.x)));
          e[k] = ex2_approx(fmaf(l, LOG2E, -m2));
          s += e[k];
        }
      }
      if (s < 1e-30f) {  // every class far below the corner-max blend: redo with the exact maximum
        mb = -INFINITY;
        for (int k = 0; k < K; ++k) {
          const float4 c = Cc[k];
          mb = fmaxf(mb, fmaf(w11, c.w, fmaf(w10, c.z, fmaf(w01, c.y, w00 * c.x))));
        }
        const float m2 = mb * LOG2E;
        s = 0.f;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const float4 c = Cc[k];
          const float l = fmaf(w11, c.w, fmaf(w10, c.z, fmaf(w01, c.y, w00 * c.x)));
          e[k] = ex2_approx(fmaf(l, LOG2E, -m2));
          s += e[k];
        }
      }
      const float rs = valid ? 1.f / s : 0.f;
#pragma unroll
      for (int k = 0; k < KS; ++k) G[lane * GS + k] = e[k] * rs;
      Wt[lane] = act ? make_float4(w00, w01, w10, w11) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        const float4 c = Cc[lab32];
        const float llab = fmaf(w11, c.w, fmaf(w10, c.z, fmaf(w01, c.y, w00 * c.x)));
        G[lane * GS + lab32] -= 1.f;
        loss += (logf(s) + mb) - llab;
        ++cnt;
      }
    }
    __syncwarp();
    if (lane < KS) {  // ---- phase B: lane = class
#pragma unroll 8
      for (int q = 0; q < CH; ++q) {
        const float g = G[q * GS + lane];
        const float4 wt = Wt[q];
        a00 = fmaf(wt.x, g, a00);
        a01 = fmaf(wt.y, g, a01);
        a10 = fmaf(wt.z, g, a10);
        a11 = fmaf(wt.w, g, a11);
      }
    }
  }
  if (lane < K && npx > 0) {
    atomicAdd(dlogits + p00 + lane, a00);
    atomicAdd(dlogits + p01 + lane, a01);
    atomicAdd(dlogits + p10 + lane, a10);
    atomicAdd(dlogits + p11 + lane, a11);
  }
  loss = warp_sum(loss);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0 && cnt > 0) {
    atomicAdd(loss_sum, (double)loss);
    atomicAdd(count, (unsigned long long)cnt);
  }
}

template <int KS>
static int launch_probe_ce(const float* logits, const int64_t* labels, const int64_t* ls, int npix, int h, int w, int Hl,
                           int Wl, int K, float* dlogits, double* loss_sum, unsigned long long* count,
                           cudaStream_t st) {
  constexpr int CH = 32, GS = KS | 1;
  constexpr int PER_WARP = KS * 4 + CH * GS + CH * 4 + (4 - (KS * 4 + CH * GS + CH * 4) % 4) % 4;
  const size_t smem = 8 * (size_t)PER_WARP * sizeof(float);
  // ATen: scale = input_size / output_size in float (area_pixel_compute_scale, align_corners=False)
  const float scale_y = (float)h / (float)Hl, scale_x = (float)w / (float)Wl;
  DG_PRE(st);
  probe_ce_kernel<KS><<<ceil_div(npix, 8), 256, smem, st>>>(logits, labels, ls[0], ls[1], ls[2], npix, h, w, Hl, Wl, K,
                                                            scale_y, scale_x, dlogits, loss_sum, count);
  DG_LAUNCH_OK("probe_ce_kernel");
  return DG_OK;
}

// Linear probe, step 3: unit gradients  dW[k][d] = (1/count) sum_pix dL[pix][k] code[pix][d],
// db[k] = (1/count) sum_pix dL[pix][k];  block 0 also writes the loss (sum / count).
// Thread = (channel d, half of the classes): 16 register accumulators, one channel load and four
// broadcast float4 loads of dL per pixel.  Channel index D is a virtual all-ones channel = the bias.
// CTAs loop over 64-pixel chunks and touch global memory with atomics once, at the end.
__global__ void __launch_bounds__(320) probe_wgrad_kernel(const float* __restrict__ code, int64_t sb, int64_t sc,
                                                          int64_t sh, int64_t sw, int D, int h, int w, int npix,
                                                          const float* __restrict__ dlogits, int K,
                                                          const double* __restrict__ loss_sum,
                                                          const unsigned long long* __restrict__ count,
                                                          float* __restrict__ dweight, float* __restrict__ dbias,
                                                          float* __restrict__ loss_out) {
  constexpr int CH = 64;
  extern __shared__ float4 gsm4[];
  const int DP = (D + 1) | 1;
  float4* dl4 = gsm4;                                          // [CH][8] = [CH][32 classes]
  float* cx = reinterpret_cast<float*>(gsm4 + CH * 8);         // [CH][DP], column D = 1
  __shared__ int64_t pixoff[CH];
  const unsigned long long c = *count;
  const float inv = 1.f / (float)c;  // count == 0 -> 0 * inf = NaN, like the reference's mean over nothing
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out) loss_out[0] = (float)(*loss_sum / (double)c);
  if (!dweight) return;
  const int TD = (int)blockDim.x >> 1;
  const int d = threadIdx.x % TD, kh = threadIdx.x / TD;       // classes [16 kh, 16 kh + 16)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (int)blockDim.x >> 5;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int p0 = blockIdx.x * CH; p0 < npix; p0 += gridDim.x * CH) {
    const int n = min(CH, npix - p0);
    __syncthreads();
    if (threadIdx.x < CH) {
      const int pix = min(p0 + (int)threadIdx.x, npix - 1);
      const int b = pix / (h * w), r = pix - b * h * w, y = r / w, x = r - y * w;
      pixoff[threadIdx.x] = b * sb + y * sh + x * sw;
    }
    for (int i = threadIdx.x; i < CH * 8; i += blockDim.x)
      dl4[i] = i < n * 8 ? reinterpret_cast<const float4*>(dlogits + (size_t)p0 * KP)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (sc == 1) {  // channels-last: a warp per pixel, lanes along the contiguous channels
      for (int p = warp; p < n; p += nwarp)
        for (int dd = lane; dd <= D; dd += 32) cx[p * DP + dd] = dd < D ? __ldg(code + pixoff[p] + dd) : 1.f;
    } else {        // NCHW: a warp per channel, lanes along the pixels
      for (int dd = warp; dd <= D; dd += nwarp)
        for (int p = lane; p < n; p += 32) cx[p * DP + dd] = dd < D ? __ldg(code + pixoff[p] + dd * sc) : 1.f;
    }
    __syncthreads();
    if (d <= D) {
      const float4* dlk = dl4 + kh * 4;
#pragma unroll 4
      for (int p = 0; p < n; ++p) {
        const float xv = cx[p * DP + d];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 g = dlk[p * 8 + q];
          acc[4 * q + 0] = fmaf(g.x, xv, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(g.y, xv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(g.z, xv, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(g.w, xv, acc[4 * q + 3]);
        }
      }
    }
  }
  if (d <= D) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = kh * 16 + i;
      if (k < K) {
        if (d < D) atomicAdd(dweight + k * D + d, acc[i] * inv);
        else if (dbias) atomicAdd(dbias + k, acc[i] * inv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster probe.  mode 0: alpha=None (one-hot assignment, loss, centre gradients);
// mode 1: softmax(inner*alpha) assignment and loss (forward only); mode 2: log_softmax(inner*alpha).
// One warp per pixel (looping), lane = cluster.  out_probs is [B,h,w,N].
__global__ void __launch_bounds__(256) cluster_probe_kernel(const float* __restrict__ code, int64_t sb, int64_t sc,
                                                            int64_t sh, int64_t sw, int D, int h, int w, int npix,
                                                            const float* __restrict__ clusters, int N, int mode,
                                                            float alpha, float* __restrict__ out_probs,
                                                            float* __restrict__ part /*[grid][N*D+1] or NULL*/,
                                                            int want_grad) {
  extern __shared__ float4 ksm4[];
  float* sm = reinterpret_cast<float*>(ksm4);
  const int DP = D | 1;
  float* chat = sm;                  // [N][DP] normalised centres
  float* dacc = chat + N * DP;       // [N][D]  sum of x_hat per assigned centre (mode 0)
  float* xs = dacc + N * D + ((4 - ((N * DP + N * D) & 3)) & 3);  // [8][128][4] per-warp normalised pixels (float4-aligned)
  __shared__ float lsum[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += 8) {  // F.normalize(clusters, dim=1), eps 1e-12
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = clusters[n * D + d];
      ss = fmaf(v, v, ss);
    }
    const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int d = lane; d < D; d += 32) chat[n * DP + d] = clusters[n * D + d] / nrm;
  }
  for (int i = threadIdx.x; i < N * D; i += blockDim.x) dacc[i] = 0.f;
  __syncthreads();
  float4* x4 = reinterpret_cast<float4*>(xs) + warp * 128;   // channel d of the warp's 4 current pixels
  float* x1 = reinterpret_cast<float*>(x4);
  const float* cr = chat + (lane < N ? lane : 0) * DP;
  float lacc = 0.f;
  for (int pix0 = (blockIdx.x * 8 + warp) * 4; pix0 < npix; pix0 += gridDim.x * 32) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pix = min(pix0 + j, npix - 1);
      const int b = pix / (h * w), r = pix - b * h * w, y = r / w, xx = r - y * w;
      const float* cp = code + b * sb + y * sh + xx * sw;
      float v[4], ss = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int d = lane + 32 * t;
        v[t] = d < D ? __ldg(cp + d * sc) : 0.f;
        ss = fmaf(v[t], v[t], ss);
      }
      const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < D) x1[(lane + 32 * t) * 4 + j] = v[t] / nrm;
    }
    __syncwarp();
    float in4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 6
    for (int d = 0; d < D; ++d) {
      const float cv = cr[d];
      const float4 xv = x4[d];
      in4[0] = fmaf(xv.x, cv, in4[0]);
      in4[1] = fmaf(xv.y, cv, in4[1]);
      in4[2] = fmaf(xv.z, cv, in4[2]);
      in4[3] = fmaf(xv.w, cv, in4[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pix = pix0 + j;
      if (pix >= npix) break;
      const float inner = lane < N ? in4[j] : -INFINITY;
      if (mode == 0) {
        const float m = warp_max(inner);
        const unsigned ball = __ballot_sync(0xffffffffu, inner == m && lane < N);
        const int am = ball ? __ffs(ball) - 1 : 0;  // first maximum, like torch.argmax
        if (out_probs && lane < N) out_probs[(size_t)pix * N + lane] = lane == am ? 1.f : 0.f;
        if (lane == 0) lacc -= m;
        if (want_grad)
          for (int d = lane; d < D; d += 32) atomicAdd(dacc + am * D + d, x1[d * 4 + j]);
      } else {
        const float z = inner * alpha;
        const float m = warp_max(z);
        const float e = lane < N ? expf(z - m) : 0.f;
        const float s = warp_sum(e);
        if (mode == 1) {
          const float p = e / s;
          if (out_probs && lane < N) out_probs[(size_t)pix * N + lane] = p;
          const float pl = warp_sum(lane < N ? p * inner : 0.f);
          if (lane == 0) lacc -= pl;
        } else if (out_probs && lane < N) {
          out_probs[(size_t)pix * N + lane] = (z - m) - logf(s);
        }
      }
    }
  }
  if (lane == 0) lsum[warp] = lacc;
  __syncthreads();
  if (!part) return;
  float* mine = part + (size_t)blockIdx.x * (N * D + 1);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += lsum[i];
    mine[N * D] = t;
  }
  if (want_grad)
    for (int i = threadIdx.x; i < N * D; i += blockDim.x) mine[i] = dacc[i];
}

// Sums the per-CTA partials (deterministically), then
// loss = sum/M; centre gradient through F.normalize:  dc = (g - c_hat <c_hat,g>) / max(|c|,eps),  g = -S_n / M
__global__ void __launch_bounds__(128) cluster_finalize_kernel(const float* __restrict__ clusters, int N, int D,
                                                               int npix, const float* __restrict__ part, int nparts,
                                                               float* __restrict__ loss_out,
                                                               float* __restrict__ dclusters) {
  __shared__ float red[4];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = N * D + 1;
  if (n == 0 && loss_out) {
    float t = 0.f;
    for (int i = tid; i < nparts; i += 128) t += part[(size_t)i * stride + N * D];
    t = warp_sum(t);
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (tid == 0) loss_out[0] = (float)(((double)red[0] + red[1] + red[2] + red[3]) / (double)npix);
    __syncthreads();
  }
  if (!dclusters) return;
  const float sgn = -1.f / (float)npix;
  float c = 0.f, gv = 0.f;
  if (tid < D) {
    float g4[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 16 <= nparts; i += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) g4[u & 3] += part[(size_t)(i + u) * stride + n * D + tid];
    }
    for (; i < nparts; ++i) g4[0] += part[(size_t)i * stride + n * D + tid];
    gv = ((g4[0] + g4[1]) + (g4[2] + g4[3])) * sgn;
    c = clusters[n * D + tid];
  }
  float ss = warp_sum(c * c);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  const float raw = sqrtf(red[0] + red[1] + red[2] + red[3]);
  const float nrm = fmaxf(raw, 1e-12f);
  __syncthreads();
  float dot = warp_sum((c / nrm) * gv);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = red[0] + red[1] + red[2] + red[3];
  // below eps the reference divides by the constant eps: no projection term
  if (tid < D) dclusters[n * D + tid] = raw > 1e-12f ? (gv - (c / nrm) * dot) / nrm : gv / nrm;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int CLUSTER_GRID = 148 * 2;  // CTAs of the cluster probe = per-CTA partial slots in the workspace

}  // namespace dg

using namespace dg;

extern "C" size_t dg_probe_workspace_bytes(int B, int h, int w, int D, int K) {
  if (B <= 0 || h <= 0 || w <= 0 || D <= 0 || K <= 0) return 0;
  const size_t npix = (size_t)B * h * w;
  return 2 * align256(npix * KP * sizeof(float)) + align256((size_t)CLUSTER_GRID * ((size_t)K * D + 1) * sizeof(float)) + 256;
}

extern "C" int dg_linear_probe_ce(const float* code, const int64_t* strides, int B, int D, int h, int w,
                                  const float* weight, const float* bias, int K, const int64_t* labels,
                                  const int64_t* label_strides, int Hl, int Wl, float* loss_out, float* dweight,
                                  float* dbias, void* ws, size_t ws_bytes, dg_stream_t stream) {
  DG_REQUIRE(code && strides && weight && labels && label_strides && loss_out && ws, DG_ERR_INVALID,
             "dg_linear_probe_ce: null pointer");
  DG_REQUIRE(B > 0 && h > 0 && w > 0 && Hl > 0 && Wl > 0, DG_ERR_INVALID, "dg_linear_probe_ce: bad sizes");
  DG_REQUIRE(D > 0 && D <= 128, DG_ERR_UNSUPPORTED, "dg_linear_probe_ce: code dim %d > 128 not supported", D);
  DG_REQUIRE(K > 0 && K <= KP, DG_ERR_UNSUPPORTED, "dg_linear_probe_ce: %d classes > %d not supported", K, KP);
  DG_REQUIRE((long long)B * h * w < (1ll << 26), DG_ERR_UNSUPPORTED, "dg_linear_probe_ce: code map too large");
  DG_REQUIRE(ws_bytes >= dg_probe_workspace_bytes(B, h, w, D, K), DG_ERR_INVALID,
             "dg_linear_probe_ce: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int npix = B * h * w;
  char* p = static_cast<char*>(ws);
  float* logits = reinterpret_cast<float*>(p);
  p += align256((size_t)npix * KP * sizeof(float));
  float* dlogits = reinterpret_cast<float*>(p);
  p += align256((size_t)npix * KP * sizeof(float));
  p += align256((size_t)CLUSTER_GRID * ((size_t)K * D + 1) * sizeof(float));
  double* loss_sum = reinterpret_cast<double*>(p);
  unsigned long long* count = reinterpret_cast<unsigned long long*>(p + 8);

  DG_PRE(st);
  probe_logits_kernel<<<min(ceil_div(npix, 32), 148 * 6), 256,
                        8 * 128 * sizeof(float4) + (size_t)K * (D | 1) * sizeof(float), st>>>(
      code, strides[0], strides[1], strides[2], strides[3], D, h, w, npix, weight, bias, K, logits, dlogits, loss_sum,
      count, dweight, dbias);
  DG_LAUNCH_OK("probe_logits_kernel");

  int rc;
  const int64_t* ls = label_strides;
  if (K <= 4) rc = launch_probe_ce<4>(logits, labels, ls, npix, h, w, Hl, Wl, K, dlogits, loss_sum, count, st);
  else if (K <= 8) rc = launch_probe_ce<8>(logits, labels, ls, npix, h, w, Hl, Wl, K, dlogits, loss_sum, count, st);
  else if (K <= 16) rc = launch_probe_ce<16>(logits, labels, ls, npix, h, w, Hl, Wl, K, dlogits, loss_sum, count, st);
  else if (K <= 28) rc = launch_probe_ce<28>(logits, labels, ls, npix, h, w, Hl, Wl, K, dlogits, loss_sum, count, st);
  else rc = launch_probe_ce<32>(logits, labels, ls, npix, h, w, Hl, Wl, K, dlogits, loss_sum, count, st);
  if (rc != DG_OK) return rc;

  const int TD = round_up(D + 1, 32);
  const size_t wg_smem = 64 * 8 * sizeof(float4) + (size_t)64 * ((D + 1) | 1) * sizeof(float);
  DG_PRE(st);
  probe_wgrad_kernel<<<dweight ? min(ceil_div(npix, 64), 148 * 2) : 1, 2 * TD, wg_smem, st>>>(
      code, strides[0], strides[1], strides[2], strides[3], D, h, w, npix, dlogits, K, loss_sum, count, dweight, dbias,
      loss_out);
  DG_LAUNCH_OK("probe_wgrad_kernel");
  return DG_OK;
}

extern "C" int dg_cluster_probe(const float* code, const int64_t* strides, int B, int D, int h, int w,
                                const float* clusters, int N, int mode, float alpha, float* loss_out, float* probs_out,
                                float* dclusters, void* ws, size_t ws_bytes, dg_stream_t stream) {
  DG_REQUIRE(code && strides && clusters && ws, DG_ERR_INVALID, "dg_cluster_probe: null pointer");
  DG_REQUIRE(B > 0 && h > 0 && w > 0, DG_ERR_INVALID, "dg_cluster_probe: bad sizes");
  DG_REQUIRE(mode >= 0 && mode <= 2, DG_ERR_INVALID, "dg_cluster_probe: mode must be 0, 1 or 2");
  DG_REQUIRE(mode != 2 || probs_out, DG_ERR_INVALID, "dg_cluster_probe: log-prob mode needs probs_out");
  DG_REQUIRE(mode == 0 || !dclusters, DG_ERR_UNSUPPORTED,
             "dg_cluster_probe: centre gradients are only provided for alpha=None (the training call)");
  DG_REQUIRE(D > 0 && D <= 128, DG_ERR_UNSUPPORTED, "dg_cluster_probe: code dim %d > 128 not supported", D);
  DG_REQUIRE(N > 0 && N <= 32, DG_ERR_UNSUPPORTED, "dg_cluster_probe: %d clusters > 32 not supported", N);
  DG_REQUIRE((long long)B * h * w < (1ll << 26), DG_ERR_UNSUPPORTED, "dg_cluster_probe: code map too large");
  DG_REQUIRE(ws_bytes >= dg_probe_workspace_bytes(B, h, w, D, N), DG_ERR_INVALID,
             "dg_cluster_probe: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int npix = B * h * w;
  char* p = static_cast<char*>(ws) + 2 * align256((size_t)npix * KP * sizeof(float));
  float* part = reinterpret_cast<float*>(p);
  const size_t smem = (size_t)(N * (D | 1) + N * D + 4 + 8 * 128 * 4) * sizeof(float);
  const int grid = min(ceil_div(npix, 32), CLUSTER_GRID);
  const bool need_part = mode != 2 && (loss_out || dclusters);
  static PerDevice attr_pd = {};
  size_t& attr_done = per_device(attr_pd);
  if (!attr_done) {  // up to ~50 KB at N = 32, D = 128
    DG_CUDA_OK(cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_done = 1;
  }
  DG_PRE(st);
  cluster_probe_kernel<<<grid, 256, smem, st>>>(code, strides[0], strides[1], strides[2], strides[3], D, h, w, npix,
                                                clusters, N, mode, alpha, probs_out, need_part ? part : nullptr,
                                                dclusters ? 1 : 0);
  DG_LAUNCH_OK("cluster_probe_kernel");
  if (need_part) {
    DG_PRE(st);
    cluster_finalize_kernel<<<dclusters ? N : 1, 128, 0, st>>>(clusters, N, D, npix, part, grid, loss_out, dclusters);
    DG_LAUNCH_OK("cluster_finalize_kernel");
  }
  return DG_OK;
}
