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
// One warp per pixel, lane = class.
__global__ void __launch_bounds__(256) probe_logits_kernel(const float* __restrict__ code, int64_t sb, int64_t sc,
                                                           int64_t sh, int64_t sw, int D, int h, int w, int npix,
                                                           const float* __restrict__ weight,
                                                           const float* __restrict__ bias, int K,
                                                           float* __restrict__ logits) {
  extern __shared__ float wsm[];  // [K][D+1]
  const int DP = D | 1;
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) wsm[(i / D) * DP + (i % D)] = weight[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= npix) return;
  const int b = pix / (h * w), r = pix - b * h * w, y = r / w, x = r - y * w;
  const float* cp = code + b * sb + y * sh + x * sw;
  const int kk = lane < K ? lane : 0;
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    const float xv = d < D ? __ldg(cp + d * sc) : 0.f;
    const int n = min(32, D - d0);
    for (int j = 0; j < n; ++j) acc = fmaf(wsm[kk * DP + d0 + j], __shfl_sync(0xffffffffu, xv, j), acc);
  }
  logits[(size_t)pix * KP + lane] = lane < K ? acc + (bias ? bias[lane] : 0.f) : 0.f;
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

// Linear probe, step 2: one warp per low-resolution CELL (y0,x0) = the label pixels whose bilinear
// footprint is the corner set {y0,y0+1}x{x0,x0+1}.  Phase A (lane = label pixel) evaluates the
// upsampled logits, log-softmax and loss; phase B (lane = class) folds the per-pixel softmax
// gradient back onto the four corners, so the only global atomics are 4 per class per cell.
__global__ void __launch_bounds__(256) probe_ce_kernel(const float* __restrict__ logits,
                                                       const int64_t* __restrict__ labels, int64_t lsb, int64_t lsh,
                                                       int64_t lsw, int ncell, int h, int w, int Hl, int Wl, int K,
                                                       float scale_y, float scale_x, float* __restrict__ dlogits,
                                                       double* __restrict__ loss_sum,
                                                       unsigned long long* __restrict__ count) {
  constexpr int CH = 64, GS = 33;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = sm + warp * (KP * 4 + CH * GS + CH * 4);
  float4* Cc = reinterpret_cast<float4*>(my);                       // [KP] corner logits per class
  float* G = my + KP * 4;                                           // [CH][GS] softmax gradient per pixel
  float4* Wt = reinterpret_cast<float4*>(my + KP * 4 + CH * GS);    // [CH] corner weights per pixel
  const int cell = blockIdx.x * 8 + warp;
  if (cell >= ncell) return;
  const int b = cell / (h * w), r = cell - b * h * w, y0 = r / w, x0 = r - y0 * w;
  const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
  const size_t p00 = ((size_t)(b * h + y0) * w + x0) * KP, p01 = p00 + (size_t)xp * KP;
  const size_t p10 = p00 + (size_t)yp * w * KP, p11 = p10 + (size_t)xp * KP;
  Cc[lane] = lane < K ? make_float4(logits[p00 + lane], logits[p01 + lane], logits[p10 + lane], logits[p11 + lane])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
  const int i0 = up_first(y0, scale_y, Hl), i1 = y0 == h - 1 ? Hl : up_first(y0 + 1, scale_y, Hl);
  const int j0 = up_first(x0, scale_x, Wl), j1 = x0 == w - 1 ? Wl : up_first(x0 + 1, scale_x, Wl);
  const int ncol = j1 - j0, npx = (i1 - i0) * ncol;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, loss = 0.f;
  int cnt = 0;
  for (int base = 0; base < npx; base += CH) {
    const int n = min(CH, npx - base);
    __syncwarp();
    for (int q = lane; q < n; q += 32) {  // ---- phase A
      const int p = base + q, i = i0 + p / ncol, j = j0 + p % ncol;
      const float ly = up_src(i, scale_y) - (float)y0, lx = up_src(j, scale_x) - (float)x0;
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
      const long long lab = labels[b * lsb + i * lsh + j * lsw];
      const bool valid = lab >= 0 && lab < K;
      float l[KP];
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (k < K) {
          const float4 c = Cc[k];
          l[k] = hy * (hx * c.x + lx * c.y) + ly * (hx * c.z + lx * c.w);  // ATen's evaluation order
          m = fmaxf(m, l[k]);
        }
      }
      float s = 0.f, llab = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (k < K) {
          if (k == (int)lab) llab = l[k];
          l[k] = expf(l[k] - m);
          s += l[k];
        }
      }
      const float rs = valid ? 1.f / s : 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k)
        if (k < K) G[q * GS + k] = valid ? l[k] * rs - (k == (int)lab ? 1.f : 0.f) : 0.f;
      Wt[q] = make_float4(w00, w01, w10, w11);
      if (valid) {
        loss += (logf(s) + m) - llab;
        ++cnt;
      }
    }
    __syncwarp();
    if (lane < K) {  // ---- phase B
      for (int q = 0; q < n; ++q) {
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

// Linear probe, step 3: unit gradients  dW[k][d] = (1/count) sum_pix dL[pix][k] code[pix][d],
// db[k] = (1/count) sum_pix dL[pix][k];  block 0 also writes the loss (sum / count).
__global__ void __launch_bounds__(256) probe_wgrad_kernel(const float* __restrict__ code, int64_t sb, int64_t sc,
                                                          int64_t sh, int64_t sw, int D, int h, int w, int npix,
                                                          const float* __restrict__ dlogits, int K,
                                                          const double* __restrict__ loss_sum,
                                                          const unsigned long long* __restrict__ count,
                                                          float* __restrict__ dweight, float* __restrict__ dbias,
                                                          float* __restrict__ loss_out) {
  constexpr int CH = 64;
  extern __shared__ float sm[];
  const int DP = D | 1;
  float* dl = sm;            // [CH][KP+1]
  float* cx = sm + CH * 33;  // [CH][DP]
  const unsigned long long c = *count;
  const float inv = 1.f / (float)c;  // count == 0 -> inf * 0 = NaN, like the reference's mean over nothing
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out) loss_out[0] = (float)(*loss_sum / (double)c);
  if (!dweight) return;
  const int p0 = blockIdx.x * CH, n = min(CH, npix - p0);
  for (int i = threadIdx.x; i < n * KP; i += blockDim.x) dl[(i / KP) * 33 + (i % KP)] = dlogits[(size_t)p0 * KP + i];
  if (sc == 1) {
    for (int i = threadIdx.x; i < n * D; i += blockDim.x) {
      const int p = i / D, d = i - p * D, pix = p0 + p;
      const int b = pix / (h * w), r = pix - b * h * w, y = r / w, x = r - y * w;
      cx[p * DP + d] = __ldg(code + b * sb + y * sh + x * sw + d);
    }
  } else {
    for (int i = threadIdx.x; i < n * D; i += blockDim.x) {
      const int d = i / n, p = i - d * n, pix = p0 + p;
      const int b = pix / (h * w), r = pix - b * h * w, y = r / w, x = r - y * w;
      cx[p * DP + d] = __ldg(code + b * sb + d * sc + y * sh + x * sw);
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < K * D + K; o += blockDim.x) {
    float acc = 0.f;
    if (o < K * D) {
      const int k = o / D, d = o - k * D;
      for (int p = 0; p < n; ++p) acc = fmaf(dl[p * 33 + k], cx[p * DP + d], acc);
      atomicAdd(dweight + o, acc * inv);
    } else if (dbias) {
      const int k = o - K * D;
      for (int p = 0; p < n; ++p) acc += dl[p * 33 + k];
      atomicAdd(dbias + k, acc * inv);
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
                                                            float* __restrict__ dchat /*[N][D] sums*/,
                                                            double* __restrict__ loss_sum) {
  extern __shared__ float sm[];
  const int DP = D | 1;
  float* chat = sm;                  // [N][DP] normalised centres
  float* dacc = chat + N * DP;       // [N][D]  sum of x_hat per assigned centre (mode 0)
  float* xs = dacc + N * D;          // [8][128] per-warp normalised pixel
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
  float* x = xs + warp * 128;
  const int nn = lane < N ? lane : 0;
  float lacc = 0.f;
  for (int pix = blockIdx.x * 8 + warp; pix < npix; pix += gridDim.x * 8) {
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
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[t] = v[t] / nrm;
      if (lane + 32 * t < D) x[lane + 32 * t] = v[t];
    }
    __syncwarp();
    float inner = 0.f;
    for (int d = 0; d < D; ++d) inner = fmaf(x[d], chat[nn * DP + d], inner);
    if (lane >= N) inner = -INFINITY;
    if (mode == 0) {
      const float m = warp_max(inner);
      const unsigned ball = __ballot_sync(0xffffffffu, inner == m && lane < N);
      const int am = ball ? __ffs(ball) - 1 : 0;  // first maximum, like torch.argmax
      if (out_probs && lane < N) out_probs[(size_t)pix * N + lane] = lane == am ? 1.f : 0.f;
      if (lane == 0) lacc -= m;
      if (dchat) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (lane + 32 * t < D) atomicAdd(dacc + am * D + lane + 32 * t, v[t]);
      }
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
  if (lane == 0 && mode != 2) atomicAdd(loss_sum, (double)lacc);
  if (mode == 0 && dchat) {
    __syncthreads();
    for (int i = threadIdx.x; i < N * D; i += blockDim.x) {
      const float a = dacc[i];
      if (a != 0.f) atomicAdd(dchat + i, a);
    }
  }
}

// loss = sum/M; centre gradient through F.normalize:  dc = (g - c_hat <c_hat,g>) / max(|c|,eps),  g = -S_n / M
__global__ void __launch_bounds__(32) cluster_finalize_kernel(const float* __restrict__ clusters, int N, int D,
                                                              int npix, const float* __restrict__ dchat,
                                                              const double* __restrict__ loss_sum,
                                                              float* __restrict__ loss_out,
                                                              float* __restrict__ dclusters) {
  const int n = blockIdx.x, lane = threadIdx.x;
  if (n == 0 && lane == 0 && loss_out) loss_out[0] = (float)(*loss_sum / (double)npix);
  if (!dclusters) return;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = clusters[n * D + d];
    ss = fmaf(v, v, ss);
  }
  const float raw = sqrtf(warp_sum(ss));
  const float nrm = fmaxf(raw, 1e-12f);
  const float sgn = -1.f / (float)npix;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(clusters[n * D + d] / nrm, dchat[n * D + d] * sgn, dot);
  dot = warp_sum(dot);
  for (int d = lane; d < D; d += 32) {
    const float g = dchat[n * D + d] * sgn;
    // below eps the reference divides by the constant eps: no projection term
    dclusters[n * D + d] = raw > 1e-12f ? (g - (clusters[n * D + d] / nrm) * dot) / nrm : g / nrm;
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace dg

using namespace dg;

extern "C" size_t dg_probe_workspace_bytes(int B, int h, int w, int D, int K) {
  if (B <= 0 || h <= 0 || w <= 0 || D <= 0 || K <= 0) return 0;
  const size_t npix = (size_t)B * h * w;
  return 2 * align256(npix * KP * sizeof(float)) + align256((size_t)K * D * sizeof(float)) + 256;
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
  p += align256((size_t)K * D * sizeof(float));
  double* loss_sum = reinterpret_cast<double*>(p);
  unsigned long long* count = reinterpret_cast<unsigned long long*>(p + 8);
  // dlogits .. count are one contiguous span
  DG_CUDA_OK(cudaMemsetAsync(dlogits, 0, (size_t)(p + 16 - reinterpret_cast<char*>(dlogits)), st));
  if (dweight) DG_CUDA_OK(cudaMemsetAsync(dweight, 0, (size_t)K * D * sizeof(float), st));
  if (dbias) DG_CUDA_OK(cudaMemsetAsync(dbias, 0, (size_t)K * sizeof(float), st));

  DG_PRE(st);
  probe_logits_kernel<<<ceil_div(npix, 8), 256, (size_t)K * (D | 1) * sizeof(float), st>>>(
      code, strides[0], strides[1], strides[2], strides[3], D, h, w, npix, weight, bias, K, logits);
  DG_LAUNCH_OK("probe_logits_kernel");

  static bool attr_done = false;
  const size_t ce_smem = 8 * (size_t)(KP * 4 + 64 * 33 + 64 * 4) * sizeof(float);
  if (!attr_done) {
    DG_CUDA_OK(cudaFuncSetAttribute(probe_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ce_smem));
    attr_done = true;
  }
  // ATen: scale = input_size / output_size in float (area_pixel_compute_scale, align_corners=False)
  const float scale_y = (float)h / (float)Hl, scale_x = (float)w / (float)Wl;
  DG_PRE(st);
  probe_ce_kernel<<<ceil_div(npix, 8), 256, ce_smem, st>>>(logits, labels, label_strides[0], label_strides[1],
                                                          label_strides[2], npix, h, w, Hl, Wl, K, scale_y, scale_x,
                                                          dlogits, loss_sum, count);
  DG_LAUNCH_OK("probe_ce_kernel");

  const size_t wg_smem = (size_t)(64 * 33 + 64 * (D | 1)) * sizeof(float);
  DG_PRE(st);
  probe_wgrad_kernel<<<dweight ? ceil_div(npix, 64) : 1, 256, wg_smem, st>>>(
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
  float* dchat = reinterpret_cast<float*>(p);
  p += align256((size_t)N * D * sizeof(float));
  double* loss_sum = reinterpret_cast<double*>(p);
  DG_CUDA_OK(cudaMemsetAsync(dchat, 0, (size_t)(p + 16 - reinterpret_cast<char*>(dchat)), st));
  const size_t smem = (size_t)(N * (D | 1) + N * D + 8 * 128) * sizeof(float);
  const int grid = min(ceil_div(npix, 8), 148 * 4);
  DG_PRE(st);
  cluster_probe_kernel<<<grid, 256, smem, st>>>(code, strides[0], strides[1], strides[2], strides[3], D, h, w, npix,
                                                clusters, N, mode, alpha, probs_out, dclusters ? dchat : nullptr,
                                                loss_sum);
  DG_LAUNCH_OK("cluster_probe_kernel");
  if (mode != 2 && (loss_out || dclusters)) {
    DG_PRE(st);
    cluster_finalize_kernel<<<dclusters ? N : 1, 32, 0, st>>>(clusters, N, D, npix, dchat, loss_sum, loss_out,
                                                               dclusters);
    DG_LAUNCH_OK("cluster_finalize_kernel");
  }
  return DG_OK;
}
