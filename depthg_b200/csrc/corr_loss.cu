// Fused correlation loss: forward values + unit gradients, no fd/cd tensor in HBM.
//
// Replaces ContrastiveCorrelationLoss.helper for every pair (intra, inter, each
// negative) and depth_feature_correlation (/root/reference/src/modules.py:1231-1278)
// together with the einsum of tensor_correlation (:797-809).
//
// Per (pair k, image b, 64-row block of sample points p) one CTA walks the 64-wide
// blocks of second-operand points q and, per 64x64 tile:
//   fd = F1n[p,:] . F2n[q,:]   (K = C, backbone features, no gradient)
//   cd = C1n[p,:] . C2n[q,:]   (K = D, code, gradient flows)
//   fd' = fd - rowmean[p] + old_mean - shift            (pointwise centring)
//   loss = -clamp(cd) * fd'    summed;      U = -fd' * 1[clamp passes] / (B P^2)
//   dC1[p,:] += U . C2n[q,:]    dC2[q,:] += U^T . C1n[p,:]
// and for the intra pair the depth term with dd = s[p] s[q] in place of fd'.
// rowmean / old_mean come from the panel mean rows (exactly mean_q fd[p,q] and
// its mean over images) so no second pass over fd is needed; the reference's extra
// "- fd.mean()" after centring is float noise (~1e-10) and is dropped.
// This generic-shape kernel computes on the fp32 CUDA cores.
#include "kernels.cuh"

namespace dg {

constexpr int TM = 64, TN = 64, KC = 16, CORR_THREADS = 256;
constexpr int ASTR = TM + 4;  // smem row pitch of the k-major operand tiles (16-byte aligned rows)
constexpr int USTR = TN + 1;

struct CorrParams {
  const float* fn;
  const float* cn;
  const float* dsign;
  const float* rowmean;  // [npairs,B,Prows]
  const float* bsum;     // [npairs,B]
  int npairs, B, P, Prows, ldf, ldc, flags, has_depth;
  float depth_shift, inv_cnt;
  float shift[DG_MAX_PAIRS];
  int32_t group[DG_MAX_PAIRS];
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];  // feature panel slots of pair k's operands (default 0 and k)
  float* dC1;
  float* dC2;
  float* partials;  // [npairs*B*n_pt][4]
  float* cd_out;
  float* loss_out;
  float* dd_out;
  float* out8;
};

// rowmean[k,b,p] = <F1n[b,p,:], mean_q F2n[k,b,q,:]>, bsum[k,b] = sum_{p<P} rowmean
struct SlotMapS {
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];
};

__global__ void __launch_bounds__(256) pair_means_kernel(const float* __restrict__ fn, const float* __restrict__ fmean,
                                                         int nsplit, int B, int P, int Prows, int ldf,
                                                         float* __restrict__ rowmean, float* __restrict__ bsum,
                                                         const __grid_constant__ SlotMapS sm) {
  extern __shared__ float mv[];  // [ldf]
  __shared__ float wsum[8];
  const int k = blockIdx.x / B, b = blockIdx.x - k * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* m = fmean + ((size_t)sm.fs2[k] * B + b) * nsplit * ldf;  // nsplit partial means of the second operand
  for (int c = threadIdx.x; c < ldf; c += blockDim.x) {
    float a = 0.f;
    for (int i = 0; i < nsplit; ++i) a += m[(size_t)i * ldf + c];
    mv[c] = a;
  }
  __syncthreads();
  const float* F1 = fn + ((size_t)sm.fs1[k] * B + b) * Prows * ldf;
  float* rm = rowmean + ((size_t)k * B + b) * Prows;
  float acc = 0.f;
  for (int p = warp; p < Prows; p += 8) {
    float s = 0.f;
    if (p < P) {
      const float4* row = reinterpret_cast<const float4*>(F1 + (size_t)p * ldf);
      for (int c4 = lane; c4 < ldf / 4; c4 += 32) {
        const float4 v = __ldg(row + c4);
        const float4 u = *reinterpret_cast<const float4*>(mv + 4 * c4);
        s += v.x * u.x + v.y * u.y + v.z * u.z + v.w * u.w;
      }
      s = warp_sum(s);
    }
    if (lane == 0) rm[p] = s;
    acc += s;
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    bsum[blockIdx.x] = t;
  }
}

// acc[i][j] (+)= sum_x Us-like products; see call sites.
__device__ __forceinline__ void grad_products(const float* __restrict__ Us, const float* __restrict__ C1s,
                                              const float* __restrict__ C2s, int ldcs, int R, float* __restrict__ d1,
                                              float* __restrict__ d2, int ldc, bool first_q, int warp, int lane) {
  float acc[8][4];
  // dC1[p0 + 8w + i][d] = sum_q U[8w+i][q] * C2[q][d]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int q = 0; q < TN; ++q) {
    float c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (j < R) ? C2s[q * ldcs + lane + 32 * j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float u = Us[(8 * warp + i) * USTR + q];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(u, c[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < R) {
        float* dst = d1 + (size_t)(8 * warp + i) * ldc + lane + 32 * j;
        *dst = first_q ? acc[i][j] : (*dst + acc[i][j]);
      }
  // dC2[q0 + 8w + i][d] += sum_p U[p][8w+i] * C1[p][d]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int p = 0; p < TM; ++p) {
    float c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (j < R) ? C1s[p * ldcs + lane + 32 * j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float u = Us[p * USTR + 8 * warp + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(u, c[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < R) atomicAdd(d2 + (size_t)(8 * warp + i) * ldc + lane + 32 * j, acc[i][j]);
}

__global__ void __launch_bounds__(CORR_THREADS) corr_tile_kernel(const __grid_constant__ CorrParams prm) {
  extern __shared__ __align__(16) float csm[];
  const int ldc = prm.ldc, ldcs = ldc + 1, R = ldc / 32;
  float* As = csm;                    // [KC][ASTR]
  float* Bs = As + KC * ASTR;         // [KC][ASTR]
  float* Us = Bs + KC * ASTR;         // [TM][USTR]
  float* C1s = Us + TM * USTR;        // [TM][ldcs]
  float* C2s = C1s + TM * ldcs;       // [TN][ldcs]
  __shared__ float s_red[8][4];
  __shared__ float s_oldmean;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int pt = blockIdx.x, kb = blockIdx.y;
  const int k = kb / prm.B, b = kb - k * prm.B;
  const int P = prm.P, Prows = prm.Prows, ldf = prm.ldf;
  const int p0 = pt * TM;
  const int n_qt = Prows / TN;

  const float* F1 = prm.fn + (((size_t)prm.fs1[k] * prm.B + b) * Prows + p0) * ldf;
  const float* F2base = prm.fn + ((size_t)prm.fs2[k] * prm.B + b) * Prows * ldf;
  const float* C1 = prm.cn + ((size_t)b * Prows + p0) * ldc;
  const float* C2base = prm.cn + ((size_t)k * prm.B + b) * Prows * ldc;
  const bool pointwise = prm.flags & DG_FLAG_POINTWISE;
  const bool depth_pair = prm.has_depth && k == 0;
  const float lo = (prm.flags & DG_FLAG_ZERO_CLAMP) ? 0.f : -9999.f;
  const float hi = (prm.flags & DG_FLAG_STABALIZE) ? 0.8f : __int_as_float(0x7f800000);
  const float shift = prm.shift[k];

  if (tid == 0) {
    float om = 0.f;
    if (pointwise) {
      for (int bb = 0; bb < prm.B; ++bb) om += prm.bsum[(size_t)k * prm.B + bb];
      om /= (float)prm.B * (float)P;
    }
    s_oldmean = om;
  }
  for (int i = tid; i < TM * ldc; i += CORR_THREADS) {
    const int r = i / ldc, d = i - r * ldc;
    C1s[r * ldcs + d] = __ldg(C1 + (size_t)r * ldc + d);
  }
  float rmean[4], sp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i;
    rmean[i] = pointwise ? __ldg(prm.rowmean + ((size_t)k * prm.B + b) * Prows + p) : 0.f;
    sp[i] = depth_pair ? __ldg(prm.dsign + (size_t)b * Prows + p) : 0.f;
  }
  __syncthreads();
  const float old_mean = s_oldmean;

  float sum_loss = 0.f, sum_cd = 0.f, sum_dloss = 0.f, sum_dd = 0.f;

  for (int qt = 0; qt < n_qt; ++qt) {
    const int q0 = qt * TN;
    const float* F2 = F2base + (size_t)q0 * ldf;
    const float* C2 = C2base + (size_t)q0 * ldc;
    for (int i = tid; i < TN * ldc; i += CORR_THREADS) {
      const int r = i / ldc, d = i - r * ldc;
      C2s[r * ldcs + d] = __ldg(C2 + (size_t)r * ldc + d);
    }
    // ---- fd tile: K = ldf (padded columns are zero) ----
    float fd[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) fd[i][j] = 0.f;
    {
      const int lr = tid >> 2, lk = (tid & 3) * 4;
      const float* ga = F1 + (size_t)lr * ldf + lk;
      const float* gb = F2 + (size_t)lr * ldf + lk;
      float4 va = __ldg(reinterpret_cast<const float4*>(ga));
      float4 vb = __ldg(reinterpret_cast<const float4*>(gb));
      for (int k0 = 0; k0 < ldf; k0 += KC) {
        __syncthreads();  // previous chunk fully consumed
        As[(lk + 0) * ASTR + lr] = va.x; As[(lk + 1) * ASTR + lr] = va.y;
        As[(lk + 2) * ASTR + lr] = va.z; As[(lk + 3) * ASTR + lr] = va.w;
        Bs[(lk + 0) * ASTR + lr] = vb.x; Bs[(lk + 1) * ASTR + lr] = vb.y;
        Bs[(lk + 2) * ASTR + lr] = vb.z; Bs[(lk + 3) * ASTR + lr] = vb.w;
        __syncthreads();
        if (k0 + KC < ldf) {  // prefetch the next chunk while computing this one
          va = __ldg(reinterpret_cast<const float4*>(ga + k0 + KC));
          vb = __ldg(reinterpret_cast<const float4*>(gb + k0 + KC));
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(As + kk * ASTR + ty * 4);
          float bv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = Bs[kk * ASTR + tx + 16 * j];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            fd[0][j] = fmaf(a.x, bv[j], fd[0][j]);
            fd[1][j] = fmaf(a.y, bv[j], fd[1][j]);
            fd[2][j] = fmaf(a.z, bv[j], fd[2][j]);
            fd[3][j] = fmaf(a.w, bv[j], fd[3][j]);
          }
        }
      }
    }
    // ---- cd tile from the resident code tiles ----
    float cd[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) cd[i][j] = 0.f;
    __syncthreads();  // C2s (and C1s on the first pass) visible
    for (int d = 0; d < ldc; ++d) {
      float a[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = C1s[(ty * 4 + i) * ldcs + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = C2s[(tx + 16 * j) * ldcs + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cd[i][j] = fmaf(a[i], bv[j], cd[i][j]);
    }
    // ---- epilogue ----
    float sq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sq[j] = depth_pair ? __ldg(prm.dsign + (size_t)b * Prows + q0 + tx + 16 * j) : 0.f;
    float ud[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + tx + 16 * j;
        const bool valid = (p < P) && (q < P);
        const float c = cd[i][j];
        const float cl = fminf(fmaxf(c, lo), hi);
        const bool pass = valid && (c >= lo) && (c <= hi);
        const float f = fd[i][j] - rmean[i] + old_mean - shift;
        const float l = -cl * f;
        float u = pass ? -f * prm.inv_cnt : 0.f;
        if (valid) {
          sum_loss += l;
          sum_cd += c;
          const size_t o = (((size_t)k * prm.B + b) * P + p) * P + q;
          if (prm.cd_out) prm.cd_out[o] = c;
          if (prm.loss_out) prm.loss_out[o] = l;
        }
        Us[(ty * 4 + i) * USTR + tx + 16 * j] = u;
        ud[i][j] = 0.f;
        if (depth_pair) {
          const float dd = sp[i] * sq[j];
          const float g = dd - prm.depth_shift;
          if (valid) {
            sum_dloss += -cl * g;
            sum_dd += dd;
            if (prm.dd_out) prm.dd_out[((size_t)b * P + p) * P + q] = dd;
          }
          ud[i][j] = pass ? -g * prm.inv_cnt : 0.f;
        }
      }
    }
    __syncthreads();
    {
      const size_t slab = (size_t)prm.B * Prows * ldc;
      float* d1 = prm.dC1 + (size_t)k * slab + ((size_t)b * Prows + p0) * ldc;
      float* d2 = prm.dC2 + (size_t)k * slab + ((size_t)b * Prows + q0) * ldc;
      grad_products(Us, C1s, C2s, ldcs, R, d1, d2, ldc, qt == 0, warp, lane);
      if (depth_pair) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) Us[(ty * 4 + i) * USTR + tx + 16 * j] = ud[i][j];
        __syncthreads();
        float* e1 = prm.dC1 + (size_t)prm.npairs * slab + ((size_t)b * Prows + p0) * ldc;
        float* e2 = prm.dC2 + (size_t)prm.npairs * slab + ((size_t)b * Prows + q0) * ldc;
        grad_products(Us, C1s, C2s, ldcs, R, e1, e2, ldc, qt == 0, warp, lane);
      }
    }
    __syncthreads();  // Us / C2s free for the next q block
  }

  // ---- deterministic block reduction of the four sums ----
  sum_loss = warp_sum(sum_loss);
  sum_cd = warp_sum(sum_cd);
  sum_dloss = warp_sum(sum_dloss);
  sum_dd = warp_sum(sum_dd);
  if (lane == 0) {
    s_red[warp][0] = sum_loss;
    s_red[warp][1] = sum_cd;
    s_red[warp][2] = sum_dloss;
    s_red[warp][3] = sum_dd;
  }
  __syncthreads();
  if (tid < 4) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][tid];
    prm.partials[((size_t)kb * gridDim.x + pt) * 4 + tid] = t;
  }
}

struct FinalizeParams {
  int npairs, B, P, has_depth, n_pt;
  int32_t group[DG_MAX_PAIRS];
};

// One block: fold the per-CTA partial sums into the 8 scalars of the output tuple.
__global__ void __launch_bounds__(256) corr_finalize_kernel(const float* __restrict__ partials,
                                                            const int* __restrict__ err, float* __restrict__ out8,
                                                            const __grid_constant__ FinalizeParams prm) {
  __shared__ float red[8][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int per_pair = prm.B * prm.n_pt;
  const int total = prm.npairs * per_pair;
  for (int e = tid; e < total; e += 256) {
    const int k = e / per_pair;
    const int g = prm.group[k];
    const float* s = partials + (size_t)e * 4;
    const float l = s[0], c = s[1];
    if (g == DG_GROUP_INTRA) { acc[0] += l; acc[1] += c; }
    else if (g == DG_GROUP_INTER) { acc[2] += l; acc[3] += c; }
    else { acc[4] += l; acc[5] += c; }
    if (k == 0) { acc[6] += s[2]; acc[7] += s[3]; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (tid < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][tid];
    int cnt[3] = {0, 0, 0};
    for (int k = 0; k < prm.npairs; ++k) cnt[prm.group[k] > 2 ? 2 : prm.group[k]]++;
    const float elems = (float)prm.B * (float)prm.P * (float)prm.P;
    const int g = tid >> 1;
    const float n = (g < 3) ? (float)cnt[g] : (prm.has_depth ? 1.f : 0.f);
    float r = n > 0.f ? t / (n * elems) : 0.f;
    if (err && *err != 0) r = __int_as_float(0x7fc00000);  // a pipeline wait timed out: poison the result
    out8[tid] = r;
  }
}

int launch_corr_finalize(const float* partials, int npairs, int B, int P, const int32_t* group, int has_depth,
                         const int* err, float* out8, int n_pt, cudaStream_t st) {
  FinalizeParams fp;
  fp.npairs = npairs; fp.B = B; fp.P = P; fp.has_depth = has_depth; fp.n_pt = n_pt;
  for (int k = 0; k < npairs; ++k) fp.group[k] = group[k];
  DG_PRE(st);
  corr_finalize_kernel<<<1, 256, 0, st>>>(partials, err, out8, fp);
  DG_LAUNCH_OK("corr_finalize_kernel");
  return DG_OK;
}

size_t corr_workspace_bytes(int npairs, int B, int P) {
  if (npairs <= 0 || B <= 0 || P <= 0) return 0;
  const size_t Prows = (size_t)round_up(P, 64);
  const size_t n_pt = Prows / TM;
  size_t floats = (size_t)npairs * B * Prows + (size_t)npairs * B + (size_t)npairs * B * n_pt * 4;
  // the tcgen05 path lays out [header][dots][partials per (pair, image, row tile, column group)][row means (dense)]
  const size_t umma = (size_t)npairs * B * (1 + (size_t)ceil_div(P, 128) * ceil_div(P, 256) * 4 + round_up(P, 256));
  if (umma > floats) floats = umma;
  const size_t pipe = corr_pipe_workspace_floats(npairs, B, P);
  if (pipe > floats) floats = pipe;
  size_t bytes = floats * sizeof(float) + 2048;  // + header and 256-byte alignment slack of each region
  return (bytes + 255) / 256 * 256;
}

int corr_loss_simt(const float* fn, const float* cn, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P,
                   int Prows, int ldf, int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift,
                   int flags, float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out,
                   void* ws, cudaStream_t st, const int32_t* fslot1, const int32_t* fslot2) {
  const int n_pt = Prows / TM;
  CorrParams prm;
  prm.fn = fn;
  prm.cn = cn;
  prm.dsign = dsign;
  float* wsf = static_cast<float*>(ws);
  float* rowmean = wsf;
  float* bsum = rowmean + (size_t)npairs * B * Prows;
  prm.rowmean = rowmean;
  prm.bsum = bsum;
  prm.partials = bsum + (size_t)npairs * B;
  prm.npairs = npairs; prm.B = B; prm.P = P; prm.Prows = Prows; prm.ldf = ldf; prm.ldc = ldc; prm.flags = flags;
  prm.has_depth = dsign != nullptr;
  prm.depth_shift = depth_shift;
  prm.inv_cnt = 1.0f / ((float)B * (float)P * (float)P);
  for (int k = 0; k < npairs; ++k) {
    prm.shift[k] = pair_shift[k];
    prm.group[k] = pair_group[k];
    prm.fs1[k] = fslot1 ? fslot1[k] : 0;
    prm.fs2[k] = fslot2 ? fslot2[k] : k;
  }
  prm.dC1 = dC1; prm.dC2 = dC2; prm.cd_out = cd_out; prm.loss_out = loss_out; prm.dd_out = dd_out; prm.out8 = out8;

  if (flags & DG_FLAG_POINTWISE) {
    DG_PRE(st);
    SlotMapS sm;
    for (int k = 0; k < npairs; ++k) { sm.fs1[k] = prm.fs1[k]; sm.fs2[k] = prm.fs2[k]; }
    pair_means_kernel<<<npairs * B, 256, (size_t)ldf * sizeof(float), st>>>(fn, fmean, nsplit, B, P, Prows, ldf, rowmean, bsum,
                                                                            sm);
    DG_LAUNCH_OK("pair_means_kernel");
  }
  const size_t slab = (size_t)B * Prows * ldc * sizeof(float);
  DG_CUDA_OK(cudaMemsetAsync(dC2, 0, slab * (npairs + 1), st));
  const size_t smem = ((size_t)2 * KC * ASTR + (size_t)TM * USTR + (size_t)2 * TM * (ldc + 1)) * sizeof(float);
  static PerDevice configured_pd = {};
  size_t& configured = per_device(configured_pd);
  if (smem > configured) {
    DG_CUDA_OK(cudaFuncSetAttribute(corr_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  DG_PRE(st);
  corr_tile_kernel<<<dim3(n_pt, npairs * B), CORR_THREADS, smem, st>>>(prm);
  DG_LAUNCH_OK("corr_tile_kernel");
  return launch_corr_finalize(prm.partials, npairs, B, P, pair_group, prm.has_depth, nullptr, out8, n_pt, st);
}

}  // namespace dg

extern "C" size_t dg_corr_loss_workspace_bytes(int npairs, int B, int P) { return dg::corr_workspace_bytes(npairs, B, P); }

extern "C" int dg_corr_loss(const dg_panels_t* pan, const float* fmean, const float* dsign, int npairs, int B, int P,
                            int Prows, int C, int ldf, int D, int ldc, const float* pair_shift,
                            const int32_t* pair_group, float depth_shift, int flags, float* out8, float* dC1,
                            float* dC2, float* cd_out, float* loss_out, float* dd_out, float* fd_dbg, void* ws,
                            size_t ws_bytes, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(pan && pan->f_hi && pan->c_hi && pair_shift && pair_group && out8 && dC1 && dC2 && ws, DG_ERR_INVALID,
             "dg_corr_loss: null pointer");
  DG_REQUIRE(npairs > 0 && npairs <= DG_MAX_PAIRS, DG_ERR_INVALID, "dg_corr_loss: npairs=%d out of range", npairs);
  DG_REQUIRE(B > 0 && P > 0 && C > 0 && D > 0, DG_ERR_INVALID, "dg_corr_loss: bad sizes");
  DG_REQUIRE(ldf >= C && ldf % 32 == 0 && ldc >= D && ldc % 32 == 0, DG_ERR_INVALID, "dg_corr_loss: bad panel pitch");
  DG_REQUIRE(ldc <= 128, DG_ERR_UNSUPPORTED, "dg_corr_loss: code dim %d > 128 not supported", D);
  DG_REQUIRE(!(flags & DG_FLAG_POINTWISE) || fmean, DG_ERR_INVALID, "dg_corr_loss: pointwise needs fmean");
  DG_REQUIRE(ws_bytes >= dg_corr_loss_workspace_bytes(npairs, B, P), DG_ERR_WORKSPACE,
             "dg_corr_loss: workspace too small");
  DG_REQUIRE(pair_group[0] == DG_GROUP_INTRA, DG_ERR_INVALID, "dg_corr_loss: pair 0 must be the intra pair");
  for (int k = 0; k < npairs; ++k)
    DG_REQUIRE(pair_group[k] >= 0 && pair_group[k] <= DG_GROUP_NEG, DG_ERR_INVALID, "dg_corr_loss: bad pair group");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  if (pan->format == DG_PANEL_FEATS_SPLIT || pan->format == DG_PANEL_CODE_SPLIT) {
    DG_REQUIRE(P <= 1024 && Prows == round_up(P, 128), DG_ERR_UNSUPPORTED,
               "dg_corr_loss: the tcgen05 path needs S*S <= 1024 and panels of round_up(S*S,128) rows (got P=%d, Prows=%d)",
               P, Prows);
    DG_REQUIRE(pan->cb_hi, DG_ERR_INVALID, "dg_corr_loss: split panel format needs the interleaved 16-bit code panel cb_hi");
    return corr_loss_pipe(pan, fmean, 1, dsign, npairs, B, P, Prows, ldf, ldc, pair_shift, pair_group, depth_shift, flags, out8,
                          dC1, dC2, cd_out, loss_out, dd_out, fd_dbg, ws, st, nullptr, nullptr, 0);
  }
  DG_REQUIRE(pan->format == DG_PANEL_F32, DG_ERR_INVALID, "dg_corr_loss: unknown panel format %d", pan->format);
  DG_REQUIRE(Prows == round_up(P, 64), DG_ERR_INVALID, "dg_corr_loss: Prows must be dg_panel_rows(P)");
  return corr_loss_simt(static_cast<const float*>(pan->f_hi), static_cast<const float*>(pan->c_hi), fmean, 1, dsign, npairs,
                        B, P, Prows, ldf, ldc, pair_shift, pair_group, depth_shift, flags, out8, dC1, dC2, cd_out,
                        loss_out, dd_out, ws, st, nullptr, nullptr);
}
