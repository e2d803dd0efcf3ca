// Cosine-similarity k-NN build: similarity GEMM fused with a running per-row top-k.
//
// Replaces the chunked  einsum("nf,mf->nm") + torch.topk(sims, 30)  loop of
// /root/reference/src/precompute_knns.py:99-113, which materialises a
// [775, 49629] similarity block per chunk on the CPU.  Here a CTA owns 64 query
// rows, streams the database in 64-row tiles, forms each 64x64 similarity tile
// in registers (fp32 FMA, so indices match the fp32 reference) and merges it
// into per-row sorted top-k lists that live in warp registers (one list entry per
// lane, insertion by ballot + shuffle).  The similarity matrix never exists.
#include <stdlib.h>

#include "kernels.cuh"

namespace dg {

constexpr int KTM = 64, KTN = 64, KKC = 16, KNN_THREADS = 256;
constexpr int KASTR = KTM + 4;
constexpr int SSTR = KTN + 1;
constexpr int FEW_CAP = 8192;  // up to this many uncertified rows go to knn_fewrows_kernel, more to the tiled kernel

__global__ void __launch_bounds__(KNN_THREADS) knn_topk_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                               int Nq, int N, int F, int k, int64_t* __restrict__ idx,
                                                               float* __restrict__ sims,
                                                               const int* __restrict__ row_list,
                                                               const int* __restrict__ row_count) {
  // row_list != null: only the query rows listed there (the ones the tensor-core pass could not certify)
  const int nrows = row_list ? min(*row_count, Nq) : Nq;
  if (blockIdx.x * KTM >= nrows) return;
  if (row_list && nrows <= FEW_CAP) return;  // few rows: knn_fewrows_kernel does them
  __shared__ __align__(16) float As[KKC * KASTR];
  __shared__ __align__(16) float Bs[KKC * KASTR];
  __shared__ float Ss[KTM * SSTR];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * KTM;

  // per-warp state: 8 rows x one sorted (descending) list spread over the 32 lanes
  float tv[8];
  int ti[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    tv[i] = -INFINITY;
    ti[i] = -1;
  }

  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int qsel = min(r0 + lr, nrows - 1);
  const int qrow = row_list ? row_list[qsel] : qsel;
  const float* ga = q + (size_t)qrow * F + lk;
  const bool vec = (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(db)) & 15) == 0;

  for (int c0 = 0; c0 < N; c0 += KTN) {
    const int drow = min(c0 + lr, N - 1);
    const float* gb = db + (size_t)drow * F + lk;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < F; k0 += KKC) {
      float va[4], vb[4];
      if (vec && k0 + lk + 3 < F) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(ga + k0));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(gb + k0));
        va[0] = a4.x; va[1] = a4.y; va[2] = a4.z; va[3] = a4.w;
        vb[0] = b4.x; vb[1] = b4.y; vb[2] = b4.z; vb[3] = b4.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool in = k0 + lk + e < F;
          va[e] = in ? __ldg(ga + k0 + e) : 0.f;
          vb[e] = in ? __ldg(gb + k0 + e) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        As[(lk + e) * KASTR + lr] = va[e];
        Bs[(lk + e) * KASTR + lr] = vb[e];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KKC; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(As + kk * KASTR + ty * 4);
        float bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[kk * KASTR + tx + 16 * j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] = fmaf(a.x, bv[j], acc[0][j]);
          acc[1][j] = fmaf(a.y, bv[j], acc[1][j]);
          acc[2][j] = fmaf(a.z, bv[j], acc[2][j]);
          acc[3][j] = fmaf(a.w, bv[j], acc[3][j]);
        }
      }
    }
    __syncthreads();  // previous tile's Ss fully merged
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ss[(ty * 4 + i) * SSTR + tx + 16 * j] = acc[i][j];
    __syncthreads();
    // merge: warp w owns rows 8w..8w+7; each lane looks at two columns of the tile
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 8 * warp + i;
      const float thr = __shfl_sync(0xffffffffu, tv[i], k - 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = lane + 32 * half;
        const float v = Ss[row * SSTR + col];
        const bool cand = (c0 + col < N) && (v > thr);
        unsigned m = __ballot_sync(0xffffffffu, cand);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float nv = __shfl_sync(0xffffffffu, v, src);
          const int ni = c0 + 32 * half + src;
          // still above the (possibly raised) k-th value?
          if (nv > __shfl_sync(0xffffffffu, tv[i], k - 1)) {
            const int pos = __popc(__ballot_sync(0xffffffffu, tv[i] >= nv));
            const float up_v = __shfl_up_sync(0xffffffffu, tv[i], 1);
            const int up_i = __shfl_up_sync(0xffffffffu, ti[i], 1);
            if (lane == pos) {
              tv[i] = nv;
              ti[i] = ni;
            } else if (lane > pos) {
              tv[i] = up_v;
              ti[i] = up_i;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int sel = r0 + 8 * warp + i;
    if (sel < nrows && lane < k) {
      const int row = row_list ? row_list[sel] : sel;
      idx[(size_t)row * k + lane] = (int64_t)ti[i];
      if (sims) sims[(size_t)row * k + lane] = tv[i];
    }
  }
}

// Exact top-k for a FEW query rows (the ones the tensor-core pass could not certify): 8 rows per CTA so that even a
// handful of rows spreads over many SMs; each warp streams database rows (coalesced, lanes over F) against the 8 query
// rows held in shared memory, keeps per-row sorted lists in its lanes, and the 8 warps' lists are merged at the end.
constexpr int FEW_ROWS = 8, FEW_THREADS = 256, FEW_WARPS = 8;

__global__ void __launch_bounds__(FEW_THREADS) knn_fewrows_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                                  int Nq, int N, int F, int k, int64_t* __restrict__ idx,
                                                                  float* __restrict__ sims, const int* __restrict__ row_list,
                                                                  const int* __restrict__ row_count) {
  extern __shared__ float fsm[];  // [FEW_ROWS][F] query rows, then [FEW_ROWS][FEW_WARPS][32] merge buffers (val, idx)
  const int nrows = min(*row_count, Nq);
  if (nrows > FEW_CAP) return;
  float* qs = fsm;
  float* mv = qs + (size_t)FEW_ROWS * F;
  int* mi = reinterpret_cast<int*>(mv + FEW_ROWS * FEW_WARPS * 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows per CTA: as few as it takes to give every CTA of the grid something to do (a handful of
  // uncertified rows then spreads over the whole GPU instead of a few SMs), at most FEW_ROWS
  const int rpb = max(1, min(FEW_ROWS, (nrows + (int)gridDim.x - 1) / (int)gridDim.x));
  for (int g0 = blockIdx.x * rpb; g0 < nrows; g0 += gridDim.x * rpb) {
    __syncthreads();
    for (int i = threadIdx.x; i < rpb * F; i += FEW_THREADS) {
      const int r = i / F, f = i - r * F;
      const int sel = min(g0 + r, nrows - 1);
      qs[i] = __ldg(q + (size_t)row_list[sel] * F + f);
    }
    __syncthreads();
    float tv[FEW_ROWS];
    int ti[FEW_ROWS];
#pragma unroll
    for (int r = 0; r < FEW_ROWS; ++r) {
      tv[r] = -INFINITY;
      ti[r] = -1;
    }
    const bool vec = (F % 128 == 0) && F <= 1024 && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    const int nv = F / 128;
    for (int n0 = warp; n0 < N; n0 += FEW_WARPS * 4) {  // four database rows per pass: 4 x nv 128-bit loads in flight
      float s4[4][FEW_ROWS];
      if (vec) {
        float4 d[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = min(n0 + u * FEW_WARPS, N - 1);
          const float4* dr = reinterpret_cast<const float4*>(db + (size_t)n * F) + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < nv) d[u][j] = __ldg(dr + 32 * j);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int r = 0; r < FEW_ROWS; ++r) s4[u][r] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nv) {
#pragma unroll
            for (int r = 0; r < FEW_ROWS; ++r) {
              if (r >= rpb) break;
              const float4 qv = *reinterpret_cast<const float4*>(qs + (size_t)r * F + 4 * lane + 128 * j);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                s4[u][r] += qv.x * d[u][j].x + qv.y * d[u][j].y + qv.z * d[u][j].z + qv.w * d[u][j].w;
            }
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = min(n0 + u * FEW_WARPS, N - 1);
          const float* dr = db + (size_t)n * F;
#pragma unroll
          for (int r = 0; r < FEW_ROWS; ++r) s4[u][r] = 0.f;
#pragma unroll 4
          for (int f = lane; f < F; f += 32) {
            const float dv = __ldg(dr + f);
#pragma unroll
            for (int r = 0; r < FEW_ROWS; ++r)
              if (r < rpb) s4[u][r] = fmaf(qs[r * F + f], dv, s4[u][r]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + u * FEW_WARPS;
        if (n >= N) break;  // warp-uniform
#pragma unroll
        for (int r = 0; r < FEW_ROWS; ++r) {
          if (r >= rpb) break;
          const float s = warp_sum(s4[u][r]);
          if (s > __shfl_sync(0xffffffffu, tv[r], k - 1)) {  // warp-uniform
            const int pos = __popc(__ballot_sync(0xffffffffu, tv[r] >= s));
            const float up_v = __shfl_up_sync(0xffffffffu, tv[r], 1);
            const int up_i = __shfl_up_sync(0xffffffffu, ti[r], 1);
            if (lane == pos) {
              tv[r] = s;
              ti[r] = n;
            } else if (lane > pos) {
              tv[r] = up_v;
              ti[r] = up_i;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < FEW_ROWS; ++r) {
      mv[(r * FEW_WARPS + warp) * 32 + lane] = lane < k ? tv[r] : -INFINITY;
      mi[(r * FEW_WARPS + warp) * 32 + lane] = ti[r];
    }
    __syncthreads();
    // warp r merges the 8 lists of row r: k rounds of "largest remaining head" (value desc, index asc)
    if (warp < rpb) {
      const int r = warp;
      const int sel = g0 + r;
      const float* lv = mv + (size_t)r * FEW_WARPS * 32;
      const int* lix = mi + (size_t)r * FEW_WARPS * 32;
      int head = 0;  // lane w < 8 tracks the head of list w
      for (int out = 0; out < k; ++out) {
        float hv = -INFINITY;
        int hi = 0x7fffffff;
        if (lane < FEW_WARPS && head < 32) {
          hv = lv[lane * 32 + head];
          hi = lix[lane * 32 + head];
        }
        float bv = hv;
        int bi = hi, bl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
            bl = ol;
          }
        }
        if (lane == bl) ++head;
        if (lane == 0 && sel < nrows) {
          const int row = row_list[sel];
          idx[(size_t)row * k + out] = (int64_t)bi;
          if (sims) sims[(size_t)row * k + out] = bv;
        }
      }
    }
  }
}

size_t knn_umma_workspace_bytes(int Nq, int N, int F);
int knn_topk_umma(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                  cudaStream_t st);

int knn_shard_begin(const float* local, int Nq, int row_lo, int N, int F, int k, void* ws, int phases, cudaStream_t st);
int knn_shard_finish(const float* db, int Nq, int row_lo, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                     int phases, int npeer_max, cudaStream_t st);
void knn_panel_layout(int N, int F, size_t* hi_off, size_t* lo_off, size_t* row_bytes);

int launch_knn_exact(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims,
                     const int* row_list, const int* row_count, cudaStream_t st) {
  if (row_list) {
    const size_t smem = ((size_t)FEW_ROWS * F + 2 * FEW_ROWS * FEW_WARPS * 32) * sizeof(float);
    DG_REQUIRE(smem <= 200 * 1024, DG_ERR_UNSUPPORTED, "knn: feature dimension %d too large for the few-rows kernel", F);
    static PerDevice configured_pd = {};
    size_t& configured = per_device(configured_pd);
    if (smem > configured) {
      DG_CUDA_OK(cudaFuncSetAttribute(knn_fewrows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    DG_PRE(st);
    knn_fewrows_kernel<<<min(ceil_div(min(Nq, FEW_CAP), FEW_ROWS), 148 * 4), FEW_THREADS, smem, st>>>(
        q, db, Nq, N, F, k, idx, sims, row_list, row_count);
    DG_LAUNCH_OK("knn_fewrows_kernel");
  }
  DG_PRE(st);
  knn_topk_kernel<<<ceil_div(Nq, KTM), KNN_THREADS, 0, st>>>(q, db, Nq, N, F, k, idx, sims, row_list, row_count);
  DG_LAUNCH_OK("knn_topk_kernel");
  return DG_OK;
}

// The tensor-core path needs two spare candidates for its certificate (k <= 30, the reference's value) and is
// only worth its set-up for non-trivial problems; DEPTHG_B200_KNN=simt forces the exact fp32 kernel.
static bool use_umma(int Nq, int N, int k) {
  const char* e = getenv("DEPTHG_B200_KNN");
  if (e && e[0] == 's') return false;
  return k <= 30 && N >= 256 && (long long)Nq * N >= (1LL << 20);
}

}  // namespace dg

extern "C" size_t dg_knn_workspace_bytes(int Nq, int N, int F, int k) {
  (void)k;
  if (Nq <= 0 || N <= 0 || F <= 0) return 256;
  return dg::knn_umma_workspace_bytes(Nq, N, F);
}

extern "C" int dg_knn_topk(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims,
                           void* ws, size_t ws_bytes, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(q && db && idx, DG_ERR_INVALID, "dg_knn_topk: null pointer");
  DG_REQUIRE(Nq > 0 && N > 0 && F > 0, DG_ERR_INVALID, "dg_knn_topk: bad sizes");
  DG_REQUIRE(k > 0 && k <= 32, DG_ERR_UNSUPPORTED, "dg_knn_topk: k=%d must be in [1,32]", k);
  DG_REQUIRE(k <= N, DG_ERR_INVALID, "dg_knn_topk: k=%d exceeds database size %d", k, N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (use_umma(Nq, N, k)) {
    DG_REQUIRE(ws && ws_bytes >= knn_umma_workspace_bytes(Nq, N, F), DG_ERR_WORKSPACE, "dg_knn_topk: workspace too small");
    return knn_topk_umma(q, db, Nq, N, F, k, idx, sims, ws, st);
  }
  if (ws && ws_bytes >= 256) DG_CUDA_OK(cudaMemsetAsync(ws, 0, 256, st));   // diagnostics header: nothing to report
  return launch_knn_exact(q, db, Nq, N, F, k, idx, sims, nullptr, nullptr, st);
}

// Two-phase form of dg_knn_topk for a query-row shard whose rows are database rows [row_lo, row_lo + Nq): see
// knn_umma.cu.  Shapes that do not take the tensor-core path do nothing in `begin` and run the plain exact build in
// `finish` (once, when DG_KNN_RERANK is among the phases).
extern "C" int dg_knn_shard_begin(const float* local, int Nq, int row_lo, int N, int F, int k, void* ws, size_t ws_bytes,
                                  int phases, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(local && Nq > 0 && N > 0 && F > 0 && row_lo >= 0 && row_lo + Nq <= N, DG_ERR_INVALID,
             "dg_knn_shard_begin: bad arguments");
  DG_REQUIRE(k > 0 && k <= 32 && k <= N, DG_ERR_INVALID, "dg_knn_shard_begin: k=%d out of range", k);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!use_umma(Nq, N, k)) return DG_OK;
  DG_REQUIRE(ws && ws_bytes >= knn_umma_workspace_bytes(Nq, N, F), DG_ERR_WORKSPACE, "dg_knn_shard_begin: workspace too small");
  return knn_shard_begin(local, Nq, row_lo, N, F, k, ws, phases, st);
}

extern "C" int dg_knn_shard_finish(const float* db, int Nq, int row_lo, int N, int F, int k, int64_t* idx, float* sims,
                                   void* ws, size_t ws_bytes, int phases, int npeer_max, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(Nq > 0 && N > 0 && F > 0 && row_lo >= 0 && row_lo + Nq <= N, DG_ERR_INVALID,
             "dg_knn_shard_finish: bad arguments");
  DG_REQUIRE(k > 0 && k <= 32 && k <= N, DG_ERR_INVALID, "dg_knn_shard_finish: k=%d out of range", k);
  DG_REQUIRE(npeer_max >= 0 && npeer_max <= 32, DG_ERR_INVALID, "dg_knn_shard_finish: npeer_max out of range");
  DG_REQUIRE(!(phases & (DG_KNN_SPLIT_REMOTE | DG_KNN_RERANK)) || db, DG_ERR_INVALID, "dg_knn_shard_finish: db missing");
  DG_REQUIRE(!(phases & DG_KNN_RERANK) || idx, DG_ERR_INVALID, "dg_knn_shard_finish: idx missing");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (use_umma(Nq, N, k)) {
    DG_REQUIRE(ws && ws_bytes >= knn_umma_workspace_bytes(Nq, N, F), DG_ERR_WORKSPACE, "dg_knn_shard_finish: workspace too small");
    return knn_shard_finish(db, Nq, row_lo, N, F, k, idx, sims, ws, phases, npeer_max, st);
  }
  if (!(phases & DG_KNN_RERANK)) return DG_OK;
  if (ws && ws_bytes >= 256) DG_CUDA_OK(cudaMemsetAsync(ws, 0, 256, st));
  return launch_knn_exact(db + (size_t)row_lo * F, db, Nq, N, F, k, idx, sims, nullptr, nullptr, st);
}

namespace dg { int knn_panel_count(); }
extern "C" int dg_knn_panel_count(void) { return dg::knn_panel_count(); }

extern "C" int dg_knn_panel_layout(int N, int F, size_t* hi_offset, size_t* lo_offset, size_t* row_bytes) {
  using namespace dg;
  DG_REQUIRE(N > 0 && F > 0 && hi_offset && lo_offset && row_bytes, DG_ERR_INVALID, "dg_knn_panel_layout: bad arguments");
  knn_panel_layout(N, F, hi_offset, lo_offset, row_bytes);
  return DG_OK;
}
