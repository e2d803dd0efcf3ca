// Cosine-similarity k-NN build: similarity GEMM fused with a running per-row top-k.
//
// Replaces the chunked  einsum("nf,mf->nm") + torch.topk(sims, 30)  loop of
// /root/reference/src/precompute_knns.py:99-113, which materialises a
// [775, 49629] similarity block per chunk on the CPU.  Here a CTA owns 64 query
// rows, streams the database in 64-row tiles, forms each 64x64 similarity tile
// in registers (fp32 FMA, so indices match the fp32 reference) and merges it
// into per-row sorted top-k lists that live in warp registers (one list entry per
// lane, insertion by ballot + shuffle).  The similarity matrix never exists.
#include "common.cuh"

namespace dg {

constexpr int KTM = 64, KTN = 64, KKC = 16, KNN_THREADS = 256;
constexpr int KASTR = KTM + 4;
constexpr int SSTR = KTN + 1;

__global__ void __launch_bounds__(KNN_THREADS) knn_topk_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                               int Nq, int N, int F, int k, int64_t* __restrict__ idx,
                                                               float* __restrict__ sims) {
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
  const int qrow = min(r0 + lr, Nq - 1);
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
    const int row = r0 + 8 * warp + i;
    if (row < Nq && lane < k) {
      idx[(size_t)row * k + lane] = (int64_t)ti[i];
      if (sims) sims[(size_t)row * k + lane] = tv[i];
    }
  }
}

}  // namespace dg

extern "C" size_t dg_knn_workspace_bytes(int Nq, int N, int F, int k) {
  (void)Nq; (void)N; (void)F; (void)k;
  return 256;  // the fp32 path keeps all state in registers / shared memory
}

extern "C" int dg_knn_topk(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims,
                           void* ws, size_t ws_bytes, dg_stream_t stream) {
  using namespace dg;
  (void)ws; (void)ws_bytes;
  DG_REQUIRE(q && db && idx, DG_ERR_INVALID, "dg_knn_topk: null pointer");
  DG_REQUIRE(Nq > 0 && N > 0 && F > 0, DG_ERR_INVALID, "dg_knn_topk: bad sizes");
  DG_REQUIRE(k > 0 && k <= 32, DG_ERR_UNSUPPORTED, "dg_knn_topk: k=%d must be in [1,32]", k);
  DG_REQUIRE(k <= N, DG_ERR_INVALID, "dg_knn_topk: k=%d exceeds database size %d", k, N);
  DG_PRE(reinterpret_cast<cudaStream_t>(stream));
  knn_topk_kernel<<<ceil_div(Nq, KTM), KNN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q, db, Nq, N, F, k,
                                                                                                  idx, sims);
  DG_LAUNCH_OK("knn_topk_kernel");
  return DG_OK;
}
