// Tensor-core cosine-similarity k-NN build (tcgen05 + TMEM + TMA) with an exact fp32 finish.
//
// Replaces the einsum + topk(…,30) loop of /root/reference/src/precompute_knns.py:99-113.
// The reference result is defined by fp32 similarities, and no tensor-core-only scheme provably
// reproduces its index order (SURVEY.md 7.1 iii), so the build is three steps that never
// materialise the similarity matrix:
//   1. split_rows_kernel   : fp32 rows -> bf16 hi/lo panels (x ~= hi + lo), K padded to 64
//   2. knn_umma_kernel     : one CTA per 128 query rows streams the database in 128-row tiles;
//                            sims = hi.hi + hi.lo + lo.hi on tcgen05 (error <~ 2e-5), two TMEM
//                            accumulators so the MMAs of tile t+1 overlap the epilogue of tile t;
//                            the epilogue (one thread per query row) keeps the 32 best candidates
//                            of its row in a sorted shared-memory list (threshold in a register); the
//                            database is cut into nseg segments (grid.y) so that small query blocks
//                            (multi-GPU shards) still fill the GPU and every row gets nseg x 32 candidates
//   3. knn_rerank_kernel   : exact fp32 dot products of the 32 candidates (a warp per query row),
//                            sort by (value desc, index asc), emit the top k, and CERTIFY the row:
//                            exact k-th value > approximate 32nd value + error bound, i.e. no row
//                            outside the candidate list can belong to the true top-k.  Rows that fail
//                            are listed and recomputed by the exact fp32 kernel (knn.cu).
#include <stdlib.h>

#include "kernels.cuh"
#include "umma.cuh"

namespace dg {

using namespace umma;

constexpr int KU_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane group)
constexpr int KU_SUB = 2;        // epilogue warps per lane group = candidate lists per (row, segment)
// BN = database rows per MMA tile (the UMMA N).  A stage holds one 64-wide K chunk of the query block (hi, lo:
// 2 x 16 KB) and of BN database rows (hi, lo: 2 x BN x 128 B).  BN = 256 loads the query chunk once per 256
// database rows instead of once per 128: 25 % less L2->SM operand traffic, which is what bounds this kernel.
// The kernel is bound by how many operand bytes an SM keeps in flight from L2, so BN = 256 uses 32-wide K chunks
// (64-byte rows, SWIZZLE_64B): 4 stages of 48 KB instead of 2 of 96 KB.
template <int BN> struct KuCfg {
  static constexpr int KC = BN == 128 ? 64 : 32;                  // K elements per stage
  static constexpr int QBYTES = 128 * KC * 2, DBYTES = BN * KC * 2;  // one bf16 panel chunk of the query / database tile
  static constexpr int STAGE = 2 * QBYTES + 2 * DBYTES;
  static constexpr int NSTAGE = BN == 128 ? 3 : 4;
  static constexpr int SMEM = NSTAGE * STAGE + 8 * 32 * 33 * 4 + 1024 + 256;
};
constexpr int KU_CAND = 32;
constexpr int KU_LSTR = KU_CAND + 1;
// Bound on |approx - exact| of the 3-product bf16 split, as a multiple of |q| * max|d| (the row norms are measured by
// split_rows_kernel, so un-normalised inputs get a proportionally wider bound instead of a silently wrong one):
//   x = hi + lo + r with |lo| <= 2^-8 |x| and |r| <= 2^-8 |lo| <= 2^-16 |x| (bf16 keeps 8 significand bits), so
//   q.d - (qh.dh + qh.dl + ql.dh) = ql.dl + rq.d + q.rd  is at most 3 * 2^-16 |q||d| by Cauchy-Schwarz;
//   the tensor core adds 3 F / 16 instruction results of 16 products each into one fp32 accumulator: every add loses
//   at most 2^-23 of a partial sum that never exceeds |q||d|  ->  (3 F / 16 + 16) * 2^-23 |q||d|.
// F = 768: 4.6e-5 + 1.9e-5 = 6.5e-5 (measured maximum on unit-norm rows: 5e-6, SURVEY 7.1 iii).
__host__ __device__ inline float ku_eps_unit(int F) {
  return 3.f * 1.52587890625e-5f + (3.f * (float)F / 16.f + 16.f) * 1.1920929e-7f;
}

struct KnnUmmaParams {
  CUtensorMap tm_qh, tm_ql, tm_dh, tm_dl;  // bf16 [rows, Fp], box 64 x 128, SWIZZLE_128B
  int Nq, N, nchunk, ntiles, nseg;          // ntiles in units of BN database rows
  int* cand_idx;    // [Nq,nseg,KU_SUB,32]
  float* cand_val;  // [Nq,nseg,KU_SUB,32] approximate sims, descending within a list
  int* err;
};

// A warp per row: bf16 hi/lo panels (K padded with zeros to Fp), the row's squared norm (rownorm2, nullable) and the
// maximum squared norm over all rows (maxnorm2_bits: the float's bit pattern, which orders like an int for x >= 0).
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, int n, int F, int Fp,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                         float* __restrict__ rownorm2, int* __restrict__ maxnorm2_bits) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (F & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  float wmax = 0.f;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps) {
    const float* xr = x + (size_t)r * F;
    __nv_bfloat16* hr = hi + (size_t)r * Fp;
    __nv_bfloat16* lr = lo + (size_t)r * Fp;
    float ss = 0.f;
    for (int c = lane * 4; c < Fp; c += 128) {
      float v[4];
      if (vec && c + 3 < F) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xr + c));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = c + e < F ? __ldg(xr + c + e) : 0.f;
      }
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ss = fmaf(v[e], v[e], ss);
        h[e] = __float2bfloat16_rn(v[e]);
        l[e] = __float2bfloat16_rn(v[e] - __bfloat162float(h[e]));
      }
      *reinterpret_cast<uint2*>(hr + c) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(lr + c) = *reinterpret_cast<uint2*>(l);
    }
    ss = warp_sum(ss);
    if (rownorm2 && lane == 0) rownorm2[r] = ss;
    wmax = fmaxf(wmax, ss);
  }
  if (maxnorm2_bits && lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_int(wmax));
}

template <int BN>
__global__ void __launch_bounds__(KU_THREADS, 1) knn_umma_kernel(const __grid_constant__ KnnUmmaParams prm) {
  constexpr int KU_STAGE = KuCfg<BN>::STAGE, KU_NSTAGE = KuCfg<BN>::NSTAGE, KC = KuCfg<BN>::KC;
  constexpr int QB = KuCfg<BN>::QBYTES, DB = KuCfg<BN>::DBYTES;
  extern __shared__ uint8_t ku_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an offset so the pointer stays in the shared address space
  // (a round trip through uintptr_t makes every later access a generic LD/ST instead of LDS/STS)
  uint8_t* ring = ku_raw + ((1024u - (smem_u32(ku_raw) & 1023u)) & 1023u);
  float* stage_tiles = reinterpret_cast<float*>(ring + KU_NSTAGE * KU_STAGE);  // [4 warps][32 rows][33] transpose staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_tiles + 8 * 32 * KU_LSTR);
  uint64_t* full = bars;               // [<=4]
  uint64_t* empty = bars + 4;          // [<=4]
  uint64_t* tfull = bars + 8;          // [2] accumulator ready
  uint64_t* tempty = bars + 10;        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  // blockIdx.y = database segment: this CTA ranks its 128 query rows against tiles [t_begin, t_end) only, which
  // gives small query blocks (multi-GPU shards) enough CTAs to fill the GPU and every row nseg x 32 candidates
  const int seg = blockIdx.y;
  const int per_seg = (prm.ntiles + prm.nseg - 1) / prm.nseg;
  const int t_begin = seg * per_seg, t_end = min(t_begin + per_seg, prm.ntiles);
  const int nchunk = prm.nchunk, ntiles = max(t_end - t_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < KU_NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 256);  // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&prm.tm_qh); prefetch_tmap(&prm.tm_ql); prefetch_tmap(&prm.tm_dh); prefetch_tmap(&prm.tm_dl);
      int job = 0;
      for (int t = 0; t < ntiles; ++t) {
        for (int c = 0; c < nchunk; ++c, ++job) {
          const int s = job % KU_NSTAGE;
          if (!mbar_wait(&empty[s], ((job / KU_NSTAGE) & 1) ^ 1)) { if (prm.err) atomicCAS(prm.err, 0, 11); return; }
          uint8_t* st = ring + s * KU_STAGE;
          mbar_arrive_expect_tx(&full[s], KU_STAGE);
          tma_load_2d(st, &prm.tm_qh, &full[s], c * KC, m0);
          tma_load_2d(st + QB, &prm.tm_ql, &full[s], c * KC, m0);
          tma_load_2d(st + 2 * QB, &prm.tm_dh, &full[s], c * KC, (t_begin + t) * BN);
          tma_load_2d(st + 2 * QB + DB, &prm.tm_dl, &full[s], c * KC, (t_begin + t) * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(FMT_BF16, 128, BN, 0, 0);
      // K-major operands: rows of KC bf16 (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row groups 8 rows apart
      const uint64_t dk128 = KC == 64 ? smem_desc(0, 16, 1024, SW_128B) : smem_desc(0, 16, 512, SW_64B);
      int job = 0;
      bool ok = true;
      for (int t = 0; t < ntiles && ok; ++t) {
        const int a = t & 1;
        ok = mbar_wait(&tempty[a], ((t >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t acc = tmem + a * BN;
        for (int c = 0; c < nchunk && ok; ++c, ++job) {
          const int s = job % KU_NSTAGE;
          ok = mbar_wait(&full[s], (job / KU_NSTAGE) & 1);
          tc_fence_after_sync();
          const uint32_t a0 = smem_u32(ring + s * KU_STAGE) >> 4;
          const uint64_t ah = dk128 + a0, al = ah + QB / 16, bh = ah + 2 * QB / 16, bl = bh + DB / 16;  // 16-byte units
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            mma_f16(acc, ah + 2 * ks, bh + 2 * ks, idesc, (c | ks) != 0);
            mma_f16(acc, ah + 2 * ks, bl + 2 * ks, idesc, 1);
            mma_f16(acc, al + 2 * ks, bh + 2 * ks, idesc, 1);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[a]);
      }
      if (!ok && prm.err) atomicCAS(prm.err, 0, 12);
    }
  } else {
    // ---- epilogue: warp-cooperative running top-32 per row.  The warp owns the 32 rows of its TMEM lane group; each
    //      row's sorted candidate list is spread over the 32 lanes (entry l in lane l), so an insertion is one
    //      ballot + two shuffles and costs the same whether the list is cold or warm.
    // Two warps share each TMEM lane group (a warp may only read lanes 32*(warp%4)..+31) and split its column
    // chunks even/odd; each keeps its own list, so a row ends up with KU_SUB lists per database segment.
    const int lg = warp & 3, sub = (warp - 2) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(32 * lg) << 16);
    float* tile = stage_tiles + (warp - 2) * (32 * KU_LSTR);
    float tv[32];
    int ti[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      tv[r] = -INFINITY;
      ti[r] = -1;
    }
    float v[32];
    float mythr = -INFINITY;  // current 32nd-best value of query row `lane` (the row this thread reads from TMEM)
    bool ok = true;
    for (int t = 0; t < ntiles && ok; ++t) {
      const int a = t & 1;
      ok = mbar_wait(&tfull[a], (t >> 1) & 1);
      tc_fence_after_sync();
      const int n0 = (t_begin + t) * BN;
#pragma unroll 1
      for (int cc = sub; cc < BN / 32; cc += KU_SUB) {
        tmem_ld_32x32(tlane + a * BN + 32 * cc, v);   // thread = query row `lane`, 32 consecutive database columns
        tmem_ld_wait();
        const int nb = n0 + 32 * cc;
        // Filter in the row-per-thread layout first: a column only matters if it beats the row's current 32nd
        // value, which this thread keeps in `mythr`.  Once the lists are warm most rows have nothing to insert,
        // and only rows that do are transposed through shared memory and visited by the list code below.
        float vmax = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) vmax = fmaxf(vmax, v[i]);
        // (columns past the end of the database are zero-filled by TMA: they can only cause a harmless visit,
        //  the list code below masks them with col_ok)
        const bool mine = vmax > mythr;
        const unsigned rows = __ballot_sync(0xffffffffu, mine);
        if (rows == 0) continue;
        if (mine) {
#pragma unroll
          for (int i = 0; i < 32; ++i) tile[lane * KU_LSTR + i] = v[i];
        }
        __syncwarp();
        const bool col_ok = nb + lane < prm.N;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          if (!((rows >> r) & 1u)) continue;                              // warp-uniform
          const float x = tile[r * KU_LSTR + lane];                       // lane = column of row r
          const float thr = __shfl_sync(0xffffffffu, tv[r], KU_CAND - 1);
          unsigned m = __ballot_sync(0xffffffffu, col_ok && x > thr);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float nv = __shfl_sync(0xffffffffu, x, src);
            if (nv > __shfl_sync(0xffffffffu, tv[r], KU_CAND - 1)) {      // still above the (possibly raised) 32nd value
              const int pos = __popc(__ballot_sync(0xffffffffu, tv[r] >= nv));
              const float up_v = __shfl_up_sync(0xffffffffu, tv[r], 1);
              const int up_i = __shfl_up_sync(0xffffffffu, ti[r], 1);
              if (lane == pos) {
                tv[r] = nv;
                ti[r] = nb + src;
              } else if (lane > pos) {
                tv[r] = up_v;
                ti[r] = up_i;
              }
            }
          }
          const float nthr = __shfl_sync(0xffffffffu, tv[r], KU_CAND - 1);
          if (lane == r) mythr = nthr;
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      mbar_arrive(&tempty[a]);
    }
    if (!ok && prm.err) atomicCAS(prm.err, 0, 13);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int q = m0 + 32 * lg + r;
      if (q < prm.Nq) {
        const size_t o = (((size_t)q * prm.nseg + seg) * KU_SUB + sub) * KU_CAND;
        prm.cand_idx[o + lane] = ti[r];
        prm.cand_val[o + lane] = tv[r];
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

// One warp per query row: exact fp32 similarities of its nseg x 32 candidates (lane l owns candidate l of every
// segment), k rounds of warp-argmax by (value desc, index asc) to emit the top k, and the certificate.
constexpr int KU_MAXSEG = 8;

__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                         int Nq, int N, int F, int k, int nseg,
                                                         const int* __restrict__ cand_idx,
                                                         const float* __restrict__ cand_val, int64_t* __restrict__ idx,
                                                         float* __restrict__ sims, int* __restrict__ fail_rows,
                                                         int* __restrict__ fail_count, const int* __restrict__ err,
                                                         const float* __restrict__ qnorm2,
                                                         const int* __restrict__ dbmax2_bits, float eps_unit) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= Nq) return;
  // error bound of the approximate similarities of this row: eps_unit * |q| * max |d|
  const float eps = eps_unit * sqrtf(__ldg(qnorm2 + row)) * sqrtf(__int_as_float(__ldg(dbmax2_bits)));
  const float* qr = q + (size_t)row * F;
  int my_idx[KU_MAXSEG];
  float my_a[KU_MAXSEG], my_e[KU_MAXSEG];
  float tau = -INFINITY;  // everything outside the candidate lists is <= tau (approximately)
#pragma unroll
  for (int sg = 0; sg < KU_MAXSEG; ++sg) {
    my_idx[sg] = -1;
    my_a[sg] = -INFINITY;
    my_e[sg] = -INFINITY;
    if (sg < nseg) {
      const size_t o = ((size_t)row * nseg + sg) * KU_CAND;
      my_idx[sg] = cand_idx[o + lane];
      my_a[sg] = my_idx[sg] >= 0 ? cand_val[o + lane] : -INFINITY;
      tau = fmaxf(tau, cand_val[o + KU_CAND - 1]);
    }
  }
  // prune: the k-th largest APPROXIMATE value a_k bounds the exact top-k from below by a_k - eps, so candidates with
  // approx < a_k - 2 eps cannot be in it and need no exact dot product (with nseg segments that is most of them)
  float a_k = -INFINITY;
  {
    float cut = INFINITY;   // values >= cut have been counted already
    int taken = 0;
    for (int it = 0; it < k && taken < k; ++it) {
      float best = -INFINITY;
#pragma unroll
      for (int sg = 0; sg < KU_MAXSEG; ++sg)
        if (my_a[sg] < cut) best = fmaxf(best, my_a[sg]);
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 16));
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 8));
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 4));
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
      best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
      int c = 0;
#pragma unroll
      for (int sg = 0; sg < KU_MAXSEG; ++sg) c += (my_a[sg] == best) ? 1 : 0;
      taken += __reduce_add_sync(0xffffffffu, c);
      cut = best;
      a_k = best;
      if (best == -INFINITY) break;
    }
  }
  const float keep = a_k - 2.f * eps;
#pragma unroll
  for (int sg = 0; sg < KU_MAXSEG; ++sg) {
    if (sg < nseg) {
      if (my_a[sg] < keep) my_idx[sg] = -1;   // cannot be in the exact top-k
      unsigned live = __ballot_sync(0xffffffffu, my_idx[sg] >= 0);
      while (live) {
        const int c = __ffs(live) - 1;
        live &= live - 1;
        const int ci = __shfl_sync(0xffffffffu, my_idx[sg], c);
        const float* dr = db + (size_t)ci * F;
        float s = 0.f;
        for (int f = lane; f < F; f += 32) s = fmaf(__ldg(qr + f), __ldg(dr + f), s);
        s = warp_sum(s);
        if (lane == c) my_e[sg] = s;
      }
    }
  }
  float ek = -INFINITY;
  for (int out = 0; out < k; ++out) {
    // this lane's best remaining candidate
    float bv = -INFINITY;
    int bi = 0x7fffffff, bs = 0;
#pragma unroll
    for (int sg = 0; sg < KU_MAXSEG; ++sg) {
      const bool better = my_idx[sg] >= 0 && (my_e[sg] > bv || (my_e[sg] == bv && my_idx[sg] < bi));
      bv = better ? my_e[sg] : bv;
      bi = better ? my_idx[sg] : bi;
      bs = better ? sg : bs;
    }
    float wv = bv;
    int wi = bi, wl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (ov > wv || (ov == wv && oi < wi)) {
        wv = ov;
        wi = oi;
        wl = ol;
      }
    }
    if (lane == wl) {
#pragma unroll
      for (int sg = 0; sg < KU_MAXSEG; ++sg)
        if (sg == bs) my_idx[sg] = -1;  // consumed
    }
    if (lane == 0 && wi != 0x7fffffff) {
      idx[(size_t)row * k + out] = (int64_t)wi;
      if (sims) sims[(size_t)row * k + out] = wv;
    }
    if (out == k - 1) ek = (wi != 0x7fffffff) ? wv : -INFINITY;
  }
  // certificate: the exact k-th best candidate must beat anything the approximate pass could have dropped
  const bool bad = !(ek > tau + eps) || (err && *err != 0);
  if (lane == 0 && bad) fail_rows[atomicAdd(fail_count, 1)] = row;  // recomputed by the exact kernel
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                    uint32_t box_cols) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  DG_REQUIRE(fn, DG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DG_REQUIRE(r == CUDA_SUCCESS, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DG_OK;
}

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// Database segments per query block.  More segments fill the SMs when there are few query blocks (a multi-GPU
// shard) but every segment restarts its candidate lists cold, and a cold list inserts far more often: measured per
// 128x128 tile 12.6 us at nseg = 1 and 22 us at nseg = 8 (N = 49 629, F = 768), i.e. ~ +11 % per extra segment.
// Pick the nseg that minimises  waves x tiles-per-segment x (1 + 0.11 (nseg - 1)).
static int knn_nseg(int Nq, int N, int k) {
  const int nblocks = ceil_div(Nq, 128), ntiles = ceil_div(N, 128);
  double best_cost = 1e300;
  // (every (row, segment) keeps KU_SUB lists of 32, so even one segment leaves >= 2 x 32 - k spare candidates)
  (void)k;
  const int lo = 1;
  int best = lo;
  for (int nseg = lo; nseg * KU_SUB <= KU_MAXSEG && nseg <= max(lo, ntiles / 8); ++nseg) {
    const double waves = (double)ceil_div(nblocks * nseg, 148);
    const double cost = waves * (double)ceil_div(ntiles, nseg) * (1.0 + 0.11 * (nseg - 1));
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = nseg;
    }
  }
  return best;
}

size_t knn_umma_workspace_bytes(int Nq, int N, int F) {
  const size_t Fp = (size_t)round_up(F, 64);
  return 256 + 2 * al256((size_t)N * Fp * 2) + 2 * al256((size_t)Nq * Fp * 2) +
         al256((size_t)Nq * KU_MAXSEG * KU_CAND * 4) * 2 + 2 * al256((size_t)Nq * 4);
}

// declared in knn.cu: exact fp32 kernel restricted to the query rows listed in row_list[0 .. *row_count)
int launch_knn_exact(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims,
                     const int* row_list, const int* row_count, cudaStream_t st);

int knn_topk_umma(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                  cudaStream_t st) {
  const int Fp = round_up(F, 64);
  uint8_t* w = static_cast<uint8_t*>(ws);
  int* err = reinterpret_cast<int*>(w);
  size_t off = 256;
  auto take = [&](size_t bytes) { uint8_t* p = w + off; off += al256(bytes); return p; };
  __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * Fp * 2));
  __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * Fp * 2));
  const bool same = (q == db && Nq == N);
  __nv_bfloat16* qh = same ? dh : reinterpret_cast<__nv_bfloat16*>(take((size_t)Nq * Fp * 2));
  __nv_bfloat16* ql = same ? dl : reinterpret_cast<__nv_bfloat16*>(take((size_t)Nq * Fp * 2));
  if (same) off += 2 * al256((size_t)Nq * Fp * 2);
  const int nseg = knn_nseg(Nq, N, k);
  int* cand_idx = reinterpret_cast<int*>(take((size_t)Nq * KU_MAXSEG * KU_CAND * 4));
  float* cand_val = reinterpret_cast<float*>(take((size_t)Nq * KU_MAXSEG * KU_CAND * 4));
  int* fail_rows = reinterpret_cast<int*>(take((size_t)Nq * 4));
  float* qnorm2 = reinterpret_cast<float*>(take((size_t)Nq * 4));
  // zeroed header: [0] error flag, [1] number of uncertified rows (dg_knn_topk's caller may read it back after the
  // stream has drained), [2] bits of the largest squared database row norm
  int* fail_count = err + 1;
  int* dbmax2 = err + 2;
  DG_CUDA_OK(cudaMemsetAsync(err, 0, 256, st));

  DG_PRE(st);
  split_rows_kernel<<<148 * 8, 256, 0, st>>>(db, N, F, Fp, dh, dl, same ? qnorm2 : nullptr, dbmax2);
  DG_LAUNCH_OK("split_rows_kernel");
  if (!same) {
    DG_PRE(st);
    split_rows_kernel<<<148 * 8, 256, 0, st>>>(q, Nq, F, Fp, qh, ql, qnorm2, nullptr);
    DG_LAUNCH_OK("split_rows_kernel");
  }
  KnnUmmaParams prm;
  int rc;
  static int bn_env = -1;  // DEPTHG_B200_KNN_BN = 128 | 256 (experiments); default 256
  if (bn_env < 0) {
    const char* e = getenv("DEPTHG_B200_KNN_BN");
    bn_env = e && atoi(e) == 128 ? 128 : 256;
  }
  const int BN = bn_env;
  const int KC = BN == 128 ? KuCfg<128>::KC : KuCfg<256>::KC;
  if ((rc = make_map(&prm.tm_qh, qh, Fp, Nq, 128, KC))) return rc;
  if ((rc = make_map(&prm.tm_ql, ql, Fp, Nq, 128, KC))) return rc;
  if ((rc = make_map(&prm.tm_dh, dh, Fp, N, BN, KC))) return rc;
  if ((rc = make_map(&prm.tm_dl, dl, Fp, N, BN, KC))) return rc;
  prm.Nq = Nq; prm.N = N; prm.nchunk = Fp / KC; prm.ntiles = ceil_div(N, BN); prm.nseg = nseg;
  prm.cand_idx = cand_idx; prm.cand_val = cand_val; prm.err = err;
  static PerDevice attr_pd = {};
  size_t& attr_set = per_device(attr_pd);
  if (!attr_set) {
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<128>::SMEM));
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<256>::SMEM));
    attr_set = 1;
  }
  DG_PRE(st);
  if (BN == 128)
    knn_umma_kernel<128><<<dim3(ceil_div(Nq, 128), nseg), KU_THREADS, KuCfg<128>::SMEM, st>>>(prm);
  else
    knn_umma_kernel<256><<<dim3(ceil_div(Nq, 128), nseg), KU_THREADS, KuCfg<256>::SMEM, st>>>(prm);
  DG_LAUNCH_OK("knn_umma_kernel");
  DG_PRE(st);
  knn_rerank_kernel<<<ceil_div(Nq * 32, 256), 256, 0, st>>>(q, db, Nq, N, F, k, nseg * KU_SUB, cand_idx, cand_val, idx, sims, fail_rows,
                                                             fail_count, err, qnorm2, dbmax2, ku_eps_unit(F));
  DG_LAUNCH_OK("knn_rerank_kernel");
  return launch_knn_exact(q, db, Nq, N, F, k, idx, sims, fail_rows, fail_count, st);
}

}  // namespace dg
