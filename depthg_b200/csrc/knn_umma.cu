// Tensor-core cosine-similarity k-NN build (tcgen05 + TMEM + TMA) with an exact fp32 finish.
//
// Replaces the einsum + topk(…,30) loop of /root/reference/src/precompute_knns.py:99-113.
// The reference result is defined by fp32 similarities, and no tensor-core-only scheme provably
// reproduces its index order (SURVEY.md 7.1 iii), so the build is three steps that never
// materialise the similarity matrix:
//   1. split_rows_kernel   : fp32 rows -> bf16 hi/lo panels (x ~= hi + lo), K padded to 64
//   2. knn_umma_kernel     : one CTA per 128 query rows streams the database in 256-row tiles;
//                            sims = hi.hi + hi.lo + lo.hi on tcgen05 (error bound: ku_eps_unit), two TMEM
//                            accumulators so the MMAs of tile t+1 overlap the epilogue of tile t;
//                            the epilogue is ONE THREAD PER QUERY ROW (the TMEM row-per-thread layout): the row's 32
//                            best candidates live in shared memory as an unsorted list with its minimum (the
//                            threshold) and the minimum's slot in registers; a value enters by overwriting the
//                            minimum, followed by a 32-entry rescan.  32 rows insert in parallel, so a cold list
//                            costs microseconds (the round-1 warp-cooperative sorted lists serialised the rows and
//                            paid ~0.5 ms per cold segment) - which is what makes database segments cheap: the
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

constexpr int KU_THREADS = 192;  // TMA warp + MMA warp + 4 epilogue warps (one per TMEM lane group)
constexpr int KU_SUB = 1;        // candidate lists per (row, segment)
constexpr int KU_CAND = 32;
// BN = database rows per MMA tile (the UMMA N).  A stage holds one K chunk of the query block (hi, lo) and of BN
// database rows (hi, lo).  BN = 256 loads the query chunk once per 256 database rows instead of once per 128 (25 % less
// L2->SM operand traffic) and uses 32-wide K chunks (64-byte rows, SWIZZLE_64B): 4 stages of 48 KB in flight.
// Besides the ring, shared memory holds the candidate lists: 128 rows x 32 entries x (value, index) = 32 KB.
// TERMS = 3: bf16 hi/lo panels, sims = hh + hl + lh (fp32-grade, error bound ku_eps_unit);
// TERMS = 1: the FAST pass of the precision ladder (SURVEY 7.1 iii): ONE fp16 panel, sims = h.h — a third of the MMAs
//            and half of the operand bytes (a stage carries 64 K elements instead of 32, so the L2->SM stream that
//            bounds the 3-term pass at ~10 TB/s halves) for an error bound 15x wider (ku_eps_unit_fast), which the
//            re-rank absorbs: more candidates are ambiguous and get an exact fp32 dot product, and the certificate
//            still decides row by row whether the candidate lists provably contain the exact top-k.
// MQ = query blocks of 128 rows per CTA.  MQ = 2 (fast pass only): every database chunk that lands in shared memory is
//      multiplied with TWO query blocks (two 128 x 256 accumulators = all 512 TMEM columns), so the L2->SM operand
//      stream per useful flop drops by a third (384 instead of 576 KB per 128 x 256 tile) - and that stream, not the
//      tensor pipe, bounds the pass (~60 GB/s per SM).  One accumulator set means the MMAs of tile t+1 wait for the
//      epilogue of tile t; measured slower than MQ = 1 (see knn_mq), so it is an experiment knob, not the default.
template <int BN, int TERMS, int MQ = 1> struct KuCfg {
  static constexpr int KC = MQ == 2 ? 32 : ((BN == 128 || TERMS == 1) ? 64 : 32);   // K elements per stage
  static constexpr int QBYTES = 128 * MQ * KC * 2, DBYTES = BN * KC * 2;  // one 16-bit panel chunk of the query / database tile
  static constexpr int STAGE = TERMS == 3 ? 2 * QBYTES + 2 * DBYTES : QBYTES + DBYTES;
  static constexpr int NSTAGE = MQ == 2 ? 5 : ((BN == 128 && TERMS == 3) ? 3 : 4);
  static constexpr int LISTS = 4 * MQ * 32 * KU_CAND * 8;
  static constexpr int SMEM = NSTAGE * STAGE + LISTS + 1024 + 256;
  static constexpr int THREADS = 64 + 128 * MQ;   // TMA warp + MMA warp + 4 MQ epilogue warps
  static constexpr int NACC = MQ == 2 ? 1 : 2;    // accumulator sets in TMEM
};
static_assert(KuCfg<256, 3>::SMEM <= 232448 && KuCfg<128, 3>::SMEM <= 232448 && KuCfg<256, 1>::SMEM <= 232448 &&
              KuCfg<256, 1, 2>::SMEM <= 232448, "knn_umma_kernel: shared memory over the CTA limit");
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

// Fast pass (one fp16 panel): x = h + r with |r| <= 2^-11 |x| for normal fp16 values (11 significand bits) and
// |r| <= 2^-25 absolutely below the normal range, so q.d - qh.dh = rq.d + qh.rd is at most
//   (2 * 2^-11 + 2^-22) |q||d|  +  2^-25 sqrt(F) (|q| + |d|)        (Cauchy-Schwarz; the second term: subnormals)
// plus the accumulation term (F / 16 + 16) * 2^-23 |q||d| of F / 16 instruction results.  Values of 65504 or more do
// not convert: the re-rank refuses the fast result (every row goes to the exact kernel) when a squared norm exceeds
// 2^30.  F = 768, unit norms: 9.9e-4.
__host__ __device__ inline float ku_eps_unit_fast(int F) {
  return 2.f * 4.8828125e-4f + 2.38418579e-7f + ((float)F / 16.f + 16.f) * 1.1920929e-7f;
}
__host__ __device__ inline float ku_eps_abs_fast(int F) { return 2.98023224e-8f * sqrtf((float)F); }

struct KnnUmmaParams {
  CUtensorMap tm_qh, tm_ql, tm_dh, tm_dl;  // bf16 [rows, Fp], box 64 x 128, SWIZZLE_128B
  int Nq, N, nchunk, ntiles, nseg;          // ntiles in units of BN database rows (VIRTUAL tiles, see tile_of)
  int list_pitch;                           // lists per row in cand_* (>= nseg: a launch may fill only the first nseg)
  // The tiles a launch visits are a virtual list mapped onto the database: virtual tile v is tile v when v < tA_end
  // and tile v + tB_shift otherwise (the sharded build visits "the local rows" and later "everything but the local
  // rows").  Columns are masked by col_mode: 0 = column < N, 1 = only [col_lo, col_hi), 2 = only outside of it.
  int tA_end, tB_shift, col_lo, col_hi, col_mode;
  int warm;                                 // 1: start from the candidate lists a previous launch left in cand_*
  int* cand_idx;    // [Nq,nseg,32]; unused entries -1
  float* cand_val;  // [Nq,nseg,32] approximate sims (unsorted); unused entries -inf
  int* err;
};

// A warp per row: bf16 hi/lo panels (K padded with zeros to Fp), the row's squared norm (rownorm2, nullable) and the
// maximum squared norm over all rows (maxnorm2_bits: the float's bit pattern, which orders like an int for x >= 0).
template <bool FAST>   // FAST: one fp16 panel (hi); lo is not written
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
      if (FAST) {
        __half h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ss = fmaf(v[e], v[e], ss);
          h[e] = __float2half_rn(v[e]);
        }
        *reinterpret_cast<uint2*>(hr + c) = *reinterpret_cast<uint2*>(h);
      } else {
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
    }
    ss = warp_sum(ss);
    if (rownorm2 && lane == 0) rownorm2[r] = ss;
    wmax = fmaxf(wmax, ss);
  }
  if (maxnorm2_bits && lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_int(wmax));
}

template <int BN, int TERMS, int MQ>
__global__ void __launch_bounds__(KuCfg<BN, TERMS, MQ>::THREADS, 1) knn_umma_kernel(const __grid_constant__ KnnUmmaParams prm) {
  using Cfg = KuCfg<BN, TERMS, MQ>;
  static_assert(MQ == 1 || (TERMS == 1 && BN == 256), "two query blocks per CTA: fast pass, 256-row database tiles");
  constexpr int NACC = Cfg::NACC;
  constexpr int KU_STAGE = Cfg::STAGE, KU_NSTAGE = Cfg::NSTAGE, KC = Cfg::KC;
  constexpr int QB = Cfg::QBYTES, DB = Cfg::DBYTES;
  extern __shared__ uint8_t ku_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an offset so the pointer stays in the shared address space
  // (a round trip through uintptr_t makes every later access a generic LD/ST instead of LDS/STS)
  uint8_t* ring = ku_raw + ((1024u - (smem_u32(ku_raw) & 1023u)) & 1023u);
  // candidate lists: per epilogue warp 32 slots x 32 rows of values, then of indices; slot e of row r lives at word
  // e * 32 + (r ^ e), which is conflict-free both for "every row scans slot e" and for "one row, all slots"
  float* lists = reinterpret_cast<float*>(ring + KU_NSTAGE * KU_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + KU_NSTAGE * KU_STAGE + Cfg::LISTS);
  uint64_t* full = bars;                          // [KU_NSTAGE]
  uint64_t* empty = bars + KU_NSTAGE;             // [KU_NSTAGE]
  uint64_t* tfull = bars + 2 * KU_NSTAGE;         // [2] accumulator ready
  uint64_t* tempty = bars + 2 * KU_NSTAGE + 2;    // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * KU_NSTAGE + 4);
  static_assert((2 * KU_NSTAGE + 5) * 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128 * MQ;
  // blockIdx.y = database segment: this CTA ranks its 128 query rows against tiles [t_begin, t_end) only, which
  // gives small query blocks (multi-GPU shards) enough CTAs to fill the GPU and every row nseg x 32 candidates
  const int seg = blockIdx.y;
  const int per_seg = (prm.ntiles + prm.nseg - 1) / prm.nseg;
  const int t_begin = seg * per_seg, t_end = min(t_begin + per_seg, prm.ntiles);
  const int nchunk = prm.nchunk, ntiles = max(t_end - t_begin, 0);
  auto tile_of = [&](int v) { return v < prm.tA_end ? v : v + prm.tB_shift; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < KU_NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128 * MQ);  // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);   // MQ = 1: two alternating sets; MQ = 2: one set of two halves
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&prm.tm_qh); prefetch_tmap(&prm.tm_dh);
      if (TERMS == 3) { prefetch_tmap(&prm.tm_ql); prefetch_tmap(&prm.tm_dl); }
      int job = 0;
      for (int t = 0; t < ntiles; ++t) {
        for (int c = 0; c < nchunk; ++c, ++job) {
          const int s = job % KU_NSTAGE;
          if (!mbar_wait(&empty[s], ((job / KU_NSTAGE) & 1) ^ 1)) { if (prm.err) atomicCAS(prm.err, 0, 11); return; }
          uint8_t* st = ring + s * KU_STAGE;
          mbar_arrive_expect_tx(&full[s], KU_STAGE);
          if (TERMS == 3) {
            tma_load_2d(st, &prm.tm_qh, &full[s], c * KC, m0);
            tma_load_2d(st + QB, &prm.tm_ql, &full[s], c * KC, m0);
            tma_load_2d(st + 2 * QB, &prm.tm_dh, &full[s], c * KC, tile_of(t_begin + t) * BN);
            tma_load_2d(st + 2 * QB + DB, &prm.tm_dl, &full[s], c * KC, tile_of(t_begin + t) * BN);
          } else {
            tma_load_2d(st, &prm.tm_qh, &full[s], c * KC, m0);
            tma_load_2d(st + QB, &prm.tm_dh, &full[s], c * KC, tile_of(t_begin + t) * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(TERMS == 3 ? FMT_BF16 : FMT_F16, 128, BN, 0, 0);
      // K-major operands: rows of KC bf16 (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row groups 8 rows apart
      const uint64_t dk128 = KC == 64 ? smem_desc(0, 16, 1024, SW_128B) : smem_desc(0, 16, 512, SW_64B);
      int job = 0;
      bool ok = true;
      for (int t = 0; t < ntiles && ok; ++t) {
        const int a = t % NACC;
        ok = mbar_wait(&tempty[a], ((t / NACC) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t acc = tmem + a * BN;
        for (int c = 0; c < nchunk && ok; ++c, ++job) {
          const int s = job % KU_NSTAGE;
          ok = mbar_wait(&full[s], (job / KU_NSTAGE) & 1);
          tc_fence_after_sync();
          const uint32_t a0 = smem_u32(ring + s * KU_STAGE) >> 4;
          if (TERMS == 3) {
            const uint64_t ah = dk128 + a0, al = ah + QB / 16, bh = ah + 2 * QB / 16, bl = bh + DB / 16;  // 16-byte units
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
              mma_f16(acc, ah + 2 * ks, bh + 2 * ks, idesc, (c | ks) != 0);
              mma_f16(acc, ah + 2 * ks, bl + 2 * ks, idesc, 1);
              mma_f16(acc, al + 2 * ks, bh + 2 * ks, idesc, 1);
            }
          } else {
            const uint64_t ah = dk128 + a0, bh = ah + QB / 16;
#pragma unroll
            for (int h = 0; h < MQ; ++h)     // query block h: rows 128 h .. of the Q chunk, accumulator columns BN h ..
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks)
                mma_f16(acc + h * BN, ah + h * (128 * KC * 2 / 16) + 2 * ks, bh + 2 * ks, idesc, (c | ks) != 0);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[a]);
      }
      if (!ok && prm.err) atomicCAS(prm.err, 0, 12);
    }
  } else {
    // ---- epilogue: thread = query row (the TMEM row-per-thread layout), see the file header
    const int lg = warp & 3;     // the TMEM lane group this warp may read: lanes 32 * (warp % 4) ..
    const int qh = (warp - 2) >> 2;   // query block of this warp (MQ = 2: warps 2-5 block 0, warps 6-9 block 1)
    const uint32_t tlane = tmem + ((uint32_t)(32 * lg) << 16) + (MQ == 2 ? qh * BN : 0);
    const int mrow0 = m0 + 128 * qh + 32 * lg;   // first query row of this warp
    float* lv = lists + (warp - 2) * (2 * 32 * KU_CAND);
    int* li = reinterpret_cast<int*>(lv + 32 * KU_CAND);
    float thr = -INFINITY;   // smallest value of the full list; -inf while the list still has room
    int minpos = 0, fill = 0;
    auto rescan = [&]() {
      float m = INFINITY;
      int mp = 0;
#pragma unroll 8
      for (int e = 0; e < KU_CAND; ++e) {
        const float x = lv[e * 32 + (lane ^ e)];
        if (x < m) { m = x; mp = e; }
      }
      thr = m;
      minpos = mp;
    };
    if (prm.warm) {   // continue the list a previous launch (same grid, same segments) left in cand_*
      // coalesced: for each row of the warp, lane e fetches entry e
      for (int r = 0; r < 32; ++r) {
        const int q = mrow0 + r;
        float x = -INFINITY;
        int xi = -1;
        if (q < prm.Nq) {
          const size_t o = ((size_t)q * prm.list_pitch + seg) * KU_CAND;
          xi = prm.cand_idx[o + lane];
          x = xi >= 0 ? prm.cand_val[o + lane] : -INFINITY;   // an untouched (memset 0xff) list is an empty list
        }
        lv[lane * 32 + (r ^ lane)] = x;    // slot e = lane of row r
        li[lane * 32 + (r ^ lane)] = xi;
      }
      __syncwarp();
      // a list is stored compacted: its first `fill` slots are taken
      for (int e = 0; e < KU_CAND; ++e) fill += li[e * 32 + (lane ^ e)] >= 0 ? 1 : 0;
      if (fill == KU_CAND) rescan();
    }
    float v[32];
    bool ok = true;
    for (int t = 0; t < ntiles && ok; ++t) {
      const int a = t % NACC;
      ok = mbar_wait(&tfull[a], (t / NACC) & 1);
      tc_fence_after_sync();
      const int n0 = tile_of(t_begin + t) * BN;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        __syncwarp();                                 // rows diverge below; the TMEM load is warp-collective
        tmem_ld_32x32(tlane + a * BN + 32 * cc, v);   // this thread's row, 32 consecutive database columns
        tmem_ld_wait();
        const int nb = n0 + 32 * cc;
        float vmax = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) vmax = fmaxf(vmax, v[i]);
        // (columns past the end of the database are zero-filled by TMA and rows a sharded build has not split yet may
        //  hold anything: the column mask below is what keeps them out)
        if (!(vmax > thr)) continue;
        // Rows with a candidate (about one in ten once the lists are warm) leave the straight-line path: the 32
        // values go to a thread-local staging array so that ONE rolled loop can walk the hit mask.  (32 unrolled
        // copies of the insertion were ~50 KB of code that every chunk had to hop through: instruction-cache misses
        // cost 4x the arithmetic.)
        unsigned hits = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) hits |= (v[i] > thr ? 1u : 0u) << i;
        float stage[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[i] = v[i];
#pragma unroll 1
        while (hits) {
          const int i = __ffs(hits) - 1;
          hits &= hits - 1;
          const float x = stage[i];
          const int col = nb + i;
          const bool inside = col >= prm.col_lo && col < prm.col_hi;
          const bool col_ok = prm.col_mode == 1 ? inside : (col < prm.N && !(prm.col_mode == 2 && inside));
          if (x > thr && col_ok) {          // thr may have risen since the mask was built
            const int slot = fill < KU_CAND ? fill : minpos;
            lv[slot * 32 + (lane ^ slot)] = x;
            li[slot * 32 + (lane ^ slot)] = col;
            if (fill < KU_CAND) ++fill;
            if (fill == KU_CAND) rescan();
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&tempty[a]);
    }
    if (!ok && prm.err) atomicCAS(prm.err, 0, 13);
    // unused slots leave as (-inf, -1); coalesced write-out: for each row of the warp, lane e stores entry e
    for (int e = fill; e < KU_CAND; ++e) {
      lv[e * 32 + (lane ^ e)] = -INFINITY;
      li[e * 32 + (lane ^ e)] = -1;
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const int q = mrow0 + r;
      if (q < prm.Nq) {
        const size_t o = ((size_t)q * prm.list_pitch + seg) * KU_CAND;
        prm.cand_val[o + lane] = lv[lane * 32 + (r ^ lane)];
        prm.cand_idx[o + lane] = li[lane * 32 + (r ^ lane)];
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

// One warp per query row.  The candidate lists carry APPROXIMATE similarities a with |a - exact| <= eps, so:
//   * candidates below (k-th largest a) - 2 eps cannot be in the exact top-k and are dropped;
//   * the survivors (about k + 1 of them) are compacted; two survivors whose a differ by more than 2 eps are ordered
//     for certain, so an exact fp32 dot product is only spent on the AMBIGUOUS ones - those within 2 eps of another
//     survivor - (all of them when the caller wants the similarities back);
//   * k rounds of warp-argmax by (key desc, index asc) emit the top k, key = exact where it was computed, a otherwise
//     (consistent with the exact order: an unambiguous key is more than eps away from every other interval);
//   * CERTIFICATE: a lower bound of the exact k-th value must beat tau + eps, tau = the largest 32nd value of the row's
//     lists, i.e. nothing outside the candidate lists can belong to the top-k.  Rows that fail (true ties, or more than
//     64 survivors) are listed and recomputed by the exact fp32 kernel (knn.cu).
constexpr int KU_MAXSEG = 8;
constexpr int KU_SURV = 64;

__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                         int Nq, int N, int F, int k, int nlists,
                                                         const int* __restrict__ cand_idx,
                                                         const float* __restrict__ cand_val, int64_t* __restrict__ idx,
                                                         float* __restrict__ sims, int* __restrict__ fail_rows,
                                                         int* __restrict__ fail_count, const int* __restrict__ err,
                                                         const float* __restrict__ qnorm2,
                                                         const int* __restrict__ dbmax2_bits, int npeer_max,
                                                         float eps_unit, float eps_abs, float max_norm2) {
  __shared__ float s_val[8][KU_SURV];
  __shared__ int s_idx[8][KU_SURV];
  const int wid = threadIdx.x >> 5;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= Nq) return;
  // error bound of the approximate similarities of this row: eps_unit * |q| * max |d|
  // (dbmax2_bits[0] = the largest squared norm this GPU measured; a sharded build that received its panels ready-made
  //  finds the other ranks' maxima in dbmax2_bits[30 .. 30 + npeer_max): header ints 32.. of the workspace)
  int mbits = __ldg(dbmax2_bits);
  for (int p = 0; p < npeer_max; ++p) mbits = max(mbits, __ldg(dbmax2_bits + 30 + p));
  // (eps_abs: the fast pass's absolute term for values below the fp16 normal range; max_norm2: above it the 16-bit
  //  panels may have overflowed and the approximate pass proves nothing - the row goes to the exact kernel)
  const float qn2 = __ldg(qnorm2 + row), dn2 = __int_as_float(mbits);
  const float eps = eps_unit * sqrtf(qn2) * sqrtf(dn2) + eps_abs * (sqrtf(qn2) + sqrtf(dn2));
  const bool overflow = !(qn2 <= max_norm2 && dn2 <= max_norm2);
  const float* qr = q + (size_t)row * F;
  int my_idx[KU_MAXSEG];
  float my_a[KU_MAXSEG];
  float tau = -INFINITY;  // everything outside the candidate lists is <= tau (approximately)
#pragma unroll
  for (int sg = 0; sg < KU_MAXSEG; ++sg) {
    my_idx[sg] = -1;
    my_a[sg] = -INFINITY;
    if (sg < nlists) {
      const size_t o = ((size_t)row * nlists + sg) * KU_CAND;
      my_idx[sg] = cand_idx[o + lane];
      my_a[sg] = my_idx[sg] >= 0 ? cand_val[o + lane] : -INFINITY;
      // the smallest entry of a (full) list bounds everything its segment dropped; a list with room dropped nothing
      float lmin = my_a[sg];
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o2));
      tau = fmaxf(tau, lmin);
    }
  }
  // a_k = the k-th largest approximate value (duplicates counted)
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
  // compact the survivors into shared memory (warp-private)
  int cnt = 0;
#pragma unroll
  for (int sg = 0; sg < KU_MAXSEG; ++sg) {
    if (sg < nlists) {
      const bool live = my_idx[sg] >= 0 && my_a[sg] >= keep;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (live && pos < KU_SURV) {
        s_val[wid][pos] = my_a[sg];
        s_idx[wid][pos] = my_idx[sg];
      }
      cnt += __popc(m);
    }
  }
  __syncwarp();
  if (cnt > KU_SURV || cnt < k || overflow || (err && *err != 0)) {   // massive ties / broken pipeline: the exact kernel does this row
    if (lane == 0) fail_rows[atomicAdd(fail_count, 1)] = row;
    return;
  }
  // lane owns survivors `lane` and `lane + 32`
  float key[2];
  int cid[2];
  bool amb[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = lane + 32 * h;
    key[h] = e < cnt ? s_val[wid][e] : -INFINITY;
    cid[h] = e < cnt ? s_idx[wid][e] : 0x7fffffff;
    amb[h] = sims != nullptr;
  }
  const float two_eps = 2.f * eps;
  const bool vec4 = (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(db)) & 15) == 0;
  for (int j = 0; j < cnt; ++j) {
    const float aj = s_val[wid][j];   // broadcast
    amb[0] |= (j != lane) && fabsf(key[0] - aj) <= two_eps;
    amb[1] |= (j != lane + 32) && fabsf(key[1] - aj) <= two_eps;
  }
  amb[0] &= lane < cnt;
  amb[1] &= lane + 32 < cnt;
  bool exact[2] = {false, false};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned live = __ballot_sync(0xffffffffu, amb[h]);
    while (live) {
      const int c = __ffs(live) - 1;
      live &= live - 1;
      const float* dr = db + (size_t)s_idx[wid][c + 32 * h] * F;
      float s = 0.f;
      if (vec4) {   // 16-byte loads: a 768-float row is 6 loads per lane, all in flight at once
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int f = lane * 4; f < F; f += 128) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(qr + f));
          const float4 b = __ldg(reinterpret_cast<const float4*>(dr + f));
          s = fmaf(a.x, b.x, s); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
        }
        s = (s + s1) + (s2 + s3);
      } else {
        for (int f = lane; f < F; f += 32) s = fmaf(__ldg(qr + f), __ldg(dr + f), s);
      }
      s = warp_sum(s);
      if (lane == c) {
        key[h] = s;
        exact[h] = true;
      }
    }
  }
  float ek = -INFINITY;   // lower bound of the exact k-th value
  for (int out = 0; out < k; ++out) {
    // this lane's best remaining candidate
    const bool second = cid[1] != 0x7fffffff && (cid[0] == 0x7fffffff || key[1] > key[0] || (key[1] == key[0] && cid[1] < cid[0]));
    float wv = second ? key[1] : key[0];
    int wi = second ? cid[1] : cid[0];
    int wl = lane * 2 + (second ? 1 : 0);
    if (wi == 0x7fffffff) wv = -INFINITY;
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
    const bool was_exact = __shfl_sync(0xffffffffu, (wl & 1) ? exact[1] : exact[0], wl >> 1);
    if (lane == (wl >> 1)) cid[wl & 1] = 0x7fffffff;   // consumed
    if (lane == 0) {
      idx[(size_t)row * k + out] = (int64_t)wi;
      if (sims) sims[(size_t)row * k + out] = wv;
    }
    if (out == k - 1) ek = was_exact ? wv : wv - eps;
  }
  // certificate: the exact k-th best candidate must beat anything the approximate pass could have dropped
  const bool bad = !(ek > tau + eps);
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

// DEPTHG_B200_KNN_PASS = split3 : the 3-term bf16 hi/lo tensor pass (round-1/2 kernel, fp32-grade similarities);
// default                        : the fast single-panel fp16 pass of the precision ladder (TERMS = 1 above).
// Read per call (tests switch it); both give the exact fp32 top-k, they differ in how much the re-rank recomputes.
static bool knn_fast() {
  const char* e = getenv("DEPTHG_B200_KNN_PASS");
  return !(e && strcmp(e, "split3") == 0);
}

// Query blocks per CTA of the tensor pass (KuCfg::MQ).  Default 1.  DEPTHG_B200_KNN_MQ=2 (fast pass only) selects the
// 256 x 256 tile; measured on B200 at N = 49 629: 5.18 ms against 4.83 ms for MQ = 1 — the two 128 x 256 accumulators
// fill TMEM, so nothing overlaps the epilogue (256 KB of TMEM reads at 64 B/cycle alone are 2 us a tile) and the
// operand stream it saves is lost again while the ring waits.  Kept for the record and covered by a GPU test.
static int knn_mq() {
  if (!knn_fast()) return 1;
  const char* e = getenv("DEPTHG_B200_KNN_MQ");
  return e && atoi(e) == 2 ? 2 : 1;
}

// Database segments per query block.  More segments fill the SMs when there are few query blocks (a multi-GPU
// shard) and even out the last wave of a large build; with thread-per-row lists a segment's cold start costs
// microseconds, so what is left is one more 32-entry list per row for the re-rank (~1 % per segment).
// Pick the nseg that minimises  waves x tiles-per-segment x (1 + 0.01 (nseg - 1)).
static int knn_nseg(int Nq, int N, int k) {
  const int nblocks = ceil_div(Nq, 128 * knn_mq()), ntiles = ceil_div(N, 128);
  double best_cost = 1e300;
  (void)k;
  int best = 1;
  for (int nseg = 1; nseg <= KU_MAXSEG && nseg <= max(1, ntiles / 8); ++nseg) {
    const double waves = (double)ceil_div(nblocks * nseg, 148);
    const double cost = waves * (double)ceil_div(ntiles, nseg) * (1.0 + 0.01 * (nseg - 1));
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = nseg;
    }
  }
  // the fast pass wants more than k + 2 candidates per row: its certificate needs the exact k-th value to clear the
  // list minimum by the (wider) error bound, which the 32nd value of a single list rarely allows
  if (knn_fast() && best < 2 && ceil_div(N, 256) >= 2) best = 2;
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

// Workspace layout shared by the one-call build and the two-phase sharded build.
struct KnnWs {
  int *err, *fail_count, *dbmax2;
  __nv_bfloat16 *dh, *dl, *qh, *ql;
  int* cand_idx;
  float* cand_val;
  int* fail_rows;
  float* qnorm2;
};

static KnnWs knn_ws_layout(void* ws, int Nq, int N, int Fp, bool q_in_db, int q_row0) {
  KnnWs k;
  uint8_t* w = static_cast<uint8_t*>(ws);
  k.err = reinterpret_cast<int*>(w);
  // zeroed header: [0] error flag, [1] number of uncertified rows (dg_knn_topk's caller may read it back after the
  // stream has drained), [2] bits of the largest squared database row norm
  k.fail_count = k.err + 1;
  k.dbmax2 = k.err + 2;
  size_t off = 256;
  auto take = [&](size_t bytes) { uint8_t* p = w + off; off += al256(bytes); return p; };
  k.dh = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * Fp * 2));
  k.dl = reinterpret_cast<__nv_bfloat16*>(take((size_t)N * Fp * 2));
  __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(take((size_t)Nq * Fp * 2));
  __nv_bfloat16* ql = reinterpret_cast<__nv_bfloat16*>(take((size_t)Nq * Fp * 2));
  k.qh = q_in_db ? k.dh + (size_t)q_row0 * Fp : qh;   // the query rows ARE database rows q_row0 .. q_row0 + Nq - 1
  k.ql = q_in_db ? k.dl + (size_t)q_row0 * Fp : ql;
  k.cand_idx = reinterpret_cast<int*>(take((size_t)Nq * KU_MAXSEG * KU_CAND * 4));
  k.cand_val = reinterpret_cast<float*>(take((size_t)Nq * KU_MAXSEG * KU_CAND * 4));
  k.fail_rows = reinterpret_cast<int*>(take((size_t)Nq * 4));
  k.qnorm2 = reinterpret_cast<float*>(take((size_t)Nq * 4));
  return k;
}

static int knn_bn() {
  static int bn_env = -1;  // DEPTHG_B200_KNN_BN = 128 | 256 (experiments, 3-term pass only); default 256
  if (bn_env < 0) {
    const char* e = getenv("DEPTHG_B200_KNN_BN");
    bn_env = e && atoi(e) == 128 ? 128 : 256;
  }
  return knn_fast() ? 256 : bn_env;
}

// One launch of the tensor pass over the virtual tile list (see KnnUmmaParams).
static int launch_knn_umma(const KnnWs& w, int Nq, int N, int Fp, int nseg, int list_pitch, int ntiles_virtual, int tA_end,
                           int tB_shift, int col_lo, int col_hi, int col_mode, int warm, cudaStream_t st) {
  KnnUmmaParams prm;
  int rc;
  const bool fast = knn_fast();
  const int BN = knn_bn(), MQ = knn_mq();
  const int KC = fast ? (MQ == 2 ? KuCfg<256, 1, 2>::KC : KuCfg<256, 1>::KC) : (BN == 128 ? KuCfg<128, 3>::KC : KuCfg<256, 3>::KC);
  if ((rc = make_map(&prm.tm_qh, w.qh, Fp, Nq, 128 * MQ, KC))) return rc;
  if ((rc = make_map(&prm.tm_ql, w.ql, Fp, Nq, 128 * MQ, KC))) return rc;
  if ((rc = make_map(&prm.tm_dh, w.dh, Fp, N, BN, KC))) return rc;
  if ((rc = make_map(&prm.tm_dl, w.dl, Fp, N, BN, KC))) return rc;
  prm.Nq = Nq; prm.N = N; prm.nchunk = Fp / KC; prm.ntiles = ntiles_virtual; prm.nseg = nseg;
  prm.list_pitch = list_pitch;
  prm.tA_end = tA_end; prm.tB_shift = tB_shift; prm.col_lo = col_lo; prm.col_hi = col_hi; prm.col_mode = col_mode;
  prm.warm = warm;
  prm.cand_idx = w.cand_idx; prm.cand_val = w.cand_val; prm.err = w.err;
  static PerDevice attr_pd = {};
  size_t& attr_set = per_device(attr_pd);
  if (!attr_set) {
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<128, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<128, 3>::SMEM));
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<256, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<256, 3>::SMEM));
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<256, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<256, 1>::SMEM));
    DG_CUDA_OK(cudaFuncSetAttribute(knn_umma_kernel<256, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, KuCfg<256, 1, 2>::SMEM));
    attr_set = 1;
  }
  DG_PRE(st);
  if (fast && MQ == 2)
    knn_umma_kernel<256, 1, 2><<<dim3(ceil_div(Nq, 256), nseg), KuCfg<256, 1, 2>::THREADS, KuCfg<256, 1, 2>::SMEM, st>>>(prm);
  else if (fast)
    knn_umma_kernel<256, 1, 1><<<dim3(ceil_div(Nq, 128), nseg), KU_THREADS, KuCfg<256, 1>::SMEM, st>>>(prm);
  else if (BN == 128)
    knn_umma_kernel<128, 3, 1><<<dim3(ceil_div(Nq, 128), nseg), KU_THREADS, KuCfg<128, 3>::SMEM, st>>>(prm);
  else
    knn_umma_kernel<256, 3, 1><<<dim3(ceil_div(Nq, 128), nseg), KU_THREADS, KuCfg<256, 3>::SMEM, st>>>(prm);
  DG_LAUNCH_OK("knn_umma_kernel");
  return DG_OK;
}

static int launch_split(const float* x, int n, int F, int Fp, __nv_bfloat16* hi, __nv_bfloat16* lo, float* rownorm2,
                        int* maxnorm2_bits, cudaStream_t st) {
  if (n <= 0) return DG_OK;
  DG_PRE(st);
  if (knn_fast())
    split_rows_kernel<true><<<min(148 * 8, ceil_div(n, 8)), 256, 0, st>>>(x, n, F, Fp, hi, lo, rownorm2, maxnorm2_bits);
  else
    split_rows_kernel<false><<<min(148 * 8, ceil_div(n, 8)), 256, 0, st>>>(x, n, F, Fp, hi, lo, rownorm2, maxnorm2_bits);
  DG_LAUNCH_OK("split_rows_kernel");
  return DG_OK;
}

static int launch_rerank(const KnnWs& w, const float* q, const float* db, int Nq, int N, int F, int k, int nseg,
                         int64_t* idx, float* sims, cudaStream_t st, int npeer_max = 0) {
  DG_PRE(st);
  const bool fast = knn_fast();
  knn_rerank_kernel<<<ceil_div(Nq * 32, 256), 256, 0, st>>>(q, db, Nq, N, F, k, nseg /* merged lists */, w.cand_idx, w.cand_val, idx,
                                                             sims, w.fail_rows, w.fail_count, w.err, w.qnorm2, w.dbmax2,
                                                             npeer_max, fast ? ku_eps_unit_fast(F) : ku_eps_unit(F),
                                                             fast ? ku_eps_abs_fast(F) : 0.f,
                                                             fast ? 1073741824.f /* 2^30 */ : INFINITY);
  DG_LAUNCH_OK("knn_rerank_kernel");
  return launch_knn_exact(q, db, Nq, N, F, k, idx, sims, w.fail_rows, w.fail_count, st);
}

int knn_topk_umma(const float* q, const float* db, int Nq, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                  cudaStream_t st) {
  const int Fp = round_up(F, 64);
  const bool same = (q == db && Nq == N);
  const KnnWs w = knn_ws_layout(ws, Nq, N, Fp, same, 0);
  const int nseg = knn_nseg(Nq, N, k);
  DG_CUDA_OK(cudaMemsetAsync(w.err, 0, 256, st));
  int rc;
  if ((rc = launch_split(db, N, F, Fp, w.dh, w.dl, same ? w.qnorm2 : nullptr, w.dbmax2, st))) return rc;
  if (!same && (rc = launch_split(q, Nq, F, Fp, w.qh, w.ql, w.qnorm2, nullptr, st))) return rc;
  const int ntiles = ceil_div(N, knn_bn());
  if ((rc = launch_knn_umma(w, Nq, N, Fp, nseg, nseg, ntiles, ntiles, 0, 0, 0, 0, 0, st))) return rc;
  return launch_rerank(w, q, db, Nq, N, F, k, nseg, idx, sims, st);
}

// ---- query-row-sharded build in two phases (the all-gather of the database is the one exchange step).
// The rank's query rows are rows [row_lo, row_lo + Nq) of the database.  Phase 1 needs only those local rows: it
// splits them and runs the tensor pass of the local queries against the LOCAL database rows, which is what a rank can
// do while the all-gather is in flight - and it leaves every candidate list warm.  Phase 2 gets the all-gathered
// database, splits the remote rows and continues the same lists over everything but the local rows (no cold-list
// penalty: the lists already hold 32 good candidates), then re-ranks exactly as the one-call build does.
// Phase 1 runs beside the all-gather, whose kernels need SMs of their own: it uses the largest segment count that
// leaves KU_COMM_SMS free (a CTA of the tensor pass owns a whole SM), and leaves the other lists of a row empty.
constexpr int KU_COMM_SMS = 36;

static int knn_nseg_local(int Nq, int nseg) {
  const int nblocks = ceil_div(Nq, 128 * knn_mq());
  int n = (148 - KU_COMM_SMS) / max(nblocks, 1);
  return max(1, min(n, nseg));
}

// `phases` selects which steps a call enqueues, so that a host that moves the panels itself (peer copies over NVLink
// instead of a collective) can put its own work between them:
//   begin : DG_KNN_SPLIT_LOCAL (reset the header, split the local rows into panel rows [row_lo, row_lo + Nq)),
//           DG_KNN_PASS_LOCAL  (tensor pass against the local rows)
//   finish: DG_KNN_SPLIT_REMOTE (split every other row of `db`; skipped when the panels were copied from the peers),
//           DG_KNN_PASS_REMOTE, DG_KNN_RERANK (needs `db`, the fp32 rows)
int knn_shard_begin(const float* local, int Nq, int row_lo, int N, int F, int k, void* ws, int phases, cudaStream_t st) {
  const int Fp = round_up(F, 64), BN = knn_bn();
  const KnnWs w = knn_ws_layout(ws, Nq, N, Fp, true, row_lo);
  const int nseg = knn_nseg(Nq, N, k);
  // (DG_KNN_ALL_SMS: the exchange runs on copy engines, so the local pass may fill the GPU)
  const int nseg_l = (phases & DG_KNN_ALL_SMS) ? nseg : knn_nseg_local(Nq, nseg);
  int rc;
  if (phases & DG_KNN_SPLIT_LOCAL) {
    DG_CUDA_OK(cudaMemsetAsync(w.err, 0, 256, st));
    if ((rc = launch_split(local, Nq, F, Fp, w.qh, w.ql, w.qnorm2, w.dbmax2, st))) return rc;
  }
  if (phases & DG_KNN_PASS_LOCAL) {
    if (nseg_l < nseg) DG_CUDA_OK(cudaMemsetAsync(w.cand_idx, 0xff, (size_t)Nq * nseg * KU_CAND * 4, st));   // empty lists
    const int t0 = row_lo / BN, t1 = ceil_div(row_lo + Nq, BN);
    if ((rc = launch_knn_umma(w, Nq, N, Fp, nseg_l, nseg, t1 - t0, 0, t0, row_lo, row_lo + Nq, 1, 0, st))) return rc;
  }
  return DG_OK;
}

int knn_shard_finish(const float* db, int Nq, int row_lo, int N, int F, int k, int64_t* idx, float* sims, void* ws,
                     int phases, int npeer_max, cudaStream_t st) {
  const int Fp = round_up(F, 64), BN = knn_bn();
  const KnnWs w = knn_ws_layout(ws, Nq, N, Fp, true, row_lo);
  const int nseg = knn_nseg(Nq, N, k);
  const int row_hi = row_lo + Nq;
  int rc;
  if (phases & DG_KNN_SPLIT_REMOTE) {
    if ((rc = launch_split(db, row_lo, F, Fp, w.dh, w.dl, nullptr, w.dbmax2, st))) return rc;
    if ((rc = launch_split(db + (size_t)row_hi * F, N - row_hi, F, Fp, w.dh + (size_t)row_hi * Fp,
                           w.dl + (size_t)row_hi * Fp, nullptr, w.dbmax2, st))) return rc;
  }
  if (phases & DG_KNN_PASS_REMOTE) {
    const int ntiles = ceil_div(N, BN);
    const int tA_end = ceil_div(row_lo, BN);                 // tiles [0, tA_end) hold rows below the local range
    const int tB_begin = max(row_hi / BN, tA_end);           // tiles [tB_begin, ntiles) hold rows above it
    if ((rc = launch_knn_umma(w, Nq, N, Fp, nseg, nseg, tA_end + (ntiles - tB_begin), tA_end, tB_begin - tA_end, row_lo,
                              row_hi, 2, 1, st))) return rc;
  }
  if (phases & DG_KNN_RERANK)
    return launch_rerank(w, db + (size_t)row_lo * F, db, Nq, N, F, k, nseg, idx, sims, st, npeer_max);
  return DG_OK;
}

int knn_panel_count() { return knn_fast() ? 1 : 2; }

// byte offsets of the bf16 hi / lo panels inside the workspace and the pitch of a panel row: what a host needs to
// copy panel rows between the workspaces of different GPUs
void knn_panel_layout(int N, int F, size_t* hi_off, size_t* lo_off, size_t* row_bytes) {
  const size_t Fp = (size_t)round_up(F, 64);
  *hi_off = 256;
  *lo_off = 256 + al256((size_t)N * Fp * 2);
  *row_bytes = Fp * 2;
}

}  // namespace dg
