// Fused correlation loss on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), P = S*S <= 1024.
//
// Same contract as corr_tile_kernel (corr_loss.cu) — replaces helper() for every pair
// and depth_feature_correlation (/root/reference/src/modules.py:1231-1278) — but one
// CTA owns a 128-row tile of a (pair k, image b) problem (the whole problem when P <= 128; two row tiles, each
// walking two 128-column tiles, when 128 < P <= 256; above that a row tile and a GROUP of two column tiles, with
// the row means precomputed by row_means_kernel) and nothing P x P ever leaves the SM:
//
//   warp 0   TMA producer : streams K-chunks of the split panels into a 3 x 64 KB smem ring
//   warp 1   MMA issuer   : one elected lane issues tcgen05.mma; accumulators live in TMEM
//              fd[128x128]  = F1 . F2^T   bf16 hi/lo split, 3 products per K-step (hh + hl + lh)
//              cd[128x128]  = C1 . C2^T   tf32 hi/lo split, 3 products (fp32-grade: the clamp
//                                          indicator must not flip, see DESIGN.md "Precision")
//   warps 2-5 epilogue    : one thread per row p reads its fd / cd rows from TMEM, forms
//              rowmean (pointwise centring), clamp, loss sums and the unit-gradient factor
//              U[p,q] = -(fd' - shift) 1[clamp passes] / (B P^2), writes U as bf16 hi/lo into
//              smem in the canonical 128B-swizzled layout, then
//   warp 1   again        : dC1[p,:] = U . C2n      (A = U K-major,          B = bf16 code rows, MN-major)
//                           dC2[q,:] = U^T . C1n    (A = same U tile, MN-major, B = bf16 code rows, MN-major)
//   warps 2-5 drain dC1 / dC2 from TMEM to HBM.  For the intra pair the depth term repeats
//   the last two steps with U_d = -(s_p s_q - depth_shift) 1[..] / (B P^2).
//
// Every mbarrier wait is bounded; on a timeout the CTA raises an error flag (the losses
// come back NaN) instead of hanging the GPU.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "umma.cuh"

namespace dg {

using namespace umma;

constexpr int UM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane group)
constexpr int UM_EPI = 256;
constexpr int UM_NSTAGE = 3;
constexpr int UM_STAGE = 65536;  // one ring stage; the U tiles (bf16 hi + lo, 2 x 32 KB) alias the stage the loads no longer need
constexpr int UM_SMEM = UM_NSTAGE * UM_STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t TM_FD = 0, TM_CD = 128, TM_D1 = 256, TM_D2 = 384;

struct UmmaParams {
  CUtensorMap tm_fhi, tm_flo;  // fp16 [npairs*B*128, ldf]   box 64 x 128, SWIZZLE_128B
  CUtensorMap tm_chi, tm_clo;  // f32  [npairs*B*128, ldc]   box 32 x 128, SWIZZLE_128B
  CUtensorMap tm_bhi, tm_blo;  // fp16 [npairs*B*128, ldc]   box 32 x 128, SWIZZLE_64B (gradient GEMM operands)
  const float* dsign;          // [B,128*ntile] or null
  const float* dots;           // [npairs,B] <mean row of F1[b], mean row of F2[k,b]> (pointwise) or null
  int npairs, B, P, ldf, ldc, flags, has_depth;
  int ntile;                   // 128-row tiles per panel: Prows / 128
  // work decomposition: a CTA owns row tile ti (of nti) and column group cg (of ncg), i.e. the NT column tiles
  // cg*NT .. cg*NT+NT-1.  S*S <= 256: nti = NT, ncg = 1 (the CTA sees every column and forms the row means itself).
  // S*S > 256 ("dense"): NT = 2, ncg > 1, row means come from row_means_kernel, dC1 is written per column group.
  int prows, nti, ncg, nti_stride, ncg_stride;
  // A dense shape with an odd number of column tiles is covered by TWO launches: groups 0 .. ncg-2 with NT = 2 and the
  // last, half-empty group with NT = 1 (cg_base = its global index).  Both write one partial-sum slot per
  // (pair, image, row tile, global group) and share the completion counter; ncg_total groups exist in all.
  int cg_base, ncg_total;
  const float* rowmean;        // [npairs,B,prows] mean_q fd[p,q] (dense + pointwise) or null
  float depth_shift, inv_cnt;
  float uscale;                // power of two the unit-gradient factor is multiplied with before its fp16 hi/lo split
  float shift[DG_MAX_PAIRS];
  int32_t group[DG_MAX_PAIRS];
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];  // FEATURE panel slots of pair k's operands (default 0 and k)
  float* out8;      // the 8 scalars of the output tuple, written by the last CTA to finish
  int* done;        // CTA completion counter (zeroed by pair_dots_kernel)
  float* dC1;       // [npairs+1,B,Prows,ldc]
  float* dC2;       // [npairs+1,ntile,B,Prows,ldc]: one partial buffer per row tile of the first operand
  float* partials;  // [npairs*B*ntile][4]
  float* cd_out;    // optional dense [npairs,B,P,P]
  float* loss_out;
  float* dd_out;
  float* fd_dbg;    // optional raw fd accumulators [npairs,B,128,128] (tests)
  int* err;         // error flag (0 = ok)
  long long* clk;   // optional per-CTA phase timestamps [grid][16] (dg_debug_set_clock_buffer)
  int dbg;          // timing experiments only (DEPTHG_B200_UMMA_DBG): 1 = no TMA loads, 2 = no MMAs
};

__device__ __forceinline__ void stamp(const UmmaParams& prm, int slot) {
  if (prm.clk) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prm.clk[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ void raise(int* err, int code) {
  if (err) atomicCAS(err, 0, code);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// Write 32 consecutive q-values of row p (as fp16 hi and lo) into the swizzled U tiles.
__device__ __forceinline__ void store_u_chunk(uint8_t* u_hi, uint8_t* u_lo, int p, int cc, const float* u) {
  const uint32_t atom = cc >> 1;
  uint8_t* row_hi = u_hi + atom * 16384 + p * 128;
  uint8_t* row_lo = u_lo + atom * 16384 + p * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = u[8 * i + 2 * e], b = u[8 * i + 2 * e + 1];
      const __half2 hh = __floats2half2_rn(a, b);  // .x = a (low half), .y = b
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
      h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const uint32_t chunk = ((cc & 1) * 4 + i) ^ (p & 7);  // Swizzle<3,4,3>: 16-byte chunk ^= row % 8
    *reinterpret_cast<uint4*>(row_hi + chunk * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row_lo + chunk * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// Drain one 32-row x 32-column accumulator block (this warp's TMEM lanes) to row-major global memory through a
// per-warp smem scratch so that every store instruction writes one full 128-byte line.
__device__ __forceinline__ void drain_block(uint32_t taddr, float* scratch /*[32][33]*/, float* dst_rows /*row 0 of the block*/,
                                            int ldc, int col0, int lane, float* v, float gscale) {
  tmem_ld_32x32(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) scratch[lane * 33 + i] = v[i] * gscale;
  __syncwarp();
#pragma unroll 8
  for (int r = 0; r < 32; ++r) dst_rows[(size_t)r * ldc + col0 + lane] = scratch[r * 33 + lane];
  __syncwarp();
}

// TMEM columns: feature / code correlations of column tile j at 256 j and 256 j + 128.  Once the epilogue has turned
// a tile into U, its columns are dead and are reused by the gradient accumulators: dC1 (summed over j) in fd_0's
// columns, dC2 of column tile j in cd_j's columns.
__device__ __forceinline__ uint32_t col_fd(int j) { return 256u * j; }
__device__ __forceinline__ uint32_t col_cd(int j) { return 256u * j + 128u; }

template <int NT>
__global__ void __launch_bounds__(UM_THREADS, 1) corr_umma_kernel(const __grid_constant__ UmmaParams prm) {
  extern __shared__ uint8_t um_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed as an offset so the pointer stays in the shared address space
  // (a round trip through uintptr_t makes every later access a generic LD/ST instead of LDS/STS)
  uint8_t* ring = um_raw + ((1024u - (smem_u32(um_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + UM_NSTAGE * UM_STAGE);
  uint64_t* full = bars;                  // [UM_NSTAGE]
  uint64_t* empty = bars + UM_NSTAGE;     // [UM_NSTAGE]
  uint64_t* acc_full = bars + 2 * UM_NSTAGE;
  uint64_t* u_ready = acc_full + 1;
  uint64_t* grad_full = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 3);
  uint64_t* g1_full = acc_full + 4;       // NT == 2: first-operand code rows landed
  uint64_t* g2_full = acc_full + 5;       //          second-operand code rows of the current column tile landed
  uint64_t* g2_empty = acc_full + 6;      //          ... and were consumed by the gradient MMAs
  __shared__ float s_red[8][4];
  __shared__ float s_rowsum[2][128];
  __shared__ float s_sign[1024];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) stamp(prm, 0);
  // work item: (pair k, image b, 128-row tile ti of the first operand, column group cg); the CTA walks the NT
  // column tiles gj0 .. gj0+NT-1 of its group itself
  const int Prows = prm.prows;
  const int cg = prm.cg_base + (int)(blockIdx.x % prm.ncg), rt = blockIdx.x / prm.ncg;
  const int ti = rt % prm.nti, kb = rt / prm.nti;
  const int gj0 = cg * 2;                  // groups are two tiles wide (cg = 0 whenever the CTA sees every column)
  const int slot_id = (kb * prm.nti + ti) * prm.ncg_total + cg;
  const int k = kb / prm.B, b = kb - k * prm.B;
  const int nfd = (prm.ldf + 63) / 64, ncd = prm.ldc / 32, nb = prm.ldc / 32;
  const int nop = nfd + ncd;                                  // operand chunks per column tile
  const int J0 = NT * nop;                                    // jobs of the correlation phase
  const int row1 = b * Prows + 128 * ti;                      // first code operand: slot 0, row tile ti
  const int row2 = (k * prm.B + b) * Prows;                   // second code operand: slot k (+ 128 tj)
  const int frow1 = (prm.fs1[k] * prm.B + b) * Prows + 128 * ti;   // feature operands may live in other slots
  const int frow2 = (prm.fs2[k] * prm.B + b) * Prows;              // (DepthContrastiveCorrelationLoss: intra pair)
  const bool fsame_slot = prm.fs1[k] == prm.fs2[k];
  const bool depth_round = prm.has_depth && k == 0;
  const int rounds = depth_round ? 2 : 1;
  // ring stages after the correlation phase: first-operand code rows, second-operand code rows, U
  // (NT == 2 uses fixed 64 KB regions 0 / 1 / 2 of the ring for them, see below)
  const int sC1 = NT == 2 ? 0 : J0 % UM_NSTAGE, sC2 = NT == 2 ? 1 : (J0 + 1) % UM_NSTAGE,
            sU = NT == 2 ? 2 : (J0 + 2) % UM_NSTAGE;
  const int nC2 = (NT == 1) ? 1 : rounds * NT;                // loads of second-operand code rows (reloaded per tile)
  uint8_t* u_hi = ring + sU * UM_STAGE;
  uint8_t* u_lo = u_hi + 32768;

  if (threadIdx.x == 0) {
    for (int s = 0; s < UM_NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(u_ready, UM_EPI);
    mbar_init(grad_full, 1);
    mbar_init(g1_full, 1);
    mbar_init(g2_full, 1);
    mbar_init(g2_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= 64) {
    for (int t = threadIdx.x - 64; t < Prows; t += UM_EPI)
      s_sign[t] = depth_round ? __ldg(prm.dsign + (size_t)b * Prows + t) : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      prefetch_tmap(&prm.tm_fhi); prefetch_tmap(&prm.tm_flo); prefetch_tmap(&prm.tm_chi);
      prefetch_tmap(&prm.tm_clo); prefetch_tmap(&prm.tm_bhi); prefetch_tmap(&prm.tm_blo);
      // fills of a stage so far: t / 3 during the correlation phase (phase = count & 1); afterwards sC1 is filled once
      // more and sC2 once per load of second-operand code rows
      bool ok = true;
      if constexpr (NT == 2) {
        // Two column tiles per CTA: every operand chunk of the first operand is loaded ONCE and multiplied with the
        // chunks of both column tiles.  Stage (96 KB, two of them) = [A hi | A lo | B0 hi | B0 lo | B1 hi | B1 lo].
        constexpr uint32_t ST2 = 98304;
        for (int c = 0; c < nop && ok; ++c) {
          const int s = c & 1;
          ok = mbar_wait(&empty[s], ((c >> 1) & 1) ^ 1);
          if (!ok) break;
          uint8_t* st = ring + s * ST2;
          const bool isf = c < nfd;
          const int c0 = isf ? c * 64 : (c - nfd) * 32;
          const CUtensorMap* mh = isf ? &prm.tm_fhi : &prm.tm_chi;
          const CUtensorMap* ml = isf ? &prm.tm_flo : &prm.tm_clo;
          const bool diag_ok = isf ? fsame_slot : k == 0;   // operands are the same panel: the diagonal tile needs no second load
          const bool same0 = diag_ok && gj0 == ti, same1 = diag_ok && gj0 + 1 == ti;
          mbar_arrive_expect_tx(&full[s], 32768u + (same0 ? 0u : 32768u) + (same1 ? 0u : 32768u));
          const int ra = isf ? frow1 : row1, rb = (isf ? frow2 : row2) + 128 * gj0;
          tma_load_2d(st, mh, &full[s], c0, ra);
          tma_load_2d(st + 16384, ml, &full[s], c0, ra);
          if (!same0) {
            tma_load_2d(st + 32768, mh, &full[s], c0, rb);
            tma_load_2d(st + 49152, ml, &full[s], c0, rb);
          }
          if (!same1) {
            tma_load_2d(st + 65536, mh, &full[s], c0, rb + 128);
            tma_load_2d(st + 81920, ml, &full[s], c0, rb + 128);
          }
        }
        // the ring is re-partitioned into three 64 KB regions (code rows 1, code rows 2, U) once every correlation
        // MMA has finished reading it
        ok = ok && mbar_wait(acc_full, 0);
        if (ok) {
          mbar_arrive_expect_tx(g1_full, (uint32_t)(2 * nb * 8192));
          for (int a = 0; a < nb; ++a) {
            tma_load_2d(ring + a * 8192, &prm.tm_bhi, g1_full, a * 32, row1);
            tma_load_2d(ring + 32768 + a * 8192, &prm.tm_blo, g1_full, a * 32, row1);
          }
        }
        for (int step = 0; step < rounds * NT && ok; ++step) {
          if (step > 0) ok = mbar_wait(g2_empty, (step - 1) & 1);
          if (!ok) break;
          const int r = row2 + 128 * (gj0 + step % NT);
          mbar_arrive_expect_tx(g2_full, (uint32_t)(2 * nb * 8192));
          for (int a = 0; a < nb; ++a) {
            tma_load_2d(ring + 65536 + a * 8192, &prm.tm_bhi, g2_full, a * 32, r);
            tma_load_2d(ring + 65536 + 32768 + a * 8192, &prm.tm_blo, g2_full, a * 32, r);
          }
        }
      } else
      for (int t = 0; t < J0 + 1 + nC2 && ok; ++t) {
        int s, fills;
        if (t < J0) { s = t % UM_NSTAGE; fills = t / UM_NSTAGE; }
        else if (t == J0) { s = sC1; fills = (J0 + 2 - sC1) / UM_NSTAGE; }
        else { s = sC2; fills = (J0 + 2 - sC2) / UM_NSTAGE + (t - J0 - 1); }
        ok = mbar_wait(&empty[s], (fills & 1) ^ 1);
        if (!ok) break;
        uint8_t* st = ring + s * UM_STAGE;
        if (prm.dbg & 1) { mbar_arrive(&full[s]); continue; }
        if (t < J0) {         // operand chunk of column tile tj: [first hi | first lo | second hi | second lo], 16 KB each
          const int tj = t / nop, c = t - tj * nop;
          const bool isf = c < nfd;
          // diagonal tile of a pair whose operands are the same panel: both operands are the same rows, load once
          const bool same = (gj0 + tj == ti) && (isf ? fsame_slot : k == 0);
          const int c0 = isf ? c * 64 : (c - nfd) * 32;
          const CUtensorMap* mh = isf ? &prm.tm_fhi : &prm.tm_chi;
          const CUtensorMap* ml = isf ? &prm.tm_flo : &prm.tm_clo;
          mbar_arrive_expect_tx(&full[s], same ? 32768u : 65536u);
          const int ra = isf ? frow1 : row1, rb = (isf ? frow2 : row2) + 128 * (gj0 + tj);
          tma_load_2d(st, mh, &full[s], c0, ra);
          tma_load_2d(st + 16384, ml, &full[s], c0, ra);
          if (!same) {
            tma_load_2d(st + 32768, mh, &full[s], c0, rb);
            tma_load_2d(st + 49152, ml, &full[s], c0, rb);
          }
        } else {              // gradient operands: bf16 code rows [128 x ldc] as ldc/32 boxes of [128 x 64 B]; hi @0, lo @32 KB
          const int r = (t == J0) ? row1 : row2 + 128 * (gj0 + (t - J0 - 1) % NT);
          mbar_arrive_expect_tx(&full[s], (uint32_t)(2 * nb * 8192));
          for (int a = 0; a < nb; ++a) {
            tma_load_2d(st + a * 8192, &prm.tm_bhi, &full[s], a * 32, r);
            tma_load_2d(st + 32768 + a * 8192, &prm.tm_blo, &full[s], a * 32, r);
          }
        }
      }
      if (!ok) raise(prm.err, 1);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t id_f = instr_desc(FMT_F16, 128, 128, 0, 0);
      const uint32_t id_c = instr_desc(FMT_TF32, 128, 128, 0, 0);
      const uint32_t id_g1 = instr_desc(FMT_F16, 128, (uint32_t)prm.ldc, 0, 1);  // A = U K-major,   B = code rows MN-major
      const uint32_t id_g2 = instr_desc(FMT_F16, 128, (uint32_t)prm.ldc, 1, 1);  // A = U^T MN-major, B = code rows MN-major
      // descriptor templates; the start-address field (bits 0..13, address >> 4) is added per tile / K-step
      const uint64_t dk128 = smem_desc(0, 16, 1024, SW_128B);       // K-major, 128-byte rows
      const uint64_t dmn128 = smem_desc(0, 16384, 1024, SW_128B);   // MN-major, 64-element atoms 16 KB apart
      const uint64_t dmn64 = smem_desc(0, 8192, 512, SW_64B);       // MN-major, 32-element atoms 8 KB apart
      bool ok = true;
      if constexpr (NT == 2) {
        constexpr uint32_t ST2 = 98304;
        for (int c = 0; c < nop && ok; ++c) {
          const int s = c & 1;
          ok = mbar_wait(&full[s], (c >> 1) & 1);
          tc_fence_after_sync();
          const bool isf = c < nfd;
          const bool diag_ok = isf ? fsame_slot : k == 0;
          const uint32_t a0 = smem_u32(ring + s * ST2) >> 4;
          const uint64_t ah = dk128 + a0, al = ah + (16384 >> 4);
#pragma unroll
          for (int tj = 0; tj < 2; ++tj) {
            const bool same = diag_ok && gj0 + tj == ti;
            const uint64_t bh = same ? ah : ah + ((32768 + tj * 32768) >> 4), bl = bh + (16384 >> 4);
            if (isf) {
              const uint32_t acc = tmem + col_fd(tj);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                mma_f16(acc, ah + 2 * ks, bh + 2 * ks, id_f, (c | ks) != 0);
                mma_f16(acc, ah + 2 * ks, bl + 2 * ks, id_f, 1);
                mma_f16(acc, al + 2 * ks, bh + 2 * ks, id_f, 1);
              }
            } else {
              const uint32_t acc = tmem + col_cd(tj);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                mma_tf32(acc, ah + 2 * ks, bh + 2 * ks, id_c, (c != nfd) || ks != 0);
                mma_tf32(acc, ah + 2 * ks, bl + 2 * ks, id_c, 1);
                mma_tf32(acc, al + 2 * ks, bh + 2 * ks, id_c, 1);
              }
            }
          }
          mma_commit(&empty[s]);
        }
      } else
      for (int t = 0; t < J0 && ok; ++t) {
        const int s = t % UM_NSTAGE;
        ok = mbar_wait(&full[s], (t / UM_NSTAGE) & 1);
        tc_fence_after_sync();
        if (prm.dbg & 2) { mbar_arrive(&empty[s]); continue; }
        const int tj = t / nop, c = t - tj * nop;
        const bool same = (gj0 + tj == ti) && (c < nfd ? fsame_slot : k == 0);
        const uint32_t a0 = smem_u32(ring + s * UM_STAGE) >> 4;
        const uint32_t b0 = same ? a0 : a0 + (32768 >> 4);
        const uint64_t ah = dk128 + a0, al = ah + (16384 >> 4), bh = dk128 + b0, bl = bh + (16384 >> 4);
        if (c < nfd) {
          const uint32_t acc = tmem + col_fd(tj);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 4 x 32 B per 128 B row: K = 16 bf16 per instruction
            mma_f16(acc, ah + 2 * ks, bh + 2 * ks, id_f, (c | ks) != 0);
            mma_f16(acc, ah + 2 * ks, bl + 2 * ks, id_f, 1);
            mma_f16(acc, al + 2 * ks, bh + 2 * ks, id_f, 1);
          }
        } else {
          const uint32_t acc = tmem + col_cd(tj);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // K = 8 tf32 per instruction
            mma_tf32(acc, ah + 2 * ks, bh + 2 * ks, id_c, (c != nfd) || ks != 0);
            mma_tf32(acc, ah + 2 * ks, bl + 2 * ks, id_c, 1);
            mma_tf32(acc, al + 2 * ks, bh + 2 * ks, id_c, 1);
          }
        }
        mma_commit(&empty[s]);
      }
      if (prm.dbg & 2) mbar_arrive(acc_full); else mma_commit(acc_full);
      // first-operand code rows (loaded once)
      uint32_t g1 = 0;
      if (ok) {
        ok = NT == 2 ? mbar_wait(g1_full, 0) : mbar_wait(&full[sC1], ((J0 + 2 - sC1) / UM_NSTAGE) & 1);
        g1 = smem_u32(ring + sC1 * UM_STAGE) >> 4;
      }
      const int c2_base = (J0 + 2 - sC2) / UM_NSTAGE;   // fills of sC2 during the correlation phase
      const uint32_t g2 = smem_u32(ring + sC2 * UM_STAGE) >> 4;
      const uint32_t uh = smem_u32(u_hi) >> 4, ul = smem_u32(u_lo) >> 4;
      int step = 0;   // (round, column tile) counter: phase of u_ready / grad_full
      int c2_loaded = 0;
      for (int rd = 0; rd < rounds && ok; ++rd) {
        for (int tj = 0; tj < NT && ok; ++tj, ++step) {
          if (NT == 2) {                                    // second-operand code rows of this column tile
            ok = mbar_wait(g2_full, step & 1);
          } else if (c2_loaded < nC2 && step == 0) {
            ok = mbar_wait(&full[sC2], (c2_base + c2_loaded) & 1);
            ++c2_loaded;
          }
          ok = ok && mbar_wait(u_ready, step & 1);
          tc_fence_after_sync();
          if (prm.dbg & 2) { mbar_arrive(grad_full); continue; }
          const uint32_t d1 = tmem + col_fd(0), d2 = tmem + col_cd(tj);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // dC1[p, d] += U[p, 16 q] . C2n[16 q, d]   (accumulates over column tiles)
            // U as K-major A: 32 B per K-step inside a 128 B row, the second 64-q atom 16 KB further;
            // code rows as MN-major B (n = d contiguous): 16 k-rows (q) per step = 1024 B
            const uint32_t aoff = ((ks >> 2) * 16384 + (ks & 3) * 32) >> 4;
            const uint64_t a_h = dk128 + uh + aoff, a_l = dk128 + ul + aoff;
            const uint64_t b_h = dmn64 + g2 + ks * 64, b_l = b_h + (32768 >> 4);
            mma_f16(d1, a_h, b_h, id_g1, (tj | ks) != 0);
            mma_f16(d1, a_h, b_l, id_g1, 1);
            mma_f16(d1, a_l, b_h, id_g1, 1);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // dC2[q, d] = U[16 p, q]^T . C1n[16 p, d]
            // the same U tile as MN-major A (m = q contiguous): 16 k-rows (p) per step = 2048 B
            const uint64_t a_h = dmn128 + uh + ks * 128, a_l = dmn128 + ul + ks * 128;
            const uint64_t b_h = dmn64 + g1 + ks * 64, b_l = b_h + (32768 >> 4);
            mma_f16(d2, a_h, b_h, id_g2, ks != 0);
            mma_f16(d2, a_h, b_l, id_g2, 1);
            mma_f16(d2, a_l, b_h, id_g2, 1);
          }
          if (NT == 2) {
            if (step + 1 < rounds * NT) mma_commit(g2_empty);   // the next column tile's code rows may overwrite the region
          } else if (c2_loaded < nC2) {
            mma_commit(&empty[sC2]);
          }
          mma_commit(grad_full);
        }
      }
      if (!ok) raise(prm.err, 2);
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // two warps per TMEM lane group; the pair splits the 128 columns of a tile in halves
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;        // 0: columns 0..63, 1: columns 64..127 of a tile
    const int ew = warp - 2;                 // 0..7
    const int row = 32 * lg + lane;          // row within the tile
    const int p = 128 * ti + row;            // sample point (row of fd / cd / dC1)
    const uint32_t tlane = tmem + ((uint32_t)(32 * lg) << 16);
    const int P = prm.P;
    const bool pointwise = prm.flags & DG_FLAG_POINTWISE;
    const float lo = (prm.flags & DG_FLAG_ZERO_CLAMP) ? 0.f : -9999.f;
    const float hi = (prm.flags & DG_FLAG_STABALIZE) ? 0.8f : __int_as_float(0x7f800000);
    float old_mean = 0.f;  // mean of fd over (b,p,q) of this pair = mean_b <mean row F1[b], mean row F2[k,b]>
    if (pointwise && prm.dots) {
      float part = 0.f;
      for (int bb = lane; bb < prm.B; bb += 32) part += __ldg(prm.dots + (size_t)k * prm.B + bb);
      old_mean = warp_sum(part) / (float)prm.B;
    }
    const float sp = s_sign[p];
    // the fp16 panels carry F16_FEAT_SCALE / F16_CODE_SCALE: fd accumulators are FS times too large, the gradient
    // accumulators uscale * F16_CODE_SCALE times; U itself is kept O(1) (no 1/(B P^2) factor) so its fp16 split is normal
    constexpr float FS = 1.f / (F16_FEAT_SCALE * F16_FEAT_SCALE);
    const float inv = (p < P) ? prm.uscale : 0.f;
    const float gscale = prm.inv_cnt / (prm.uscale * F16_CODE_SCALE);
    const float dsh = prm.depth_shift;
    float sum_loss = 0.f, sum_cd = 0.f, sum_dloss = 0.f, sum_dd = 0.f;
    float v[32], c[32];
    uint32_t passmask[NT][2];                 // clamp indicator bits of this thread's row: [column tile][32-column chunk]

    if (threadIdx.x == 64) stamp(prm, 1);
    bool ok = mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    if (threadIdx.x == 64) stamp(prm, 2);
    float c0 = prm.shift[k] - old_mean;      // fd' - shift = fd - rowmean + old_mean - shift = fd - c0
    if (pointwise && prm.rowmean) {          // dense: this CTA only sees its column group, the row means were precomputed
      c0 += __ldg(prm.rowmean + (size_t)kb * Prows + p);
    } else if (pointwise) {                  // rowmean over all q: padded columns are exactly zero, no mask needed
      float s = 0.f;
      for (int tj = 0; tj < NT; ++tj) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          tmem_ld_32x32(tlane + col_fd(tj) + 32 * (2 * half + h2), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s += v[i];
        }
      }
      s_rowsum[half][row] = s * FS;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      c0 += (s_rowsum[0][row] + s_rowsum[1][row]) / (float)P;
    }
    if (threadIdx.x == 64) stamp(prm, 3);
    float* scratch = reinterpret_cast<float*>(u_hi) + ew * (32 * 33);  // per-warp transpose buffer, aliases U while it is dead
    const size_t slab = (size_t)prm.B * Prows * prm.ldc;
    int step = 0;
    for (int rd = 0; rd < rounds; ++rd) {
#pragma unroll
      for (int tj = 0; tj < NT; ++tj) {
        {
          if (step > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // every warp is done with its scratch before U is rewritten
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int cc = 2 * half + h2;
            const int q0 = 128 * (gj0 + tj) + 32 * cc;
            if (rd == 0) {
              // ---- main pass: branch-free.  Padded rows/columns have fd = cd = 0 and depth sign 0, so they add nothing
              //      to the sums; only U needs the explicit mask (inv = 0 for padded rows, tail zeroing for padded columns).
              tmem_ld_32x32(tlane + col_fd(tj) + 32 * cc, v);
              tmem_ld_32x32(tlane + col_cd(tj) + 32 * cc, c);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= FS;
              if (prm.cd_out || prm.loss_out || prm.dd_out || prm.fd_dbg) {   // optional dense outputs (materialize_cd / tests)
                const size_t obase = (((size_t)k * prm.B + b) * P + p) * P;
                if (prm.fd_dbg) {
                  float* dst = prm.fd_dbg + ((size_t)kb * Prows + p) * Prows + q0;
#pragma unroll
                  for (int i = 0; i < 32; ++i) dst[i] = v[i];
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const int q = q0 + i;
                  if (p < P && q < P) {
                    const float cl = fminf(fmaxf(c[i], lo), hi);
                    if (prm.cd_out) prm.cd_out[obase + q] = c[i];
                    if (prm.loss_out) prm.loss_out[obase + q] = -cl * (v[i] - c0);
                    if (prm.dd_out && depth_round) prm.dd_out[((size_t)b * P + p) * P + q] = sp * s_sign[q];
                  }
                }
              }
              if (depth_round) {   // intra pair with the depth term: also the dd sums and the clamp-indicator bits
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float cdv = c[i];
                  const float cl = fminf(fmaxf(cdv, lo), hi);
                  const float f = v[i] - c0;
                  const float dd = sp * s_sign[q0 + i];
                  sum_loss = fmaf(-cl, f, sum_loss);
                  sum_cd += cdv;
                  sum_dloss = fmaf(-cl, dd - dsh, sum_dloss);
                  sum_dd += dd;
                  const bool pass = (cdv >= lo) && (cdv <= hi) && (q0 + i < P);
                  bits |= pass ? (1u << i) : 0u;
                  v[i] = pass ? -f * inv : 0.f;
                }
                passmask[tj][h2] = bits;
              } else {             // lean path of the other 6 pairs
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float cdv = c[i];
                  const float cl = fminf(fmaxf(cdv, lo), hi);
                  const float f = v[i] - c0;
                  sum_loss = fmaf(-cl, f, sum_loss);
                  sum_cd += cdv;
                  v[i] = ((cdv >= lo) && (cdv <= hi)) ? -f * inv : 0.f;
                }
                if (q0 + 32 > P) {   // only the chunk that crosses P has padded columns to zero
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (q0 + i >= P) v[i] = 0.f;
                }
              }
            } else {
              // ---- depth term: U_d = -(s_p s_q - depth_shift) 1[clamp passes] / (B P^2), indicator from the main pass
              const uint32_t bits = passmask[tj][h2];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((bits >> i) & 1u) ? -(sp * s_sign[q0 + i] - dsh) * inv : 0.f;
            }
            store_u_chunk(u_hi, u_lo, row, cc, v);
          }
          if (threadIdx.x == 64 && step == 0) stamp(prm, 4);
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(u_ready);
          ok = ok && mbar_wait(grad_full, step & 1);
          tc_fence_after_sync();
          if (threadIdx.x == 64 && step == 0) stamp(prm, 5);
          // U is dead until it is rewritten: drain through it.  The two warps of a lane group split the column chunks.
          const size_t which = (rd == 0) ? (size_t)k : (size_t)prm.npairs;
          {  // dC2 rows of column tile tj, partial buffer of row tile ti
            float* dbase = prm.dC2 + (which * prm.nti_stride + ti) * slab + ((size_t)b * Prows + 128 * (gj0 + tj) + 32 * lg) * prm.ldc;
            for (int cc = half; cc < ncd; cc += 2) drain_block(tlane + col_cd(tj) + 32 * cc, scratch, dbase, prm.ldc, 32 * cc, lane, v, gscale);
          }
          if (tj == NT - 1) {  // dC1 rows of row tile ti (complete after the last column tile)
            float* dbase = prm.dC1 + (which * prm.ncg_stride + cg) * slab + ((size_t)b * Prows + 128 * ti + 32 * lg) * prm.ldc;
            for (int cc = 1 - half; cc < ncd; cc += 2) drain_block(tlane + col_fd(0) + 32 * cc, scratch, dbase, prm.ldc, 32 * cc, lane, v, gscale);
          }
          if (threadIdx.x == 64 && step == 0) stamp(prm, 6);
          ++step;
        }
      }
    }
    if (!ok) raise(prm.err, 3);
    sum_loss = warp_sum(sum_loss);
    sum_cd = warp_sum(sum_cd);
    sum_dloss = warp_sum(sum_dloss);
    sum_dd = warp_sum(sum_dd);
    if (lane == 0) {
      s_red[ew][0] = sum_loss; s_red[ew][1] = sum_cd; s_red[ew][2] = sum_dloss; s_red[ew][3] = sum_dd;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (ew == 0) {
      if (lane < 4) {
        float t = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += s_red[w8][lane];
        prm.partials[(size_t)slot_id * 4 + lane] = t;
      }
      // ---- fused finalize: the last CTA to get here folds all partial sums into the 8 output scalars.
      //      __syncwarp orders the four partial stores before lane 0's acq_rel counter increment, which publishes them
      //      at gpu scope without a full __threadfence (that would also invalidate L1)
      __syncwarp();
      int ticket = 0;
      if (lane == 0)
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(prm.done) : "memory");
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      if (ticket == prm.npairs * prm.B * prm.nti * prm.ncg_total - 1) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const int per_pair = prm.B * prm.nti * prm.ncg_total;
        const int total = prm.npairs * per_pair;
        for (int e = lane; e < total; e += 32) {   // fixed order -> deterministic
          const int kk = e / per_pair;
          const int g = prm.group[kk];
          const float l = __ldcg(prm.partials + (size_t)e * 4), c2 = __ldcg(prm.partials + (size_t)e * 4 + 1);
          if (g == DG_GROUP_INTRA) { acc[0] += l; acc[1] += c2; }
          else if (g == DG_GROUP_INTER) { acc[2] += l; acc[3] += c2; }
          else { acc[4] += l; acc[5] += c2; }
          if (kk == 0) {
            acc[6] += __ldcg(prm.partials + (size_t)e * 4 + 2);
            acc[7] += __ldcg(prm.partials + (size_t)e * 4 + 3);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
        int cnt[3] = {0, 0, 0};
        for (int kk = 0; kk < prm.npairs; ++kk) cnt[prm.group[kk] > 2 ? 2 : prm.group[kk]]++;
        const float elems = (float)prm.B * (float)prm.P * (float)prm.P;
        if (lane < 8) {
          const int g = lane >> 1;
          const float n = (g < 3) ? (float)cnt[g] : (prm.has_depth ? 1.f : 0.f);
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == lane) t = acc[i];
          float r = n > 0.f ? t / (n * elems) : 0.f;
          if (*reinterpret_cast<volatile int*>(prm.err) != 0) r = __int_as_float(0x7fc00000);  // pipeline timeout
          prm.out8[lane] = r;
        }
      }
    }
  }
  if (threadIdx.x == 64) stamp(prm, 9);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (threadIdx.x == 64) stamp(prm, 10);
}

// dots[k,b] = < mean_p F1n[b,p,:], mean_q F2n[k,b,q,:] >  (one 256-thread block each, all loads in flight at once);
// block 0 also clears the error flag.  The one-call loss path runs the same body as extra CTAs of the code gather
// (kernels.cuh: DotsJob) and skips this launch.
struct SlotMap {
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];
};

__global__ void __launch_bounds__(256) pair_dots_kernel(const __grid_constant__ DotsJob job) {
  pair_dots_body(job, blockIdx.x);
}

void umma_ws_layout(void* ws, int** err, float** dots) {
  *err = static_cast<int*>(ws);
  *dots = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256);
}

// Dense shapes (S*S > 256): rowmean[k,b,p] = mean_q fd[p,q] = < F1n[p,:], mean row of F2n >, the same identity
// pair_dots_kernel uses for the pair mean.  One block per (pair, image); the F1 rows are read back from the bf16
// hi/lo panels (x = hi + lo is what the tensor cores see).
__global__ void __launch_bounds__(256) row_means_kernel(const __half* __restrict__ fhi,
                                                        const __half* __restrict__ flo,
                                                        const float* __restrict__ fmean, int nsplit, int B, int P,
                                                        int prows, int ldf, float* __restrict__ rowmean,
                                                        const __grid_constant__ SlotMap sm, int interleave) {
  extern __shared__ float m2[];  // [ldf] mean row of the second operand
  const int kb = blockIdx.x, k = kb / B, b = kb - k * B;
  const float* mp = fmean + ((size_t)sm.fs2[k] * B + b) * nsplit * ldf;
  for (int c = threadIdx.x; c < ldf; c += 256) {
    float a = 0.f;
    for (int i = 0; i < nsplit; ++i) a += __ldg(mp + (size_t)i * ldf + c);
    m2[c] = a;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = ((size_t)sm.fs1[k] * B + b) * prows;
  for (int p = warp; p < prows; p += 8) {
    if (p >= P) {   // padded rows must read as exact zeros (their U entries are masked by multiplying with 0)
      if (lane == 0) rowmean[(size_t)kb * prows + p] = 0.f;
      continue;
    }
    // split panels: hi / lo rows of pitch ldf; interleaved: one row of pitch 2 ldf, 32-channel chunks [32 hi | 32 lo]
    const __half* hrow = interleave ? fhi + (base + p) * 2 * ldf : fhi + (base + p) * ldf;
    const __half* lrow = interleave ? hrow + 32 : flo + (base + p) * ldf;
    float s = 0.f;
    for (int c2 = lane; c2 < ldf / 2; c2 += 32) {
      const size_t off = interleave ? il_col(2 * c2) : (size_t)(2 * c2);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(hrow + off));
      const float2 d = __half22float2(*reinterpret_cast<const __half2*>(lrow + off));
      s = fmaf(a.x + d.x, m2[2 * c2], s);
      s = fmaf(a.y + d.y, m2[2 * c2 + 1], s);
    }
    s = warp_sum(s) * (1.f / F16_FEAT_SCALE);
    if (lane == 0) rowmean[(size_t)kb * prows + p] = s;
  }
}

static long long* g_clk = nullptr;  // debug: device buffer for per-CTA phase timestamps
void set_clock_buffer(long long* p) { g_clk = p; }
long long* get_clock_buffer() { return g_clk; }

// dots[k,b] (+ reset of the error flag / completion counter) as its own launch: the per-stage dg_corr_loss entry point
int launch_pair_dots(const DotsJob& job, cudaStream_t st) {
  DG_PRE(st);
  pair_dots_kernel<<<job.npairs * job.B, 256, 0, st>>>(job);
  DG_LAUNCH_OK("pair_dots_kernel");
  return DG_OK;
}

int launch_row_means(const void* f_hi, const void* f_lo, const float* fmean, int nsplit, int npairs, int B, int P, int Prows,
                     int ldf, float* rowmean, const int32_t* fs1, const int32_t* fs2, cudaStream_t st, int interleave) {
  SlotMap sm;
  for (int k = 0; k < npairs; ++k) { sm.fs1[k] = fs1[k]; sm.fs2[k] = fs2[k]; }
  DG_PRE(st);
  row_means_kernel<<<npairs * B, 256, (size_t)ldf * sizeof(float), st>>>(
      static_cast<const __half*>(f_hi), static_cast<const __half*>(f_lo), fmean, nsplit, B, P, Prows, ldf, rowmean, sm,
      interleave);
  DG_LAUNCH_OK("row_means_kernel");
  return DG_OK;
}

// ------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t cols,
                       uint64_t rows, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  DG_REQUIRE(fn, DG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (uint64_t)elt_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DG_REQUIRE(r == CUDA_SUCCESS, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DG_OK;
}

int corr_loss_umma(const dg_panels_t* pan, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P, int Prows, int ldf,
                   int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift, int flags,
                   float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out, float* fd_dbg,
                   void* ws, cudaStream_t st, const int32_t* fslot1, const int32_t* fslot2, int nfslots, bool dots_done) {
  UmmaParams prm;
  const int ntile = Prows / 128;
  const bool dense = P > 256;   // more column tiles than TMEM holds at once: column groups of two tiles
  const int nti = ceil_div(P, 128), ncg = dense ? ceil_div(P, 256) : 1;
  const uint64_t rows = (uint64_t)npairs * B * Prows;
  const uint64_t frows = (uint64_t)(nfslots > 0 ? nfslots : npairs) * B * Prows;
  int rc;
  if ((rc = make_map_2d(&prm.tm_fhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->f_hi, ldf, frows, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_flo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->f_lo, ldf, frows, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_chi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pan->c_hi, ldc, rows, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_clo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pan->c_lo, ldc, rows, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->cb_hi, ldc, rows, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_map_2d(&prm.tm_blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->cb_lo, ldc, rows, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  // workspace: [err int (256 B)][dots npairs*B floats, padded to 256 B][partials npairs*B*4 floats]
  int* err;
  float* dots;
  umma_ws_layout(ws, &err, &dots);
  const size_t dots_bytes = ((size_t)npairs * B * sizeof(float) + 255) / 256 * 256;
  float* partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256 + dots_bytes);
  const size_t part_bytes = ((size_t)npairs * B * nti * ncg * 4 * sizeof(float) + 255) / 256 * 256;
  float* rowmean = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256 + dots_bytes + part_bytes);
  prm.dsign = dsign;
  prm.dots = (flags & DG_FLAG_POINTWISE) ? dots : nullptr;
  prm.npairs = npairs; prm.B = B; prm.P = P; prm.ldf = ldf; prm.ldc = ldc; prm.flags = flags;
  prm.ntile = ntile;
  // odd number of column tiles: the last group holds one real tile -> its own NT = 1 launch (no all-padding tile)
  const bool odd_tail = dense && (nti & 1);
  prm.prows = Prows; prm.nti = nti; prm.ncg = odd_tail ? ncg - 1 : ncg;
  prm.cg_base = 0; prm.ncg_total = ncg;
  prm.nti_stride = ntile; prm.ncg_stride = dense ? Prows / 256 : 1;
  prm.rowmean = (dense && (flags & DG_FLAG_POINTWISE)) ? rowmean : nullptr;
  prm.has_depth = dsign != nullptr;
  prm.depth_shift = depth_shift;
  prm.inv_cnt = 1.0f / ((float)B * (float)P * (float)P);
  {  // |fd' - shift| <= 3 + |shift| must stay inside fp16 after scaling: 2^4 for every sane shift, smaller for huge ones
    float big = fabsf(depth_shift);
    for (int k = 0; k < npairs; ++k) big = fmaxf(big, fabsf(pair_shift[k]));
    int e = 4;
    while (e > -12 && ldexpf(1.f, e) * (big + 4.f) > 16384.f) --e;
    DG_REQUIRE(ldexpf(1.f, e) * (big + 4.f) <= 16384.f, DG_ERR_UNSUPPORTED, "corr_loss: shift %g too large", (double)big);
    prm.uscale = ldexpf(1.f, e);
  }
  for (int k = 0; k < npairs; ++k) {
    prm.shift[k] = pair_shift[k];
    prm.group[k] = pair_group[k];
    prm.fs1[k] = fslot1 ? fslot1[k] : 0;
    prm.fs2[k] = fslot2 ? fslot2[k] : k;
  }
  prm.out8 = out8;
  prm.done = err + 1;
  prm.dC1 = dC1; prm.dC2 = dC2; prm.partials = partials;
  prm.cd_out = cd_out; prm.loss_out = loss_out; prm.dd_out = dd_out; prm.fd_dbg = fd_dbg; prm.err = err;
  prm.clk = g_clk;
  {
    const char* e = getenv("DEPTHG_B200_UMMA_DBG");
    prm.dbg = e ? atoi(e) : 0;
  }
  if (!dots_done) {
    DotsJob job;
    job.fmean = (flags & DG_FLAG_POINTWISE) ? fmean : nullptr;
    job.dots = dots; job.err = err; job.nsplit = nsplit; job.npairs = npairs; job.B = B; job.ldf = ldf;
    for (int k = 0; k < DG_MAX_PAIRS; ++k) { job.fs1[k] = k < npairs ? prm.fs1[k] : 0; job.fs2[k] = k < npairs ? prm.fs2[k] : 0; }
    DG_PRE(st);
    pair_dots_kernel<<<npairs * B, 256, 0, st>>>(job);
    DG_LAUNCH_OK("pair_dots_kernel");
  }
  if (prm.rowmean) {
    SlotMap sm;
    for (int k = 0; k < npairs; ++k) { sm.fs1[k] = prm.fs1[k]; sm.fs2[k] = prm.fs2[k]; }
    DG_PRE(st);
    row_means_kernel<<<npairs * B, 256, (size_t)ldf * sizeof(float), st>>>(
        static_cast<const __half*>(pan->f_hi), static_cast<const __half*>(pan->f_lo), fmean, nsplit, B, P,
        Prows, ldf, rowmean, sm, 0);
    DG_LAUNCH_OK("row_means_kernel");
  }
  static PerDevice attr_pd = {};
  size_t& attr_set = per_device(attr_pd);
  if (!attr_set) {
    DG_CUDA_OK(cudaFuncSetAttribute(corr_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM));
    DG_CUDA_OK(cudaFuncSetAttribute(corr_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM));
    attr_set = 1;
  }
  DG_PRE(st);
  if (ntile == 1) corr_umma_kernel<1><<<npairs * B, UM_THREADS, UM_SMEM, st>>>(prm);
  else corr_umma_kernel<2><<<npairs * B * nti * prm.ncg, UM_THREADS, UM_SMEM, st>>>(prm);
  DG_LAUNCH_OK("corr_umma_kernel");
  if (odd_tail) {
    prm.ncg = 1;
    prm.cg_base = ncg - 1;
    DG_PRE(st);
    corr_umma_kernel<1><<<npairs * B * nti, UM_THREADS, UM_SMEM, st>>>(prm);
    DG_LAUNCH_OK("corr_umma_kernel");
  }
  return DG_OK;  // out8 is written by the last CTA of corr_umma_kernel
}

}  // namespace dg
