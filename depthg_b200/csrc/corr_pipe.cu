// Fused correlation loss on tcgen05 / TMEM / TMA as a PERSISTENT, double-buffered pipeline (P = S*S <= 1024).
//
// Same contract as corr_umma_kernel (corr_umma.cu) and corr_tile_kernel (corr_loss.cu): replaces helper() for every
// pair and depth_feature_correlation (/root/reference/src/modules.py:1231-1278) with nothing P x P leaving the SM.
// What changed against the round-1 kernel, and why (profiles/r02_phases.txt): a CTA there ran load -> MMA -> epilogue
// -> gradient GEMMs -> drain strictly in series (18 us of operand streaming, then 13 us in which the SM's TMA and
// tensor pipes idle), 224 CTAs made 1.5 waves, and the operand streams of all SMs stalled and resumed in lockstep.
// Here one CTA per SM walks a list of work items
//     item = (pair k, image b, 128-row tile ti of the first operand, 128-column tile tj of the second)
// and TMEM holds TWO accumulator sets (2 x [fd 128 cols | cd 128 cols]), so the operand stream and the MMAs of item
// n+1 run underneath the epilogue, the gradient GEMMs and the drain of item n:
//
//   warp 0   TMA producer : one in-order stream of 32 KB stages [A | B] (128 rows x [32 hi | 32 lo] fp16 each: the
//              16-bit panels are interleaved per 32-channel chunk, so a 128-byte SWIZZLE_128B box row carries both
//              halves) through a 3-stage ring: the code chunks of the item (cd operands), then its feature chunks;
//              interleaved at fixed positions: the gradient-GEMM operand fills of the PREVIOUS item (code rows -> codebuf)
//   warp 1   MMA issuer   : cd = C1.C2^T and fd = F1.F2^T as 3-term fp16 hi/lo products (hh + hl + lh, fp32-grade:
//              panels are scaled fp16, kernels.cuh) into TMEM set n&1; at fixed positions inside item n+1's chunk
//              stream it waits for item n's U tile and issues dC1 = U.C2n, dC2 = U^T.C1n into the columns fd/cd vacated
//   warps 2-9 epilogue    : item n: row means (pointwise), clamp, loss sums, U = -(fd' - shift) 1[clamp passes] as fp16
//              hi/lo into the swizzled U tile; after the gradient GEMMs: drain dC1 / dC2 (scaled by 1/(B P^2)) to HBM;
//              for the intra pair the depth term repeats the U / gradient / drain steps
//
// Gradients of different column tiles (dC1) / row tiles (dC2) go to per-tile partial buffers that
// gather_norm_bwd_kernel sums.  Every mbarrier wait is bounded; a timeout raises the error flag (losses come back NaN).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "umma.cuh"
#include "fps_body.cuh"

namespace dg {

using namespace umma;

constexpr int CP_THREADS = 352;  // warp 0 TMA (operand chunks), warp 1 MMA, warps 2..9 epilogue (two per TMEM lane group), warp 10 TMA (codebuf fills)
constexpr int CP_EPI = 256;
constexpr int CP_NSTAGE = 3;
constexpr int CP_STAGE = 32768;     // [A | B], 16 KB each: 128 rows x (32 hi | 32 lo) fp16 = 128-byte swizzled rows
constexpr int CP_U = 65536;         // U hi (32 KB) + U lo (32 KB); doubles as the drain's transpose scratch
constexpr int CP_CODE = 32768;      // gradient-GEMM operand: ONE half (hi or lo) of 128 code rows at a time (ldc <= 128)
constexpr int CP_NCODE = 2;         // ... double-buffered: the fill of step s+1 is in flight while step s multiplies (a
                                    // single buffer made the 4-8 fills of an item a chain of L2 round trips: 9-15 us)
constexpr int CP_SMEM = CP_NSTAGE * CP_STAGE + CP_U + CP_NCODE * CP_CODE + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(CP_SMEM + 1664 /*static: s_red, s_rowsum, s_sign*/ <= 232448, "corr_pipe_kernel: shared memory over the 227 KB CTA limit");

struct PipeParams {
  // 16-bit panels are interleaved per 32-channel chunk [32 hi | 32 lo] (kernels.cuh: il_col), so one 128-byte TMA
  // box row carries both halves of a chunk
  CUtensorMap tm_f;    // fp16 [nfslots*B*Prows, 2 ldf]  box 64 x 128, SWIZZLE_128B : feature chunk (hi | lo)
  CUtensorMap tm_c;    // fp16 [npairs*B*Prows, 2 ldc]   box 64 x 128, SWIZZLE_128B : code chunk (cd operands)
  CUtensorMap tm_cg;   // same panel, box 32 x 128, SWIZZLE_64B : one half (hi or lo) of a chunk (gradient-GEMM operands)
  const float* dsign;          // [B,Prows] or null
  const float* dots;           // [npairs,B] <mean row of F1[b], mean row of F2[k,b]> (pointwise) or null
  const float* rowmean;        // [npairs,B,Prows] mean_q fd[p,q] when an item does not see every column (nt > 1), else null
  int npairs, B, P, prows, ldf, ldc, flags, has_depth;
  int nt;                      // 128-wide tiles that contain real points: ceil(P / 128)
  int ntile;                   // partial-buffer pitch: Prows / 128
  int nitems;                  // npairs * B * nt * nt
  float depth_shift, inv_cnt, uscale;
  float shift[DG_MAX_PAIRS];
  int32_t group[DG_MAX_PAIRS];
  int32_t fs1[DG_MAX_PAIRS], fs2[DG_MAX_PAIRS];  // FEATURE panel slots of pair k's operands (default 0 and k)
  float* out8;
  int* done;        // item completion counter (zeroed by the pair_dots work)
  float* dC1;       // [npairs+1, ntile(tj), B, Prows, ldc]
  float* dC2;       // [npairs+1, ntile(ti), B, Prows, ldc]
  float* partials;  // [nitems][4]
  float* cd_out;    // optional dense [npairs,B,P,P]
  float* loss_out;
  float* dd_out;
  float* fd_dbg;    // optional raw fd [npairs,B,Prows,Prows] (tests)
  int* err;
  long long* clk;   // optional phase stamps of each CTA's first item [grid][16]
  int l2_hints;              // L2 eviction-priority hints on the operand loads (DEPTHG_B200_L2HINTS=1; measured: slower, off by default)
  int dbg_mma;               // timing experiment (DEPTHG_B200_PIPE_MMA): 1 = hi.hi product only, 2 = no chunk MMAs (results garbage)
  const uint8_t* dbg_bulk;   // timing experiment (DEPTHG_B200_PIPE_BULK): stream stages as 1-D bulk copies from here (results garbage)
  // The NEXT step's sampling riding along (dg_loss_io_t::next_*): CTAs nmain .. nmain + nride - 1 of the grid run the
  // FPS launch of the next step (one image each, + one CTA for its permutations).  They become resident when the
  // first one-item CTAs of the item list exit - 224 items on 148 CTAs leave 72 SMs idle for the second half of this
  // kernel - so the next step starts at its gathers: FPS leaves the step's critical path without a launch, a side
  // stream or an SM taken from anybody.
  int nmain, nride;
  FpsArgs ride;
};

struct Item {
  int k, b, ti, tj, kb, pidx;
  bool fsame, csame, depth;   // feature operands identical / code operands identical (diagonal tile of a self pair) / depth round
};

__device__ __forceinline__ Item decode_item(const PipeParams& prm, int it) {
  Item w;
  const int nt2 = prm.nt * prm.nt;
  // Pair-major (the seven CTAs that share an image's first operand are spread over time: image-major order, where
  // they fetch it at the same moment, measured 5 us slower), LAST image first: the gathers wrote the panels image by
  // image, so the last images' panels are the ones still in L2.
  const int g = it / nt2;                        // (pair, image) group
  const int r = it - g * nt2;
  w.ti = r / prm.nt;
  w.tj = r - w.ti * prm.nt;
  w.k = g / prm.B;
  w.b = prm.B - 1 - (g - w.k * prm.B);
  w.kb = w.k * prm.B + w.b;
  w.pidx = w.kb * nt2 + r;                       // canonical (pair-major) index of the item's partial sums
  w.fsame = prm.fs1[w.k] == prm.fs2[w.k] && w.ti == w.tj;
  w.csame = w.k == 0 && w.ti == w.tj;
  w.depth = prm.has_depth && w.k == 0;
  return w;
}

__device__ __forceinline__ void pstamp(const PipeParams& prm, int slot) {
  if (prm.clk) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prm.clk[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ void praise(int* err, int code) {
  if (err) atomicCAS(err, 0, code);
}

// where, inside item n+1's feature-chunk stream, the previous item's gradient GEMM groups are issued (chunk index they
// precede) and where the producer issues the second code fill: fractions of the chunk count, see the file header
// A ride-along CTA (PipeParams::ride).  Deliberately NOT inlined: with the FPS body inside corr_pipe_kernel the
// compiler's code for the correlation roles came out 8 us slower (79 vs 71 us at cfg2).
__device__ __noinline__ void ride_cta(const FpsArgs& a, int c, uint8_t* smem, long long* clk) {
  pdl_trigger();
  pdl_wait();
  if (clk && threadIdx.x == 0) {   // phase stamps (dg_debug_set_clock_buffer): slot 1 = start, slot 9 = end of the CTA
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    clk[(size_t)blockIdx.x * 16 + 1] = t;
  }
  if (c == a.nimg) {           // (only present when a.pj.n > 0) all threads of the CTA draw the permutations
    super_perms_block(a.pj.seed, a.pj.offset, a.pj.n, a.pj.B, a.pj.out, reinterpret_cast<int*>(smem));
  } else if (threadIdx.x < FPS_THREADS) {
    if (a.stage)
      fps_cta<4, 256, 4, true>(a, c, reinterpret_cast<float*>(smem));
    else
      fps_cta<4, 256, 4, false>(a, c, reinterpret_cast<float*>(smem));
  }
  if (clk && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    clk[(size_t)blockIdx.x * 16 + 9] = t;
  }
}

__global__ void __launch_bounds__(CP_THREADS, 1) corr_pipe_kernel(const __grid_constant__ PipeParams prm) {
  extern __shared__ uint8_t cp_raw[];
  if ((int)blockIdx.x >= prm.nmain) {   // ride-along CTA: the next step's FPS (see PipeParams::ride)
    ride_cta(prm.ride, (int)blockIdx.x - prm.nmain, cp_raw, prm.clk);
    return;
  }
  uint8_t* ring = cp_raw + ((1024u - (smem_u32(cp_raw) & 1023u)) & 1023u);   // 1024-byte aligned, stays an smem pointer
  uint8_t* u_hi = ring + CP_NSTAGE * CP_STAGE;
  uint8_t* u_lo = u_hi + 32768;
  uint8_t* codebuf = u_hi + CP_U;
  uint64_t* bars = reinterpret_cast<uint64_t*>(codebuf + CP_NCODE * CP_CODE);
  uint64_t* full = bars;                     // [CP_NSTAGE] stage landed
  uint64_t* empty = bars + CP_NSTAGE;        // [CP_NSTAGE] stage consumed by the MMAs
  uint64_t* acc_full = bars + 2 * CP_NSTAGE; // [2] fd/cd of TMEM set ready
  uint64_t* acc_free = acc_full + 2;         // [2] TMEM set drained by the epilogue
  uint64_t* u_ready = acc_full + 4;          // U tile written
  uint64_t* grad_full = acc_full + 5;        // gradient accumulators of a round ready
  uint64_t* g_full = acc_full + 6;           // [CP_NCODE] codebuf fill landed
  uint64_t* g_free = acc_full + 8;           // [CP_NCODE] codebuf fill consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 10);
  __shared__ float s_red[8][4];
  __shared__ float s_rowsum[2][128];
  __shared__ float s_sign[128];   // depth signs of the item's column tile (the row's own sign is read from global)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nfd = prm.ldf / 32, ncd = prm.ldc / 32, nb = prm.ldc / 32;
  const int Prows = prm.prows;
  const int first = blockIdx.x, stride = prm.nmain;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CP_NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_free[a], 1);
    }
    mbar_init(u_ready, CP_EPI);
    mbar_init(grad_full, 1);
    for (int c = 0; c < CP_NCODE; ++c) {
      mbar_init(&g_full[c], 1);
      mbar_init(&g_free[c], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // barriers and TMEM are set up while the gathers drain; the panels are only read from here on

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      prefetch_tmap(&prm.tm_f); prefetch_tmap(&prm.tm_c); prefetch_tmap(&prm.tm_cg);
      bool ok = true;
      int job = 0;      // ring stages issued so far
      // one ring stage: 32-channel chunk c (panel columns 64 c .. 64 c + 63) of rows ra (A) and rb (B)
      auto chunk = [&](const CUtensorMap* m, int c, int ra, int rb, bool same) {
        const int s = job % CP_NSTAGE;
        ok = ok && mbar_wait(&empty[s], ((job / CP_NSTAGE) & 1) ^ 1);
        if (!ok) return;
        uint8_t* st = ring + s * CP_STAGE;
        mbar_arrive_expect_tx(&full[s], same ? 16384u : 32768u);
        if (prm.dbg_bulk) {
          const size_t off = ((size_t)(blockIdx.x * 977 + job) * 32768) % (size_t)(64u << 20);
          bulk_load(st, prm.dbg_bulk + off, same ? 16384u : 32768u, &full[s]);
        } else {
          // the first operand's panel (slot 0 / the image's own code) is read by every pair of the image: keep it;
          // the second operand's is read by this item only: let it go first
          if (prm.l2_hints) {
            tma_load_2d_hint(st, m, &full[s], 64 * c, ra, L2_EVICT_LAST);
            if (!same) tma_load_2d_hint(st + 16384, m, &full[s], 64 * c, rb, L2_EVICT_FIRST);
          } else {
            tma_load_2d(st, m, &full[s], 64 * c, ra);
            if (!same) tma_load_2d(st + 16384, m, &full[s], 64 * c, rb);
          }
        }
        ++job;
      };
      for (int it = first; it < prm.nitems && ok; it += stride) {
        const Item w = decode_item(prm, it);
        const int row1 = w.b * Prows + 128 * w.ti;                       // first code operand: slot 0
        const int row2 = (w.k * prm.B + w.b) * Prows + 128 * w.tj;       // second code operand: slot k
        const int frow1 = (prm.fs1[w.k] * prm.B + w.b) * Prows + 128 * w.ti;
        const int frow2 = (prm.fs2[w.k] * prm.B + w.b) * Prows + 128 * w.tj;
        for (int c = 0; c < ncd && ok; ++c) chunk(&prm.tm_c, c, row1, row2, w.csame);
        for (int c = 0; c < nfd && ok; ++c) chunk(&prm.tm_f, c, frow1, frow2, w.fsame);
      }
      if (!ok) praise(prm.err, 1);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t id_c = instr_desc(FMT_F16, 128, 128, 0, 0);
      const uint32_t id_g1 = instr_desc(FMT_F16, 128, (uint32_t)prm.ldc, 0, 1);  // A = U K-major,    B = code rows MN-major
      const uint32_t id_g2 = instr_desc(FMT_F16, 128, (uint32_t)prm.ldc, 1, 1);  // A = U^T MN-major, B = code rows MN-major
      const uint64_t dk128 = smem_desc(0, 16, 1024, SW_128B);       // K-major, 128-byte rows: operand chunks, and U as A
      const uint64_t dmn128 = smem_desc(0, 16384, 1024, SW_128B);   // U as MN-major A (64-element atoms 16 KB apart)
      const uint64_t dmn64 = smem_desc(0, 8192, 512, SW_64B);       // code rows as MN-major B (32-element atoms 8 KB apart)
      const uint32_t uh = smem_u32(u_hi) >> 4, ul = smem_u32(u_lo) >> 4, gb = smem_u32(codebuf) >> 4;
      bool ok = true;
      int job = 0, fills = 0, nu = 0;   // ring stages / codebuf fills / U tiles consumed so far
      long long w_full = 0, w_free = 0, w_steps = 0, w_issue = 0, w_commit = 0, w_poll = 0, w_grad = 0, t_begin = clock64();   // where this thread's time goes (phase-stamp mode only)
      // 3-term product of one ring stage into accumulator `acc`
      auto chunk_mma = [&](uint32_t acc, bool same, bool first_chunk) {
        const int s = job % CP_NSTAGE;
        const long long tw = prm.clk ? clock64() : 0;
        ok = ok && mbar_wait(&full[s], (job / CP_NSTAGE) & 1);
        if (prm.clk) w_full += clock64() - tw;
        tc_fence_after_sync();
        if (!ok) return;
        const uint32_t a0 = smem_u32(ring + s * CP_STAGE) >> 4;
        const uint64_t ah = dk128 + a0, al = ah + (64 >> 4);        // a 128-byte row = [32 hi | 32 lo] fp16
        const uint64_t bh = same ? ah : ah + (16384 >> 4), bl = bh + (64 >> 4);
        const long long ti0 = prm.clk ? clock64() : 0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // each half: 2 x 32 B, K = 16 fp16 per instruction
          if (prm.dbg_mma == 2) continue;
          mma_f16(acc, ah + 2 * ks, bh + 2 * ks, id_c, !(first_chunk && ks == 0));
          if (prm.dbg_mma == 1) continue;
          mma_f16(acc, ah + 2 * ks, bl + 2 * ks, id_c, 1);
          mma_f16(acc, al + 2 * ks, bh + 2 * ks, id_c, 1);
        }
        const long long ti1 = prm.clk ? clock64() : 0;
        mma_commit(&empty[s]);
        if (prm.clk) { w_issue += ti1 - ti0; w_commit += clock64() - ti1; }
        ++job;
      };
      // One gradient-GEMM group of the item that owns TMEM set `set` against the half of the code rows now in codebuf:
      //   which = 1: dC1 += U . Cn   (A = U K-major, accumulator in the fd columns)
      //   which = 2: dC2 += U^T . Cn (A = the same U tile MN-major, accumulator in the cd columns)
      //   both_u: A runs over U hi and U lo (the code half is `hi`), else over U hi only (the code half is `lo`);
      //   fresh: the first MMA overwrites the accumulator.  K = 128 U columns / rows = 8 steps of 16.
      auto grad_mma = [&](int set, int which, bool both_u, bool fresh, int cb) {
        const uint32_t d = tmem + 256u * set + (which == 1 ? 0u : 128u);
        const uint32_t idd = which == 1 ? id_g1 : id_g2;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint64_t a_h, a_l;
          if (which == 1) {
            const uint32_t aoff = ((ks >> 2) * 16384 + (ks & 3) * 32) >> 4;
            a_h = dk128 + uh + aoff; a_l = dk128 + ul + aoff;
          } else {
            a_h = dmn128 + uh + ks * 128; a_l = dmn128 + ul + ks * 128;
          }
          const uint64_t bq = dmn64 + gb + (uint32_t)cb * (CP_CODE >> 4) + ks * 64;
          mma_f16(d, a_h, bq, idd, !(fresh && ks == 0));
          if (both_u) mma_f16(d, a_l, bq, idd, 1);
        }
      };
      // The previous item's pending gradient work: per round (1, or 2 with the depth term) a sequence of steps, each
      // consuming one codebuf fill: self pair on the diagonal [hi: dC1, dC2][lo: dC1, dC2]; otherwise
      // [C2 hi: dC1][C2 lo: dC1][C1 hi: dC2][C1 lo: dC2].  The last step of a round publishes its accumulators.
      int p_set = 0, p_rounds = 0, p_round = 0, p_step = 0;
      bool p_csame = false, p_live = false;
      // are the inputs of the pending item's next step there?  (non-blocking: polled between operand chunks, so the
      // gradient GEMMs slot into the chunk stream whenever the epilogue and the fill producer are ready for them)
      auto step_ready = [&]() {
        return (p_step != 0 || mbar_test(u_ready, nu & 1)) && mbar_test(&g_full[fills % CP_NCODE], (fills / CP_NCODE) & 1);
      };
      auto step = [&]() {   // issue the next step of the pending item (blocking on its inputs)
        if (p_step == 0) { ok = ok && mbar_wait(u_ready, nu & 1); ++nu; }
        const int cb = fills % CP_NCODE;
        ok = ok && mbar_wait(&g_full[cb], (fills / CP_NCODE) & 1);
        ++fills;
        tc_fence_after_sync();
        if (!ok) return;
        const int nsteps = p_csame ? 2 : 4;
        if (p_csame) {
          grad_mma(p_set, 1, p_step == 0, p_step == 0, cb);
          grad_mma(p_set, 2, p_step == 0, p_step == 0, cb);
        } else {
          grad_mma(p_set, p_step < 2 ? 1 : 2, (p_step & 1) == 0, (p_step & 1) == 0, cb);
        }
        mma_commit(&g_free[cb]);                 // this fill is consumed
        if (++p_step == nsteps) {
          mma_commit(grad_full);
          p_step = 0;
          if (++p_round == p_rounds) p_live = false;
        }
      };
      int n = 0;
      for (int it = first; it < prm.nitems && ok; it += stride, ++n) {
        const Item w = decode_item(prm, it);
        const int set = n & 1;
        {
          const long long tw = prm.clk ? clock64() : 0;
          ok = ok && mbar_wait(&acc_free[set], ((n >> 1) & 1) ^ 1);
          if (prm.clk) w_free += clock64() - tw;
        }
        tc_fence_after_sync();
        const uint32_t acc_fd = tmem + 256u * set, acc_cd = acc_fd + 128u;
        for (int c = 0; c < ncd && ok; ++c) chunk_mma(acc_cd, w.csame, c == 0);
        for (int c = 0; c < nfd && ok; ++c) {
          const long long tp0 = prm.clk ? clock64() : 0;
          const bool go = p_live && step_ready();
          const long long tp1 = prm.clk ? clock64() : 0;
          if (go) step();
          if (prm.clk) { w_poll += tp1 - tp0; w_grad += clock64() - tp1; }
          chunk_mma(acc_fd, w.fsame, c == 0);
        }
        mma_commit(&acc_full[set]);
        {
          const long long tw = prm.clk ? clock64() : 0;
          while (p_live && ok) step();           // whatever the previous item still owes (incl. its depth round)
          if (prm.clk) w_steps += clock64() - tw;
        }
        p_live = true; p_set = set; p_csame = w.csame; p_rounds = w.depth ? 2 : 1; p_round = 0; p_step = 0;
      }
      while (p_live && ok) step();
      if (prm.clk) {
        long long* c = prm.clk + (size_t)blockIdx.x * 16;
        c[10] = w_full; c[11] = w_free; c[12] = w_steps; c[15] = clock64() - t_begin; c[0] = n;
        c[3] = w_issue; c[4] = w_commit; c[5] = w_poll; c[6] = w_grad;
      }
      if (!ok) praise(prm.err, 2);
    }
  } else if (warp == 10) {
    // ================================ TMA producer of the gradient-GEMM operands ================================
    // The gradient GEMMs read their code-row operand from ONE 32 KB buffer that holds a single half (hi or lo) of 128
    // code rows, so an item needs a sequence of fills, each consumed by one MMA step (see the MMA warp):
    //   self pair on the diagonal (same rows for both GEMMs), per round: [rows hi][rows lo]
    //   otherwise, per round: [C2 hi][C2 lo] (dC1 = U . C2n)  then  [C1 hi][C1 lo] (dC2 = U^T . C1n)
    // A warp of its own: its waits (codebuf free again) depend on the epilogue's pace and must not hold up the
    // operand stream of warp 0.
    if (lane == 0) {
      prefetch_tmap(&prm.tm_cg);
      bool ok = true;
      int fills = 0;
      for (int it = first; it < prm.nitems && ok; it += stride) {
        const Item w = decode_item(prm, it);
        const int row1 = w.b * Prows + 128 * w.ti, row2 = (w.k * prm.B + w.b) * Prows + 128 * w.tj;
        const int per_round = w.csame ? 2 : 4, total = per_round * (w.depth ? 2 : 1);
        for (int f = 0; f < total && ok; ++f, ++fills) {
          const int j = f % per_round;
          const int r = j < 2 ? row2 : row1, half2 = j & 1;
          const int cb = fills % CP_NCODE;
          ok = ok && mbar_wait(&g_free[cb], ((fills / CP_NCODE) & 1) ^ 1);
          if (!ok) break;
          mbar_arrive_expect_tx(&g_full[cb], (uint32_t)(nb * 8192));
          for (int a = 0; a < nb; ++a)
            tma_load_2d(codebuf + cb * CP_CODE + a * 8192, &prm.tm_cg, &g_full[cb], 64 * a + 32 * half2, r);
        }
      }
      if (!ok) praise(prm.err, 4);
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // All hot loops walk 8 TMEM columns at a time with `#pragma unroll 1`: the round-1 epilogue was ~250 KB of fully
    // unrolled code whose instruction fetches had to compete with the operand stream for L2 (ncu: "no instruction"
    // was its second largest stall); these bodies are ~2 KB and run out of the instruction cache.
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;        // 0: columns 0..63, 1: columns 64..127 of the tile
    const int ew = warp - 2;                 // 0..7
    const int row = 32 * lg + lane;          // row within the tile
    const int P = prm.P;
    const bool pointwise = prm.flags & DG_FLAG_POINTWISE;
    const float lo = (prm.flags & DG_FLAG_ZERO_CLAMP) ? 0.f : -9999.f;
    const float hi = (prm.flags & DG_FLAG_STABALIZE) ? 0.8f : __int_as_float(0x7f800000);
    constexpr float FS = 1.f / (F16_FEAT_SCALE * F16_FEAT_SCALE);
    constexpr float CS = 1.f / (F16_CODE_SCALE * F16_CODE_SCALE);
    const float gscale = prm.inv_cnt / (prm.uscale * F16_CODE_SCALE);
    const float dsh = prm.depth_shift;
    const bool dense_out = prm.cd_out || prm.loss_out || prm.dd_out || prm.fd_dbg;
    float* scratch = reinterpret_cast<float*>(u_hi) + ew * (32 * 33);  // per-warp transpose buffer, aliases U while it is dead
    const size_t slab = (size_t)prm.B * Prows * prm.ldc;
    uint8_t* urow_hi = u_hi + half * 16384 + row * 128;   // this thread's row of its 64-column U atom
    uint8_t* urow_lo = u_lo + half * 16384 + row * 128;
    bool ok = true;
    long long e_acc = 0, e_grad = 0;
    int n = 0, ng = 0;   // items / gradient rounds seen so far (barrier phases)
    int sign_key = -1;
    for (int it = first; it < prm.nitems; it += stride, ++n) {
      const Item w = decode_item(prm, it);
      const int set = n & 1;
      const int p = 128 * w.ti + row;          // sample point (row of fd / cd / dC1)
      const uint32_t tlane = tmem + ((uint32_t)(32 * lg) << 16) + 256u * set;
      const uint32_t t_fd = tlane + 64u * half, t_cd = t_fd + 128u;   // this thread's 64 columns
      const int qbase = 128 * w.tj + 64 * half;
      if (w.depth && sign_key != w.b * 8 + w.tj) {   // depth signs of the column tile (only the intra pair's items need them)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x - 64 < 128) s_sign[threadIdx.x - 64] = __ldg(prm.dsign + (size_t)w.b * Prows + 128 * w.tj + threadIdx.x - 64);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        sign_key = w.b * 8 + w.tj;
      }
      float old_mean = 0.f;  // mean of fd over (b,p,q) of this pair = mean_b <mean row F1[b], mean row F2[k,b]>
      if (pointwise && prm.dots) {
        float part = 0.f;
        for (int bb = lane; bb < prm.B; bb += 32) part += __ldg(prm.dots + (size_t)w.k * prm.B + bb);
        old_mean = warp_sum(part) / (float)prm.B;
      }
      const float sp = w.depth ? __ldg(prm.dsign + (size_t)w.b * Prows + p) : 0.f;
      const float um = (p < P) ? prm.uscale : 0.f;   // U row factor: padded rows are masked by multiplying with 0
      const float* ssign = s_sign + 64 * half;
      float sum_loss = 0.f, sum_cd = 0.f, sum_dloss = 0.f, sum_dd = 0.f;

      if (threadIdx.x == 64 && n == 0) pstamp(prm, 1);
      {
        const long long tw = prm.clk ? clock64() : 0;
        ok = ok && mbar_wait(&acc_full[set], (n >> 1) & 1);
        if (prm.clk) e_acc += clock64() - tw;
      }
      tc_fence_after_sync();
      if (threadIdx.x == 64 && n == 0) pstamp(prm, 2);
      if (threadIdx.x == 64 && n == 1) pstamp(prm, 7);
      float c0 = prm.shift[w.k] - old_mean;      // fd' - shift = fd - rowmean + old_mean - shift = fd - c0
      if (pointwise && prm.rowmean) {            // several column tiles: the row means were precomputed
        c0 += __ldg(prm.rowmean + (size_t)w.kb * Prows + p);
      } else if (pointwise) {                    // the item sees every column: padded columns are exactly zero
        float s = 0.f;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          float v[8];
          tmem_ld_32x8(t_fd + 8 * j, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) s += v[i];
        }
        s_rowsum[half][row] = s * FS;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        c0 += (s_rowsum[0][row] + s_rowsum[1][row]) / (float)P;
      }
      if (threadIdx.x == 64 && n == 0) pstamp(prm, 3);
      if (dense_out) {   // optional dense outputs (materialize_cd / tests): a separate cold pass
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          float v[8], c[8];
          tmem_ld_32x8(t_fd + 8 * j, v);
          tmem_ld_32x8(t_cd + 8 * j, c);
          tmem_ld_wait();
          const size_t obase = (((size_t)w.k * prm.B + w.b) * P + p) * P;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int q = qbase + 8 * j + i;
            const float fdv = v[i] * FS, cdv = c[i] * CS;
            if (prm.fd_dbg) prm.fd_dbg[((size_t)w.kb * Prows + p) * Prows + q] = fdv;
            if (p < P && q < P) {
              const float cl = fminf(fmaxf(cdv, lo), hi);
              if (prm.cd_out) prm.cd_out[obase + q] = cdv;
              if (prm.loss_out) prm.loss_out[obase + q] = -cl * (fdv - c0);
              if (prm.dd_out && w.depth) prm.dd_out[((size_t)w.b * P + p) * P + q] = sp * ssign[8 * j + i];
            }
          }
        }
      }
      uint32_t passmask[2] = {0u, 0u};
      const int rounds = w.depth ? 2 : 1;
      for (int rd = 0; rd < rounds; ++rd, ++ng) {
        if (rd > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // every warp is done with its drain scratch before U is rewritten
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {      // 8 columns per step: q = qbase + 8 j .. + 7
          float u[8];
          if (rd == 0) {
            // main pass.  Padded rows/columns have fd = cd = 0 and depth sign 0, so they add nothing to the sums; only
            // U needs the explicit mask (um = 0 for padded rows, column test for the tile that crosses P).
            float v[8], c[8];
            tmem_ld_32x8(t_fd + 8 * j, v);
            tmem_ld_32x8(t_cd + 8 * j, c);
            tmem_ld_wait();
            const int qlim = P - (qbase + 8 * j);   // columns i >= qlim are padding
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float cdv = c[i] * CS;
              const float cl = fminf(fmaxf(cdv, lo), hi);
              const float f = fmaf(v[i], FS, -c0);
              sum_loss = fmaf(-cl, f, sum_loss);
              sum_cd += cdv;
              const bool pass = (cdv >= lo) && (cdv <= hi) && (i < qlim);
              if (w.depth) {
                const float dd = sp * ssign[8 * j + i];
                sum_dloss = fmaf(-cl, dd - dsh, sum_dloss);
                sum_dd += dd;
                bits |= pass ? (1u << i) : 0u;
              }
              u[i] = pass ? -f * um : 0.f;
            }
            if (w.depth) passmask[j >> 2] |= bits << (8 * (j & 3));
          } else {
            // depth term: U_d = -(s_p s_q - depth_shift) 1[clamp passes], indicator from the main pass
            const uint32_t bits = passmask[j >> 2] >> (8 * (j & 3));
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = ((bits >> i) & 1u) ? -(sp * ssign[8 * j + i] - dsh) * um : 0.f;
          }
          // fp16 hi / lo of the 8 values: one 16-byte store each into the 128B-swizzled U row (chunk j ^ row % 8)
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 hh = __floats2half2_rn(u[2 * e], u[2 * e + 1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(u[2 * e] - hf.x, u[2 * e + 1] - hf.y);
            h[e] = *reinterpret_cast<const uint32_t*>(&hh);
            l[e] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t chunk = (uint32_t)(j ^ (row & 7));
          *reinterpret_cast<uint4*>(urow_hi + chunk * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(urow_lo + chunk * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        if (threadIdx.x == 64 && n == 0 && rd == 0) pstamp(prm, 4);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(u_ready);
        {
          const long long tw = prm.clk ? clock64() : 0;
          ok = ok && mbar_wait(grad_full, ng & 1);
          if (prm.clk) e_grad += clock64() - tw;
        }
        tc_fence_after_sync();
        if (threadIdx.x == 64 && n == 0 && rd == 0) pstamp(prm, 5);
        // U is dead until it is rewritten: drain through it.  The two warps of a lane group split the column chunks:
        // 32-column blocks 0, 2 (dC2) / 1, 3 (dC1) to half 0 and the others to half 1.
        const size_t which = (rd == 0) ? (size_t)w.k : (size_t)prm.npairs;
        float* d2 = prm.dC2 + (which * prm.ntile + w.ti) * slab + ((size_t)w.b * Prows + 128 * w.tj + 32 * lg) * prm.ldc;
        float* d1 = prm.dC1 + (which * prm.ntile + w.tj) * slab + ((size_t)w.b * Prows + 128 * w.ti + 32 * lg) * prm.ldc;
#pragma unroll 1
        for (int blk = 0; blk < 2 * ncd; ++blk) {       // blocks 0..ncd-1: dC2 (cd columns), ncd..2ncd-1: dC1 (fd columns)
          if ((blk & 1) != half) continue;
          const bool is2 = blk < ncd;
          const int cc = is2 ? blk : blk - ncd;
          const uint32_t src = tlane + (is2 ? 128u : 0u) + 32u * cc;
          float* dst = (is2 ? d2 : d1) + 32 * cc;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            float v[8];
            tmem_ld_32x8(src + 8 * j, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) scratch[lane * 33 + 8 * j + i] = v[i] * gscale;
          }
          __syncwarp();
#pragma unroll 4
          for (int r = 0; r < 32; ++r) dst[(size_t)r * prm.ldc + lane] = scratch[r * 33 + lane];
          __syncwarp();
        }
        if (threadIdx.x == 64 && n == 0 && rd == 0) pstamp(prm, 6);
        if (threadIdx.x == 64 && n == 1 && rd == 0) pstamp(prm, 8);
      }
      // this TMEM set and the U tile are free again
      tc_fence_before_sync();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) mbar_arrive(&acc_free[set]);

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
          prm.partials[(size_t)w.pidx * 4 + lane] = t;
        }
        // fused finalize: the CTA that completes the last item folds all partial sums into the 8 output scalars.
        // __syncwarp orders the four partial stores before lane 0's acq_rel counter increment, which publishes them
        // at gpu scope without a full __threadfence
        __syncwarp();
        int ticket = 0;
        if (lane == 0)
          asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(prm.done) : "memory");
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == prm.nitems - 1) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          const int per_pair = prm.B * prm.nt * prm.nt;
          for (int e = lane; e < prm.nitems; e += 32) {   // fixed order -> deterministic
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
            const float nn = (g < 3) ? (float)cnt[g] : (prm.has_depth ? 1.f : 0.f);
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i == lane) t = acc[i];
            float r = nn > 0.f ? t / (nn * elems) : 0.f;
            if (*reinterpret_cast<volatile int*>(prm.err) != 0) r = __int_as_float(0x7fc00000);  // pipeline timeout
            prm.out8[lane] = r;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // s_red / s_rowsum reuse by the next item
    }
    if (prm.clk && threadIdx.x == 64) {
      prm.clk[(size_t)blockIdx.x * 16 + 13] = e_acc;
      prm.clk[(size_t)blockIdx.x * 16 + 14] = e_grad;
    }
    if (!ok) praise(prm.err, 3);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 64) pstamp(prm, 9);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------ host side
int make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t cols, uint64_t rows,
                uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw);                    // corr_umma.cu
int launch_pair_dots(const DotsJob& job, cudaStream_t st);
int launch_row_means(const void* f_hi, const void* f_lo, const float* fmean, int nsplit, int npairs, int B, int P, int Prows,
                     int ldf, float* rowmean, const int32_t* fs1, const int32_t* fs2, cudaStream_t st, int interleave);
long long* get_clock_buffer();

// workspace: [err int, done int (256 B)][dots npairs*B floats, padded][partials nitems*4 floats, padded][row means]
size_t corr_pipe_workspace_floats(int npairs, int B, int P) {
  const size_t nt = (size_t)ceil_div(P, 128);
  return (size_t)npairs * B * (1 + nt * nt * 4 + (size_t)round_up(P, 128)) + 512;
}

// Can this FPS launch run as ride-along CTAs of the correlation kernel?  (the 28 x 28 grid class of fps_kernel<4, 256, 4>,
// shared memory within this kernel's allocation, permutation draw within one CTA)
bool corr_pipe_can_ride(const FpsArgs& a, size_t smem) {
  return a.H * a.W <= 896 && smem <= (size_t)CP_SMEM && a.pj.n <= CP_THREADS;
}

int corr_loss_pipe(const dg_panels_t* pan, const float* fmean, int nsplit, const float* dsign, int npairs, int B, int P,
                   int Prows, int ldf, int ldc, const float* pair_shift, const int32_t* pair_group, float depth_shift,
                   int flags, float* out8, float* dC1, float* dC2, float* cd_out, float* loss_out, float* dd_out,
                   float* fd_dbg, void* ws, cudaStream_t st, const int32_t* fslot1, const int32_t* fslot2, int nfslots,
                   bool dots_done, const FpsArgs* ride, size_t ride_smem) {
  DG_REQUIRE(Prows % 128 == 0 && Prows >= P && P <= 1024, DG_ERR_INVALID, "corr_loss_pipe: bad panel rows %d for %d points", Prows, P);
  DG_REQUIRE(!ride || corr_pipe_can_ride(*ride, ride_smem), DG_ERR_INVALID, "corr_loss_pipe: this FPS job cannot ride along");
  DG_REQUIRE(ldf % 32 == 0 && ldc % 32 == 0 && ldc <= 128, DG_ERR_UNSUPPORTED, "corr_loss_pipe: bad panel pitch");
  PipeParams prm;
  const int nt = ceil_div(P, 128);
  const uint64_t rows = (uint64_t)npairs * B * Prows;
  const uint64_t frows = (uint64_t)(nfslots > 0 ? nfslots : npairs) * B * Prows;
  int rc;
  if ((rc = make_map_2d(&prm.tm_f, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->f_hi, 2 * (uint64_t)ldf, frows, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->cb_hi, 2 * (uint64_t)ldc, rows, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&prm.tm_cg, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, pan->cb_hi, 2 * (uint64_t)ldc, rows, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  int* err;
  float* dots;
  umma_ws_layout(ws, &err, &dots);
  const size_t dots_bytes = ((size_t)npairs * B * sizeof(float) + 255) / 256 * 256;
  float* partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256 + dots_bytes);
  const size_t part_bytes = ((size_t)npairs * B * nt * nt * 4 * sizeof(float) + 255) / 256 * 256;
  float* rowmean = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256 + dots_bytes + part_bytes);
  const bool pointwise = flags & DG_FLAG_POINTWISE;
  prm.dsign = dsign;
  prm.dots = pointwise ? dots : nullptr;
  prm.rowmean = (nt > 1 && pointwise) ? rowmean : nullptr;
  prm.npairs = npairs; prm.B = B; prm.P = P; prm.prows = Prows; prm.ldf = ldf; prm.ldc = ldc; prm.flags = flags;
  prm.has_depth = dsign != nullptr;
  prm.nt = nt; prm.ntile = Prows / 128; prm.nitems = npairs * B * nt * nt;
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
  prm.clk = get_clock_buffer();
  prm.l2_hints = getenv("DEPTHG_B200_L2HINTS") && getenv("DEPTHG_B200_L2HINTS")[0] == '1';
  prm.dbg_mma = getenv("DEPTHG_B200_PIPE_MMA") ? atoi(getenv("DEPTHG_B200_PIPE_MMA")) : 0;
  prm.dbg_bulk = getenv("DEPTHG_B200_PIPE_BULK") ? static_cast<const uint8_t*>(pan->f_hi) : nullptr;
  if (!dots_done) {
    DotsJob job;
    job.fmean = pointwise ? fmean : nullptr;
    job.dots = dots; job.err = err; job.nsplit = nsplit; job.npairs = npairs; job.B = B; job.ldf = ldf;
    for (int k = 0; k < DG_MAX_PAIRS; ++k) { job.fs1[k] = k < npairs ? prm.fs1[k] : 0; job.fs2[k] = k < npairs ? prm.fs2[k] : 0; }
    if ((rc = launch_pair_dots(job, st))) return rc;
  }
  if (prm.rowmean) {
    if ((rc = launch_row_means(pan->f_hi, nullptr, fmean, nsplit, npairs, B, P, Prows, ldf, rowmean, prm.fs1, prm.fs2, st, 1))) return rc;
  }
  static PerDevice attr_pd = {}, sm_pd = {};
  size_t& attr_set = per_device(attr_pd);
  size_t& sms = per_device(sm_pd);
  if (!attr_set) {
    DG_CUDA_OK(cudaFuncSetAttribute(corr_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM));
    int dev = 0, n = 0;
    DG_CUDA_OK(cudaGetDevice(&dev));
    DG_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms = (size_t)(n > 0 ? n : 148);
    attr_set = 1;
  }
  const int grid = prm.nitems < (int)sms ? prm.nitems : (int)sms;   // persistent: one CTA per SM walks the item list
  prm.nmain = grid;
  prm.nride = 0;
  memset(&prm.ride, 0, sizeof prm.ride);
  if (ride) {
    prm.ride = *ride;
    prm.ride.clk = prm.clk;
    prm.nride = ride->nimg + (ride->pj.n > 0 ? 1 : 0);
  }
  DG_PRE(st);
  launch_pdl(corr_pipe_kernel, dim3(grid + prm.nride), dim3(CP_THREADS), (size_t)CP_SMEM, st, prm);
  DG_LAUNCH_OK("corr_pipe_kernel");
  return DG_OK;  // out8 is written by the CTA that completes the last item
}

}  // namespace dg
