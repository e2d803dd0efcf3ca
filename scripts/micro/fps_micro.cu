// Latency microbenchmarks behind the FPS round design (one CTA, dependent chains, cycles per op).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fps_micro fps_micro.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 2000

__global__ void k_redux(int* out, long long* cyc) {
  int v = threadIdx.x * 7 + 3;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
    int m = __reduce_max_sync(0xffffffffu, v);
    v = (v ^ m) + threadIdx.x;  // depends on the result
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_redux_pair(int* out, long long* cyc) {   // max then min-of-index among maxima (the FPS warp argmax)
  int v = threadIdx.x * 7 + 3, idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
    int m = __reduce_max_sync(0xffffffffu, v);
    int w = __reduce_min_sync(0xffffffffu, v == m ? idx : 0x7fffffff);
    v = (v ^ w) + threadIdx.x + i;
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ballot_pair(int* out, long long* cyc) {  // max + ballot/ffs + shfl
  int v = threadIdx.x * 7 + 3, idx = threadIdx.x * 3;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
    int m = __reduce_max_sync(0xffffffffu, v);
    unsigned b = __ballot_sync(0xffffffffu, v == m);
    int w = __shfl_sync(0xffffffffu, idx, __ffs(b) - 1);
    v = (v ^ w) + threadIdx.x + i;
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(int* out, long long* cyc) {
  int v = threadIdx.x * 7 + 3;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + i;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bfly64(int* out, long long* cyc) {  // 5-level butterfly max of a 64-bit composite
  long long c = ((long long)(threadIdx.x * 7 + 3) << 32) | threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long x = __shfl_xor_sync(0xffffffffu, c, o);
      c = x > c ? x : c;
    }
    c = (c ^ (long long)i) + threadIdx.x;
  }
  long long t1 = clock64();
  out[threadIdx.x] = (int)c;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(int* out, long long* cyc) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) v = s[v];
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NT>
__global__ void k_bar_xchg(int* out, long long* cyc) {  // STS -> named barrier -> LDS round trip, NT threads
  __shared__ int s[2][8];
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
    if ((threadIdx.x & 31) == 0) s[i & 1][threadIdx.x >> 5] = v;
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    v = s[i & 1][(threadIdx.x + i) & (NT / 32 - 1)] + 1;
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// mbarrier-free flag exchange between 2..8 warps: volatile smem flag spin (no bar.sync)
template <int NT>
__global__ void k_flag_xchg(int* out, long long* cyc) {
  __shared__ volatile int s[8][2];
  if (threadIdx.x < 16) ((volatile int*)s)[threadIdx.x] = -1;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = NT / 32;
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; ++i) {
    if (lane == 0) { s[w][0] = v; __threadfence_block(); s[w][1] = i; }
    int acc = 0;
    if (lane < NW) { while (s[lane][1] < i) {} acc = s[lane][0]; }
    acc = __reduce_max_sync(0xffffffffu, acc);
    v = acc + 1;
    // a second phase so no warp overwrites a slot before everybody has read it (double use of parity would do too)
    asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// Full FPS round, single warp, PR points per lane (blocked), min trick, tree max, index via ballot.
template <int PR>
__global__ void k_round1w(const float* pts, int npts, int nsel, int* out, long long* cyc) {
  extern __shared__ float smem_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sX = smem_dyn + warp * 3 * 800; float* sY = sX + 800; float* sZ = sY + 800;
  for (int i = lane; i < 800; i += 32) {
    sX[i] = i < npts ? pts[3 * i] : 0.f; sY[i] = i < npts ? pts[3 * i + 1] : 0.f; sZ[i] = i < npts ? pts[3 * i + 2] : 0.f;
  }
  __syncwarp();
  float X[PR], Y[PR], Z[PR]; int key[PR];
#pragma unroll
  for (int j = 0; j < PR; ++j) {
    const int i = lane * PR + j;
    X[j] = sX[i]; Y[j] = sY[i]; Z[j] = sZ[i];
    key[j] = (i < npts && i != 0) ? 0x7f800000 : -1;
  }
  int last = 0;
  long long t0 = clock64();
  for (int r = 1; r < nsel; ++r) {
    const float lx = sX[last], ly = sY[last], lz = sZ[last];
    int bk = -1;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const float dx = __fsub_rn(lx, X[j]), dy = __fsub_rn(ly, Y[j]), dz = __fsub_rn(lz, Z[j]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      key[j] = min(key[j], __float_as_int(d));
      bk = max(bk, key[j]);
    }
    const int wk = __reduce_max_sync(0xffffffffu, bk);
    const unsigned b = __ballot_sync(0xffffffffu, bk == wk);
    const int wl = __ffs(b) - 1;
    int bj = PR;
#pragma unroll
    for (int j = PR - 1; j >= 0; --j) bj = key[j] == wk ? j : bj;
    last = __shfl_sync(0xffffffffu, lane * PR + bj, wl);
    if (lane == wl) {
#pragma unroll
      for (int j = 0; j < PR; ++j) if (j == bj) key[j] = -1;
    }
    if (lane == 0) out[warp * 128 + r] = last;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}


// ---- V0: the shipped round loop (fps.cu fps_rounds<4,256>) with optional phase stamps -------------------------
constexpr int FPS_WARPS = 8;
template <int PR, int RT, bool STAMP>
__device__ __forceinline__ void fps_rounds_v0(const float* sX, const float* sY, const float* sZ, unsigned char* sTaken,
                                           int npts, int nsel, int (*s_key)[FPS_WARPS], int (*s_idx)[FPS_WARPS], int* order, long long* ph) {
  constexpr int RW = RT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float X[PR], Y[PR], Z[PR];
  int key[PR];
#pragma unroll
  for (int j = 0; j < PR; ++j) {
    const int i = j * RT + tid;
    const bool ok = i < npts;
    X[j] = ok ? sX[i] : 0.f; Y[j] = ok ? sY[i] : 0.f; Z[j] = ok ? sZ[i] : 0.f;
    key[j] = ok && i != 0 ? 0x7f800000 : -1;
  }
  int last = 0;
  long long a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (int r = 1; r < nsel; ++r) {
    long long t0, t1, t2, t3, t4, t5;
    if (STAMP) t0 = clock64();
    const float lx = sX[last], ly = sY[last], lz = sZ[last];
    int bk = -1, bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const float dx = __fsub_rn(lx, X[j]), dy = __fsub_rn(ly, Y[j]), dz = __fsub_rn(lz, Z[j]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const int kj = key[j];
      const int k = kj < 0 ? -1 : min(kj, __float_as_int(d));
      key[j] = k;
      const bool better = k > bk;
      bk = better ? k : bk;
      bi = better ? j * RT + tid : bi;
    }
    if (STAMP) t1 = clock64();
    const int wk = __reduce_max_sync(0xffffffffu, bk);
    const int wi = __reduce_min_sync(0xffffffffu, bk == wk ? bi : 0x7fffffff);
    if (STAMP) t2 = clock64();
    if (RW == 1) {
      last = wi;
    } else {
      const int buf = r & 1;
      if (lane == 0) { s_key[buf][warp] = wk; s_idx[buf][warp] = wi; }
      asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
      const int ck = lane < RW ? s_key[buf][lane] : -1;
      const int ci = lane < RW ? s_idx[buf][lane] : 0x7fffffff;
      if (STAMP) t3 = clock64() + (ck & 0);
      const int gk = __reduce_max_sync(0xffffffffu, ck);
      last = __reduce_min_sync(0xffffffffu, ck == gk ? ci : 0x7fffffff);
      if (STAMP) t4 = clock64() + (last & 0);
    }
    if (((last % RT) >> 5) == warp) {
      if ((last % RT) == tid) {
        const int j = last / RT;
#pragma unroll
        for (int jj = 0; jj < PR; ++jj)
          if (jj == j) key[jj] = -1;
        sTaken[last] = 1;
      }
    }
    if (STAMP) { t5 = clock64(); a0 += t1 - t0; a1 += t2 - t1; a2 += t3 - t2; a3 += t4 - t3; a4 += t5 - t4; }
    if (tid == 0) order[r] = last;
  }
  if (STAMP && tid == 0) { ph[0] = a0; ph[1] = a1; ph[2] = a2; ph[3] = a3; ph[4] = a4; }
}
template <int PR, int RT, bool STAMP>
__global__ void __launch_bounds__(256) k_v0(const float* pts, int npts, int nsel, int* out, long long* cyc) {
  __shared__ float sX[896], sY[896], sZ[896];
  __shared__ unsigned char sTaken[896];
  __shared__ int s_key[2][FPS_WARPS], s_idx[2][FPS_WARPS];
  for (int i = threadIdx.x; i < 896; i += blockDim.x) {
    sX[i] = i < npts ? pts[3 * i] : 0.f; sY[i] = i < npts ? pts[3 * i + 1] : 0.f; sZ[i] = i < npts ? pts[3 * i + 2] : 0.f;
    sTaken[i] = 0;
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x < RT) fps_rounds_v0<PR, RT, STAMP>(sX, sY, sZ, sTaken, npts, nsel, s_key, s_idx, out, cyc + 1);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---- V1: min trick (no taken test), no marking (a taken point's key becomes 0 by itself), 64-bit exchange ----------
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
template <int PR, int RT, bool PACK>
__device__ __forceinline__ void fps_rounds_v1(const float* sX, const float* sY, const float* sZ, unsigned char* sTaken,
                                              int npts, int nsel, int2 (*s_kv)[16], int* order) {
  constexpr int RW = RT / 32;
  static_assert(!PACK || PR % 2 == 0, "packed arithmetic works on point pairs");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float X[PR], Y[PR], Z[PR];
  unsigned long long X2[PR / 2 + 1], Y2[PR / 2 + 1], Z2[PR / 2 + 1];
  int key[PR];
#pragma unroll
  for (int j = 0; j < PR; ++j) {
    const int i = j * RT + tid;
    const bool ok = i < npts;
    X[j] = ok ? sX[i] : 0.f; Y[j] = ok ? sY[i] : 0.f; Z[j] = ok ? sZ[i] : 0.f;
    key[j] = ok ? 0x7f800000 : -1;   // point 0 drops to key 0 in round 1 (its distance to itself)
  }
  if (PACK) {
#pragma unroll
    for (int j = 0; j < PR / 2; ++j) {
      X2[j] = pack2(X[2 * j], X[2 * j + 1]); Y2[j] = pack2(Y[2 * j], Y[2 * j + 1]); Z2[j] = pack2(Z[2 * j], Z[2 * j + 1]);
    }
  }
  int last = 0;
  for (int r = 1; r < nsel; ++r) {
    const float lx = sX[last], ly = sY[last], lz = sZ[last];
    int bk = -1;
    if (PACK) {
      const unsigned long long lx2 = pack2(lx, lx), ly2 = pack2(ly, ly), lz2 = pack2(lz, lz);
#pragma unroll
      for (int j = 0; j < PR / 2; ++j) {
        const unsigned long long dx = sub2(lx2, X2[j]), dy = sub2(ly2, Y2[j]), dz = sub2(lz2, Z2[j]);
        // the two additions stay scalar: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding)
        float xx0, xx1, yy0, yy1, zz0, zz1;
        unpack2(mul2(dx, dx), xx0, xx1); unpack2(mul2(dy, dy), yy0, yy1); unpack2(mul2(dz, dz), zz0, zz1);
        const float d0 = __fadd_rn(__fadd_rn(xx0, yy0), zz0), d1 = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
        key[2 * j] = min(key[2 * j], __float_as_int(d0));
        key[2 * j + 1] = min(key[2 * j + 1], __float_as_int(d1));
        bk = max(bk, max(key[2 * j], key[2 * j + 1]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < PR; ++j) {
        const float dx = __fsub_rn(lx, X[j]), dy = __fsub_rn(ly, Y[j]), dz = __fsub_rn(lz, Z[j]);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        key[j] = min(key[j], __float_as_int(d));
        bk = max(bk, key[j]);
      }
    }
    const int wk = __reduce_max_sync(0xffffffffu, bk);
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = PR - 1; j >= 0; --j) bi = key[j] == wk ? j * RT + tid : bi;
    const int wi = __reduce_min_sync(0xffffffffu, bi);
    const int buf = r & 1;
    if (lane == 0) s_kv[buf][warp] = make_int2(wk, wi);
    asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
    const int2 c = s_kv[buf][lane & (RW - 1)];
    const int gk = __reduce_max_sync(0xffffffffu, c.x);
    last = __reduce_min_sync(0xffffffffu, c.x == gk ? c.y : 0x7fffffff);
    if (gk == 0) {   // every remaining distance is zero (coincident points): first index not taken yet wins
      int fi = 0x7fffffff;
#pragma unroll
      for (int j = PR - 1; j >= 0; --j) {
        const int i = j * RT + tid;
        if (key[j] == 0 && !sTaken[i]) fi = i;
      }
      fi = __reduce_min_sync(0xffffffffu, fi);
      asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");   // everybody has read slot `buf`
      if (lane == 0) s_kv[buf][warp].y = fi;
      asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
      last = __reduce_min_sync(0xffffffffu, s_kv[buf][lane & (RW - 1)].y);
    }
    if (tid == 0) { sTaken[last] = 1; order[r] = last; }
  }
}
template <int PR, int RT, bool PACK>
__global__ void __launch_bounds__(512) k_v1(const float* pts, int npts, int nsel, int* out, long long* cyc) {
  __shared__ float sX[1024], sY[1024], sZ[1024];
  __shared__ unsigned char sTaken[1024];
  __shared__ int2 s_kv[2][16];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    sX[i] = i < npts ? pts[3 * i] : 0.f; sY[i] = i < npts ? pts[3 * i + 1] : 0.f; sZ[i] = i < npts ? pts[3 * i + 2] : 0.f;
    sTaken[i] = i == 0;
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x < RT) fps_rounds_v1<PR, RT, PACK>(sX, sY, sZ, sTaken, npts, nsel, s_kv, out);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  int* out; long long* cyc; float* pts;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 128); cudaMalloc(&pts, 3 * 1024 * 4);
  float h[3 * 1024];
  unsigned s = 12345;
  for (int i = 0; i < 3 * 1024; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.0f / 16777216.0f); }
  cudaMemcpy(pts, h, sizeof h, cudaMemcpyHostToDevice);
  long long c;
#define RUN(name, launch, div)                                                       \
  do {                                                                               \
    launch; launch;                                                                  \
    cudaDeviceSynchronize();                                                         \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                  \
    printf("%-28s %8.1f cycles/iter  (%s)\n", name, (double)c / (div), cudaGetErrorString(cudaGetLastError())); \
  } while (0)
  RUN("redux.max chain", (k_redux<<<1, 32>>>(out, cyc)), N_IT);
  RUN("redux max+min pair", (k_redux_pair<<<1, 32>>>(out, cyc)), N_IT);
  RUN("redux max+ballot+shfl", (k_ballot_pair<<<1, 32>>>(out, cyc)), N_IT);
  RUN("shfl chain", (k_shfl<<<1, 32>>>(out, cyc)), N_IT);
  RUN("butterfly64 (5 lvl)", (k_bfly64<<<1, 32>>>(out, cyc)), N_IT);
  RUN("lds chain", (k_lds<<<1, 32>>>(out, cyc)), N_IT);
  RUN("sts+bar+lds 64 thr", (k_bar_xchg<64><<<1, 64>>>(out, cyc)), N_IT);
  RUN("sts+bar+lds 128 thr", (k_bar_xchg<128><<<1, 128>>>(out, cyc)), N_IT);
  RUN("sts+bar+lds 256 thr", (k_bar_xchg<256><<<1, 256>>>(out, cyc)), N_IT);
  RUN("flag xchg 128 thr", (k_flag_xchg<128><<<1, 128>>>(out, cyc)), N_IT);
  cudaFuncSetAttribute(k_round1w<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9600);
  RUN("round 1 warp/img, 1 warp/CTA", (k_round1w<25><<<1, 32, 9600>>>(pts, 784, 121, out, cyc)), 120);
  RUN("round 1 warp/img, 4 warps/CTA", (k_round1w<25><<<1, 128, 4 * 9600>>>(pts, 784, 121, out, cyc)), 120);
  RUN("round 1 warp/img, 8 warps/CTA", (k_round1w<25><<<1, 256, 8 * 9600>>>(pts, 784, 121, out, cyc)), 120);
  long long ph[6];
  RUN("V0 RT=256 PR=4", (k_v0<4, 256, false><<<1, 256>>>(pts, 784, 121, out, cyc)), 120);
  RUN("V0 RT=256 PR=4 x64 CTAs", (k_v0<4, 256, false><<<64, 256>>>(pts, 784, 121, out, cyc)), 120);
  RUN("V0 RT=128 PR=7", (k_v0<7, 128, false><<<1, 256>>>(pts, 784, 121, out, cyc)), 120);
  RUN("V0 RT=256 PR=4 stamped", (k_v0<4, 256, true><<<1, 256>>>(pts, 784, 121, out, cyc)), 120);
  cudaMemcpy(ph, cyc, 48, cudaMemcpyDeviceToHost);
  printf("   phases/round: coords+dist %.1f | warp pair %.1f | sts+bar+lds %.1f | block pair %.1f | mark %.1f\n",
         ph[1] / 120.0, ph[2] / 120.0, ph[3] / 120.0, ph[4] / 120.0, ph[5] / 120.0);
  static int ref[3][128], got[128];
  float* ptsv[3]; const char* pname[3] = {"random", "ties (flat grid)", "all coincident"};
  for (int v = 0; v < 3; ++v) {
    cudaMalloc(&ptsv[v], 3 * 1024 * 4);
    for (int i = 0; i < 1024; ++i) {
      if (v == 0) { h[3 * i] = h[3 * i]; }
      if (v == 1) { h[3 * i] = (float)(i % 28) * 0.25f; h[3 * i + 1] = (float)(i / 28) * 0.25f; h[3 * i + 2] = -1.5f; }
      if (v == 2) { h[3 * i] = 0.f; h[3 * i + 1] = 0.f; h[3 * i + 2] = 0.f; }
    }
    cudaMemcpy(ptsv[v], h, sizeof h, cudaMemcpyHostToDevice);
    k_v0<4, 256, false><<<1, 256>>>(ptsv[v], 784, 121, out, cyc);
    cudaMemcpy(ref[v], out, 121 * 4, cudaMemcpyDeviceToHost);
  }
#define CHECK(name, launchv)                                                          \
  for (int v = 0; v < 3; ++v) {                                                       \
    const float* P = ptsv[v]; cudaMemset(out, 0, 512); launchv; cudaMemcpy(got, out, 121 * 4, cudaMemcpyDeviceToHost);    \
    int bad = 0; for (int i = 1; i < 121; ++i) bad += got[i] != ref[v][i];            \
    printf("   %-22s vs V0 on %-18s: %d of 120 picks differ (%s)\n", name, pname[v], bad, cudaGetErrorString(cudaGetLastError())); \
  }
  RUN("V1 RT=256 PR=4", (k_v1<4, 256, false><<<1, 256>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 256/4", (k_v1<4, 256, false><<<1, 256>>>(P, 784, 121, out, cyc)));
  RUN("V1 RT=256 PR=4 packed", (k_v1<4, 256, true><<<1, 256>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 256/4 packed", (k_v1<4, 256, true><<<1, 256>>>(P, 784, 121, out, cyc)));
  RUN("V1 RT=128 PR=7", (k_v1<7, 128, false><<<1, 128>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 128/7", (k_v1<7, 128, false><<<1, 128>>>(P, 784, 121, out, cyc)));
  RUN("V1 RT=128 PR=8 packed", (k_v1<8, 128, true><<<1, 128>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 128/8 packed", (k_v1<8, 128, true><<<1, 128>>>(P, 784, 121, out, cyc)));
  RUN("V1 RT=512 PR=2 packed", (k_v1<2, 512, true><<<1, 512>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 512/2 packed", (k_v1<2, 512, true><<<1, 512>>>(P, 784, 121, out, cyc)));
  RUN("V1 RT=512 PR=2", (k_v1<2, 512, false><<<1, 512>>>(pts, 784, 121, out, cyc)), 120);
  RUN("V1 RT=256 PR=4 packed x64", (k_v1<4, 256, true><<<64, 256>>>(pts, 784, 121, out, cyc)), 120);
  RUN("V1 RT=64 PR=14 packed", (k_v1<14, 64, true><<<1, 64>>>(pts, 784, 121, out, cyc)), 120);
  CHECK("V1 64/14 packed", (k_v1<14, 64, true><<<1, 64>>>(P, 784, 121, out, cyc)));
  return 0;
}
