// Block-level negative-pair permutation draw (see sampler.cu), shared with the FPS kernel, which runs it in one extra
// CTA so that the draw costs no launch and no time on the step's critical path.
#pragma once
#include <curand_kernel.h>

#include "common.cuh"

namespace dg {

// n permutations of 0..B-1 (Fisher-Yates on a Philox4x32-10 stream keyed by (seed, subsequence k, offset)), fixed points
// bumped by one, mod B - /root/reference/src/modules.py:1184-1188.  Needs >= n threads and n * B ints of shared memory.
__device__ __forceinline__ void super_perms_block(unsigned long long seed, unsigned long long offset, int n, int B,
                                                  int64_t* __restrict__ out, int* sp_buf) {
  const int k = threadIdx.x;
  if (k < n) {               // thread k owns column k of sp_buf[i * n + k] (conflict-free)
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)k, offset, &st);
    for (int i = 0; i < B; ++i) sp_buf[i * n + k] = i;
    for (int i = B - 1; i > 0; --i) {
      const unsigned int r = curand(&st);
      const int j = (int)(((unsigned long long)r * (unsigned long long)(i + 1)) >> 32);
      const int t = sp_buf[i * n + k];
      sp_buf[i * n + k] = sp_buf[j * n + k];
      sp_buf[j * n + k] = t;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * B; e += blockDim.x) {     // coalesced write-out with the fixed-point bump
    const int kk = e / B, i = e - kk * B;
    int v = sp_buf[i * n + kk];
    if (v == i) v += 1;    // perm[perm == arange] += 1
    out[e] = (int64_t)(v % B);   // perm % size
  }
}

}  // namespace dg
