// Negative-pair permutations in one launch.
//
// super_perm (/root/reference/src/modules.py:1184-1188) draws torch.randperm(B), bumps fixed
// points by one and reduces mod B; the loss calls it neg_samples times per step (:1340-1341),
// which costs ~6 library kernels per draw.  This kernel draws all neg_samples permutations at
// once: one thread per permutation runs a Fisher-Yates shuffle on a Philox4x32-10 stream keyed
// by the caller's (seed, offset) — the caller advances its generator, so runs are reproducible
// under torch.manual_seed — then applies the same fixed-point bump.  Same distribution as the
// reference, different random stream (the module keeps the exact torch stream as an option).
#include "kernels.cuh"
#include "sampler.cuh"

namespace dg {

// `use_smem`: the shuffle runs in shared memory (n * B ints) - a Fisher-Yates walk is a chain of dependent loads and
// stores, ~0.4 us a step in global memory and ~50 ns in shared memory; permutations that do not fit shuffle in place.
__global__ void super_perms_kernel(unsigned long long seed, unsigned long long offset, int n, int B,
                                   int64_t* __restrict__ out, int use_smem) {
  extern __shared__ int sp_buf[];
  pdl_trigger();
  pdl_wait();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (use_smem) {
    super_perms_block(seed, offset, n, B, out, sp_buf);
    return;
  }
  if (k >= n) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)k, offset, &st);
  int64_t* p = out + (size_t)k * B;
  for (int i = 0; i < B; ++i) p[i] = i;
  for (int i = B - 1; i > 0; --i) {
    const unsigned int r = curand(&st);
    const int j = (int)(((unsigned long long)r * (unsigned long long)(i + 1)) >> 32);
    const int64_t t = p[i];
    p[i] = p[j];
    p[j] = t;
  }
  for (int i = 0; i < B; ++i) {
    int64_t v = p[i];
    if (v == i) v += 1;  // perm[perm == arange] += 1
    p[i] = v % B;        // perm % size
  }
}

}  // namespace dg

extern "C" int dg_super_perms(unsigned long long seed, unsigned long long offset, int n, int B, int64_t* out,
                              dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(out, DG_ERR_INVALID, "dg_super_perms: null pointer");
  DG_REQUIRE(n > 0 && B > 0, DG_ERR_INVALID, "dg_super_perms: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)n * B * sizeof(int);
  const bool use_smem = n <= 128 && smem <= 48 * 1024;
  DG_PRE(st);
  if (use_smem)
    launch_pdl(super_perms_kernel, dim3(1), dim3(128), smem, st, seed, offset, n, B, out, 1);
  else
    launch_pdl(super_perms_kernel, dim3(ceil_div(n, 32)), dim3(32), 0, st, seed, offset, n, B, out, 0);
  DG_LAUNCH_OK("super_perms_kernel");
  return DG_OK;
}
