// Timing probe (test infrastructure): cycles of one 3xTF32 MMA batch (24 tcgen05.mma M128 x N x K8, issue -> commit ->
// mbarrier wait) for N = 64 / 128 / 256 with the no-swizzle K-major operand tiles of csrc/tc_dev.cuh, one CTA per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cirs_codes_b200/csrc/tc_dev.cuh"
using namespace cirs_tc;

__global__ void __launch_bounds__(128)
time_kernel(int N, int reps, long long* out) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  char* a_hi = smem; char* a_lo = a_hi + 128 * 64 * 4; char* b_hi = a_lo + 128 * 64 * 4; char* b_lo = b_hi + N * 64 * 4;
  for (int i = tid; i < (2 * 128 * 64 + 2 * N * 64); i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tb = tmem_base, idesc = idesc_tf32(128, N, 0, 0);
  long long t0 = 0, t1 = 0;
  uint32_t ph = 0;
  for (int r = 0; r < reps + 1; ++r) {
    if (r == 1) t0 = clock64();
    if (tid == 0) {
      mma_3xtf32(tb, smem_u32(a_hi), smem_u32(a_lo), 2 * 128 * 16, 128 * 16, 128, smem_u32(b_hi), smem_u32(b_lo), 2 * N * 16,
                 N * 16, 128, idesc, 8, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, ph); ph ^= 1;
    fence_after_sync();
    __syncthreads();
  }
  t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / reps;
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  for (int N : {64, 128, 256}) {
    const size_t smem = 2 * 128 * 64 * 4 + 2 * (size_t)N * 64 * 4;
    cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {1, 148}) {
      time_kernel<<<grid, 128, smem>>>(N, 200, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d grid=%3d: %lld cycles per batch of 24 MMAs (%.0f cycles per MMA, %.1f GMAC/s-per-SM-eq) %s\n", N, grid, c,
             c / 24.0, 24.0 * 128 * N * 8 / (double)c * 1.965, cudaGetErrorString(e));
    }
  }
  return 0;
}
