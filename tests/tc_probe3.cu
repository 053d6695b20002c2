// Probe (test infrastructure): tcgen05.mma with the A operand in TMEM (written with tcgen05.st), 3xTF32, against FP64.
// D[128, 64] = A[128, 64] . B^T, B given as Bt[n][k] (K-major shared-memory tile).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../cirs_codes_b200/csrc/tc_dev.cuh"
using namespace cirs_tc;

__global__ void __launch_bounds__(128)
probe_kernel(const float* __restrict__ A, const float* __restrict__ Bt, float* __restrict__ D, int* err) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, N = 64, K = 64;
  char* b_hi = smem; char* b_lo = smem + N * K * 4;
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tile_stage(b_hi, b_lo, N, K, tid, 128, [&](int r, int c4) { return *reinterpret_cast<const float4*>(Bt + (size_t)r * K + 4 * c4); });
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tb = tmem_base;
  // A row tid -> TMEM lane tid: hi at columns [64, 128), lo at [128, 192); D at [0, 64)
  for (int c0 = 0; c0 < K; c0 += 16) {
    float hi[16], lo[16];
    for (int j = 0; j < 16; ++j) { const float x = A[(size_t)tid * K + c0 + j]; hi[j] = tf32_hi(x); lo[j] = tf32_hi(x - hi[j]); }
    tmem_st16(tmem_addr(tb, warp * 32, 64 + c0), hi);
    tmem_st16(tmem_addr(tb, warp * 32, 128 + c0), lo);
  }
  tmem_st_wait();
  fence_before_sync(); __syncthreads(); fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, N, 0, 0);
    uint32_t acc = 0;
    for (int j = 0; j < K / 8; ++j) {
      const uint64_t bh = smem_desc(smem_u32(b_hi) + j * 2 * N * 16, N * 16, 128), bl = smem_desc(smem_u32(b_lo) + j * 2 * N * 16, N * 16, 128);
      mma_tf32_ts(tb, tb + 128 + 8 * j, bh, idesc, acc);      // lo . hi
      mma_tf32_ts(tb, tb + 64 + 8 * j, bl, idesc, 1u);        // hi . lo
      mma_tf32_ts(tb, tb + 64 + 8 * j, bh, idesc, 1u);        // hi . hi
      acc = 1u;
    }
    mma_commit(&bar);
  }
  if (!mbar_wait(&bar, 0)) { if (tid == 0) *err = 1; }
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem_addr(tb, warp * 32, c0), v);
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = v[j];
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
  const int M = 128, N = 64, K = 64;
  std::vector<float> A(M * K), Bt(N * K);
  srand(7);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : Bt) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD; int* dErr;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bt.size() * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dErr, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4); cudaMemset(dErr, 0, 4);
  const size_t smem = 2 * N * K * 4;
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(M * N); int err = 0;
  cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * Bt[n * K + k];
    maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
  }
  printf("A-in-TMEM 3xTF32: cuda=%s timeout=%d max_abs_err=%.3e (max |ref| %.2f)  D[0][0..3]= %.4f %.4f %.4f %.4f\n",
         cudaGetErrorString(e), err, maxerr, maxref, D[0], D[1], D[2], D[3]);
  return 0;
}
