// Stand-alone probe of the tcgen05 building blocks in csrc/tc_dev.cuh (test infrastructure, not product code):
// D[128, N] = A[128, K] . B[K, N] through every operand mode the actor-head kernels use, checked against an FP64 CPU
// product.  usage: tc_probe modeA modeB swapA swapB [N] [K]
//   modeA 0: A K-major  (source A[m][k]),  1: A MN-major (source At[k][m])
//   modeB 0: B K-major  (source Bt[n][k]), 1: B MN-major (source B[k][n])
//   swapX 1: exchange LBO and SBO of that operand's descriptor (hypothesis test)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../cirs_codes_b200/csrc/tc_dev.cuh"

using namespace cirs_tc;

__global__ void __launch_bounds__(128)
probe_kernel(const float* __restrict__ srcA, const float* __restrict__ srcB, float* __restrict__ D, int N, int K,
             int modeA, int modeB, int swapA, int swapB, int passes, int* err) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128;
  const int RA = modeA == 0 ? M : K, CA = modeA == 0 ? K : M;
  const int RB = modeB == 0 ? N : K, CB = modeB == 0 ? K : N;
  char* a_hi = smem;
  char* a_lo = a_hi + RA * CA * 4;
  char* b_hi = a_lo + RA * CA * 4;
  char* b_lo = b_hi + RB * CB * 4;
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tile_stage(a_hi, a_lo, RA, CA, tid, 128, [&](int r, int c4) { return *reinterpret_cast<const float4*>(srcA + (size_t)r * CA + 4 * c4); });
  tile_stage(b_hi, b_lo, RB, CB, tid, 128, [&](int r, int c4) { return *reinterpret_cast<const float4*>(srcB + (size_t)r * CB + 4 * c4); });
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    uint32_t a_lbo = modeA == 0 ? RA * 16 : 128, a_sbo = modeA == 0 ? 128 : RA * 16, a_step = modeA == 0 ? 2 * RA * 16 : 128;
    uint32_t b_lbo = modeB == 0 ? RB * 16 : 128, b_sbo = modeB == 0 ? 128 : RB * 16, b_step = modeB == 0 ? 2 * RB * 16 : 128;
    if (swapA) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
    if (swapB) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t idesc = idesc_tf32(M, N, modeA, modeB);
    if (passes == 3) {
      mma_3xtf32(tb, smem_u32(a_hi), smem_u32(a_lo), a_step, a_lbo, a_sbo, smem_u32(b_hi), smem_u32(b_lo), b_step, b_lbo,
                 b_sbo, idesc, K / 8, false);
    } else {
      for (int j = 0; j < K / 8; ++j)
        mma_tf32(tb, smem_desc(smem_u32(a_hi) + j * a_step, a_lbo, a_sbo),
                 smem_desc(smem_u32(b_hi) + j * b_step, b_lbo, b_sbo), idesc, j > 0);
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
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

int main(int argc, char** argv) {
  const int modeA = argc > 1 ? atoi(argv[1]) : 0, modeB = argc > 2 ? atoi(argv[2]) : 1;
  const int swapA = argc > 3 ? atoi(argv[3]) : 0, swapB = argc > 4 ? atoi(argv[4]) : 0;
  const int N = argc > 5 ? atoi(argv[5]) : 64, K = argc > 6 ? atoi(argv[6]) : 64, M = 128;
  std::vector<float> A(M * K), B(K * N);   // logical A[m][k], B[k][n]
  srand(1234);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  std::vector<float> sA(M * K), sB(K * N);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) sA[modeA == 0 ? m * K + k : k * M + m] = A[m * K + k];
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) sB[modeB == 0 ? n * K + k : k * N + n] = B[k * N + n];
  float *dA, *dB, *dD; int* dErr;
  cudaMalloc(&dA, sA.size() * 4); cudaMalloc(&dB, sB.size() * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dErr, 4);
  cudaMemcpy(dA, sA.data(), sA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, sB.data(), sB.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * (size_t)(M * K + K * N) * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int passes = 1; passes <= 3; passes += 2) {
    cudaMemset(dD, 0, M * N * 4); cudaMemset(dErr, 0, 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, modeA, modeB, swapA, swapB, passes, dErr);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(M * N); int err = 0;
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[k * N + n];
      maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
    }
    printf("modeA=%d modeB=%d swapA=%d swapB=%d N=%d K=%d passes=%d cuda=%s timeout=%d max_abs_err=%.3e (max |ref| %.2f)\n",
           modeA, modeB, swapA, swapB, N, K, passes, cudaGetErrorString(e), err, maxerr, maxref);
    if (maxerr > 1e-2) {
      printf("   D[0][0..7]:"); for (int n = 0; n < 8; ++n) printf(" %9.4f", D[n]); printf("\n ref[0][0..7]:");
      for (int n = 0; n < 8; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[k] * B[k * N + n]; printf(" %9.4f", r); }
      printf("\n   D[1][0..7]:"); for (int n = 0; n < 8; ++n) printf(" %9.4f", D[N + n]); printf("\n ref[1][0..7]:");
      for (int n = 0; n < 8; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[K + k] * B[k * N + n]; printf(" %9.4f", r); }
      printf("\n");
    }
    if (e != cudaSuccess) return 2;
  }
  return 0;
}
