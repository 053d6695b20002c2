// Shared device helpers for the CIRS B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define CIRS_OK 0
#define CIRS_ERR_ARG 1
#define CIRS_ERR_CUDA 2

#define CIRS_CHECK_LAUNCH()                              \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) {                            \
      cirs_set_error(cudaGetErrorString(e__));           \
      return CIRS_ERR_CUDA;                              \
    }                                                    \
  } while (0)

void cirs_set_error(const char* msg);
void cirs_note_launch(void);  // counts kernel launches issued by this library (bench.py's gpu_launches)
// optional per-kernel timing with CUDA events on the launching stream (cirs_profile_enable, error.cu)
bool cirs_profile_begin(const char* name, cudaStream_t st);
void cirs_profile_end(cudaStream_t st);

// every kernel launch of the library goes through this macro (grid, block, dynamic smem, stream, args...)
#define CIRS_LAUNCH(kernel, grid, block, smem, stream, ...)       \
  do {                                                            \
    const bool prof__ = cirs_profile_begin(#kernel, (stream));    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
    cirs_note_launch();                                           \
    if (prof__) cirs_profile_end((stream));                       \
  } while (0)

#define FULL_MASK 0xffffffffu
#define CATEGORICAL_EPS 1.1920928955078125e-07f  // torch.finfo(float32).eps, Categorical clamp_probs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

// ---- Philox4x32-10 (counter-based RNG; one call -> 4 x 32 random bits) ----
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in (0, 1]
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 8) + 1) * (1.0f / 16777216.0f); }

// ---- cp.async (LDGSTS): global -> shared without staging registers; completion by commit / wait groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
