// Small index kernels of the update's front end (no library calls on the path):
//   cirs_update_plan  -- VectorReplayBuffer.sample_index(0) (tianshou/data/buffer/manager.py:144-169) on the device:
//                        from the per-environment transition counts the rollout kernel wrote, the env-major list of
//                        stored buffer slots and the exclusive prefix sum of the counts
//   cirs_gather_i32   -- dst[i] = src[idx[i]]: the minibatch order indices[perm] of one repeat
//                        (tianshou/data/batch.py:733-744 applied to the sampled indices)
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace {

// one CTA: chunked block scan over n_env counts, then every thread expands the environments of its chunk
__global__ void __launch_bounds__(1024)
update_plan_kernel(int n_env, int traj_len, const int32_t* __restrict__ n_slot, int32_t* __restrict__ tok_slot,
                   int32_t* __restrict__ env_off) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_env; base += 1024) {
    const int e = base + tid;
    const int cnt = e < n_env ? min(max(n_slot[e], 0), traj_len) : 0;
    int v = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL_MASK, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;   // inclusive totals of the warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + v - cnt;
    if (e < n_env) {
      env_off[e] = excl;
      if (tok_slot)
        for (int t = 0; t < cnt; ++t) tok_slot[excl + t] = e * traj_len + t;
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) env_off[n_env] = carry_s;
}

__global__ void __launch_bounds__(256)
gather_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// coverage of a collect (evaluation.py:286-371): distinct recommended items (bitset + popcount) and the weighted count
// sum_i weight[act_i] (dominated-category metrics: weight = the per-item value the reference's integer arithmetic yields)
__global__ void __launch_bounds__(256)
coverage_mark_kernel(int n, const int32_t* __restrict__ idx, const int32_t* __restrict__ act, int n_item,
                     const int32_t* __restrict__ weight, uint32_t* __restrict__ bits, unsigned long long* out) {
  __shared__ unsigned long long sh[8];
  unsigned long long w = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int a = act[idx ? idx[i] : i];
    if (a >= 0 && a < n_item) {
      atomicOr(bits + (a >> 5), 1u << (a & 31));
      if (weight) w += (unsigned long long)weight[a];
    }
  }
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(FULL_MASK, w, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    if (t) atomicAdd(out + 2, t);
  }
}
__global__ void __launch_bounds__(256)
coverage_count_kernel(int words, const uint32_t* __restrict__ bits, unsigned long long* out) {
  unsigned long long c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) c += __popc(bits[i]);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace

extern "C" int cirs_coverage_count(int32_t n, const int32_t* idx, const int32_t* act, int32_t n_item,
                                   const int32_t* item_weight, uint32_t* bits, int64_t* out3, void* stream) {
  if (n < 0 || n_item <= 0 || !act || !bits || !out3) {
    cirs_set_error("cirs_coverage_count: bad argument");
    return CIRS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (n_item + 31) / 32;
  cudaMemsetAsync(bits, 0, sizeof(uint32_t) * words, st);
  cudaMemsetAsync(out3, 0, 3 * sizeof(int64_t), st);
  unsigned long long* out = reinterpret_cast<unsigned long long*>(out3);
  if (n > 0) {
    int grid = (n + 255) / 256;
    if (grid > 148 * 4) grid = 148 * 4;
    CIRS_LAUNCH(coverage_mark_kernel, grid, 256, 0, st, n, idx, act, n_item, item_weight, bits, out);
    CIRS_CHECK_LAUNCH();
    CIRS_LAUNCH(coverage_count_kernel, (words + 255) / 256, 256, 0, st, words, bits, out);
    CIRS_CHECK_LAUNCH();
  }
  return CIRS_OK;
}

// stream-ordered zero fill (cudaMemsetAsync): gradient buffers are cleared without a library fill kernel
extern "C" int cirs_zero(void* ptr, int64_t bytes, void* stream) {
  if (bytes < 0 || (bytes > 0 && !ptr)) {
    cirs_set_error("cirs_zero: bad argument");
    return CIRS_ERR_ARG;
  }
  if (bytes == 0) return CIRS_OK;
  const cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    cirs_set_error(cudaGetErrorString(e));
    return CIRS_ERR_CUDA;
  }
  return CIRS_OK;
}

extern "C" int cirs_update_plan(int32_t n_env, int32_t traj_len, const int32_t* n_slot, int32_t* tok_slot,
                                int32_t* env_off, void* stream) {
  if (n_env < 0 || traj_len <= 0 || !n_slot || !env_off) {
    cirs_set_error("cirs_update_plan: bad argument");
    return CIRS_ERR_ARG;
  }
  CIRS_LAUNCH(update_plan_kernel, 1, 1024, 0, (cudaStream_t)stream, n_env, traj_len, n_slot, tok_slot, env_off);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_gather_i32(int32_t* dst, const int32_t* src, const int32_t* idx, int32_t n, void* stream) {
  if (n < 0 || (n > 0 && (!dst || !src || !idx))) {
    cirs_set_error("cirs_gather_i32: bad argument");
    return CIRS_ERR_ARG;
  }
  if (n == 0) return CIRS_OK;
  CIRS_LAUNCH(gather_i32_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, dst, src, idx, n);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
