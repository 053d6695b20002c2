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

}  // namespace

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
