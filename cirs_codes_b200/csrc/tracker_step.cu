// K2: StateTracker rollout step -- one warp per environment, per-environment K/V cache.
// Replaces core/state_tracker.py:188-250 (build_state) + :170-186 (forward) for one new token per call.
//
// Each warp builds the new token (user token at position 0, reward-gated item token afterwards), runs it
// through the post-norm encoder layers attending to the cached keys/values of its own environment, appends
// its key/value to the cache, and writes the decoded state.  Weights are k-major (Wt[in][ldo]) so that lane o
// reads W[., o]: every weight load is a coalesced 128-byte line shared by the whole warp and served from L1/L2
// (the full weight set is 113 KB at d=32).  Algorithmic HBM bytes per env-step: embedding row d*4
// + K/V append 2*nlayers*d*4 + K/V read 2*nlayers*(p+1)*d*4 + state S*4.
#include "tracker_dev.cuh"

namespace {
using namespace cirs_tracker;
constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
tracker_step_kernel(cirs_tracker_weights W, int n_env, int n_rows, const int32_t* __restrict__ env_id,
                    const uint8_t* __restrict__ active, const int32_t* __restrict__ pos_arr, int expect_pos,
                    const int32_t* __restrict__ idx, const float* __restrict__ dense,
                    const float* __restrict__ rew, float* __restrict__ kcache, float* __restrict__ vcache,
                    float* __restrict__ state_out, int64_t state_stride, float* __restrict__ cur_state,
                    int traj_len, float* __restrict__ traj_obs, float* __restrict__ traj_obs_next,
                    int scratch_per_warp) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  const int e = env_id ? env_id[k] : k;
  if (active && !active[e]) return;
  const int p = pos_arr[e];
  if (expect_pos >= 0 && p != expect_pos) return;
  if (p < 0 || p >= W.max_len) return;
  const int n_in = p == 0 ? W.d_user_in : W.d_item_in;
  tracker_token_warp(W, n_env, e, k, p, idx ? idx[k] : 0, dense ? dense + (size_t)k * n_in : nullptr,
                     (p > 0 && rew) ? rew[k] : 0.f, kcache, vcache, smem + (size_t)warp * scratch_per_warp, lane,
                     state_out, state_stride, cur_state, traj_len, traj_obs, traj_obs_next);
}

}  // namespace

extern "C" int cirs_tracker_step(const cirs_tracker_weights* w, int32_t n_env, int32_t n_rows,
                                 const int32_t* env_id, const uint8_t* active, const int32_t* pos,
                                 int32_t expect_pos, const int32_t* idx, const float* dense, const float* rew,
                                 float* kcache, float* vcache, float* state_out, int64_t state_stride,
                                 float* cur_state, int32_t traj_len, float* traj_obs, float* traj_obs_next,
                                 void* stream) {
  if (!w || n_rows < 0 || !pos || !kcache || !vcache) {
    cirs_set_error("cirs_tracker_step: null argument");
    return CIRS_ERR_ARG;
  }
  if (w->d % w->nhead != 0 || w->nlayers > CIRS_MAX_LAYERS || w->d_item_in != w->d || w->dim_state > 32 * 4) {
    cirs_set_error("cirs_tracker_step: unsupported shape (d % nhead, nlayers, d_item_in != d)");
    return CIRS_ERR_ARG;
  }
  if ((!w->emb_user || !w->emb_item) && !dense && !idx) {
    cirs_set_error("cirs_tracker_step: need idx or dense input");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  auto up = [](int v) { return (v + 31) & ~31; };
  const int ldd = up(w->d), ld3 = up(3 * w->d), ldh = up(w->d_hid);
  const int yv = ldd > up((w->d_user_in > w->d_item_in + 1 ? w->d_user_in : w->d_item_in + 1))
                     ? ldd : up((w->d_user_in > w->d_item_in + 1 ? w->d_user_in : w->d_item_in + 1));
  const int per_warp = ldd + yv + ld3 + (ldh > ldd ? ldh : ldd) + up(w->nhead * w->max_len);
  const size_t smem = (size_t)per_warp * WARPS_PER_CTA * sizeof(float);
  if (smem > 200 * 1024) {
    cirs_set_error("cirs_tracker_step: shared memory budget exceeded");
    return CIRS_ERR_ARG;
  }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(tracker_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(tracker_step_kernel, grid, WARPS_PER_CTA * 32, smem, (cudaStream_t)stream, 
      *w, n_env, n_rows, env_id, active, pos, expect_pos, idx, dense, rew, kcache, vcache, state_out,
      state_stride, cur_state, traj_len, traj_obs, traj_obs_next, per_warp);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
