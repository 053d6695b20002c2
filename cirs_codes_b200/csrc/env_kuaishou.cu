// K1: vectorised KuaishouEnv / SimulatedEnv step -- one warp per environment.
// Replaces core/env/simulatedEnv/simulated_env.py:111-168, environments/KuaishouRec/env/kuaishouEnv.py:161-218
// and core/util.py:21-54 (per-environment Python objects stepped in a for-loop, tianshou/env/venvs.py:212-220).
//
// Lane j of the warp owns history slot j: it loads hist[e][j] (coalesced int32), gathers that item's category
// bitmask, and contributes (a) its term of the exposure sum  exp(-(t-j) * dist(a, a_j) / tau), (b) its vote in
// the windowed category-overlap exit test, (c) its vote in the repeat count.  Warp shuffles / ballots reduce.
// Algorithmic HBM bytes per env-step:  57 + 20*w(t,N) + 8*t  (gathered-distance variant)
// or 57 + 20*w + 20*t (distance recomputed from the category masks, the default).
#include "env_dev.cuh"

namespace {
using namespace cirs_env;

constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
kuaishou_reset_kernel(cirs_kuaishou_env E, int n_rows, const int32_t* __restrict__ env_id,
                      const int32_t* __restrict__ users, uint8_t* __restrict__ active) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  kuaishou_reset_warp(E, env_id ? env_id[k] : k, users[k], lane, active);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
kuaishou_step_kernel(cirs_kuaishou_env E, int n_rows, const int32_t* __restrict__ env_id,
                     uint8_t* __restrict__ active, const int32_t* __restrict__ act, float* __restrict__ rew,
                     uint8_t* __restrict__ done, int traj_len, int32_t* __restrict__ traj_act,
                     float* __restrict__ traj_rew, uint8_t* __restrict__ traj_done, int32_t* __restrict__ ep_len,
                     int force_length) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  const int e = env_id ? env_id[k] : k;
  if (active && !active[e]) return;
  kuaishou_step_warp(E, e, k, act[k], lane, active, rew, done, traj_len, traj_act, traj_rew, traj_done, ep_len,
                     force_length, nullptr);
}

}  // namespace

extern "C" int cirs_kuaishou_reset(const cirs_kuaishou_env* env, int32_t n_rows, const int32_t* env_id,
                                   const int32_t* users, uint8_t* active, void* stream) {
  if (!env || !users || n_rows < 0) {
    cirs_set_error("cirs_kuaishou_reset: null argument");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(kuaishou_reset_kernel, grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream, *env, n_rows, env_id, users,
                                                                                   active);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_kuaishou_step(const cirs_kuaishou_env* env, int32_t n_rows, const int32_t* env_id,
                                  uint8_t* active, const int32_t* act, float* rew, uint8_t* done,
                                  int32_t traj_len, int32_t* traj_act, float* traj_rew, uint8_t* traj_done,
                                  int32_t* ep_len, int32_t force_length, void* stream) {
  if (!env || !act || !rew || !done || n_rows < 0) {
    cirs_set_error("cirs_kuaishou_step: null argument");
    return CIRS_ERR_ARG;
  }
  if (env->simulated ? !env->normed_mat : !env->mat) {
    cirs_set_error("cirs_kuaishou_step: reward table missing");
    return CIRS_ERR_ARG;
  }
  if ((env->alpha_u == nullptr) != (env->beta_i == nullptr)) {
    cirs_set_error("cirs_kuaishou_step: alpha_u and beta_i must be given together");
    return CIRS_ERR_ARG;
  }
  if (traj_act && (!traj_rew || !traj_done)) {
    cirs_set_error("cirs_kuaishou_step: trajectory outputs must be given together");
    return CIRS_ERR_ARG;
  }
  if (force_length > env->max_turn || (traj_act && force_length > traj_len)) {
    // the history has max_turn slots and the trajectory traj_len: a longer forced episode would read / write past them
    cirs_set_error("cirs_kuaishou_step: force_length exceeds env->max_turn or traj_len");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(kuaishou_step_kernel, grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream, 
      *env, n_rows, env_id, active, act, rew, done, traj_len, traj_act, traj_rew, traj_done, ep_len,
      force_length);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
