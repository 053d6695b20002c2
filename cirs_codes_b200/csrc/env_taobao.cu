// K1': vectorised SimulatedEnv(VirtualTB) step -- one warp per environment.
// Replaces core/env/simulatedEnv/simulated_env.py:77-168 over environments/VirtualTaobao/virtualTB/envs/virtualTB.py:74-133
// and core/util.py:21-46 (one Python object per environment, stepped in a for-loop, tianshou/env/venvs.py:212-220),
// including the reward model evaluated inside every step, UserModel_MMOE.forward (core/user_model_mmoe.py:144-220).
//
// Per warp: lanes 0..26 hold the action; the exit test is min(t, N-1) warp-reduced float32 distances; the exposure
// sum gives lane j history slot j (float64 like the reference); the reward model is four k-major matvecs
// (118 -> 64 -> 64 -> {32 experts, 4 gates}) through shared memory.  Algorithmic HBM bytes per env-step:
// 705 + 108 * (min(t, N-1) + t); the reward model's 57 KB of weights are L1/L2 resident.
#include "taobao_dev.cuh"

namespace {
using namespace cirs_taobao;
constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
taobao_reset_kernel(cirs_taobao_env E, int n_rows, const int32_t* __restrict__ env_id,
                    const float* __restrict__ users, uint8_t* __restrict__ active) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  taobao_reset_warp(E, env_id ? env_id[k] : k, users + (size_t)k * NU, lane, active);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
taobao_step_kernel(cirs_taobao_env E, int n_rows, const int32_t* __restrict__ env_id, uint8_t* __restrict__ active,
                   const float* __restrict__ act, float* __restrict__ act_env, float* __restrict__ rew,
                   uint8_t* __restrict__ done, int traj_len, float* __restrict__ traj_act,
                   float* __restrict__ traj_act_env, float* __restrict__ traj_rew, uint8_t* __restrict__ traj_done, int32_t* __restrict__ ep_len,
                   int force_length) {
  __shared__ float smem[WARPS_PER_CTA * STEP_SCRATCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  const int e = env_id ? env_id[k] : k;
  if (active && !active[e]) return;
  taobao_step_warp(E, e, k, act + (size_t)k * NI, lane, smem + warp * STEP_SCRATCH, active, act_env, rew, done,
                   traj_len, traj_act, traj_act_env, traj_rew, traj_done, ep_len, force_length);
}

bool bad_model(const cirs_mmoe_weights& m) {
  return !m.lin_w || !m.w1t || !m.b1 || !m.w2t || !m.b2 || !m.wet || !m.be || !m.wgt || !m.bg || !m.tower ||
         m.n_in != NU + 3 + NI || m.h1 < 1 || m.h1 > 128 || m.h2 < 1 || m.h2 > 128 || m.n_expert < 1 ||
         m.n_expert > 32 || m.expert_dim < 1 || m.n_expert * m.expert_dim > 64;
}

}  // namespace

extern "C" int cirs_taobao_reset(const cirs_taobao_env* env, int32_t n_rows, const int32_t* env_id,
                                 const float* users, uint8_t* active, void* stream) {
  if (!env || !users || n_rows < 0 || !env->user || !env->turn || !env->prev_rew || !env->cum_rew) {
    cirs_set_error("cirs_taobao_reset: null argument");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(taobao_reset_kernel, grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream, *env, n_rows, env_id, users,
              active);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_taobao_step(const cirs_taobao_env* env, int32_t n_rows, const int32_t* env_id, uint8_t* active,
                                const float* act, float* act_env, float* rew, uint8_t* done, int32_t traj_len,
                                float* traj_act, float* traj_act_env, float* traj_rew, uint8_t* traj_done,
                                int32_t* ep_len, int32_t force_length, void* stream) {
  if (!env || !act || !rew || !done || n_rows < 0 || !env->hist || !env->user || !env->turn) {
    cirs_set_error("cirs_taobao_step: null argument");
    return CIRS_ERR_ARG;
  }
  if (bad_model(env->um)) {
    cirs_set_error("cirs_taobao_step: unsupported reward model (n_in 118, hidden <= 128, experts*dim <= 64)");
    return CIRS_ERR_ARG;
  }
  if (traj_rew && !traj_done) {
    cirs_set_error("cirs_taobao_step: trajectory outputs must be given together");
    return CIRS_ERR_ARG;
  }
  if (force_length > env->max_turn || (traj_rew && force_length > traj_len)) {
    cirs_set_error("cirs_taobao_step: force_length exceeds env->max_turn or traj_len");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(taobao_step_kernel, grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream, *env, n_rows, env_id, active,
              act, act_env, rew, done, traj_len, traj_act, traj_act_env, traj_rew, traj_done, ep_len, force_length);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
