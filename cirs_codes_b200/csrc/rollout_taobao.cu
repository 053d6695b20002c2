// A whole Collector.collect(n_episode = B) on SimulatedEnv(VirtualTB) as ONE kernel (core/collector.py:147-367 with
// the fork's semantics: everything reset at the start, no reset on done, finished environments dropped).
//
// Unlike KuaishouEnv (a 10728-wide softmax head that is tiled ACROSS environments), nothing in a VirtualTaobao turn
// couples environments: the policy head is 27 wide, the environment step and the tracker token are per environment.
// So one warp owns one environment for its entire episode and the kernel needs no grid-wide barrier at all:
//   user token -> [ trunk + mu head + Normal sample -> map_action -> exit test / exposure / MMOE reward ->
//                   gated action token through the encoder (K/V cache) -> replay-buffer slots ] x turns.
// All intermediate vectors live in the warp's shared-memory scratch; HBM sees the K/V cache, the history and the
// replay-buffer rows.  Device code shared with the stand-alone kernels (taobao_dev.cuh, tracker_dev.cuh).
#include "taobao_dev.cuh"

namespace {
using namespace cirs_taobao;
constexpr int WARPS_PER_CTA = 4;

struct Args {
  cirs_taobao_env E;
  cirs_tracker_weights T;
  cirs_policy_weights P;
  const float* users;
  uint8_t* active;
  float* cur_state;
  int traj_len;
  float *traj_obs, *traj_obs_next, *traj_act, *traj_act_env, *traj_rew;
  uint8_t* traj_done;
  int32_t* ep_len;
  float *kcache, *vcache;
  uint64_t seed;
  const unsigned long long* rng_counter;
  int mode, max_steps, force_length, scratch_per_warp, tracker_floats;
};

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) rollout_taobao_kernel(Args A) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * WARPS_PER_CTA + warp;
  const int B = A.E.n_env;
  if (e >= B) return;
  float* trk = smem + (size_t)warp * A.scratch_per_warp;   // tracker scratch
  float* sc = trk + A.tracker_floats;                      // env-step scratch (sc[0..27) = mapped action)
  float* sp = sc + STEP_SCRATCH;                           // actor scratch; sp[224..256) = raw action
  const int S = A.T.dim_state, nA = A.P.n_action;
  const uint64_t off0 = A.rng_counter ? *A.rng_counter : 0ull;
  taobao_reset_warp(A.E, e, A.users + (size_t)e * NU, lane, A.active);
  if (lane == 0) A.ep_len[e] = 0;
  __syncwarp();
  cirs_tracker::tracker_token_warp<false>(A.T, B, e, e, 0, 0, A.users + (size_t)e * NU, 0.f, A.kcache, A.vcache, trk,
                                          lane, nullptr, 0, A.cur_state, A.traj_len, A.traj_obs, A.traj_obs_next);
  __syncwarp();
  float* raw = sp + ACTOR_SCRATCH;
  for (int t = 0; t < A.max_steps; ++t) {
    const float a = actorprob_warp(A.P, A.cur_state + (size_t)e * S, lane, sp, nullptr, A.seed, off0 + (uint64_t)t, e,
                                   A.mode, nullptr, nullptr, nullptr, nullptr);
    raw[lane] = lane < nA ? a : 0.f;
    __syncwarp();
    const bool d = taobao_step_warp(A.E, e, e, raw, lane, sc, A.active, nullptr, nullptr, nullptr, A.traj_len,
                                    A.traj_act, A.traj_act_env, A.traj_rew, A.traj_done, A.ep_len, A.force_length);
    const float r = (float)A.E.prev_rew[e];
    cirs_tracker::tracker_token_warp<false>(A.T, B, e, e, t + 1, 0, sc, r, A.kcache, A.vcache, trk, lane, nullptr, 0,
                                            A.cur_state, A.traj_len, A.traj_obs, A.traj_obs_next);
    __syncwarp();
    if (d) break;
  }
}

__global__ void advance_counter_kernel(unsigned long long* c, unsigned long long n) { *c += n; }

}  // namespace

extern "C" int cirs_rollout_taobao(const cirs_taobao_env* env, const cirs_tracker_weights* tw,
                                   const cirs_policy_weights* pw, const float* users, uint8_t* active,
                                   float* cur_state, int32_t traj_len, float* traj_obs, float* traj_obs_next,
                                   float* traj_act, float* traj_act_env, float* traj_rew, uint8_t* traj_done,
                                   int32_t* ep_len, float* kcache, float* vcache, int32_t kv_n_env, uint64_t seed,
                                   uint64_t* rng_counter, int32_t mode, int32_t max_steps, int32_t force_length,
                                   void* stream) {
  if (!env || !tw || !pw || !users || !active || !cur_state || !traj_obs || !traj_obs_next || !traj_act || !traj_act_env || !traj_rew ||
      !traj_done || !ep_len || !kcache || !vcache || max_steps < 1) {
    cirs_set_error("cirs_rollout_taobao: null argument");
    return CIRS_ERR_ARG;
  }
  if (tw->emb_user || tw->emb_item || tw->d_user_in != NU || tw->d_item_in != NI || tw->d != NI ||
      tw->d % tw->nhead != 0 || tw->nlayers > CIRS_MAX_LAYERS || max_steps + 1 > tw->max_len || !pw->sigma ||
      pw->n_action != NI || pw->dim_state != tw->dim_state || pw->dim_state > 32 || !env->map_action) {
    cirs_set_error("cirs_rollout_taobao: unsupported shapes (dense 88 / 27 tracker inputs, d = 27, continuous actor "
                   "with 27 actions, env->map_action = 1, max_steps < max_len)");
    return CIRS_ERR_ARG;
  }
  if (kv_n_env != env->n_env) {
    cirs_set_error("cirs_rollout_taobao: the K/V caches are sized for a different number of environments "
                   "(kv_n_env != env->n_env): rebuild them (build_state(dim_batch, reset=True)) before the collect");
    return CIRS_ERR_ARG;
  }
  if (force_length > env->max_turn || max_steps > env->max_turn || max_steps > traj_len || force_length > traj_len) {
    cirs_set_error("cirs_rollout_taobao: max_steps / force_length exceed env->max_turn or the trajectory length");
    return CIRS_ERR_ARG;
  }
  Args A{};
  A.E = *env; A.T = *tw; A.P = *pw; A.users = users; A.active = active; A.cur_state = cur_state;
  A.traj_len = traj_len; A.traj_obs = traj_obs; A.traj_obs_next = traj_obs_next; A.traj_act = traj_act; A.traj_act_env = traj_act_env;
  A.traj_rew = traj_rew; A.traj_done = traj_done; A.ep_len = ep_len; A.kcache = kcache; A.vcache = vcache;
  A.seed = seed; A.rng_counter = reinterpret_cast<const unsigned long long*>(rng_counter); A.mode = mode;
  A.max_steps = max_steps; A.force_length = force_length;
  A.tracker_floats = cirs_tracker::tracker_scratch_floats(*tw);
  A.scratch_per_warp = A.tracker_floats + STEP_SCRATCH + ACTOR_SCRATCH + 32;
  const size_t smem = (size_t)A.scratch_per_warp * WARPS_PER_CTA * sizeof(float);
  if (smem > 200 * 1024) {
    cirs_set_error("cirs_rollout_taobao: shared memory budget exceeded");
    return CIRS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(rollout_taobao_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = (env->n_env + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(rollout_taobao_kernel, grid, WARPS_PER_CTA * 32, smem, st, A);
  CIRS_CHECK_LAUNCH();
  if (rng_counter) {
    CIRS_LAUNCH(advance_counter_kernel, 1, 1, 0, st, reinterpret_cast<unsigned long long*>(rng_counter),
                (unsigned long long)max_steps);
    CIRS_CHECK_LAUNCH();
  }
  return CIRS_OK;
}
