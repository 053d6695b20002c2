// Persistent rollout kernel: a whole Collector.collect(n_episode = B) in ONE cooperative launch.
//
// Replaces the reference's while-loop over turns (core/collector.py:219-320), i.e. per turn policy.forward
// (core/policy/ppo.py:111-163) -> env.step (simulated_env.py:111-168, kuaishouEnv.py:161-218) -> build_state
// (core/state_tracker.py:225-248) -> buffer.add (tianshou/data/buffer/manager.py:91-142), for all B environments.
//
// Why one kernel: a turn's work is tiny (~0.3 KB of algorithmic HBM traffic per env-step), so a
// launch-per-component rollout is bound by launch latency and dependent-kernel drain (measured: 2.1 ms of GPU time
// per collect as a 150-node CUDA graph, of which most is idle turns).  Here the grid stays resident; each turn is
// two phases separated by grid-wide barriers:
//   A  actor head: (64-row tile x catalogue split) work items, trunk + logits tiles + online softmax / race partials
//   B  per environment (one warp): merge the split partials -> action, log-prob; environment step; tracker token
//      against the K/V cache; trajectory written straight into the replay buffer's slots
// and the loop ends as soon as the device-side count of running environments reaches zero.
// Device code is shared with the stand-alone kernels (actor_dev.cuh, env_dev.cuh, tracker_dev.cuh), so both paths
// compute bit-identical results.
#include <cooperative_groups.h>

#include "actor_dev.cuh"
#include "env_dev.cuh"
#include "tracker_dev.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace cirs_actor;

struct RolloutArgs {
  cirs_kuaishou_env E;
  cirs_tracker_weights T;
  HeadArgs H;               // policy weights, mode, seed / rng counter, partial workspace; state = cur_state
  int n_env, max_steps, force_length, traj_len;
  const int32_t* users;
  uint8_t* active;
  int32_t* act;             // [B] last action per slot
  float *logp, *value;      // [B]
  float* cur_state;         // [B, S]
  float *rew;               // [B]
  uint8_t* done;            // [B]
  float *traj_obs, *traj_obs_next;
  int32_t* traj_act;
  float* traj_rew;
  uint8_t* traj_done;
  int32_t* ep_len;
  float *kcache, *vcache;
  int* n_active;            // device counter of running environments
  int scratch_per_warp;
};

__global__ void __launch_bounds__(NT, 2) rollout_kuaishou_kernel(RolloutArgs A) {
  extern __shared__ __align__(16) float smem_dyn[];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warps_per_cta = NT / 32;
  const int gwarp = blockIdx.x * warps_per_cta + warp, n_warps = gridDim.x * warps_per_cta;
  const int B = A.n_env, S = A.T.dim_state;
  float* scratch = smem_dyn + (size_t)warp * A.scratch_per_warp;

  // ---- reset + user token (position 0)
  if (blockIdx.x == 0 && tid == 0) *A.n_active = B;
  for (int e = gwarp; e < B; e += n_warps) {
    const int u = A.users[e];
    cirs_env::kuaishou_reset_warp(A.E, e, u, lane, A.active);
    if (lane == 0) A.ep_len[e] = 0;
    cirs_tracker::tracker_token_warp(A.T, B, e, e, 0, u, nullptr, 0.f, A.kcache, A.vcache, scratch, lane, nullptr, 0,
                                     A.cur_state, A.traj_len, A.traj_obs, A.traj_obs_next);
  }
  __threadfence();
  grid.sync();

  const int row_tiles = (B + BM - 1) / BM, n_items = row_tiles * A.H.n_split;
  for (int t = 0; t < A.max_steps; ++t) {
    // ---- phase A: actor head partials
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      actor_head_body(A.H, w % row_tiles, w / row_tiles, smem_dyn);
      __syncthreads();
    }
    __threadfence();
    grid.sync();
    // ---- phase B: one warp per environment
    for (int e = gwarp; e < B; e += n_warps) {
      if (!A.active[e]) continue;
      int a = 0;
      if (lane == 0) a = actor_combine_row(A.H, e, A.act, A.logp);
      a = __shfl_sync(FULL_MASK, a, 0);
      cirs_env::kuaishou_step_warp(A.E, e, e, a, lane, A.active, A.rew, A.done, A.traj_len, A.traj_act, A.traj_rew,
                                   A.traj_done, A.ep_len, A.force_length, A.n_active);
      __syncwarp();
      const float r = A.rew[e];
      cirs_tracker::tracker_token_warp(A.T, B, e, e, t + 1, a, nullptr, r, A.kcache, A.vcache, scratch, lane, nullptr,
                                       0, A.cur_state, A.traj_len, A.traj_obs, A.traj_obs_next);
    }
    if (blockIdx.x == 0 && tid == 0 && A.H.rng_counter) *A.H.rng_counter += 1ull;
    __threadfence();
    grid.sync();
    if (*reinterpret_cast<volatile int*>(A.n_active) <= 0) break;
  }
}

}  // namespace

// workspace: head partials for (n_split + 1) * n_env rows + the running-environment counter
extern "C" int64_t cirs_rollout_workspace_bytes(int32_t n_env, int32_t n_action) {
  return cirs_actor_workspace_bytes(n_env, n_action) + (int64_t)sizeof(Partial) * 64 * (int64_t)n_env + 256;
}

extern "C" int cirs_rollout_kuaishou(const cirs_kuaishou_env* env, const cirs_tracker_weights* tw,
                                     const cirs_policy_weights* pw, const int32_t* users, uint8_t* active,
                                     int32_t* act, float* logp, float* value, float* cur_state, float* rew,
                                     uint8_t* done, int32_t traj_len, float* traj_obs, float* traj_obs_next,
                                     int32_t* traj_act, float* traj_rew, uint8_t* traj_done, int32_t* ep_len,
                                     float* kcache, float* vcache, uint64_t seed, uint64_t* rng_counter, int32_t mode,
                                     int32_t max_steps, int32_t force_length, void* workspace, void* stream) {
  if (!env || !tw || !pw || !users || !active || !act || !logp || !value || !cur_state || !rew || !done || !ep_len ||
      !kcache || !vcache || !workspace || max_steps < 0) {
    cirs_set_error("cirs_rollout_kuaishou: null argument");
    return CIRS_ERR_ARG;
  }
  if (!tw->emb_user || !tw->emb_item || tw->d % tw->nhead != 0 || tw->nlayers > CIRS_MAX_LAYERS ||
      tw->d_item_in != tw->d || tw->d_user_in != tw->d || max_steps + 1 > tw->max_len) {
    cirs_set_error("cirs_rollout_kuaishou: unsupported tracker shape (needs embedding tables, max_steps < max_len)");
    return CIRS_ERR_ARG;
  }
  if (env->n_env <= 0) return CIRS_OK;
  static int max_ctas = 0, n_sm = 0;
  const int per_warp = cirs_tracker::tracker_scratch_floats(*tw);
  size_t smem = SMEM_BYTES;
  if ((size_t)per_warp * (NT / 32) * sizeof(float) > smem) smem = (size_t)per_warp * (NT / 32) * sizeof(float);
  if (smem > 200 * 1024) {
    cirs_set_error("cirs_rollout_kuaishou: shared memory budget exceeded");
    return CIRS_ERR_ARG;
  }
  if (!max_ctas) {
    int dev = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(rollout_kuaishou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel, NT, smem);
    if (per_sm < 1) {
      cirs_set_error("cirs_rollout_kuaishou: kernel does not fit on an SM");
      return CIRS_ERR_CUDA;
    }
    max_ctas = per_sm * n_sm;
  }
  RolloutArgs A{};
  A.E = *env; A.T = *tw;
  A.H.W = *pw; A.H.n_rows = env->n_env; A.H.gather = nullptr; A.H.state_by_k = 0; A.H.out_by_k = 1;
  A.H.active = active; A.H.state = cur_state; A.H.state_stride = tw->dim_state; A.H.noise_q = nullptr;
  A.H.seed = seed; A.H.offset = 1ull << 40; A.H.rng_counter = reinterpret_cast<unsigned long long*>(rng_counter);
  A.H.mode = mode; A.H.seen = nullptr; A.H.act_in = nullptr; A.H.value = value;
  int grid = max_ctas;
  if (!plan_head(A.H, grid) || A.H.n_split > 64) {
    cirs_set_error("cirs_rollout_kuaishou: unsupported head shape");
    return CIRS_ERR_ARG;
  }
  A.H.part = reinterpret_cast<Partial*>(workspace);
  A.n_active = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) +
                                      sizeof(Partial) * (size_t)(A.H.n_split + 1) * env->n_env);
  A.n_active = reinterpret_cast<int*>(((uintptr_t)A.n_active + 63) & ~(uintptr_t)63);
  A.n_env = env->n_env; A.max_steps = max_steps; A.force_length = force_length; A.traj_len = traj_len;
  A.users = users; A.active = active; A.act = act; A.logp = logp; A.value = value; A.cur_state = cur_state;
  A.rew = rew; A.done = done; A.traj_obs = traj_obs; A.traj_obs_next = traj_obs_next; A.traj_act = traj_act;
  A.traj_rew = traj_rew; A.traj_done = traj_done; A.ep_len = ep_len; A.kcache = kcache; A.vcache = vcache;
  A.scratch_per_warp = per_warp;
  // no more CTAs than there is work for (warps for phase B, work items for phase A)
  const int row_tiles = (env->n_env + BM - 1) / BM;
  int need = row_tiles * A.H.n_split;
  const int need_b = (env->n_env + NT / 32 - 1) / (NT / 32);
  if (need_b > need) need = need_b;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  void* params[] = {&A};
  const bool prof = cirs_profile_begin("rollout_kuaishou_kernel", (cudaStream_t)stream);
  cudaError_t err = cudaLaunchCooperativeKernel((void*)rollout_kuaishou_kernel, dim3(grid), dim3(NT), params, smem,
                                                (cudaStream_t)stream);
  cirs_note_launch();
  if (prof) cirs_profile_end((cudaStream_t)stream);
  if (err != cudaSuccess) {
    cirs_set_error(cudaGetErrorString(err));
    return CIRS_ERR_CUDA;
  }
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
