// K3': continuous actor (tianshou ActorProb + Independent(Normal), VirtualTaobao) -- one warp per row.
// Replaces core/policy/ppo.py:144-156 (forward: actor -> dist -> sample) with tianshou/utils/net/continuous.py:179-199
// and, for stored transitions, PPOPolicy.process_fn's old log-prob and A2CPolicy._compute_returns' critic calls
// (core/policy/ppo.py:104-108, tianshou/policy/modelfree/a2c.py:89-90).
// The 27-wide head is a warp matvec (lane c owns action component c, k-major weights -> coalesced loads); trunk and
// critic use the same per-row code as the persistent Kuaishou rollout, bit-identical to the tile kernels.
#include "taobao_dev.cuh"

namespace {
using namespace cirs_taobao;
constexpr int WARPS_PER_CTA = 4;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
actorprob_sample_kernel(cirs_policy_weights W, int n_rows, const int32_t* __restrict__ env_id,
                        const uint8_t* __restrict__ active, const float* __restrict__ state, int64_t state_stride,
                        const float* __restrict__ noise, uint64_t seed, uint64_t offset,
                        const unsigned long long* __restrict__ rng_counter, int mode, float* __restrict__ act,
                        float* __restrict__ logp, float* __restrict__ value, float* __restrict__ mu_out) {
  __shared__ float smem[WARPS_PER_CTA * ACTOR_SCRATCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS_PER_CTA + warp;
  if (k >= n_rows) return;
  const int e = env_id ? env_id[k] : k;
  if (active && !active[e]) return;
  const int nA = W.n_action;
  const uint64_t off = offset + (rng_counter ? *rng_counter : 0ull);
  // compact rows (env_id given) read state row k, per-slot calls read row e -- as cirs_actor_sample
  const float* s = state + (size_t)(env_id ? k : e) * state_stride;
  const float a = actorprob_warp(W, s, lane, smem + warp * ACTOR_SCRATCH, noise ? noise + (size_t)k * nA : nullptr, seed,
                                 off, e, mode, nullptr, value ? value + k : nullptr, logp ? logp + k : nullptr,
                                 mu_out ? mu_out + (size_t)k * nA : nullptr);
  if (lane < nA) act[(size_t)k * nA + lane] = a;
}

__global__ void bump_counter_kernel(unsigned long long* c) { *c += 1ull; }

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
actorprob_eval_kernel(cirs_policy_weights W, int n_rows, const int32_t* __restrict__ row_idx,
                      const float* __restrict__ obs, const float* __restrict__ act, float* __restrict__ value,
                      float* __restrict__ logp) {
  __shared__ float smem[WARPS_PER_CTA * ACTOR_SCRATCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_CTA + warp;
  if (r >= n_rows) return;
  const int slot = row_idx ? row_idx[r] : r;
  actorprob_warp(W, obs + (size_t)slot * W.dim_state, lane, smem + warp * ACTOR_SCRATCH, nullptr, 0, 0, 0, 1,
                 act ? act + (size_t)slot * W.n_action : nullptr, value ? value + slot : nullptr,
                 (act && logp) ? logp + slot : nullptr, nullptr);
}

bool bad_weights(const cirs_policy_weights* w) {
  return !w || !w->sigma || !w->w3t || !w->b3 || w->n_action < 1 || w->n_action > 32 || w->dim_state > 32 ||
         w->ld_action < w->n_action;
}
}  // namespace

extern "C" int cirs_actorprob_sample(const cirs_policy_weights* w, int32_t n_rows, const int32_t* env_id,
                                     const uint8_t* active, const float* state, int64_t state_stride,
                                     const float* noise_eps, uint64_t seed, uint64_t offset, uint64_t* rng_counter,
                                     int32_t mode, float* act, float* logp, float* value, float* mu_out,
                                     void* stream) {
  if (bad_weights(w) || !state || !act || n_rows < 0) {
    cirs_set_error("cirs_actorprob_sample: bad argument (continuous actor needs sigma, n_action <= 32)");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(actorprob_sample_kernel, grid, WARPS_PER_CTA * 32, 0, st, *w, n_rows, env_id, active, state, state_stride,
              noise_eps, seed, offset, reinterpret_cast<const unsigned long long*>(rng_counter), mode, act, logp,
              value, mu_out);
  CIRS_CHECK_LAUNCH();
  if (rng_counter) {
    CIRS_LAUNCH(bump_counter_kernel, 1, 1, 0, st, reinterpret_cast<unsigned long long*>(rng_counter));
    CIRS_CHECK_LAUNCH();
  }
  return CIRS_OK;
}

extern "C" int cirs_actorprob_eval(const cirs_policy_weights* w, int32_t n_rows, const int32_t* row_idx,
                                   const float* obs, const float* act, float* value, float* logp, void* stream) {
  if (bad_weights(w) || !obs || n_rows < 0 || (!value && !logp)) {
    cirs_set_error("cirs_actorprob_eval: bad argument");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  const int grid = (n_rows + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CIRS_LAUNCH(actorprob_eval_kernel, grid, WARPS_PER_CTA * 32, 0, (cudaStream_t)stream, *w, n_rows, row_idx, obs, act,
              value, logp);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
