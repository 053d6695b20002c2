// K3: policy.forward for the discrete actor -- trunk MLP + Linear(64 -> n_action) + softmax + Categorical.sample
// fused, the [rows, n_action] probabilities are never written.
// Replaces core/policy/ppo.py:111-163, tianshou/utils/net/discrete.py:56-67,109-114, utils/net/common.py:87-92,
// 178-197 and torch.distributions.Categorical.sample() == argmax_j p_j / q_j, q ~ Exp(1).
//
// Grid: (row tiles of 64) x (splits of the catalogue).  Each CTA computes the trunk for its 64 rows (h2 kept
// k-major in shared memory), then walks its share of the 128-column tiles of W3t: the tile is staged through
// shared memory with coalesced float4 loads and every thread accumulates an 8 x 4 block of logits in
// registers (FP32 FFMA).  The epilogue never stores logits: per row it keeps the online-softmax pair (max, sum)
// and the running winner of the exponential race  argmax_j  logit_j - log q_j.  Partials per (split, row) go to
// the workspace and a second tiny kernel merges the splits and emits act / log_prob / value.
#include "actor_dev.cuh"
#include "head_tc.cuh"

namespace cirs_head_tc {
int policy_eval_tc(const cirs_policy_weights* w, int32_t n, const int32_t* row_idx, const float* obs,
                   const int32_t* act, float* value, float* logp, void* workspace, cudaStream_t st,
                   const int32_t* n_dev = nullptr);
}

namespace {
using namespace cirs_actor;

__global__ void __launch_bounds__(NT) actor_head_kernel(HeadArgs P) {
  extern __shared__ __align__(16) float smem_dyn[];
  actor_head_body(P, blockIdx.x, blockIdx.y, smem_dyn);
}

__global__ void actor_combine_kernel(HeadArgs P, int32_t* __restrict__ act, float* __restrict__ logp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0 && P.rng_counter) *P.rng_counter += 1ull;  // the head kernel of this call has already read it
  if (k >= P.n_rows) return;
  actor_combine_row(P, k, act, logp);
}

int run_head(HeadArgs& P, int32_t* act, float* logp, void* workspace, cudaStream_t st) {
  if (!plan_head(P)) {
    cirs_set_error("actor head: dim_state > 32 or ld_action not a multiple of 128");
    return CIRS_ERR_ARG;
  }
  P.part = reinterpret_cast<Partial*>(workspace);
  dim3 grid((P.n_rows + BM - 1) / BM, P.n_split);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(actor_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    attr_set = true;
  }
  CIRS_LAUNCH(actor_head_kernel, grid, NT, SMEM_BYTES, st, P);
  CIRS_CHECK_LAUNCH();
  if (act || logp) {
    CIRS_LAUNCH(actor_combine_kernel, (P.n_rows + 127) / 128, 128, 0, st, P, act, logp);
    CIRS_CHECK_LAUNCH();
  }
  return CIRS_OK;
}

}  // namespace

extern "C" int64_t cirs_actor_workspace_bytes(int32_t n_rows, int32_t n_action) {
  const int n_tiles = (n_action + BN - 1) / BN;
  int64_t splits = pick_split(n_rows, n_action);
  if (splits > n_tiles) splits = n_tiles;
  const int64_t ffma = (int64_t)sizeof(Partial) * (splits + 1) * (int64_t)(n_rows > 0 ? n_rows : 1) + 256;
  const int64_t ldA = ((int64_t)n_action + 127) & ~127LL;
  const int64_t tc = cirs_head_tc::policy_eval_tc_workspace_bytes(n_rows > 0 ? n_rows : 1) +
                     cirs_head_tc::policy_eval_tc_image_bytes(ldA);
  return ffma > tc ? ffma : tc;
}

extern "C" int cirs_actor_sample(const cirs_policy_weights* w, int32_t n_rows, const int32_t* env_id,
                                 const uint8_t* active, const float* state, int64_t state_stride,
                                 const float* noise_q, uint64_t seed, uint64_t offset, uint64_t* rng_counter,
                                 int32_t mode, const uint32_t* seen, int32_t* act, float* logp, float* value,
                                 void* workspace, void* stream) {
  if (!w || !state || !act || !workspace || n_rows < 0 || (mode != MODE_SAMPLE && mode != MODE_ARGMAX)) {
    cirs_set_error("cirs_actor_sample: bad argument");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  HeadArgs P{};
  P.W = *w; P.n_rows = n_rows; P.gather = env_id; P.state_by_k = env_id != nullptr; P.out_by_k = 1;
  P.active = active; P.state = state; P.state_stride = state_stride; P.noise_q = noise_q; P.seed = seed;
  P.offset = offset; P.rng_counter = reinterpret_cast<unsigned long long*>(rng_counter); P.mode = mode; P.seen = seen;
  P.act_in = nullptr; P.value = value;
  return run_head(P, act, logp, workspace, (cudaStream_t)stream);
}

extern "C" int cirs_policy_eval(const cirs_policy_weights* w, int32_t n_rows, const int32_t* row_idx,
                                const float* obs, const int32_t* act, float* value, float* logp, void* workspace,
                                void* stream) {
  if (!w || !obs || !workspace || n_rows < 0 || (act && !logp)) {
    cirs_set_error("cirs_policy_eval: bad argument");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  // tensor-core path; a value-only evaluation (act == NULL: V(obs_next)) is just the trunk kernel + scatter
  if (w->sigma == nullptr && w->dim_state <= 32 && cirs_head_tc::head_tc_enabled(n_rows, w->n_action, w->ld_action))
    return cirs_head_tc::policy_eval_tc(w, n_rows, row_idx, obs, act, value, logp, workspace, (cudaStream_t)stream);
  HeadArgs P{};
  P.W = *w; P.n_rows = n_rows; P.gather = row_idx; P.state_by_k = 0; P.out_by_k = 0; P.active = nullptr;
  P.state = obs; P.state_stride = w->dim_state; P.noise_q = nullptr; P.mode = MODE_EVAL; P.seen = nullptr;
  P.act_in = act; P.value = value;
  return run_head(P, nullptr, act ? logp : nullptr, workspace, (cudaStream_t)stream);
}

// cirs_policy_eval with the row count still ON THE DEVICE: n_cap sizes grids / layouts / the workspace, rows >=
// min(n_cap, *n_dev) are skipped.  Lets the host queue process_fn's evaluations behind the rollout kernel BEFORE it has
// read the collect's transition count back (PPOPolicy.post_collect), so they run while the host is still waking up.
// Tensor-core path only (discrete actor, 128-column pass F).
extern "C" int cirs_policy_eval_dev(const cirs_policy_weights* w, int32_t n_cap, const int32_t* n_dev,
                                    const int32_t* row_idx, const float* obs, const int32_t* act, float* value,
                                    float* logp, void* workspace, void* stream) {
  if (!w || !obs || !workspace || !n_dev || n_cap <= 0 || (act && !logp)) {
    cirs_set_error("cirs_policy_eval_dev: bad argument");
    return CIRS_ERR_ARG;
  }
  if (!(w->sigma == nullptr && w->dim_state <= 32 && cirs_head_tc::head_tc_enabled(n_cap, w->n_action, w->ld_action))) {
    cirs_set_error("cirs_policy_eval_dev: needs the tensor-core head (discrete actor, n_action >= 64, CIRS_NO_TC unset)");
    return CIRS_ERR_ARG;
  }
  return cirs_head_tc::policy_eval_tc(w, n_cap, row_idx, obs, act, value, logp, workspace, (cudaStream_t)stream, n_dev);
}
