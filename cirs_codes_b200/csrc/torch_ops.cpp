// torch.ops.cirs_b200.* -- the C ABI of include/cirs_b200.h registered as PyTorch operators (TORCH_LIBRARY), the form
// SURVEY 8b proposes for the six per-step operators of the path:
//     env_step_kuaishou   SimulatedEnv.step + KuaishouEnv.step for a whole vector env     (simulated_env.py:111-168)
//     tracker_step        StateTrackerTransformer.build_state                              (core/state_tracker.py:188-250)
//     actor_sample        PPOPolicy.forward                                                (core/policy/ppo.py:111-163)
//     gae                 A2CPolicy._compute_returns / compute_episodic_return            (a2c.py:80-109, base.py:272-313)
//     ppo_minibatch       one minibatch of PPOPolicy.learn                                 (ppo.py:181-220)
//     adam_clip           clip_grad_norm_ + optim.step                                     (ppo.py:221-226)
// Every operator takes / returns at::Tensor on the CURRENT CUDA stream of the tensors' device, never synchronises, and
// reports errors through TORCH_CHECK (-> Python RuntimeError carrying cirs_last_error()).  The descriptor structs of the
// C ABI (cirs_kuaishou_env, cirs_tracker_weights, cirs_policy_weights, cirs_ppo_config: tables of device pointers and
// shapes that the host classes own) travel as opaque int64 handles = the address of the caller's struct.
// This file contains no kernel: it is a thin layer over libcirs_b200.so, which stays the drop-in boundary.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/cirs_b200.h"

namespace {

using at::Tensor;
using OptTensor = c10::optional<Tensor>;

void* stream_of(const Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

void check(int rc, const char* what) { TORCH_CHECK(rc == 0, "cirs_b200::", what, ": ", cirs_last_error()); }

template <class T>
T* ptr(const Tensor& t, at::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == dt, name, " has dtype ", t.scalar_type(), ", expected ", dt);
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  return reinterpret_cast<T*>(t.data_ptr());
}
template <class T>
T* optr(const OptTensor& t, at::ScalarType dt, const char* name) {
  return (t.has_value() && t->defined()) ? ptr<T>(*t, dt, name) : nullptr;
}
template <class S>
const S* handle(int64_t h, const char* name) {
  TORCH_CHECK(h != 0, name, ": null descriptor handle");
  return reinterpret_cast<const S*>(static_cast<intptr_t>(h));
}

// act i32[n] (+ env_id i32[n], active u8[B] updated in place) -> (rew f32[n], done u8[n])
std::tuple<Tensor, Tensor> env_step_kuaishou(int64_t env, const Tensor& act, const OptTensor& env_id,
                                             const OptTensor& active, int64_t force_length) {
  const int32_t* a = ptr<int32_t>(act, at::kInt, "act");
  const int64_t n = act.numel();
  Tensor rew = at::empty({n}, act.options().dtype(at::kFloat));
  Tensor done = at::empty({n}, act.options().dtype(at::kByte));
  check(cirs_kuaishou_step(handle<cirs_kuaishou_env>(env, "env"), (int32_t)n, optr<int32_t>(env_id, at::kInt, "env_id"),
                           optr<uint8_t>(active, at::kByte, "active"), a, rew.data_ptr<float>(),
                           done.data_ptr<uint8_t>(), 0, nullptr, nullptr, nullptr, nullptr, (int32_t)force_length,
                           stream_of(act)),
        "env_step_kuaishou");
  return {rew, done};
}

// one token per row appended to the K/V caches -> state f32[n_rows, dim_state]
Tensor tracker_step(int64_t w, int64_t n_env, int64_t dim_state, const OptTensor& env_id, const OptTensor& active,
                    const OptTensor& pos, int64_t expect_pos, const OptTensor& idx, const OptTensor& dense,
                    const OptTensor& rew, const Tensor& kcache, const Tensor& vcache) {
  float* kc = ptr<float>(kcache, at::kFloat, "kcache");
  float* vc = ptr<float>(vcache, at::kFloat, "vcache");
  const bool has_idx = idx.has_value() && idx->defined();
  TORCH_CHECK(has_idx || (dense.has_value() && dense->defined()), "tracker_step: idx or dense is required");
  const int64_t n_rows = has_idx ? idx->size(0) : dense->size(0);
  Tensor out = at::empty({n_rows, dim_state}, kcache.options());
  check(cirs_tracker_step(handle<cirs_tracker_weights>(w, "tracker weights"), (int32_t)n_env, (int32_t)n_rows,
                          optr<int32_t>(env_id, at::kInt, "env_id"), optr<uint8_t>(active, at::kByte, "active"),
                          optr<int32_t>(pos, at::kInt, "pos"), (int32_t)expect_pos, optr<int32_t>(idx, at::kInt, "idx"),
                          optr<float>(dense, at::kFloat, "dense"), optr<float>(rew, at::kFloat, "rew"), kc, vc,
                          out.data_ptr<float>(), dim_state, nullptr, 0, nullptr, nullptr, stream_of(kcache)),
        "tracker_step");
  return out;
}

// state f32[n, >= dim_state] -> (act i32[n], logp f32[n], value f32[n]); mode 0 sample / 1 argmax
std::tuple<Tensor, Tensor, Tensor> actor_sample(int64_t w, const Tensor& state, const OptTensor& env_id,
                                                const OptTensor& active, const OptTensor& noise_q, int64_t seed,
                                                int64_t offset, int64_t mode, const Tensor& workspace) {
  const float* s = ptr<float>(state, at::kFloat, "state");
  TORCH_CHECK(state.dim() == 2, "state must be [n_rows, stride]");
  const int64_t n = state.size(0);
  Tensor act = at::empty({n}, state.options().dtype(at::kInt));
  Tensor logp = at::empty({n}, state.options());
  Tensor value = at::empty({n}, state.options());
  check(cirs_actor_sample(handle<cirs_policy_weights>(w, "policy weights"), (int32_t)n,
                          optr<int32_t>(env_id, at::kInt, "env_id"), optr<uint8_t>(active, at::kByte, "active"), s,
                          state.size(1), optr<float>(noise_q, at::kFloat, "noise_q"), (uint64_t)seed, (uint64_t)offset,
                          nullptr, (int32_t)mode, nullptr, act.data_ptr<int32_t>(), logp.data_ptr<float>(),
                          value.data_ptr<float>(), ptr<uint8_t>(workspace, at::kByte, "workspace"), stream_of(state)),
        "actor_sample");
  return {act, logp, value};
}

// env-major slots [B, L]: -> (returns f32[B * L], adv f32[B * L]); ret_rms f64[3] optional; moments f64[3] optional (out)
std::tuple<Tensor, Tensor> gae(const Tensor& n_slot, const Tensor& v_s, const Tensor& v_next, const Tensor& rew,
                               const Tensor& done, double gamma, double gae_lambda, const OptTensor& ret_rms,
                               const OptTensor& moments) {
  const int32_t* ns = ptr<int32_t>(n_slot, at::kInt, "n_slot");
  const int64_t B = n_slot.numel();
  TORCH_CHECK(B > 0 && v_s.numel() % B == 0, "v_s must hold B * L slots");
  const int64_t L = v_s.numel() / B;
  Tensor returns = at::zeros_like(v_s), adv = at::zeros_like(v_s);
  Tensor scratch = at::empty({2 * B}, v_s.options().dtype(at::kDouble));
  check(cirs_compute_returns((int32_t)B, (int32_t)L, ns, ptr<float>(v_s, at::kFloat, "v_s"),
                             ptr<float>(v_next, at::kFloat, "v_next"), ptr<float>(rew, at::kFloat, "rew"),
                             ptr<uint8_t>(done, at::kByte, "done"), gamma, gae_lambda,
                             optr<double>(ret_rms, at::kDouble, "ret_rms"), scratch.data_ptr<double>(),
                             optr<double>(moments, at::kDouble, "moments"), returns.data_ptr<float>(),
                             adv.data_ptr<float>(), stream_of(v_s)),
        "gae");
  return {returns, adv};
}

// forward + loss + backward of one minibatch: gradients accumulate into the struct behind `grads`, -> losses f32[4]
Tensor ppo_minibatch(int64_t w, int64_t grads, int64_t cfg, int64_t n_global, const Tensor& idx, const Tensor& obs,
                     const Tensor& act, const Tensor& adv, const Tensor& returns, const Tensor& v_old,
                     const Tensor& logp_old, const OptTensor& adv_stat, const OptTensor& d_obs,
                     const Tensor& workspace) {
  const int64_t n = idx.numel();
  Tensor losses = at::zeros({4}, obs.options());
  TORCH_CHECK(act.is_cuda() && act.is_contiguous(), "act must be a contiguous CUDA tensor");
  check(cirs_ppo_minibatch(handle<cirs_policy_weights>(w, "policy weights"),
                           handle<cirs_policy_weights>(grads, "policy gradients"), handle<cirs_ppo_config>(cfg, "config"),
                           (int32_t)n, (int32_t)n_global, ptr<int32_t>(idx, at::kInt, "idx"),
                           ptr<float>(obs, at::kFloat, "obs"), act.data_ptr(), ptr<float>(adv, at::kFloat, "adv"),
                           ptr<float>(returns, at::kFloat, "returns"), ptr<float>(v_old, at::kFloat, "v_old"),
                           ptr<float>(logp_old, at::kFloat, "logp_old"), optr<double>(adv_stat, at::kDouble, "adv_stat"),
                           optr<float>(d_obs, at::kFloat, "d_obs"), losses.data_ptr<float>(),
                           ptr<uint8_t>(workspace, at::kByte, "workspace"), stream_of(obs)),
        "ppo_minibatch");
  return losses;
}

// in place on params / grads / exp_avg / exp_avg_sq (flat f32 buffers), state i32[2] step counters, scratch f64[16]
void adam_clip(Tensor params, Tensor grads, Tensor exp_avg, Tensor exp_avg_sq, int64_t n_dup, int64_t cfg, Tensor state,
               Tensor scratch) {
  const int64_t n = params.numel();
  TORCH_CHECK(grads.numel() == n && exp_avg.numel() == n && exp_avg_sq.numel() == n, "flat buffers differ in size");
  check(cirs_clip_adam(ptr<float>(params, at::kFloat, "params"), ptr<float>(grads, at::kFloat, "grads"),
                       ptr<float>(exp_avg, at::kFloat, "exp_avg"), ptr<float>(exp_avg_sq, at::kFloat, "exp_avg_sq"), n,
                       n_dup, handle<cirs_ppo_config>(cfg, "config"), ptr<int32_t>(state, at::kInt, "state"),
                       ptr<double>(scratch, at::kDouble, "scratch"), stream_of(params)),
        "adam_clip");
}

}  // namespace

TORCH_LIBRARY(cirs_b200, m) {
  m.def("env_step_kuaishou(int env, Tensor act, Tensor? env_id, Tensor(a!)? active, int force_length=0) -> (Tensor, Tensor)");
  m.def("tracker_step(int weights, int n_env, int dim_state, Tensor? env_id, Tensor? active, Tensor? pos, int expect_pos, "
        "Tensor? idx, Tensor? dense, Tensor? rew, Tensor(a!) kcache, Tensor(b!) vcache) -> Tensor");
  m.def("actor_sample(int weights, Tensor state, Tensor? env_id, Tensor? active, Tensor? noise_q, int seed, int offset, "
        "int mode, Tensor(a!) workspace) -> (Tensor, Tensor, Tensor)");
  m.def("gae(Tensor n_slot, Tensor v_s, Tensor v_next, Tensor rew, Tensor done, float gamma, float gae_lambda, "
        "Tensor? ret_rms, Tensor(a!)? moments) -> (Tensor, Tensor)");
  m.def("ppo_minibatch(int weights, int grads, int cfg, int n_global, Tensor idx, Tensor obs, Tensor act, Tensor adv, "
        "Tensor returns, Tensor v_old, Tensor logp_old, Tensor? adv_stat, Tensor(a!)? d_obs, Tensor(b!) workspace) -> Tensor");
  m.def("adam_clip(Tensor(a!) params, Tensor(b!) grads, Tensor(c!) exp_avg, Tensor(d!) exp_avg_sq, int n_dup, int cfg, "
        "Tensor(e!) state, Tensor(f!) scratch) -> ()");
}

TORCH_LIBRARY_IMPL(cirs_b200, CUDA, m) {
  m.impl("env_step_kuaishou", env_step_kuaishou);
  m.impl("tracker_step", tracker_step);
  m.impl("actor_sample", actor_sample);
  m.impl("gae", gae);
  m.impl("ppo_minibatch", ppo_minibatch);
  m.impl("adam_clip", adam_clip);
}
