// K7: clip_grad_norm_ + Adam over a flat parameter buffer (core/policy/ppo.py:221-226; torch.optim.Adam,
// single-tensor CPU semantics: betas .9/.999, eps 1e-8, no weight decay, no amsgrad).
//
// Duplicate-parameter semantics: the reference builds optim_RL and the clip list from
// list(actor.parameters()) + list(critic.parameters()) where actor.preprocess IS critic.preprocess
// (CIRS-RL-kuaishou.py:245-258), so every trunk tensor occurs twice: it is counted twice in the total norm, its
// gradient is multiplied by the clip coefficient twice, and Adam.step() updates it twice in a row with the same
// (already clipped) gradient, advancing its step counter by two.  The first n_dup floats of the buffer are those
// tensors.  HBM traffic: 16 B read + 12 B written per parameter (+ 4 B read for the norm).
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, int64_t n_dup, double* __restrict__ scratch) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += (i < n_dup ? 2.0 : 1.0) * v * v;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(scratch, t);
  }
}

// bias corrections of one Adam step, computed once per launch by bump_kernel (FP64 pow / sqrt like torch's scalar math)
struct StepConst { float step_size, bc2_sqrt; };

__device__ __forceinline__ void adam_once(float& p, float g, float& m, float& v, const StepConst k,
                                          const cirs_ppo_config& c) {
  m = m + (g - m) * (1.0f - c.beta1);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = v * c.beta2 + (1.0f - c.beta2) * g * g;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / k.bc2_sqrt + c.adam_eps;
  p = p - k.step_size * (m / denom);                        // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            int64_t n_dup, cirs_ppo_config c, const double* __restrict__ scratch) {
  float coef = 1.0f;
  if (c.max_grad_norm > 0.f) {
    const float total = (float)sqrt(scratch[0]);             // torch.linalg.vector_norm of the per-tensor norms
    coef = fminf(c.max_grad_norm / (total + 1e-6f), 1.0f);   // clip_grad_norm_: clamp(max_norm / (total + 1e-6), max=1)
  }
  // scratch[2..7]: (step_size, sqrt(bias_correction2)) for the steps s1 (other tensors), s2 - 1 and s2 (trunk)
  const StepConst k1{(float)scratch[2], (float)scratch[3]}, k2a{(float)scratch[4], (float)scratch[5]},
      k2b{(float)scratch[6], (float)scratch[7]};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], gi = g[i], mi = m[i], vi = v[i];
    if (i < n_dup) {
      gi = gi * coef * coef;
      adam_once(pi, gi, mi, vi, k2a, c);
      adam_once(pi, gi, mi, vi, k2b, c);
    } else {
      gi = gi * coef;
      adam_once(pi, gi, mi, vi, k1, c);
    }
    p[i] = pi; g[i] = gi; m[i] = mi; v[i] = vi;
  }
}

__global__ void bump_kernel(int32_t* state, double* scratch, cirs_ppo_config c) {
  state[0] += 1;
  state[1] += 2;
  scratch[0] = 0.0;
  const int steps[3] = {state[0], state[1] - 1, state[1]};
  for (int i = 0; i < 3; ++i) {
    const double bc1 = 1.0 - pow((double)c.beta1, (double)steps[i]);
    const double bc2 = 1.0 - pow((double)c.beta2, (double)steps[i]);
    scratch[2 + 2 * i] = (double)(float)((double)c.lr / bc1);
    scratch[3 + 2 * i] = (double)(float)sqrt(bc2);
  }
}

}  // namespace

extern "C" int cirs_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              int64_t n_dup, const cirs_ppo_config* cfg, int32_t* state, double* scratch,
                              void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !cfg || !state || !scratch || n < 0 || n_dup < 0 ||
      n_dup > n) {
    cirs_set_error("cirs_clip_adam: bad argument");
    return CIRS_ERR_ARG;
  }
  if (n == 0) return CIRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CIRS_LAUNCH(bump_kernel, 1, 1, 0, st, state, scratch, *cfg);
  CIRS_CHECK_LAUNCH();
  if (cfg->max_grad_norm > 0.f) {
    CIRS_LAUNCH(sumsq_kernel, blocks, 256, 0, st, grads, n, n_dup, scratch);
    CIRS_CHECK_LAUNCH();
  }
  CIRS_LAUNCH(adam_kernel, blocks, 256, 0, st, params, grads, exp_avg, exp_avg_sq, n, n_dup, *cfg, scratch);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
