// K4: returns / GAE -- A2CPolicy._compute_returns (tianshou/policy/modelfree/a2c.py:80-109),
// BasePolicy.compute_episodic_return + _gae_return (tianshou/policy/base.py:272-313, 380-396) and
// RunningMeanStd.update (tianshou/utils/statistics.py:80-95).  Float64 like the reference (numba f64 scan).
//
// The reference scans ONE flat env-major array backwards; `end_flag` is set at every episode end and at the last
// stored slot of every sub-buffer, so the recursion never crosses an environment boundary: one thread per
// environment runs its own reverse scan over its n_slot[e] transitions (exactly the same arithmetic, in the same
// order within the environment).  HBM traffic: 13 B read + 8 B written per transition -- negligible.
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace {

// scratch layout: double[2 * n_env] per-environment {sum, sumsq} of unnormalised returns, then int64 count
__global__ void gae_kernel(int n_env, int L, const int32_t* __restrict__ n_slot, const float* __restrict__ v_s,
                           const float* __restrict__ v_next, const float* __restrict__ rew,
                           const uint8_t* __restrict__ done, double gamma, double lam,
                           const double* __restrict__ ret_rms, float* __restrict__ returns,
                           float* __restrict__ adv, double* __restrict__ part) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_env) return;
  const int n = n_slot[e];
  const double scale = ret_rms ? sqrt(ret_rms[1] + 1e-8) : 1.0;  // a2c.py:97-99, _eps = 1e-8
  double gae = 0.0, s = 0.0, ss = 0.0;
  for (int t = n - 1; t >= 0; --t) {
    const size_t i = (size_t)e * L + t;
    const bool dn = done[i] != 0;
    const double vs = (double)v_s[i] * scale;
    const double vn = dn ? 0.0 : (double)v_next[i] * scale;           // value_mask, base.py:246-269
    const double delta = (double)rew[i] + vn * gamma - vs;            // base.py:389
    const bool end = dn || (t == n - 1);                              // done | unfinished_index, base.py:308-309
    gae = delta + (end ? 0.0 : gamma * lam) * gae;                    // base.py:390-394
    const double ret = gae + vs;                                      // base.py:312
    returns[i] = (float)(ret / scale);                                // a2c.py:103-104
    adv[i] = (float)gae;
    s += ret;
    ss += ret * ret;
  }
  if (part) {
    part[2 * e] = s;
    part[2 * e + 1] = ss;
  }
}

// the same kernel followed, in the LAST block to finish, by the moments reduction (one launch instead of two)
__device__ unsigned int g_gae_ticket = 0;
__global__ void __launch_bounds__(128)
gae_moments_kernel(int n_env, int L, const int32_t* __restrict__ n_slot, const float* __restrict__ v_s,
                   const float* __restrict__ v_next, const float* __restrict__ rew, const uint8_t* __restrict__ done,
                   double gamma, double lam, const double* __restrict__ ret_rms, float* __restrict__ returns,
                   float* __restrict__ adv, double* __restrict__ part, double* __restrict__ moments) {
  __shared__ bool last;
  __shared__ double sh[3][4];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_env) {
    const int n = n_slot[e];
    const double scale = ret_rms ? sqrt(ret_rms[1] + 1e-8) : 1.0;
    double gae = 0.0, s = 0.0, ss = 0.0;
    for (int t = n - 1; t >= 0; --t) {
      const size_t i = (size_t)e * L + t;
      const bool dn = done[i] != 0;
      const double vs = (double)v_s[i] * scale;
      const double vn = dn ? 0.0 : (double)v_next[i] * scale;
      const double delta = (double)rew[i] + vn * gamma - vs;
      const bool end = dn || (t == n - 1);
      gae = delta + (end ? 0.0 : gamma * lam) * gae;
      const double ret = gae + vs;
      returns[i] = (float)(ret / scale);
      adv[i] = (float)gae;
      s += ret;
      ss += ret * ret;
    }
    part[2 * e] = s;
    part[2 * e + 1] = ss;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&g_gae_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the per-environment partials in environment order per thread, then a fixed-order tree: deterministic
  double s = 0.0, ss = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < n_env; i += blockDim.x) {
    s += *(volatile double*)(part + 2 * i);
    ss += *(volatile double*)(part + 2 * i + 1);
    cnt += (double)n_slot[i];
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = ss = cnt = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += sh[0][i]; ss += sh[1][i]; cnt += sh[2][i]; }
    moments[0] = s; moments[1] = ss; moments[2] = cnt;
    g_gae_ticket = 0;
  }
}

// raw moments {sum, sumsq, count} of the unnormalised returns of this rank (ranks all-reduce them before merging)
__global__ void __launch_bounds__(1024) moments_kernel(int n_env, const int32_t* __restrict__ n_slot,
                                                        const double* __restrict__ part, double* moments) {
  __shared__ double sh[3][32];
  double s = 0.0, ss = 0.0, cnt = 0.0;
  for (int e = threadIdx.x; e < n_env; e += blockDim.x) {
    s += part[2 * e];
    ss += part[2 * e + 1];
    cnt += (double)n_slot[e];
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = cnt; }
  __syncthreads();
  if (w == 0) {
    s = l < (blockDim.x >> 5) ? sh[0][l] : 0.0;
    ss = l < (blockDim.x >> 5) ? sh[1][l] : 0.0;
    cnt = l < (blockDim.x >> 5) ? sh[2][l] : 0.0;
    s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt);
    if (l == 0) { moments[0] = s; moments[1] = ss; moments[2] = cnt; }
  }
}

// RunningMeanStd.update (statistics.py:80-95) from the raw batch moments {sum, sumsq, count}
__global__ void rms_merge_kernel(const double* __restrict__ moments, double* ret_rms) {
  const double s = moments[0], ss = moments[1], cnt = moments[2];
  if (cnt <= 0.0) return;
  const double bm = s / cnt, bv = fmax(ss / cnt - bm * bm, 0.0);
  const double mean = ret_rms[0], var = ret_rms[1], c0 = ret_rms[2];
  const double delta = bm - mean, tot = c0 + cnt;
  const double new_mean = mean + delta * cnt / tot;
  const double m2 = var * c0 + bv * cnt + delta * delta * c0 * cnt / tot;
  ret_rms[0] = new_mean;
  ret_rms[1] = m2 / tot;
  ret_rms[2] = tot;
}

}  // namespace

extern "C" int cirs_compute_returns(int32_t n_env, int32_t traj_len, const int32_t* n_slot, const float* v_s,
                                    const float* v_next, const float* rew, const uint8_t* done, double gamma,
                                    double gae_lambda, const double* ret_rms, double* scratch, double* moments,
                                    float* returns, float* adv, void* stream) {
  if (n_env < 0 || !n_slot || !v_s || !v_next || !rew || !done || !returns || !adv || (moments && !scratch)) {
    cirs_set_error("cirs_compute_returns: null argument");
    return CIRS_ERR_ARG;
  }
  if (n_env == 0) return CIRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (moments) {
    CIRS_LAUNCH(gae_moments_kernel, (n_env + 127) / 128, 128, 0, st, n_env, traj_len, n_slot, v_s, v_next, rew, done,
                gamma, gae_lambda, ret_rms, returns, adv, scratch, moments);
    CIRS_CHECK_LAUNCH();
    return CIRS_OK;
  }
  CIRS_LAUNCH(gae_kernel, (n_env + 127) / 128, 128, 0, st, n_env, traj_len, n_slot, v_s, v_next, rew, done, gamma,
                                                  gae_lambda, ret_rms, returns, adv, nullptr);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_rms_update(double* ret_rms, const double* moments, void* stream) {
  if (!ret_rms || !moments) {
    cirs_set_error("cirs_rms_update: null argument");
    return CIRS_ERR_ARG;
  }
  CIRS_LAUNCH(rms_merge_kernel, 1, 1, 0, (cudaStream_t)stream, moments, ret_rms);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
