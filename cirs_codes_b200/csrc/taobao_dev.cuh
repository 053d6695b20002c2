// Device code of K1' (VirtualTaobao / SimulatedEnv step incl. the MMOE reward model) and of the continuous actor
// (ActorProb + Independent(Normal)), shared by the stand-alone kernels (env_taobao.cu, actor_gauss.cu) and the
// one-kernel rollout (rollout_taobao.cu).  One warp per environment / row.
#pragma once
#include "actor_dev.cuh"
#include "tracker_dev.cuh"

namespace cirs_taobao {

constexpr int NU = CIRS_TB_USER, NI = CIRS_TB_ITEM;
// shared-memory floats one warp needs for taobao_step_warp: action[32] + x[128] + h1[128] + h2[128] + experts/gates[128]
constexpr int STEP_SCRATCH = 32 + 128 + 128 + 128 + 128;
// ... and for actorprob_warp: actor_trunk_warp's 160 + h2 copy 64
constexpr int ACTOR_SCRATCH = 160 + 64;

__device__ __forceinline__ void taobao_reset_warp(const cirs_taobao_env& E, int e, const float* __restrict__ user,
                                                  int lane, uint8_t* __restrict__ active) {
  for (int i = lane; i < NU; i += 32) E.user[(size_t)e * NU + i] = user[i];
  if (lane == 0) {
    E.turn[e] = 0;
    E.prev_rew[e] = 0.0;
    E.cum_rew[e] = 0.0;
    if (active) active[e] = 1;
  }
}

// UserModel_MMOE.forward of one feature vector x[n_in] held in shared memory (every lane returns y).
// h1 / h2 / ex: shared scratch of 128 floats each.
__device__ __forceinline__ float mmoe_forward_warp(const cirs_mmoe_weights& M, const float* x, float* h1, float* h2,
                                                   float* ex, int lane) {
  using cirs_tracker::matvec;
  const int ld1 = (M.h1 + 31) & ~31, ld2 = (M.h2 + 31) & ~31;
  const int nE = M.n_expert * M.expert_dim, lde = (nE + 31) & ~31, ldg = (M.n_expert + 31) & ~31;
  float lin = 0.f;  // linear part: x @ w[n_in, 1]  (core/layers.py:67-70)
  for (int i = lane; i < M.n_in; i += 32) lin = fmaf(x[i], __ldg(M.lin_w + i), lin);
  lin = warp_sum(lin);
  matvec<false>(M.w1t, M.b1, x, M.n_in, M.h1, ld1, h1, lane, 1);   // deepctr DNN: relu(W x + b), layers/core.py:120-134
  matvec<false>(M.w2t, M.b2, h1, M.h1, M.h2, ld2, h2, lane, 1);
  matvec<false>(M.wet, M.be, h2, M.h2, nE, lde, ex, lane, 0);      // experts, viewed [expert_dim][n_expert]
  matvec<false>(M.wgt, M.bg, h2, M.h2, M.n_expert, ldg, ex + 64, lane, 0);  // gate logits
  float mx = -INFINITY;
  for (int k = 0; k < M.n_expert; ++k) mx = fmaxf(mx, ex[64 + k]);
  float z = 0.f;
  for (int k = 0; k < M.n_expert; ++k) z += expf(ex[64 + k] - mx);
  float dnn = 0.f;
  for (int dd = 0; dd < M.expert_dim; ++dd) {
    float m = 0.f;  // bmm(expert_out[dim, :], gate): core/layers.py:113-114
    for (int k = 0; k < M.n_expert; ++k) m = fmaf(ex[dd * M.n_expert + k], expf(ex[64 + k] - mx) / z, m);
    dnn = fmaf(m, __ldg(M.tower + dd), dnn);
  }
  __syncwarp();
  return (lin + dnn) + M.out_bias;  // user_model_mmoe.py:202-207; PredictionLayer 'regression' adds the bias
}

// One SimulatedEnv(VirtualTB).step of environment slot e by one warp.  act: 27 floats (global or shared memory).
// ``sc``: STEP_SCRATCH floats of this warp's shared memory; on return sc[0..27) holds the action the environment
// used (the tracker's next token input).  Returns done.
__device__ __forceinline__ bool taobao_step_warp(const cirs_taobao_env& E, int e, int k, const float* act, int lane,
                                                 float* sc, uint8_t* __restrict__ active, float* __restrict__ act_env,
                                                 float* __restrict__ rew, uint8_t* __restrict__ done, int traj_len,
                                                 float* __restrict__ traj_act, float* __restrict__ traj_act_env,
                                                 float* __restrict__ traj_rew,
                                                 uint8_t* __restrict__ traj_done, int32_t* __restrict__ ep_len,
                                                 int force_length) {
  const int T = E.max_turn, t = E.turn[e];
  float* sa = sc;
  float* sx = sc + 32;
  float* h1 = sx + 128;
  float* h2 = h1 + 128;
  float* ex = h2 + 128;
  float* hist = E.hist + (size_t)e * T * NI;
  float a = 0.f, a_raw = 0.f;
  if (lane < NI) {
    a_raw = act[lane];
    a = a_raw;
    if (E.map_action) {  // tianshou/policy/base.py:164-172, float32 arithmetic like numpy
      a = fminf(fmaxf(a, -1.0f), 1.0f);
      a = __fadd_rn(E.act_low, __fdiv_rn(__fmul_rn(__fsub_rn(E.act_high, E.act_low), __fadd_rn(a, 1.0f)), 2.0f));
    }
  }
  sa[lane] = a;
  __syncwarp();
  // exit test: Euclidean distance to the last min(t, N-1) actions (virtualTB.py:126-133), float32 like numpy
  bool leave = false;
  for (int l = t - 1; l > max(-1, t - E.num_leave_compute); --l) {
    const float df = lane < NI ? __fsub_rn(a, hist[(size_t)l * NI + lane]) : 0.f;
    const float dist = sqrtf(warp_sum(__fmul_rn(df, df)));
    if ((double)dist <= E.leave_threshold) leave = true;
  }
  // exposure effect: gamma * sum_j exp(-(t-j) * ||a - a_j|| / tau), float64 (simulated_env.py:147-168, util.py:24-46);
  // lane j owns history slot j
  double expo = 0.0;
  if (t > 0 && E.tau > 0.0) {
    for (int j0 = 0; j0 < t; j0 += 32) {
      const int j = j0 + lane;
      if (j < t) {
        const float* hj = hist + (size_t)j * NI;
        double s = 0.0;
#pragma unroll 9
        for (int c = 0; c < NI; ++c) {
          const double df = (double)sa[c] - (double)hj[c];
          s = fma(df, df, s);
        }
        expo += exp(-(double)(t - j) * sqrt(s) / E.tau);
      }
    }
    expo = warp_sum_d(expo) * E.gamma_exposure;
  }
  __syncwarp();
  if (t < T && lane < NI) hist[(size_t)t * NI + lane] = a;   // simulated_env.py:123-124
  // reward model input [user 88, prev reward, 0, turn, action 27]  (simulated_env.py:79-80)
  for (int i = lane; i < NU; i += 32) sx[i] = E.user[(size_t)e * NU + i];
  if (lane == 0) {
    sx[NU] = (float)E.prev_rew[e];
    sx[NU + 1] = 0.f;
    sx[NU + 2] = (float)t;
  }
  if (lane < NI) sx[NU + 3 + lane] = a;
  __syncwarp();
  float y = mmoe_forward_warp(E.um, sx, h1, h2, ex, lane);
  y = fminf(fmaxf(y, 0.f), 10.f);                               // simulated_env.py:83-86
  const double r64 = (E.version == 1) ? (double)y / (1.0 + expo) : ((double)y - expo);   // :102-107, clip0 = identity
  bool d = leave || (t >= T - 1);                                // virtualTB.py:78-80
  if (force_length > 0) d = (t + 1 >= force_length);             // collector.py:253-258
  if (act_env && lane < NI) act_env[(size_t)k * NI + lane] = a;
  if (traj_act && t < traj_len && lane < NI) traj_act[((size_t)e * traj_len + t) * NI + lane] = a_raw;
  if (traj_act_env && t < traj_len && lane < NI) traj_act_env[((size_t)e * traj_len + t) * NI + lane] = a;
  if (lane == 0) {
    E.prev_rew[e] = r64;
    E.cum_rew[e] += r64;
    E.turn[e] = t + 1;
    if (rew) rew[k] = (float)r64;
    if (done) done[k] = d ? 1 : 0;
    if (traj_rew && t < traj_len) {
      traj_rew[(size_t)e * traj_len + t] = (float)r64;
      traj_done[(size_t)e * traj_len + t] = d ? 1 : 0;
    }
    if (ep_len && d) ep_len[e] = t + 1;
    if (active && d) active[e] = 0;
  }
  __syncwarp();
  return d;
}

// N(0,1) draw of (row id, component c) at Philox offset `off` (Box-Muller on two uniforms in (0,1])
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t off, int id, int c) {
  const uint4 r = philox4x32(make_uint4((uint32_t)id, (uint32_t)c, (uint32_t)off, (uint32_t)(off >> 32)),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u1 = u01(r.x), u2 = u01(r.y);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

#define CIRS_LOG_SQRT_2PI 0.9189385332046727f

// Continuous actor for ONE row by one warp: trunk + critic (bit-identical to the tile kernels), mu head, sample,
// log-prob.  s: the row's state; sc: ACTOR_SCRATCH floats of shared memory.  eps: this row's N(0,1) draws or NULL
// (Philox).  act_in: evaluate log-prob of this action instead of sampling.  Lane c < n_action returns a_c.
__device__ __forceinline__ float actorprob_warp(const cirs_policy_weights& W, const float* s, int lane, float* sc,
                                                const float* eps, uint64_t seed, uint64_t off, int id, int mode,
                                                const float* act_in, float* __restrict__ value_out, float* logp_out,
                                                float* __restrict__ mu_out) {
  const int nA = W.n_action;
  float* h2 = sc + 160;
  cirs_actor::actor_trunk_warp(W, s, lane, sc, h2, value_out);
  float a = 0.f, lp = 0.f;
  if (lane < nA) {
    float z = __ldg(W.b3 + lane);
#pragma unroll 16
    for (int k = 0; k < cirs_actor::HID; ++k) z = fmaf(h2[k], __ldg(W.w3t + (size_t)k * W.ld_action + lane), z);
    const float mu = W.max_action * tanhf(z);               // continuous.py:186-187
    const float sg = expf(__ldg(W.sigma + lane));           // :194-196
    if (act_in) a = act_in[lane];
    else if (mode == 1) a = mu;                              // deterministic_eval, ppo.py:150-151
    else {
      const float n = eps ? eps[lane] : philox_normal(seed, off, id, lane);
      a = __fadd_rn(__fmul_rn(n, sg), mu);                   // torch.normal(mean, std): N(0,1) * std + mean
    }
    const float df = a - mu;
    lp = -(df * df) / (2.0f * (sg * sg)) - logf(sg) - CIRS_LOG_SQRT_2PI;   // torch/distributions/normal.py log_prob
    if (mu_out) mu_out[lane] = mu;
  }
  lp = warp_sum(lp);                                          // Independent(., 1): sum over the action dims
  if (lane == 0 && logp_out) *logp_out = lp;
  __syncwarp();
  return a;
}

}  // namespace cirs_taobao
