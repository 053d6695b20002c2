// Device code of K1 (KuaishouEnv / SimulatedEnv step), shared by the stand-alone kernels (env_kuaishou.cu) and the
// persistent rollout kernel (rollout.cu).  See env_kuaishou.cu for the description.
#pragma once
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace cirs_env {

// reset of environment slot e by one warp
__device__ __forceinline__ void kuaishou_reset_warp(const cirs_kuaishou_env& E, int e, int user, int lane,
                                                    uint8_t* __restrict__ active) {
  if (lane == 0) {
    E.user[e] = user;
    E.turn[e] = 0;
    E.cum_rew[e] = 0.0;
    if (active) active[e] = 1;
  }
  for (int j = lane; j < E.max_turn; j += 32) E.hist[(size_t)e * E.max_turn + j] = 0;
  if (E.seen) {
    const int nw = (E.n_item + 31) >> 5;
    for (int j = lane; j < nw; j += 32) E.seen[(size_t)e * nw + j] = 0u;
  }
}

// The part of a transition that does not depend on the action, as loads a caller can issue EARLY (the persistent
// rollout issues them before it merges the head's partials, so that their L2 round trips overlap that merge): turn,
// user, the lane's slot of the first 32 history entries with its category mask, the user's alpha, the running return.
struct StepPre {
  int t, u, hj;
  uint32_t mj;
  float alpha;
  double cum;
};
__device__ __forceinline__ StepPre kuaishou_step_pre(const cirs_kuaishou_env& E, int e, int lane) {
  StepPre P;
  P.t = E.turn[e];
  P.u = E.user[e];
  P.hj = lane < E.max_turn ? E.hist[(size_t)e * E.max_turn + lane] : 0;   // slots >= t hold 0 (reset) or stale ids: valid indices
  P.cum = E.cum_rew[e];
  P.mj = __ldg(E.cat_mask + P.hj);
  P.alpha = E.alpha_u ? __ldg(E.alpha_u + P.u) : 1.f;
  return P;
}

// one transition of environment slot e by one warp (row k of the call's rew / done outputs); action a.
// Returns true when the episode ended.  n_active (optional) is decremented when it does.  pre (optional): the result of
// kuaishou_step_pre for this slot, r_out (optional): the reward (lane 0).
__device__ __forceinline__ bool kuaishou_step_warp(const cirs_kuaishou_env& E, int e, int k, int a, int lane,
                                                   uint8_t* __restrict__ active, float* __restrict__ rew,
                                                   uint8_t* __restrict__ done, int traj_len,
                                                   int32_t* __restrict__ traj_act, float* __restrict__ traj_rew,
                                                   uint8_t* __restrict__ traj_done, int32_t* __restrict__ ep_len,
                                                   int force_length, int* __restrict__ n_active,
                                                   const StepPre* pre = nullptr, float* r_out = nullptr) {
  const int T = E.max_turn, N = E.num_leave_compute;
  const int t = pre ? pre->t : E.turn[e], u = pre ? pre->u : E.user[e];
  // everything that depends on the action, issued together (one round trip)
  const uint32_t ma = __ldg(E.cat_mask + a);
  const float tab = __ldg((E.simulated ? E.normed_mat : E.mat) + (size_t)u * E.n_item + a);
  const float beta = (E.simulated && E.alpha_u) ? __ldg(E.beta_i + a) : 1.f;
  const int32_t* hist = E.hist + (size_t)e * T;

  // window of the exit test: seq[t-N : t] with Python's negative-slice wrap when t < N (kuaishouEnv.py:203)
  int w_lo = t - N;
  if (w_lo < 0) w_lo = max(0, 2 * t - N);

  double expo = 0.0;   // exposure partial sum
  int n_prev = 0;      // previous occurrences of a   (num_actions[a] - 1, simulated_env.py:129-132)
  bool leave = false;
  uint32_t bits = ma;  // categories of a still to be checked
  for (int j0 = 0; j0 < t; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = j < t;
    int hj = 0;
    uint32_t mj = 0u;
    if (pre && j0 == 0) {
      hj = pre->hj;
      mj = pre->mj;
    } else if (valid) {
      hj = hist[j];
      mj = __ldg(E.cat_mask + hj);
    }
    n_prev += __popc(__ballot_sync(FULL_MASK, valid && hj == a));
    if (E.simulated && valid && E.tau > 0.f) {
      float dist;
      if (E.dist) {
        dist = __ldg(E.dist + (size_t)a * E.n_item + hj);  // df_dist_small.iloc[a, hist], util.py:33-36
      } else {
        const int inter = __popc(ma & mj), uni = __popc(ma | mj);
        dist = inter ? (float)uni / (float)inter : INFINITY;  // 1 / Jaccard, util.py:234-268
      }
      expo += (double)expf(-(float)(t - j) * dist / E.tau);  // util.py:45
    }
  }
  // exit test: count_c over the window items for every category c of the action (kuaishouEnv.py:204-213).
  // The window holds at most N items (usually inside one 32-slot chunk); its masks are L1 hits from pass one.
  if (t > 0) {
    while (bits) {
      const int c = __ffs(bits) - 1;
      bits &= bits - 1;
      int cnt = 0;
      for (int j0 = (w_lo & ~31); j0 < t; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j >= w_lo && j < t) hit = ((pre && j0 == 0 ? pre->mj : __ldg(E.cat_mask + hist[j])) >> c) & 1u;
        cnt += __popc(__ballot_sync(FULL_MASK, hit));
      }
      if ((float)cnt > E.leave_threshold) leave = true;
    }
  }
  expo = warp_sum_d(expo);

  bool d = leave || (t >= T - 1);  // kuaishouEnv.py:166-168
  if (force_length > 0) d = (t + 1 >= force_length);  // collector.py:253-258
  if (lane == 0) {
    float r;
    if (!E.simulated) {
      r = tab;  // kuaishouEnv.py:171
    } else {
      double ex = (t == 0 || E.tau <= 0.f) ? 0.0 : expo;
      if (E.alpha_u) ex = ex * (double)(pre ? pre->alpha : __ldg(E.alpha_u + u)) * (double)beta;
      ex *= (double)E.gamma_exposure;
      const double pr = (double)tab;
      double rr = (E.version == 1) ? pr / (1.0 + ex) : (pr - ex);  // clip0 is an identity, util.py:53-54
      if (E.r_decay != 1.0f) rr *= pow((double)E.r_decay, (double)n_prev);
      r = (float)rr;
    }
    if (t < T) E.hist[(size_t)e * T + t] = a;
    E.turn[e] = t + 1;
    E.cum_rew[e] = (pre ? pre->cum : E.cum_rew[e]) + (double)r;
    if (E.seen) E.seen[(size_t)e * ((E.n_item + 31) >> 5) + (a >> 5)] |= (1u << (a & 31));
    rew[k] = r;
    if (r_out) *r_out = r;
    done[k] = d ? 1 : 0;
    if (traj_act && t < traj_len) {
      traj_act[(size_t)e * traj_len + t] = a;
      traj_rew[(size_t)e * traj_len + t] = r;
      traj_done[(size_t)e * traj_len + t] = d ? 1 : 0;
    }
    if (ep_len && d) ep_len[e] = t + 1;
    if (active && d) active[e] = 0;
    if (n_active && d) atomicSub(n_active, 1);
  }
  return d;
}

}  // namespace cirs_env
