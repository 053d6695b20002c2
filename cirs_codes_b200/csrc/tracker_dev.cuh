// Device code of K2 (StateTracker rollout step), shared by the stand-alone kernel (tracker_step.cu) and the
// persistent rollout kernel (rollout.cu).  See tracker_step.cu for the description.
#pragma once
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace cirs_tracker {



template <int NK>
__device__ __forceinline__ void matvec_chunk(const float* __restrict__ wp, const float* __restrict__ x, int n_in,
                                             int ldo, float (&acc)[4]) {
#pragma unroll 4
  for (int i = 0; i < n_in; ++i) {
    const float xi = x[i];
    const float* w = wp + (size_t)i * ldo;
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] = fmaf(__ldg(w + 32 * k), xi, acc[k]);
  }
}

// y[o] = act(b[o] + sum_i Wt[i][o] * x[i]) for o < n_out.  x, y in shared memory (y != x).
// act: 0 identity, 1 relu, 2 sigmoid
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, const float* __restrict__ b, const float* x,
                                       int n_in, int n_out, int ldo, float* y, int lane, int act) {
  for (int o0 = 0; o0 < n_out; o0 += 128) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int nk = min(4, (ldo - o0) >> 5);
    const float* wp = Wt + o0 + lane;
    switch (nk) {
      case 1: matvec_chunk<1>(wp, x, n_in, ldo, acc); break;
      case 2: matvec_chunk<2>(wp, x, n_in, ldo, acc); break;
      case 3: matvec_chunk<3>(wp, x, n_in, ldo, acc); break;
      default: matvec_chunk<4>(wp, x, n_in, ldo, acc); break;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = o0 + lane + 32 * k;
      if (k < nk && o < n_out) {
        float v = acc[k] + __ldg(b + o);
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = 1.f / (1.f + expf(-v));
        y[o] = v;
      }
    }
  }
  __syncwarp();
}

// x <- LayerNorm(x + y) * w + b  over d elements held in shared memory (biased variance, eps 1e-5)
__device__ __forceinline__ void add_layernorm(float* x, const float* y, const float* __restrict__ w,
                                              const float* __restrict__ b, int d, int lane) {
  float s = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[c] + y[c];
    x[c] = v;
    s += v;
  }
  const float mu = warp_sum(s) / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = x[c] - mu;
    q += v * v;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / d + 1e-5f);
  for (int c = lane; c < d; c += 32) x[c] = (x[c] - mu) * rstd * __ldg(w + c) + __ldg(b + c);
  __syncwarp();
}


__host__ __device__ inline int trk_up32(int v) { return (v + 31) & ~31; }
// floats of shared-memory scratch one warp needs
__host__ __device__ inline int tracker_scratch_floats(const cirs_tracker_weights& W) {
  const int ldd = trk_up32(W.d), ld3 = trk_up32(3 * W.d), ldh = trk_up32(W.d_hid);
  const int win = trk_up32(W.d_user_in > W.d_item_in + 1 ? W.d_user_in : W.d_item_in + 1);
  return ldd + (ldd > win ? ldd : win) + ld3 + (ldh > ldd ? ldh : ldd) + trk_up32(W.nhead * W.max_len);
}

// One new token of environment slot e (sequence position p) by one warp.  ``x`` = this warp's scratch
// (tracker_scratch_floats(W) floats of shared memory).  id: user / item id when the table exists, else src_dense.
__device__ __forceinline__ void tracker_token_warp(const cirs_tracker_weights& W, int n_env, int e, int k, int p,
                                                   int id, const float* __restrict__ src_dense, float rew_k,
                                                   float* __restrict__ kcache, float* __restrict__ vcache,
                                                   float* x, int lane, float* __restrict__ state_out,
                                                   int64_t state_stride, float* __restrict__ cur_state, int traj_len,
                                                   float* __restrict__ traj_obs, float* __restrict__ traj_obs_next) {
  const int d = W.d, nh = W.nhead, dh = d / nh, dhid = W.d_hid;
  const int ldd = (d + 31) & ~31, ld3 = (3 * d + 31) & ~31, ldh = (dhid + 31) & ~31, lds = (W.dim_state + 31) & ~31;
  // x: [d] running activation
  float* y = x + ldd;                                 // [max(d, d_in+1)] scratch vector
  float* qkv = y + max(ldd, ((max(W.d_user_in, W.d_item_in + 1) + 31) & ~31));  // [3d]
  float* hid = qkv + ld3;                             // [d_hid] (also attention output o[d])
  float* prob = hid + max(ldh, ldd);                  // [nhead][max_len]

  // ---- token -------------------------------------------------------------------------------------------
  if (p == 0) {
    const int n_in = W.d_user_in;
    const float* src = W.emb_user ? W.emb_user + (size_t)id * d : src_dense;
    for (int c = lane; c < n_in; c += 32) y[c] = __ldg(src + c);
    __syncwarp();
    matvec(W.user_wt, W.user_b, y, n_in, d, ldd, x, lane, 0);  // ffn_user, state_tracker.py:212
  } else {
    const int n_in = W.d_item_in;  // == d (the gate multiplies the item vector elementwise)
    const float* src = W.emb_item ? W.emb_item + (size_t)id * d : src_dense;
    if (lane == 0) y[0] = rew_k;
    for (int c = lane; c < n_in; c += 32) y[1 + c] = __ldg(src + c);
    __syncwarp();
    matvec(W.gate_wt, W.gate_b, y, 1 + n_in, d, ldd, x, lane, 2);  // g = sigmoid(W_g [r;a] + b_g), :239
    for (int c = lane; c < d; c += 32) x[c] *= y[1 + c];            // a' = g * a, :240
    __syncwarp();
  }
  const float sq = sqrtf((float)d);
  for (int c = lane; c < d; c += 32) x[c] = x[c] * sq + __ldg(W.pe + (size_t)p * d + c);  // :180-181
  __syncwarp();

  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < W.nlayers; ++l) {
    const cirs_encoder_layer& L = W.layer[l];
    matvec(L.in_wt, L.in_b, x, d, 3 * d, ld3, qkv, lane, 0);
    float* kc = kcache + ((size_t)l * n_env + e) * W.max_len * d;
    float* vc = vcache + ((size_t)l * n_env + e) * W.max_len * d;
    for (int c = lane; c < d; c += 32) {
      kc[(size_t)p * d + c] = qkv[d + c];
      vc[(size_t)p * d + c] = qkv[2 * d + c];
    }
    // scores: lane j handles cached position j
    for (int h = 0; h < nh; ++h) {
      const float* q = qkv + h * dh;
      float mx = -INFINITY;
      for (int j0 = 0; j0 <= p; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j <= p) {
          const float* kr = (j == p) ? (qkv + d + h * dh) : (kc + (size_t)j * d + h * dh);
          float a = 0.f;
          for (int c = 0; c < dh; ++c) a = fmaf(q[c] * scale, kr[c], a);
          s = a;
          prob[h * W.max_len + j] = s;
        }
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      __syncwarp();
      float sum = 0.f;
      for (int j = lane; j <= p; j += 32) {
        const float ex = expf(prob[h * W.max_len + j] - mx);
        prob[h * W.max_len + j] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j <= p; j += 32) prob[h * W.max_len + j] *= inv;
    }
    __syncwarp();
    // o[c] = sum_j prob[head(c)][j] * V[j][c]
    for (int c = lane; c < d; c += 32) {
      const float* pr = prob + (c / dh) * W.max_len;
      float a = 0.f;
      for (int j = 0; j < p; ++j) a = fmaf(pr[j], vc[(size_t)j * d + c], a);
      a = fmaf(pr[p], qkv[2 * d + c], a);
      hid[c] = a;
    }
    __syncwarp();
    matvec(L.out_wt, L.out_b, hid, d, d, ldd, y, lane, 0);
    add_layernorm(x, y, L.n1_w, L.n1_b, d, lane);
    matvec(L.l1_wt, L.l1_b, x, d, dhid, ldh, hid, lane, 1);
    matvec(L.l2_wt, L.l2_b, hid, dhid, d, ldd, y, lane, 0);
    add_layernorm(x, y, L.n2_w, L.n2_b, d, lane);
  }
  // decoder -> state
  matvec(W.dec_wt, W.dec_b, x, d, W.dim_state, lds, y, lane, 0);
  const int S = W.dim_state;
  for (int c = lane; c < S; c += 32) {
    const float v = y[c];
    if (state_out) state_out[(size_t)k * state_stride + c] = v;
    if (cur_state) cur_state[(size_t)e * S + c] = v;
    if (traj_obs && p < traj_len) traj_obs[((size_t)e * traj_len + p) * S + c] = v;
    if (traj_obs_next && p >= 1 && p - 1 < traj_len) traj_obs_next[((size_t)e * traj_len + p - 1) * S + c] = v;
  }
}

}  // namespace cirs_tracker
