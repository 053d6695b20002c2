// Device code of K2 (StateTracker rollout step), shared by the stand-alone kernel (tracker_step.cu) and the
// persistent rollout kernel (rollout.cu).  See tracker_step.cu for the description.
#pragma once
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace cirs_tracker {



// weight loads: global memory through the read-only path, or plain loads when the caller staged the weights in
// shared memory (persistent rollout kernel, when they fit)
template <bool SM>
__device__ __forceinline__ float ldw(const float* p) {
  if (SM) return *p;
  return __ldg(p);
}

template <bool SM>
__device__ __forceinline__ float4 ldw4(const float* p) {
  if (SM) return *reinterpret_cast<const float4*>(p);
  return __ldg(reinterpret_cast<const float4*>(p));
}

// y[o] = act(b[o] + sum_i Wt[i][o] * x[i]) for o < n_out.  x, y in shared memory (y != x).
// act: 0 identity, 1 relu, 2 sigmoid.
// Lane mapping: a chunk of up to 128 outputs is 32 "quads" of 4 consecutive outputs; with nq quads in the chunk the
// warp forms G = 32 / P groups (P = nq rounded up to a power of two): lane (g, q) accumulates quad q over the
// inputs i = g, g + G, ... with ONE 16-byte weight load per 4 FMAs, and the groups are summed with xor-shuffles.
// This cuts the instruction count per token ~5x against one-output-per-lane scalar loads (the single resident warp
// per scheduler is issue/latency bound, not bandwidth bound).
template <bool SM>
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, const float* __restrict__ b, const float* x,
                                       int n_in, int n_out, int ldo, float* y, int lane, int act) {
  for (int o0 = 0; o0 < n_out; o0 += 128) {
    const int nq = min(32, (n_out - o0 + 3) >> 2);
    const int P = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
    const int G = 32 / P, q = lane & (P - 1), g = lane / P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
      const float* w = Wt + o0 + 4 * q + (size_t)g * ldo;
      const int step = G * ldo;
#pragma unroll 8
      for (int i = g; i < n_in; i += G) {
        const float xi = x[i];
        const float4 w4 = ldw4<SM>(w);
        acc.x = fmaf(w4.x, xi, acc.x);
        acc.y = fmaf(w4.y, xi, acc.y);
        acc.z = fmaf(w4.z, xi, acc.z);
        acc.w = fmaf(w4.w, xi, acc.w);
        w += step;
      }
    }
    for (int off = P; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(FULL_MASK, acc.x, off);
      acc.y += __shfl_xor_sync(FULL_MASK, acc.y, off);
      acc.z += __shfl_xor_sync(FULL_MASK, acc.z, off);
      acc.w += __shfl_xor_sync(FULL_MASK, acc.w, off);
    }
    if (g == 0 && q < nq) {
      const float v4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = o0 + 4 * q + j;
        if (o < n_out) {
          float v = v4[j] + ldw<SM>(b + o);
          if (act == 1) v = fmaxf(v, 0.f);
          else if (act == 2) v = 1.f / (1.f + expf(-v));
          y[o] = v;
        }
      }
    }
  }
  __syncwarp();
}

// x <- LayerNorm(x + y) * w + b  over d elements held in shared memory (biased variance, eps 1e-5)
template <bool SM>
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
  for (int c = lane; c < d; c += 32) x[c] = (x[c] - mu) * rstd * ldw<SM>(w + c) + ldw<SM>(b + c);
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
template <bool SM = false>
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
    // dense inputs may live in shared memory or have been written by this kernel: plain loads
    for (int c = lane; c < n_in; c += 32) y[c] = W.emb_user ? __ldg(src + c) : src[c];
    __syncwarp();
    matvec<SM>(W.user_wt, W.user_b, y, n_in, d, ldd, x, lane, 0);  // ffn_user, state_tracker.py:212
  } else {
    const int n_in = W.d_item_in;  // == d (the gate multiplies the item vector elementwise)
    const float* src = W.emb_item ? W.emb_item + (size_t)id * d : src_dense;
    if (lane == 0) y[0] = rew_k;
    for (int c = lane; c < n_in; c += 32) y[1 + c] = W.emb_item ? __ldg(src + c) : src[c];
    __syncwarp();
    matvec<SM>(W.gate_wt, W.gate_b, y, 1 + n_in, d, ldd, x, lane, 2);  // g = sigmoid(W_g [r;a] + b_g), :239
    for (int c = lane; c < d; c += 32) x[c] *= y[1 + c];            // a' = g * a, :240
    __syncwarp();
  }
  const float sq = sqrtf((float)d);
  for (int c = lane; c < d; c += 32) x[c] = x[c] * sq + __ldg(W.pe + (size_t)p * d + c);  // :180-181
  __syncwarp();

  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < W.nlayers; ++l) {
    const cirs_encoder_layer& L = W.layer[l];
    matvec<SM>(L.in_wt, L.in_b, x, d, 3 * d, ld3, qkv, lane, 0);
    float* kc = kcache + ((size_t)l * n_env + e) * W.max_len * d;
    float* vc = vcache + ((size_t)l * n_env + e) * W.max_len * d;
    for (int c = lane; c < d; c += 32) {
      kc[(size_t)p * d + c] = qkv[d + c];
      vc[(size_t)p * d + c] = qkv[2 * d + c];
    }
    // scores of ALL heads: lane j (chunks of 32) owns cached position j and reads that key row once; the loads of
    // one row are independent, so they are issued as a batch (one L2 round trip per 32 positions)
    const bool vec4 = ((d & 3) == 0) && ((dh & 3) == 0);
    for (int j0 = 0; j0 <= p; j0 += 32) {
      const int j = j0 + lane;
      if (j <= p) {
        const float* kr = (j == p) ? (qkv + d) : (kc + (size_t)j * d);
        if (vec4) {
#pragma unroll 4
          for (int h = 0; h < nh; ++h) {
            const float* q = qkv + h * dh;
            const float4* k4 = reinterpret_cast<const float4*>(kr + h * dh);
            float a = 0.f;
#pragma unroll 4
            for (int c4 = 0; c4 < (dh >> 2); ++c4) {
              const float4 kv = k4[c4];
              a = fmaf(q[4 * c4] * scale, kv.x, a);
              a = fmaf(q[4 * c4 + 1] * scale, kv.y, a);
              a = fmaf(q[4 * c4 + 2] * scale, kv.z, a);
              a = fmaf(q[4 * c4 + 3] * scale, kv.w, a);
            }
            prob[h * W.max_len + j] = a;
          }
        } else {
          for (int h = 0; h < nh; ++h) {
            const float* q = qkv + h * dh;
            float a = 0.f;
#pragma unroll 4
            for (int c = 0; c < dh; ++c) a = fmaf(q[c] * scale, kr[h * dh + c], a);
            prob[h * W.max_len + j] = a;
          }
        }
      }
    }
    __syncwarp();
    for (int h = 0; h < nh; ++h) {
      float* ph = prob + h * W.max_len;
      float mx = -INFINITY;
      for (int j = lane; j <= p; j += 32) mx = fmaxf(mx, ph[j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j <= p; j += 32) {
        const float ex = expf(ph[j] - mx);
        ph[j] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j <= p; j += 32) ph[j] *= inv;
    }
    __syncwarp();
    // o[c] = sum_j prob[head(c)][j] * V[j][c]   (row j of V is one coalesced line; loads batched by the unroll)
    for (int c = lane; c < d; c += 32) {
      const float* pr = prob + (c / dh) * W.max_len;
      float a = 0.f;
#pragma unroll 8
      for (int j = 0; j < p; ++j) a = fmaf(pr[j], vc[(size_t)j * d + c], a);
      a = fmaf(pr[p], qkv[2 * d + c], a);
      hid[c] = a;
    }
    __syncwarp();
    matvec<SM>(L.out_wt, L.out_b, hid, d, d, ldd, y, lane, 0);
    add_layernorm<SM>(x, y, L.n1_w, L.n1_b, d, lane);
    matvec<SM>(L.l1_wt, L.l1_b, x, d, dhid, ldh, hid, lane, 1);
    matvec<SM>(L.l2_wt, L.l2_b, hid, dhid, d, ldd, y, lane, 0);
    add_layernorm<SM>(x, y, L.n2_w, L.n2_b, d, lane);
  }
  // decoder -> state
  matvec<SM>(W.dec_wt, W.dec_b, x, d, W.dim_state, lds, y, lane, 0);
  const int S = W.dim_state;
  for (int c = lane; c < S; c += 32) {
    const float v = y[c];
    if (state_out) state_out[(size_t)k * state_stride + c] = v;
    if (cur_state) cur_state[(size_t)e * S + c] = v;
    if (traj_obs && p < traj_len) traj_obs[((size_t)e * traj_len + p) * S + c] = v;
    if (traj_obs_next && p >= 1 && p - 1 < traj_len) traj_obs_next[((size_t)e * traj_len + p - 1) * S + c] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Group version: G warps (G = 2, 4, 8; all in one CTA) cooperate on ONE token.  A single warp walking the 2-layer
// transformer step is a ~30 us chain of dependent shared-memory loads and FMAs; with few environments still running
// most of the SM's warps idle, so the matrix-vector products are split over the group's warps (each warp owns a
// contiguous range of output quads), heads are split for the attention scores, and cheap element-wise stages are
// computed redundantly by every warp into write-only buffers (no in-place updates, so no barrier is needed for
// them).  Warps meet at a named barrier (bar.sync id, 32 G) after every split stage: 12 barriers per token.
// Scratch: ``x`` = the group's first warp's slice, ``xtra`` = at least 2 * round_up(d, 32) more floats.

__device__ __forceinline__ void group_sync(int bar_id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_threads) : "memory");
}

// this warp's share of y[o] = act(b[o] + sum_i Wt[i][o] x[i]); no trailing synchronisation
template <bool SM>
__device__ __forceinline__ void matvec_part(const float* __restrict__ Wt, const float* __restrict__ b, const float* x,
                                            int n_in, int n_out, int ldo, float* y, int lane, int act, int wg, int G) {
  const int nq_total = (n_out + 3) >> 2;
  const int qpw = min(32, (nq_total + G - 1) / G);          // quads per warp (<= 32: n_out <= 128 G)
  for (int q0 = wg * qpw; q0 < nq_total; q0 += G * qpw) {
    const int nq = min(qpw, nq_total - q0);
    const int P = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
    const int Gi = 32 / P, q = lane & (P - 1), g = lane / P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
      const float* w = Wt + 4 * (q0 + q) + (size_t)g * ldo;
      const int step = Gi * ldo;
#pragma unroll 8
      for (int i = g; i < n_in; i += Gi) {
        const float xi = x[i];
        const float4 w4 = ldw4<SM>(w);
        acc.x = fmaf(w4.x, xi, acc.x);
        acc.y = fmaf(w4.y, xi, acc.y);
        acc.z = fmaf(w4.z, xi, acc.z);
        acc.w = fmaf(w4.w, xi, acc.w);
        w += step;
      }
    }
    for (int off = P; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(FULL_MASK, acc.x, off);
      acc.y += __shfl_xor_sync(FULL_MASK, acc.y, off);
      acc.z += __shfl_xor_sync(FULL_MASK, acc.z, off);
      acc.w += __shfl_xor_sync(FULL_MASK, acc.w, off);
    }
    if (g == 0 && q < nq) {
      const float v4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = 4 * (q0 + q) + j;
        if (o < n_out) {
          float v = v4[j] + ldw<SM>(b + o);
          if (act == 1) v = fmaxf(v, 0.f);
          else if (act == 2) v = 1.f / (1.f + expf(-v));
          y[o] = v;
        }
      }
    }
  }
}

// out <- LayerNorm(xin + yin) * w + b, all d elements by THIS warp (out is neither xin nor yin)
template <bool SM>
__device__ __forceinline__ void add_layernorm_to(float* out, const float* xin, const float* yin,
                                                 const float* __restrict__ w, const float* __restrict__ b, int d,
                                                 int lane) {
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += xin[c] + yin[c];
  const float mu = warp_sum(s) / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = xin[c] + yin[c] - mu;
    q += v * v;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / d + 1e-5f);
  for (int c = lane; c < d; c += 32) out[c] = (xin[c] + yin[c] - mu) * rstd * ldw<SM>(w + c) + ldw<SM>(b + c);
  __syncwarp();
}

template <bool SM>
__device__ __forceinline__ void tracker_token_group(const cirs_tracker_weights& W, int n_env, int e, int p, int id,
                                                    float rew_k, float* __restrict__ kcache,
                                                    float* __restrict__ vcache, float* x, float* xtra, int lane, int wg,
                                                    int G, int bar_id, float* __restrict__ cur_state, int traj_len,
                                                    float* __restrict__ traj_obs, float* __restrict__ traj_obs_next,
                                                    long long* tq = nullptr) {
  // tq (optional, one group only): %globaltimer stamps after token / per layer: in-proj, attention, out-proj+LN, FFN+LN / decoder
  int tqi = 0;
  auto stamp = [&]() {
    if (tq && wg == 0 && lane == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); tq[tqi++] = t_; }
  };
  stamp();
  const int d = W.d, nh = W.nhead, dh = d / nh, dhid = W.d_hid;
  const int ldd = (d + 31) & ~31, ld3 = (3 * d + 31) & ~31, ldh = (dhid + 31) & ~31, lds = (W.dim_state + 31) & ~31;
  const int nthr = 32 * G;
  float* y = x + ldd;                                 // token inputs [1 + d]
  float* qkv = y + max(ldd, ((max(W.d_user_in, W.d_item_in + 1) + 31) & ~31));  // [3d]
  float* hid = qkv + ld3;                             // [d_hid] (also attention output o[d])
  float* prob = hid + max(ldh, ldd);                  // [nhead][max_len]
  float* x2 = xtra;                                   // [d] second activation buffer (ping-pong with x)
  float* yb = xtra + ldd;                             // [d] output of gate / out_proj / linear2 / decoder

  // ---- token (embedding tables only: the Kuaishou rollout); every warp fills y redundantly (write-only)
  if (p == 0) {
    const float* src = W.emb_user + (size_t)id * d;
    for (int c = lane; c < W.d_user_in; c += 32) y[c] = __ldg(src + c);
    __syncwarp();
    matvec_part<SM>(W.user_wt, W.user_b, y, W.d_user_in, d, ldd, yb, lane, 0, wg, G);
    group_sync(bar_id, nthr);
    const float sq = sqrtf((float)d);
    for (int c = lane; c < d; c += 32) x[c] = yb[c] * sq + __ldg(W.pe + c);
  } else {
    const float* src = W.emb_item + (size_t)id * d;
    if (lane == 0) y[0] = rew_k;
    for (int c = lane; c < W.d_item_in; c += 32) y[1 + c] = __ldg(src + c);
    __syncwarp();
    matvec_part<SM>(W.gate_wt, W.gate_b, y, 1 + W.d_item_in, d, ldd, yb, lane, 2, wg, G);
    group_sync(bar_id, nthr);
    const float sq = sqrtf((float)d);
    for (int c = lane; c < d; c += 32) x[c] = (yb[c] * y[1 + c]) * sq + __ldg(W.pe + (size_t)p * d + c);
  }
  __syncwarp();
  stamp();

  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < W.nlayers; ++l) {
    const cirs_encoder_layer& L = W.layer[l];
    matvec_part<SM>(L.in_wt, L.in_b, x, d, 3 * d, ld3, qkv, lane, 0, wg, G);
    group_sync(bar_id, nthr);
    stamp();
    float* kc = kcache + ((size_t)l * n_env + e) * W.max_len * d;
    float* vc = vcache + ((size_t)l * n_env + e) * W.max_len * d;
    if (wg == G - 1) {   // the last warp (it owns no head when G > nhead) stores this position's K / V
      for (int c = lane; c < d; c += 32) {
        kc[(size_t)p * d + c] = qkv[d + c];
        vc[(size_t)p * d + c] = qkv[2 * d + c];
      }
    }
    // scores + softmax of head h by warp h % G: lane j (chunks of 32) owns cached position j
    for (int h = wg; h < nh; h += G) {
      const float* q = qkv + h * dh;
      float* ph = prob + h * W.max_len;
      for (int j0 = 0; j0 <= p; j0 += 32) {
        const int j = j0 + lane;
        if (j <= p) {
          const float* kr = (j == p) ? (qkv + d) : (kc + (size_t)j * d);
          float a = 0.f;
#pragma unroll 4
          for (int c = 0; c < dh; ++c) a = fmaf(q[c] * scale, kr[h * dh + c], a);
          ph[j] = a;
        }
      }
      __syncwarp();
      float mx = -INFINITY;
      for (int j = lane; j <= p; j += 32) mx = fmaxf(mx, ph[j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j <= p; j += 32) {
        const float ex = expf(ph[j] - mx);
        ph[j] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j <= p; j += 32) ph[j] *= inv;
    }
    group_sync(bar_id, nthr);
    // o[c] = sum_j prob[head(c)][j] V[j][c]: every warp computes all of o (write-only ``hid``)
    for (int c = lane; c < d; c += 32) {
      const float* pr = prob + (c / dh) * W.max_len;
      float a = 0.f;
#pragma unroll 8
      for (int j = 0; j < p; ++j) a = fmaf(pr[j], vc[(size_t)j * d + c], a);
      a = fmaf(pr[p], qkv[2 * d + c], a);
      hid[c] = a;
    }
    __syncwarp();
    stamp();
    matvec_part<SM>(L.out_wt, L.out_b, hid, d, d, ldd, yb, lane, 0, wg, G);
    group_sync(bar_id, nthr);
    add_layernorm_to<SM>(x2, x, yb, L.n1_w, L.n1_b, d, lane);
    stamp();
    matvec_part<SM>(L.l1_wt, L.l1_b, x2, d, dhid, ldh, hid, lane, 1, wg, G);
    group_sync(bar_id, nthr);
    matvec_part<SM>(L.l2_wt, L.l2_b, hid, dhid, d, ldd, yb, lane, 0, wg, G);
    group_sync(bar_id, nthr);
    add_layernorm_to<SM>(x, x2, yb, L.n2_w, L.n2_b, d, lane);
    stamp();
  }
  // decoder -> state
  matvec_part<SM>(W.dec_wt, W.dec_b, x, d, W.dim_state, lds, yb, lane, 0, wg, G);
  group_sync(bar_id, nthr);
  stamp();
  if (wg == 0) {
    const int S = W.dim_state;
    for (int c = lane; c < S; c += 32) {
      const float v = yb[c];
      if (cur_state) cur_state[(size_t)e * S + c] = v;
      if (traj_obs && p < traj_len) traj_obs[((size_t)e * traj_len + p) * S + c] = v;
      if (traj_obs_next && p >= 1 && p - 1 < traj_len) traj_obs_next[((size_t)e * traj_len + p - 1) * S + c] = v;
    }
  }
}

}  // namespace cirs_tracker
