// K2 + trunk for the persistent rollout kernel, latency form (d = 32, d_hid = 128: the reference's default tracker).
//
// A turn's phase B is a chain of ~20 dependent matrix-vector stages per environment, and from the third turn on a CTA
// holds ONE row: what matters is the latency of a stage, not its throughput.  tracker_cta_dev.cuh gives every
// (row, output) to one thread, i.e. a 32 .. 128 long dependent FMA chain behind 32 .. 128 scalar shared-memory loads
// per stage (measured 0.5 - 1.5 us per stage, 15 us per token).  Here the weights are staged TRANSPOSED ([out][in],
// once per launch) and a (row, output) pair is shared by 8 consecutive lanes: each lane reads one 16-byte vector of
// the weight row and one of the input (a quarter-warp reads 32 consecutive floats of each: conflict-free), runs 4
// FMAs per 32 inputs, and three xor-shuffles finish the sum.  One barrier per stage, ~20 instructions per thread and
// stage, all 256 threads busy even for a single row.  Summation order differs from the stand-alone kernels' (FP32
// rounding level, like the cooperative form before it).
#pragma once
#include "tracker_dev.cuh"

namespace cirs_tfast {
using cirs_tracker::trk_up32;

constexpr int NT = 256;
constexpr int D = 32, DHID = 128, HIDP = 64;   // model width, FFN width, policy trunk width (CIRS_HIDDEN)

struct Layer { int in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b; };
struct Layout {   // float offsets of the transposed copies inside the staged block
  int user_w, user_b, gate_w, gate_w0, gate_b, dec_w, dec_b;
  Layer layer[CIRS_MAX_LAYERS];
  int total;
};
// trunk image (one contiguous block, rebuilt per launch, re-fetched per turn): w1T [64][32] | b1 [64] | w2T [64][64] |
// b2 [64] | wv [64] | bv [4]
constexpr int TR_W1 = 0, TR_B1 = TR_W1 + HIDP * 32, TR_W2 = TR_B1 + HIDP, TR_B2 = TR_W2 + HIDP * HIDP,
              TR_WV = TR_B2 + HIDP, TR_BV = TR_WV + HIDP, TR_FLOATS = TR_BV + 4;

__host__ inline bool supported(const cirs_tracker_weights& W, const cirs_policy_weights& P) {
  return W.d == D && W.d_hid == DHID && W.d_user_in == D && W.d_item_in == D && W.dim_state <= 32 && W.nhead > 0 &&
         D % W.nhead == 0 && ((D / W.nhead) & 3) == 0 && W.emb_user && W.emb_item && P.dim_state == W.dim_state;
}
__host__ inline Layout layout(int nlayers) {
  Layout L{};
  int off = 0;
  auto take = [&](int n) { const int r = off; off += n; return r; };
  L.user_w = take(D * D); L.user_b = take(D);
  L.gate_w = take(D * D); L.gate_w0 = take(D); L.gate_b = take(D);
  L.dec_w = take(32 * D); L.dec_b = take(32);
  for (int l = 0; l < nlayers; ++l) {
    Layer& Y = L.layer[l];
    Y.in_w = take(3 * D * D); Y.in_b = take(3 * D);
    Y.out_w = take(D * D); Y.out_b = take(D);
    Y.l1_w = take(DHID * D); Y.l1_b = take(DHID);
    Y.l2_w = take(D * DHID); Y.l2_b = take(D);
    Y.n1_w = take(D); Y.n1_b = take(D); Y.n2_w = take(D); Y.n2_b = take(D);
  }
  L.total = off;
  return L;
}

// dst[o][k] (row length NI, NO rows) = src[k * ldo + o] for o < n_out, k < n_in, else 0.  One warp per 32 x 32 block:
// coalesced global reads along o, 16-byte shared stores.
__device__ __forceinline__ void stage_T(float* dst, const float* __restrict__ src, int n_in, int n_out, int ldo, int NI,
                                        int NO) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = NI >> 5, nb = (NO >> 5) * kb;
  for (int blk = warp; blk < nb; blk += NT / 32) {
    const int o = (blk / kb) * 32 + lane, k0 = (blk % kb) * 32;
    float v[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) v[kk] = (o < n_out && k0 + kk < n_in) ? __ldg(src + (size_t)(k0 + kk) * ldo + o) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(dst + (size_t)o * NI + k0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}
__device__ __forceinline__ void stage_vec(float* dst, const float* __restrict__ src, int n, int N) {
  for (int i = threadIdx.x; i < N; i += NT) dst[i] = i < n ? __ldg(src + i) : 0.f;
}

// every thread of the CTA; ws = the staged block (Layout offsets).  Ends with a barrier.
__device__ __forceinline__ void stage_tracker(const cirs_tracker_weights& W, const Layout& L, float* ws) {
  const int ldd = trk_up32(W.d), ld3 = trk_up32(3 * W.d), ldh = trk_up32(W.d_hid), lds = trk_up32(W.dim_state);
  stage_T(ws + L.user_w, W.user_wt, D, D, ldd, D, D);
  stage_vec(ws + L.user_b, W.user_b, D, D);
  stage_T(ws + L.gate_w, W.gate_wt + ldd, D, D, ldd, D, D);   // rows 1 .. d of [1 + d][ldd]: the item inputs
  stage_vec(ws + L.gate_w0, W.gate_wt, D, D);                  // row 0: the reward input
  stage_vec(ws + L.gate_b, W.gate_b, D, D);
  stage_T(ws + L.dec_w, W.dec_wt, D, W.dim_state, lds, D, 32);
  stage_vec(ws + L.dec_b, W.dec_b, W.dim_state, 32);
  for (int l = 0; l < W.nlayers; ++l) {
    const cirs_encoder_layer& Y = W.layer[l];
    const Layer& Z = L.layer[l];
    stage_T(ws + Z.in_w, Y.in_wt, D, 3 * D, ld3, D, 3 * D);
    stage_vec(ws + Z.in_b, Y.in_b, 3 * D, 3 * D);
    stage_T(ws + Z.out_w, Y.out_wt, D, D, ldd, D, D);
    stage_vec(ws + Z.out_b, Y.out_b, D, D);
    stage_T(ws + Z.l1_w, Y.l1_wt, D, DHID, ldh, D, DHID);
    stage_vec(ws + Z.l1_b, Y.l1_b, DHID, DHID);
    stage_T(ws + Z.l2_w, Y.l2_wt, DHID, D, ldd, DHID, D);
    stage_vec(ws + Z.l2_b, Y.l2_b, D, D);
    stage_vec(ws + Z.n1_w, Y.n1_w, D, D); stage_vec(ws + Z.n1_b, Y.n1_b, D, D);
    stage_vec(ws + Z.n2_w, Y.n2_w, D, D); stage_vec(ws + Z.n2_b, Y.n2_b, D, D);
  }
  __syncthreads();
}
// the policy trunk's image into shared memory (ts, TR_FLOATS floats); ends with a barrier
__device__ __forceinline__ void stage_trunk(const cirs_policy_weights& P, float* ts) {
  stage_T(ts + TR_W1, P.w1t, P.dim_state, HIDP, HIDP, 32, HIDP);
  stage_vec(ts + TR_B1, P.b1, HIDP, HIDP);
  stage_T(ts + TR_W2, P.w2t, HIDP, HIDP, HIDP, HIDP, HIDP);
  stage_vec(ts + TR_B2, P.b2, HIDP, HIDP);
  stage_vec(ts + TR_WV, P.wv, HIDP, HIDP);
  stage_vec(ts + TR_BV, P.bv, 1, 4);
  __syncthreads();
}

// ep(r, o, b[o] + sum_k W[o][k] x[r][k]) for r < R, o < NO; W [NO][NI] transposed in shared memory, x[r] at
// xin + r * stride (16-byte aligned, NI floats).  NI, NO multiples of 32.  Ends with a barrier.
// (Unrolling several passes per trip for instruction-level parallelism was measured SLOWER, 453 vs 429 us per rollout:
// the token step is bound by the number of instructions on its critical path, not by their latencies.)
template <int NI, int NO, class Ep>
__device__ __forceinline__ void mv(const float* W, const float* b, const float* xin, int stride, int R, Ep ep) {
  const int tid = threadIdx.x, ks = tid & 7;
  const int tasks = R * NO;   // a multiple of 32: every pass is full
  for (int t0 = 0; t0 < tasks; t0 += NT / 8) {
    const int t = t0 + (tid >> 3);
    const int r = t / NO, o = t - r * NO;
    const float* w = W + (size_t)o * NI + 4 * ks;
    const float* x = xin + (size_t)r * stride + 4 * ks;
    float acc[NI / 32];
#pragma unroll
    for (int i = 0; i < NI / 32; ++i) {
      const float4 wv = *reinterpret_cast<const float4*>(w + 32 * i);
      const float4 xv = *reinterpret_cast<const float4*>(x + 32 * i);
      acc[i] = fmaf(wv.w, xv.w, fmaf(wv.z, xv.z, fmaf(wv.y, xv.y, wv.x * xv.x)));
    }
    float a = acc[0];
#pragma unroll
    for (int i = 1; i < NI / 32; ++i) a += acc[i];
    a += __shfl_xor_sync(FULL_MASK, a, 1);
    a += __shfl_xor_sync(FULL_MASK, a, 2);
    a += __shfl_xor_sync(FULL_MASK, a, 4);
    if (ks == 0) ep(r, o, a + b[o]);
  }
  __syncthreads();
}

// out[r] = LayerNorm(x[r] + y[r]) * w + b over D = 32 elements, one warp per row (R <= 8).  Ends with a barrier.
__device__ __forceinline__ void ln(float* sc, int stride, int out_off, int x_off, int y_off, const float* w,
                                   const float* b, int R) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < R) {
    float* row = sc + (size_t)warp * stride;
    const float v = row[x_off + lane] + row[y_off + lane];
    const float mu = warp_sum(v) * (1.0f / D);
    const float dv = v - mu;
    const float rstd = 1.0f / sqrtf(warp_sum(dv * dv) * (1.0f / D) + 1e-5f);
    row[out_off + lane] = dv * rstd * w[lane] + b[lane];
  }
  __syncthreads();
}

// One new token (position p, the same for all rows) for R <= 8 rows, then the policy trunk + critic of the new state.
// Row r: environment slot row_e[r], id row_id[r] (user at p == 0, item otherwise), reward row_rew[r], position
// row_kn[r] in the next turn's row list (< 0: none).  sc: R rows of `stride` floats (layout of tracker_cta_dev.cuh,
// the last 128 floats of a row = h1 | h2).  ws / L: staged tracker weights; ts: staged trunk image.  kv_s: the rows'
// cached positions prefetched by the caller (as in tracker_cta_dev.cuh), or NULL.  h2_img: tensor-core head's tile
// images (actor_tc_dev.cuh); h2_out [n_env][64] otherwise.
// emb_ready: the caller already put the ids' embedding rows at sc[r * stride + EMB_OFF .. + 32) (and synchronised).
// before_trunk(): called by every thread after the decoder stage (the caller's last chance to publish row_kn).
constexpr int EMB_OFF = 100;   // Y + 4
template <class H2Store, class Hook>
__device__ __forceinline__ void token_and_trunk(const cirs_tracker_weights& W, const Layout& L, const float* ws,
                                                const float* ts, int n_env, int R, int p, const int* row_e,
                                                const int* row_id, const float* row_rew, const int* row_kn,
                                                float* __restrict__ kcache, float* __restrict__ vcache, float* sc,
                                                int stride, float* __restrict__ cur_state, int traj_len,
                                                float* __restrict__ traj_obs, float* __restrict__ traj_obs_next,
                                                const float* kv_s, int kv_ld, float* __restrict__ value_out,
                                                H2Store h2_store, bool emb_ready, Hook before_trunk,
                                                long long* tq = nullptr) {
  int tqi = 0;
  auto stamp = [&]() {
    if (tq && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); tq[tqi] = t_; }
    ++tqi;
  };
  stamp();
  const int nh = W.nhead, dh = D / nh, S = W.dim_state, max_len = W.max_len;
  // per-row layout of tracker_cta_dev.cuh for d = 32, d_hid = 128: X | X2 | YB | Y (64) | QKV (96) | HID (128) | PROB
  constexpr int X = 0, X2 = 32, YB = 64, Y = 96, QKV = 160, HID = 256, PROB = 384;
  const int H1 = stride - 128, H2 = stride - 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float sq = 5.656854249492380f;   // sqrt(32)

  // ---- token input: the id's embedding row -> Y + 4 (16-byte aligned), then the gate / user projection
  static_assert(EMB_OFF == Y + 4, "embedding row offset");
  if (!emb_ready) {
    if (tid < R * 8) {
      const int r = tid >> 3, c4 = tid & 7;
      const float* src = (p == 0 ? W.emb_user : W.emb_item) + (size_t)row_id[r] * D + 4 * c4;
      *reinterpret_cast<float4*>(sc + (size_t)r * stride + Y + 4 + 4 * c4) = __ldg(reinterpret_cast<const float4*>(src));
    }
    __syncthreads();
  }
  if (p == 0) {
    mv<D, D>(ws + L.user_w, ws + L.user_b, sc + Y + 4, stride, R, [&](int r, int o, float a) {
      sc[(size_t)r * stride + X + o] = a * sq + __ldg(W.pe + o);
    });
  } else {
    const float* w0 = ws + L.gate_w0;
    const float* pe = W.pe + (size_t)p * D;
    mv<D, D>(ws + L.gate_w, ws + L.gate_b, sc + Y + 4, stride, R, [&](int r, int o, float a) {
      float* row = sc + (size_t)r * stride;
      const float g = 1.f / (1.f + expf(-fmaf(w0[o], row_rew[r], a)));
      row[X + o] = (g * row[Y + 4 + o]) * sq + __ldg(pe + o);
    });
  }
  stamp();

  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < W.nlayers; ++l) {
    const Layer& Z = L.layer[l];
    mv<D, 3 * D>(ws + Z.in_w, ws + Z.in_b, sc + X, stride, R, [&](int r, int o, float a) {
      sc[(size_t)r * stride + QKV + o] = a;
      if (o >= D) {   // this position's K / V -> cache (read again by later turns)
        const size_t base = (((size_t)l * n_env + row_e[r]) * max_len + p) * D;
        if (o < 2 * D) kcache[base + o - D] = a;
        else vcache[base + o - 2 * D] = a;
      }
    });
    stamp();
    // attention: one warp per (row, head); lane j owns cached position j
    for (int task = warp; task < R * nh; task += NT / 32) {
      const int r = task / nh, h = task - r * nh;
      float* row = sc + (size_t)r * stride;
      const float* q = row + QKV + h * dh;
      float* ph = row + PROB + h * max_len;
      const float* kc = kcache + ((size_t)l * n_env + row_e[r]) * max_len * D + h * dh;
      const float* vc = vcache + ((size_t)l * n_env + row_e[r]) * max_len * D + h * dh;
      int kvs = D;
      if (kv_s) {
        kc = kv_s + (size_t)((r * W.nlayers + l) * 2) * p * kv_ld + h * dh;
        vc = kc + (size_t)p * kv_ld;
        kvs = kv_ld;
      }
      float mx = -INFINITY;
      for (int j = lane; j <= p; j += 32) {
        const float* kr = (j == p) ? (row + QKV + D + h * dh) : (kc + (size_t)j * kvs);
        float a = 0.f;
        for (int c = 0; c < dh; c += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(kr + c);
          const float4 qv = *reinterpret_cast<const float4*>(q + c);
          a = fmaf(qv.x * scale, kv.x, a);
          a = fmaf(qv.y * scale, kv.y, a);
          a = fmaf(qv.z * scale, kv.z, a);
          a = fmaf(qv.w * scale, kv.w, a);
        }
        ph[j] = a;
        mx = fmaxf(mx, a);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j <= p; j += 32) {
        const float ex = expf(ph[j] - mx);
        ph[j] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      __syncwarp();
      // o[c] = sum_j prob[j] V[j][c]: lane = (position group g, channel c); groups summed by xor-shuffles
      const int c = lane % dh, g = lane / dh, G = 32 / dh;
      float a = 0.f;
      for (int j = g; j < p; j += G) a = fmaf(ph[j], vc[(size_t)j * kvs + c], a);
      for (int off = dh; off < 32; off <<= 1) a += __shfl_xor_sync(FULL_MASK, a, off);
      if (g == 0) row[HID + h * dh + c] = fmaf(ph[p], row[QKV + 2 * D + h * dh + c], a) * inv;
    }
    __syncthreads();
    stamp();
    mv<D, D>(ws + Z.out_w, ws + Z.out_b, sc + HID, stride, R,
             [&](int r, int o, float a) { sc[(size_t)r * stride + YB + o] = a; });
    stamp();
    ln(sc, stride, X2, X, YB, ws + Z.n1_w, ws + Z.n1_b, R);
    stamp();
    mv<D, DHID>(ws + Z.l1_w, ws + Z.l1_b, sc + X2, stride, R,
                [&](int r, int o, float a) { sc[(size_t)r * stride + HID + o] = fmaxf(a, 0.f); });
    stamp();
    mv<DHID, D>(ws + Z.l2_w, ws + Z.l2_b, sc + HID, stride, R,
                [&](int r, int o, float a) { sc[(size_t)r * stride + YB + o] = a; });
    stamp();
    ln(sc, stride, X, X2, YB, ws + Z.n2_w, ws + Z.n2_b, R);
    stamp();
  }
  // decoder -> state (kept in YB, zero beyond dim_state, for the trunk), buffers
  mv<D, 32>(ws + L.dec_w, ws + L.dec_b, sc + X, stride, R, [&](int r, int o, float a) {
    sc[(size_t)r * stride + YB + o] = a;
    if (o < S) {
      const int e = row_e[r];
      if (cur_state) cur_state[(size_t)e * S + o] = a;
      if (traj_obs && p < traj_len) traj_obs[((size_t)e * traj_len + p) * S + o] = a;
      if (traj_obs_next && p >= 1 && p - 1 < traj_len) traj_obs_next[((size_t)e * traj_len + p - 1) * S + o] = a;
    }
  });
  stamp();
  stamp();
  before_trunk();
  // ---- policy trunk + critic of the new state (core/policy/ppo.py:122-126: preprocess Net, Critic head)
  mv<32, HIDP>(ts + TR_W1, ts + TR_B1, sc + YB, stride, R,
               [&](int r, int o, float a) { sc[(size_t)r * stride + H1 + o] = fmaxf(a, 0.f); });
  mv<HIDP, HIDP>(ts + TR_W2, ts + TR_B2, sc + H1, stride, R, [&](int r, int o, float a) {
    a = fmaxf(a, 0.f);
    sc[(size_t)r * stride + H2 + o] = a;
    h2_store(r, o, a);
  });
  if (value_out && warp < R) {
    const float* h2 = sc + (size_t)warp * stride + H2;
    const float v = warp_sum(fmaf(h2[lane], ts[TR_WV + lane], h2[lane + 32] * ts[TR_WV + lane + 32]));
    if (lane == 0) value_out[row_e[warp]] = v + ts[TR_BV];
  }
  __syncthreads();
}

}  // namespace cirs_tfast
