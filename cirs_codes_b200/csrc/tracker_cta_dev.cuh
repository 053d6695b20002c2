// K2 for the persistent rollout kernel, CTA-cooperative: the new token of up to RB environments is carried through
// the tracker by ALL threads of a CTA at once -- one thread per (row, output) of every matrix-vector stage, one warp per
// (row, head) of the attention, one warp per row for LayerNorm.
//
// Why: a turn of the rollout is a chain of ~14 dependent stages per environment.  With one warp (or a group of warps)
// per environment each stage is a 32 .. 128 long dependent FMA chain behind shuffles and named barriers, ~1.5 us per
// stage = 23 us per token whatever the number of environments (measured, profiles/r1_*); yet a CTA only ever holds 1-4
// environments of the 512 (148 CTAs), so 7 of its 8 warps idle.  Here the R rows of a CTA advance together: a stage is
// R x n_out independent outputs spread over the 256 threads (conflict-free weight reads: consecutive threads read
// consecutive outputs of the k-major weights, the input vector is a broadcast), one __syncthreads per stage.
// Same arithmetic as tracker_token_warp (tracker_dev.cuh): bias first, inputs ascending, fmaf -- FP32 rounding level
// agreement with the stand-alone kernel (the tests' 1e-6 bar between the fused and the generic collect).
#pragma once
#include "tracker_dev.cuh"

namespace cirs_tracker {

constexpr int CTA_NT = 256;
constexpr int CTA_RB = 8;   // rows per cooperative pass (one warp per row in the warp-level steps)

// per-row scratch of the cooperative token step (floats)
__host__ __device__ inline int cta_row_floats(const cirs_tracker_weights& W) {
  const int ldd = trk_up32(W.d), ld3 = trk_up32(3 * W.d), ldh = trk_up32(W.d_hid);
  const int win = trk_up32(W.d_user_in > W.d_item_in + 1 ? W.d_user_in : W.d_item_in + 1);
  return 3 * ldd + win + ld3 + (ldh > ldd ? ldh : ldd) + trk_up32(W.nhead * W.max_len);
}

// out[r][o] = act(b[o] + sum_i Wt[i][o] in[r][i]),  r < R, o < n_out;  act: 0 identity, 1 relu, 2 sigmoid
template <bool SM>
__device__ __forceinline__ void mv_rows(const float* __restrict__ Wt, const float* __restrict__ b, int ldo, int n_in,
                                        int n_out, const float* in, float* out, int row_stride, int in_off,
                                        int out_off, int R, int act) {
  for (int idx = threadIdx.x; idx < R * n_out; idx += CTA_NT) {
    const int r = idx / n_out, o = idx - r * n_out;
    const float* x = in + (size_t)r * row_stride + in_off;
    const float* w = Wt + o;
    float a = ldw<SM>(b + o);
    int i = 0;
    // 16 inputs per batch: every load of the batch is issued before the dependent FMA chain starts (one shared-memory
    // latency per 16 inputs instead of per 8; measured on B200 with scratch/mv_bench.cu)
    for (; i + 16 <= n_in; i += 16) {
      float4 xv[4];
      float wv[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = *reinterpret_cast<const float4*>(x + i + 4 * u);
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = ldw<SM>(w + (size_t)(i + u) * ldo);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a = fmaf(wv[4 * u], xv[u].x, a);
        a = fmaf(wv[4 * u + 1], xv[u].y, a);
        a = fmaf(wv[4 * u + 2], xv[u].z, a);
        a = fmaf(wv[4 * u + 3], xv[u].w, a);
      }
    }
    for (; i < n_in; ++i) a = fmaf(ldw<SM>(w + (size_t)i * ldo), x[i], a);
    if (act == 1) a = fmaxf(a, 0.f);
    else if (act == 2) a = 1.f / (1.f + expf(-a));
    out[(size_t)r * row_stride + out_off + o] = a;
  }
  __syncthreads();
}

// the same product with the inputs split into `parts` ranges handled by different threads (part-major thread order, so
// the lanes of a warp still read consecutive outputs of one weight row): partial sums -> part_buf[(part * R + r) * n_out + o]
// (bias in part 0), then one pass adds them in part order.  Shortens the dependent chain of the 128-input FFN product.
template <bool SM>
__device__ __forceinline__ void mv_rows_split(const float* __restrict__ Wt, const float* __restrict__ b, int ldo,
                                              int n_in, int n_out, const float* in, float* out, int row_stride,
                                              int in_off, int out_off, int R, int parts, float* part_buf) {
  const int per = n_in / parts;   // n_in % (4 * parts) == 0 is the caller's business
  for (int idx = threadIdx.x; idx < parts * R * n_out; idx += CTA_NT) {
    const int part = idx / (R * n_out), rem = idx - part * (R * n_out);
    const int r = rem / n_out, o = rem - r * n_out;
    const float* x = in + (size_t)r * row_stride + in_off + part * per;
    const float* w = Wt + o + (size_t)part * per * ldo;
    float a = part == 0 ? ldw<SM>(b + o) : 0.f;
    int i = 0;
    for (; i + 16 <= per; i += 16) {
      float4 xv[4];
      float wv[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = *reinterpret_cast<const float4*>(x + i + 4 * u);
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = ldw<SM>(w + (size_t)(i + u) * ldo);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a = fmaf(wv[4 * u], xv[u].x, a);
        a = fmaf(wv[4 * u + 1], xv[u].y, a);
        a = fmaf(wv[4 * u + 2], xv[u].z, a);
        a = fmaf(wv[4 * u + 3], xv[u].w, a);
      }
    }
    for (; i < per; ++i) a = fmaf(ldw<SM>(w + (size_t)i * ldo), x[i], a);
    part_buf[idx] = a;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * n_out; idx += CTA_NT) {
    const int r = idx / n_out, o = idx - r * n_out;
    float a = part_buf[idx];
    for (int q = 1; q < parts; ++q) a += part_buf[q * R * n_out + idx];
    out[(size_t)r * row_stride + out_off + o] = a;
  }
  __syncthreads();
}

// out[r] = LayerNorm(xin[r] + yin[r]) * w + b, one warp per row (R <= 8 warps)
template <bool SM>
__device__ __forceinline__ void ln_rows(float* sc, int row_stride, int out_off, int x_off, int y_off,
                                        const float* __restrict__ w, const float* __restrict__ b, int d, int R) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < R) {
    float* row = sc + (size_t)warp * row_stride;
    add_layernorm_to<SM>(row + out_off, row + x_off, row + y_off, w, b, d, lane);
  }
  __syncthreads();
}

// One new token for R <= CTA_RB rows.  Row r: environment slot row_e[r], sequence position p (the same for all rows of
// a call: 0 = user token, t >= 1 = action token), id row_id[r] (user / item), reward row_rew[r].  sc: R rows of
// ``stride`` >= cta_row_floats floats of shared memory.  On return sc[r * stride + state_off .. + dim_state) holds the decoded state of row r
// (state_off is returned).  Every thread of the CTA calls it.
template <bool SM>
__device__ __forceinline__ int tracker_token_cta(const cirs_tracker_weights& W, int n_env, int R, int p, const int* row_e,
                                                 const int* row_id, const float* row_rew, float* __restrict__ kcache,
                                                 float* __restrict__ vcache, float* sc, int stride,
                                                 float* __restrict__ cur_state,
                                                 int traj_len, float* __restrict__ traj_obs,
                                                 float* __restrict__ traj_obs_next, long long* tq = nullptr,
                                                 const float* kv_s = nullptr, int kv_ld = 0,
                                                 float* part_buf = nullptr) {
  // kv_s (optional): the cached positions 0 .. p-1 of the R rows prefetched into shared memory by the caller,
  //   kv_s[(((r * nlayers + l) * 2 + {0: K, 1: V}) * p + j) * kv_ld + c]   (kv_ld = d + 4: conflict-free float4 rows)
  // part_buf (optional): CTA_RB * 128 floats for the split FFN product
  // tq (optional, one CTA): %globaltimer stamps at the stage boundaries (profiling aid, scratch/host_profile.py)
  int tqi = 0;
  auto stamp = [&]() {
    if (tq && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); tq[tqi] = t_; }
    ++tqi;
  };
  stamp();
  const int d = W.d, nh = W.nhead, dh = d / nh, dhid = W.d_hid, S = W.dim_state;
  const int ldd = trk_up32(d), ld3 = trk_up32(3 * d), ldh = trk_up32(dhid), lds = trk_up32(S);
  const int win = trk_up32(W.d_user_in > W.d_item_in + 1 ? W.d_user_in : W.d_item_in + 1);
  // per-row layout (stride >= cta_row_floats(W): the caller may append its own per-row scratch)
  const int X = 0, X2 = ldd, YB = 2 * ldd, Y = 3 * ldd, QKV = Y + win, HID = QKV + ld3, PROB = HID + max(ldh, ldd);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float sq = sqrtf((float)d);

  // ---- token inputs
  if (p == 0) {
    const int n_in = W.d_user_in;
    for (int i = tid; i < R * n_in; i += CTA_NT) {
      const int r = i / n_in, c = i - r * n_in;
      sc[(size_t)r * stride + Y + c] = __ldg(W.emb_user + (size_t)row_id[r] * d + c);
    }
    __syncthreads();
    mv_rows<SM>(W.user_wt, W.user_b, ldd, n_in, d, sc, sc, stride, Y, YB, R, 0);
    for (int i = tid; i < R * d; i += CTA_NT) {
      const int r = i / d, c = i - r * d;
      float* row = sc + (size_t)r * stride;
      row[X + c] = row[YB + c] * sq + __ldg(W.pe + c);
    }
  } else {
    const int n_in = W.d_item_in;
    for (int i = tid; i < R * (1 + n_in); i += CTA_NT) {
      const int r = i / (1 + n_in), c = i - r * (1 + n_in);
      sc[(size_t)r * stride + Y + c] = c == 0 ? row_rew[r] : __ldg(W.emb_item + (size_t)row_id[r] * d + c - 1);
    }
    __syncthreads();
    mv_rows<SM>(W.gate_wt, W.gate_b, ldd, 1 + n_in, d, sc, sc, stride, Y, YB, R, 2);
    for (int i = tid; i < R * d; i += CTA_NT) {
      const int r = i / d, c = i - r * d;
      float* row = sc + (size_t)r * stride;
      row[X + c] = (row[YB + c] * row[Y + 1 + c]) * sq + __ldg(W.pe + (size_t)p * d + c);
    }
  }
  __syncthreads();
  stamp();

  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < W.nlayers; ++l) {
    const cirs_encoder_layer& L = W.layer[l];
    mv_rows<SM>(L.in_wt, L.in_b, ld3, d, 3 * d, sc, sc, stride, X, QKV, R, 0);
    stamp();
    // this position's K / V -> cache (read again by later turns)
    for (int i = tid; i < R * d; i += CTA_NT) {
      const int r = i / d, c = i - r * d;
      const float* row = sc + (size_t)r * stride;
      const size_t base = (((size_t)l * n_env + row_e[r]) * W.max_len + p) * d + c;
      kcache[base] = row[QKV + d + c];
      vcache[base] = row[QKV + 2 * d + c];
    }
    // attention: one warp per (row, head); lane j owns cached position j
    for (int task = warp; task < R * nh; task += CTA_NT / 32) {
      const int r = task / nh, h = task - r * nh;
      float* row = sc + (size_t)r * stride;
      const float* q = row + QKV + h * dh;
      float* ph = row + PROB + h * W.max_len;
      const float* kc = kcache + ((size_t)l * n_env + row_e[r]) * W.max_len * d + h * dh;
      const float* vc = vcache + ((size_t)l * n_env + row_e[r]) * W.max_len * d + h * dh;
      int kvs = d;   // row stride of the cached K / V rows
      if (kv_s) {
        kc = kv_s + (size_t)((r * W.nlayers + l) * 2) * p * kv_ld + h * dh;
        vc = kc + (size_t)p * kv_ld;
        kvs = kv_ld;
      }
      float mx = -INFINITY;
      for (int j = lane; j <= p; j += 32) {
        const float* kr = (j == p) ? (row + QKV + d + h * dh) : (kc + (size_t)j * kvs);
        float a = 0.f;
        if ((dh & 3) == 0 && (d & 3) == 0) {   // 16-byte reads: conflict-free on the padded shared rows
#pragma unroll 2
          for (int c = 0; c < dh; c += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + c);
            const float4 qv = *reinterpret_cast<const float4*>(q + c);
            a = fmaf(qv.x * scale, kv.x, a);
            a = fmaf(qv.y * scale, kv.y, a);
            a = fmaf(qv.z * scale, kv.z, a);
            a = fmaf(qv.w * scale, kv.w, a);
          }
        } else {
#pragma unroll 4
          for (int c = 0; c < dh; ++c) a = fmaf(q[c] * scale, kr[c], a);
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
      for (int j = lane; j <= p; j += 32) ph[j] *= inv;
      __syncwarp();
      for (int c = lane; c < dh; c += 32) {
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < p; ++j) a = fmaf(ph[j], vc[(size_t)j * kvs + c], a);
        a = fmaf(ph[p], row[QKV + 2 * d + h * dh + c], a);
        row[HID + h * dh + c] = a;
      }
    }
    __syncthreads();
    stamp();
    mv_rows<SM>(L.out_wt, L.out_b, ldd, d, d, sc, sc, stride, HID, YB, R, 0);
    stamp();
    ln_rows<SM>(sc, stride, X2, X, YB, L.n1_w, L.n1_b, d, R);
    stamp();
    mv_rows<SM>(L.l1_wt, L.l1_b, ldh, d, dhid, sc, sc, stride, X2, HID, R, 1);
    stamp();
    {
      const int parts = (part_buf && d <= 32 && dhid % 64 == 0) ? 4 : ((part_buf && d <= 64 && dhid % 32 == 0) ? 2 : 1);
      if (parts > 1) mv_rows_split<SM>(L.l2_wt, L.l2_b, ldd, dhid, d, sc, sc, stride, HID, YB, R, parts, part_buf);
      else mv_rows<SM>(L.l2_wt, L.l2_b, ldd, dhid, d, sc, sc, stride, HID, YB, R, 0);
    }
    stamp();
    ln_rows<SM>(sc, stride, X, X2, YB, L.n2_w, L.n2_b, d, R);
    stamp();
  }
  // decoder -> state (kept in YB for the caller's trunk), buffers
  mv_rows<SM>(W.dec_wt, W.dec_b, lds, d, S, sc, sc, stride, X, YB, R, 0);
  for (int i = tid; i < R * S; i += CTA_NT) {
    const int r = i / S, c = i - r * S, e = row_e[r];
    const float v = sc[(size_t)r * stride + YB + c];
    if (cur_state) cur_state[(size_t)e * S + c] = v;
    if (traj_obs && p < traj_len) traj_obs[((size_t)e * traj_len + p) * S + c] = v;
    if (traj_obs_next && p >= 1 && p - 1 < traj_len) traj_obs_next[((size_t)e * traj_len + p - 1) * S + c] = v;
  }
  __syncthreads();
  stamp();
  return YB;
}

}  // namespace cirs_tracker
