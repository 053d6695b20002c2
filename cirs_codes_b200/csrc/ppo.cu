// K5: one PPO minibatch -- forward, losses and backward of the actor / critic heads (core/policy/ppo.py:181-220).
//
//   forward   trunk (20 -> 64 -> 64, ReLU) on the minibatch's observations, critic value, logits = h2 W3t + b3
//             (FP32 tile GEMM, gemm.cuh), per-row softmax statistics, Categorical log_prob / entropy
//             (log o clamp on the renormalised softmax), clipped surrogate, clipped value loss.
//   backward  d logits is never materialised: the two GEMMs that consume it (dW3t = h2^T dl, dh2 = dl W3) rebuild
//             each element from the stored logits and the row statistics while staging their operand tiles.
//             The trunk backward runs on the same tile GEMM; d loss / d obs is scattered to the buffer slots of the
//             minibatch (the upstream gradient of the tracker's training pass).
// Tie semantics follow torch: min / max route half of the gradient to each argument when they are equal, clamp
// passes the gradient on its closed interval.
#include "gemm.cuh"
#include "head_tc.cuh"
#include "../../include/cirs_b200.h"

namespace {
using namespace cirs;
constexpr int HID = CIRS_HIDDEN;

struct Workspace {
  float *h1, *h2, *value, *logits, *dh2, *dz2, *dz1, *rowm, *rinvz, *coef, *rowG, *dv, *terms;
  int32_t* acta;
  // tensor-core head (head_tc.cu): per-row partials over the catalogue splits
  float *pm, *ps, *la, *ent_part, *dh2_part, *w3img, *h2img;
};
constexpr int TC_SPLIT = cirs_head_tc::MAX_SPLIT;

__host__ __device__ inline int64_t align64(int64_t x) { return (x + 63) & ~(int64_t)63; }

Workspace carve(void* base, int64_t n, int64_t ldA) {
  float* p = reinterpret_cast<float*>(base);
  Workspace w;
  auto take = [&](int64_t cnt) { float* r = p; p += align64(cnt); return r; };
  w.h1 = take(n * HID); w.h2 = take(n * HID); w.value = take(n); w.logits = take(n * ldA);
  w.dh2 = take(n * HID); w.dz2 = take(n * HID); w.dz1 = take(n * HID);
  w.rowm = take(n); w.rinvz = take(n); w.coef = take(n); w.rowG = take(n); w.dv = take(n); w.terms = take(n * 4);
  w.acta = reinterpret_cast<int32_t*>(take(n));
  w.pm = take(n * TC_SPLIT); w.ps = take(n * TC_SPLIT); w.la = take(n); w.ent_part = take(n * TC_SPLIT);
  w.dh2_part = take(n * TC_SPLIT * HID);
  w.w3img = take(cirs_head_tc::head_tc_image_floats(ldA < 128 ? 128 : ldA));
  w.h2img = take(cirs_head_tc::head_tc_h2_image_floats(n));
  return w;
}

// ---- trunk forward on gathered observation rows: h1, h2 (row-major) and the critic value.
// 32 rows per CTA, thread = (row, group of 8 outputs); the weights are read as warp-uniform float4 (one broadcast load
// per 4 FMAs).  Accumulation order (bias first, k ascending, fmaf) is the one of actor_trunk_warp / actor_head_body,
// so all three produce bit-identical h2.
__global__ void __launch_bounds__(256)
trunk_fwd_kernel(cirs_policy_weights W, int n, const int32_t* __restrict__ idx, const float* __restrict__ obs,
                 float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ value) {
  constexpr int R = 32;
  __shared__ float s_in[R][33];
  __shared__ float hT[HID][R + 1];
  __shared__ __align__(16) float s_w1[32 * HID];    // W1t [dim_state <= 32][64]
  __shared__ __align__(16) float s_w2[HID * HID];   // W2t [64][64]
  const int tid = threadIdx.x, r0 = blockIdx.x * R, S = W.dim_state;
  // stage the trunk's weights once per CTA: every load of a thread is in flight at the same time (one L2 round trip)
  // instead of a dependent L1-miss per k-step of the loops below
  for (int i = tid; i < S * HID / 4; i += 256) reinterpret_cast<float4*>(s_w1)[i] = __ldg(reinterpret_cast<const float4*>(W.w1t) + i);
  for (int i = tid; i < HID * HID / 4; i += 256) reinterpret_cast<float4*>(s_w2)[i] = __ldg(reinterpret_cast<const float4*>(W.w2t) + i);
  for (int i = tid; i < R * S; i += 256) {
    const int r = i / S, c = i % S;
    s_in[r][c] = (r0 + r < n) ? obs[(int64_t)(idx ? idx[r0 + r] : r0 + r) * S + c] : 0.f;
  }
  __syncthreads();
  const int row = tid % R, og = tid / R;   // og: outputs og*8 .. og*8+7 (uniform per warp)
  const bool ok = r0 + row < n;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(W.b1 + og * 8 + j);
#pragma unroll 4
  for (int k = 0; k < S; ++k) {
    const float x = s_in[row][k];
    const float4 w0 = *reinterpret_cast<const float4*>(s_w1 + k * HID + og * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(s_w1 + k * HID + og * 8 + 4);
    acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]); acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
    acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]); acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] = fmaxf(acc[j], 0.f);
    hT[og * 8 + j][row] = acc[j];
  }
  if (ok) {
    float4* d = reinterpret_cast<float4*>(h1 + (int64_t)(r0 + row) * HID + og * 8);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(W.b2 + og * 8 + j);
#pragma unroll 8
  for (int k = 0; k < HID; ++k) {
    const float x = hT[k][row];
    const float4 w0 = *reinterpret_cast<const float4*>(s_w2 + k * HID + og * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(s_w2 + k * HID + og * 8 + 4);
    acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]); acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
    acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]); acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] = fmaxf(acc[j], 0.f);
    hT[og * 8 + j][row] = acc[j];
  }
  if (ok) {
    float4* d = reinterpret_cast<float4*>(h2 + (int64_t)(r0 + row) * HID + og * 8);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  if (tid < R && r0 + tid < n) {
    float v = __ldg(W.bv);
    for (int k = 0; k < HID; ++k) v = fmaf(hT[k][tid], __ldg(W.wv + k), v);
    value[r0 + tid] = v;
  }
}

// ---- trunk + critic backward in ONE kernel (replaces critic grad, dz2, dW2, dz1, dW1, d obs: six launches).
// 32 rows per CTA.  dh2: [n, 64] when n_split == 0, else the tensor-core pass's partials [n_split, n, 64].
// Parameter gradients are accumulated with atomicAdd into the (zeroed) flat gradient buffer; d_obs rows are stored
// at the minibatch's buffer slots (the upstream gradient of the tracker's training pass).
__global__ void __launch_bounds__(256)
trunk_bwd_kernel(cirs_policy_weights W, cirs_policy_weights G, int n, const int32_t* __restrict__ idx,
                 const float* __restrict__ obs, const float* __restrict__ dh2, int n_split,
                 const float* __restrict__ dv, const float* __restrict__ h1, const float* __restrict__ h2,
                 float* __restrict__ d_obs) {
  constexpr int R = 16, LD = HID + 1;   // 16 rows per CTA: ~100 CTAs per 1.6k-row minibatch, every phase is short
  __shared__ float s_dz2[R][LD], s_h1[R][LD], s_t[R][LD];   // s_t: h2, later dz1
  __shared__ float s_obs[R][33], s_dv[R];
  __shared__ float s_w1c[HID * 33];                 // W1t transposed: [64][dim_state (+ pad)], conflict-free for d obs
  __shared__ __align__(16) float s_w2[HID * HID];   // W2t [64][64]
  static_assert(256 % R == 0 && HID % (256 / R) == 0, "row tile");
  const int tid = threadIdx.x, r0 = blockIdx.x * R, S = W.dim_state;
  for (int i = tid; i < S * HID; i += 256) s_w1c[(i % HID) * 33 + i / HID] = __ldg(W.w1t + i);
  for (int i = tid; i < HID * HID / 4; i += 256) reinterpret_cast<float4*>(s_w2)[i] = __ldg(reinterpret_cast<const float4*>(W.w2t) + i);
  {
    // each thread owns 8 (row, column) elements; the split partials of all 8 are loaded together so that their
    // latencies overlap (the partials are [n_split, n, 64])
    constexpr int E = R * HID / 256;
    float d[E];
    int64_t off[E];
    bool okv[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = tid + e * 256, gr = r0 + i / HID;
      okv[e] = gr < n;
      off[e] = (int64_t)gr * HID + (i % HID);
      d[e] = 0.f;
    }
    if (n_split > 0) {
      const int64_t stride = (int64_t)n * HID;
      int sp = 0;
      for (; sp + 4 <= n_split; sp += 4) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (!okv[e]) continue;
          const float* q = dh2 + (int64_t)sp * stride + off[e];
          d[e] += (q[0] + q[stride]) + (q[2 * stride] + q[3 * stride]);
        }
      }
      for (; sp < n_split; ++sp) {
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (okv[e]) d[e] += dh2[(int64_t)sp * stride + off[e]];
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (okv[e]) d[e] = dh2[off[e]];
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = tid + e * 256, r = i / HID, c = i % HID;
      float z = 0.f, a1 = 0.f, a2 = 0.f;
      if (okv[e]) {
        a2 = h2[off[e]];
        a1 = h1[off[e]];
        z = a2 > 0.f ? d[e] + dv[r0 + r] * __ldg(W.wv + c) : 0.f;      // dz2 = (dh2 + dv wv) [h2 > 0]
      }
      s_dz2[r][c] = z; s_h1[r][c] = a1; s_t[r][c] = a2;
    }
  }
  for (int i = tid; i < R * S; i += 256) {
    const int r = i / S, c = i % S;
    s_obs[r][c] = (r0 + r < n) ? obs[(int64_t)idx[r0 + r] * S + c] : 0.f;
  }
  if (tid < R) s_dv[tid] = (r0 + tid < n) ? dv[r0 + tid] : 0.f;
  __syncthreads();
  // critic.last and b2 gradients
  if (tid < HID) {
    float gw = 0.f, gb = 0.f;
#pragma unroll 8
    for (int r = 0; r < R; ++r) { gw = fmaf(s_dv[r], s_t[r][tid], gw); gb += s_dz2[r][tid]; }
    atomicAdd(G.wv + tid, gw);
    atomicAdd(G.b2 + tid, gb);
  } else if (tid == HID) {
    float g = 0.f;
    for (int r = 0; r < R; ++r) g += s_dv[r];
    atomicAdd(G.bv, g);
  }
  // dW2t[k][c] += sum_r h1[r][k] dz2[r][c]:  thread = (k, 16 columns)
  {
    const int k = tid >> 2, cq = (tid & 3) * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const float x = s_h1[r][k];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(x, s_dz2[r][cq + j], acc[j]);
    }
    // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the atomic traffic of scalar adds -- every CTA
    // adds into the same 4096 addresses, so this phase is bound by the L2's atomic throughput
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      atomicAdd(reinterpret_cast<float4*>(G.w2t + (size_t)k * HID + cq + j),
                make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
  }
  __syncthreads();   // s_t (h2) is consumed
  // dz1[r][k] = (sum_c dz2[r][c] W2t[k][c]) [h1 > 0]:  thread = (row, KPT consecutive k)
  {
    constexpr int KPT = HID / (256 / R);
    const int r = tid % R, kg = (tid / R) * KPT;
    float acc[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int c4 = 0; c4 < HID / 4; ++c4) {
      const float z0 = s_dz2[r][4 * c4], z1 = s_dz2[r][4 * c4 + 1], z2 = s_dz2[r][4 * c4 + 2], z3 = s_dz2[r][4 * c4 + 3];
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(s_w2 + (kg + j) * HID + 4 * c4);
        acc[j] = fmaf(z0, w.x, acc[j]); acc[j] = fmaf(z1, w.y, acc[j]);
        acc[j] = fmaf(z2, w.z, acc[j]); acc[j] = fmaf(z3, w.w, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < KPT; ++j) s_t[r][kg + j] = s_h1[r][kg + j] > 0.f ? acc[j] : 0.f;
  }
  __syncthreads();
  // b1, dW1t[s][c] += sum_r obs[r][s] dz1[r][c]
  if (tid < HID) {
    float gb = 0.f;
    for (int r = 0; r < R; ++r) gb += s_t[r][tid];
    atomicAdd(G.b1 + tid, gb);
  }
  for (int o = tid; o < S * HID / 4; o += 256) {
    const int sI = o / (HID / 4), c = (o % (HID / 4)) * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int r = 0; r < R; ++r) {
      const float x = s_obs[r][sI];
      a0 = fmaf(x, s_t[r][c], a0); a1 = fmaf(x, s_t[r][c + 1], a1);
      a2 = fmaf(x, s_t[r][c + 2], a2); a3 = fmaf(x, s_t[r][c + 3], a3);
    }
    atomicAdd(reinterpret_cast<float4*>(G.w1t + (size_t)sI * HID + c), make_float4(a0, a1, a2, a3));
  }
  // d_obs[slot][s] = sum_c dz1[r][c] W1t[s][c]
  if (d_obs) {
    for (int o = tid; o < R * S; o += 256) {
      const int r = o / S, sI = o % S;
      if (r0 + r >= n) continue;
      float a = 0.f;
#pragma unroll 8
      for (int c = 0; c < HID; ++c) a = fmaf(s_t[r][c], s_w1c[c * 33 + sI], a);
      d_obs[(int64_t)idx[r0 + r] * S + sI] = a;
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];  // fixed order: deterministic
  return r;
}

// per-row tail shared by the FFMA and the tensor-core paths: Categorical.log_prob of the taken action, clipped
// surrogate, clipped value loss and the row's d loss / d logp, d loss / d value  (ppo.py:183-207)
__device__ __forceinline__ void row_finish(int r, int slot, int a, float mx, float invz, float la_logit, float ent,
                                           float rowG, const cirs_ppo_config& cfg, int n_global,
                                           const float* __restrict__ adv, const float* __restrict__ returns,
                                           const float* __restrict__ v_old, const float* __restrict__ logp_old,
                                           const double* __restrict__ adv_stat, const Workspace& ws) {
  const float pa = expf(la_logit - mx) * invz;
  const bool inr_a = pa >= CATEGORICAL_EPS && pa <= 1.0f - CATEGORICAL_EPS;
  const float logp = logf(fminf(fmaxf(pa, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS));
  const float inv_n = 1.0f / (float)n_global;
  // advantage normalisation with the minibatch's mean / unbiased std (ppo.py:185-186)
  float A = adv[slot];
  if (cfg.norm_adv) {
    const double cnt = adv_stat[0], mean = adv_stat[1] / cnt;
    const double var = (adv_stat[2] - cnt * mean * mean) / (cnt - 1.0);
    A = (float)(((double)A - mean) / sqrt(fmax(var, 0.0)));
  }
  const float ratio = expf(logp - logp_old[slot]);                               // ppo.py:187
  const float lo = 1.0f - cfg.eps_clip, hi = 1.0f + cfg.eps_clip;
  const float surr1 = ratio * A, surr2 = fminf(fmaxf(ratio, lo), hi) * A;        // ppo.py:189-190
  const float clip_i = -fminf(surr1, surr2);                                     // ppo.py:196
  const float g1 = surr1 < surr2 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
  const float g2 = surr2 < surr1 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
  const bool in_clip = ratio >= lo && ratio <= hi;
  const float dmin_dratio = g1 * A + (in_clip ? g2 * A : 0.f);
  ws.coef[r] = inr_a ? -dmin_dratio * ratio * inv_n : 0.f;                       // d loss / d logp_r
  // critic (ppo.py:199-207)
  const float v = ws.value[r], vo = v_old[slot], R = returns[slot];
  float vf_i, dvf;
  if (cfg.value_clip) {
    const float dvc = v - vo;
    const float v_clip = vo + fminf(fmaxf(dvc, -cfg.eps_clip), cfg.eps_clip);
    const float vf1 = (R - v) * (R - v), vf2 = (R - v_clip) * (R - v_clip);
    vf_i = fmaxf(vf1, vf2);
    const float w1 = vf1 > vf2 ? 1.f : (vf1 == vf2 ? 0.5f : 0.f);
    const float w2 = vf2 > vf1 ? 1.f : (vf1 == vf2 ? 0.5f : 0.f);
    const bool in_v = dvc >= -cfg.eps_clip && dvc <= cfg.eps_clip;
    dvf = w1 * (-2.f * (R - v)) + (in_v ? w2 * (-2.f * (R - v_clip)) : 0.f);
  } else {
    vf_i = (R - v) * (R - v);
    dvf = -2.f * (R - v);
  }
  ws.dv[r] = cfg.vf_coef * dvf * inv_n;
  ws.rowm[r] = mx;
  ws.rinvz[r] = invz;
  ws.rowG[r] = rowG;   // sum_j g_j p_j with g_j = -(log clamp(p_j) + [p_j in range])
  ws.acta[r] = a;
  ws.terms[4 * r] = clip_i;
  ws.terms[4 * r + 1] = vf_i;
  ws.terms[4 * r + 2] = ent;
}

// ---- per-row softmax statistics, losses and d loss / d logp, d loss / d value   (one CTA per row)
__global__ void __launch_bounds__(256)
row_loss_kernel(int nA, int64_t ldA, cirs_ppo_config cfg, int n_global, const int32_t* __restrict__ idx,
                const int32_t* __restrict__ act, const float* __restrict__ adv, const float* __restrict__ returns,
                const float* __restrict__ v_old, const float* __restrict__ logp_old,
                const double* __restrict__ adv_stat, Workspace ws) {
  __shared__ float sh[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* L = ws.logits + (int64_t)r * ldA;
  float mx = -INFINITY;
  for (int c = tid; c < nA; c += 256) mx = fmaxf(mx, L[c]);
  mx = block_reduce(mx, true, sh);
  float z = 0.f;
  for (int c = tid; c < nA; c += 256) z += expf(L[c] - mx);
  z = block_reduce(z, false, sh);
  const float invz = 1.0f / z;
  // Categorical(probs = softmax): probs <- p / sum(p); logits = log(clamp(probs, eps, 1 - eps))
  float ent = 0.f, pin = 0.f;
  for (int c = tid; c < nA; c += 256) {
    const float p = expf(L[c] - mx) * invz;
    const bool inr = p >= CATEGORICAL_EPS && p <= 1.0f - CATEGORICAL_EPS;
    const float pc = fminf(fmaxf(p, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
    ent -= p * logf(pc);
    if (inr) pin += p;
  }
  ent = block_reduce(ent, false, sh);
  pin = block_reduce(pin, false, sh);
  if (tid != 0) return;
  const int slot = idx[r];
  row_finish(r, slot, act[slot], mx, invz, L[act[slot]], ent, ent - pin, cfg, n_global, adv, returns, v_old, logp_old,
             adv_stat, ws);
}

// ---- tensor-core path: merge the per-split online-softmax partials of pass F, then the same per-row tail
__global__ void __launch_bounds__(128)
row_loss_tc_kernel(int n, int n_split, cirs_ppo_config cfg, int n_global, const int32_t* __restrict__ idx,
                   const int32_t* __restrict__ act, const float* __restrict__ adv, const float* __restrict__ returns,
                   const float* __restrict__ v_old, const float* __restrict__ logp_old,
                   const double* __restrict__ adv_stat, Workspace ws) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) {
    // rows up to the next multiple of 64: neutral statistics, so that pass B3 can bulk-copy whole 64-row slices
    if (r < ((n + 63) & ~63)) { ws.rowm[r] = 0.f; ws.rinvz[r] = 0.f; ws.coef[r] = 0.f; ws.acta[r] = -1; }
    return;
  }
  const float* pm = ws.pm + (int64_t)r * n_split;
  const float* ps = ws.ps + (int64_t)r * n_split;
  float mx = -INFINITY;
#pragma unroll 8
  for (int s = 0; s < n_split; ++s) mx = fmaxf(mx, pm[s]);
  float z = 0.f;
#pragma unroll 8
  for (int s = 0; s < n_split; ++s) z += ps[s] * expf(pm[s] - mx);
  const int slot = idx[r];
  // entropy (terms[4r+2]) and rowG are filled after pass B2 (ent_merge_kernel); rowG is only used when ent_coef != 0
  row_finish(r, slot, act[slot], mx, 1.0f / z, ws.la[r], 0.f, 0.f, cfg, n_global, adv, returns, v_old, logp_old,
             adv_stat, ws);
}

__global__ void ent_merge_kernel(int n, int n_split, const float* __restrict__ ent_part, float* __restrict__ terms) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float e = 0.f;
  for (int s = 0; s < n_split; ++s) e += ent_part[(int64_t)r * n_split + s];
  terms[4 * r + 2] = e;
}

// policy evaluation (process_fn): merge pass F's partials -> Categorical.log_prob of the stored action; scatter the
// critic value; outputs are indexed by buffer slot like the FFMA path (actor_combine_row, out_by_k = 0)
__global__ void eval_merge_kernel(int n, int n_split, const int32_t* __restrict__ idx, const float* __restrict__ pm,
                                  const float* __restrict__ ps, const float* __restrict__ la,
                                  const float* __restrict__ vtmp, float* __restrict__ value, float* __restrict__ logp,
                                  const int32_t* __restrict__ n_dev) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  if (r >= n) return;
  const int o = idx ? idx[r] : r;
  if (value) value[o] = vtmp[r];
  if (!logp) return;
  float mx = -INFINITY;
  for (int s = 0; s < n_split; ++s) mx = fmaxf(mx, pm[(int64_t)r * n_split + s]);
  float z = 0.f;
  for (int s = 0; s < n_split; ++s) z += ps[(int64_t)r * n_split + s] * expf(pm[(int64_t)r * n_split + s] - mx);
  float pa = expf(la[r] - mx) / z;
  pa = fminf(fmaxf(pa, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
  logp[o] = logf(pa);
}

// ---- continuous actor (ActorProb + Independent(Normal)): per-row head forward, losses and d loss / d z, one warp per
// row.  z = h2 W3t + b3, mu = max_action tanh(z), std = exp(sigma);  logp = sum_c Normal.log_prob(a_c)
// (core/policy/ppo.py:183-187 with dist_fn = Independent(Normal), CIRS-RL-taobao.py:228-232).
// Writes dz[r][32] = d loss / d z (zero padded) and dsg[r][32] = d loss / d sigma_param contributions of row r.
#define CIRS_LOG_SQRT_2PI 0.9189385332046727f
__global__ void __launch_bounds__(256)
gauss_row_kernel(cirs_policy_weights W, int n, cirs_ppo_config cfg, int n_global, const int32_t* __restrict__ idx,
                 const float* __restrict__ act, const float* __restrict__ adv, const float* __restrict__ returns,
                 const float* __restrict__ v_old, const float* __restrict__ logp_old,
                 const double* __restrict__ adv_stat, Workspace ws, float* __restrict__ dz, float* __restrict__ dsg) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  const int nA = W.n_action, slot = idx[r];
  const float* h2 = ws.h2 + (int64_t)r * HID;
  float lp = 0.f, en = 0.f, df = 0.f, var = 1.f, t = 0.f;
  if (lane < nA) {
    float z = __ldg(W.b3 + lane);
#pragma unroll 16
    for (int k = 0; k < HID; ++k) z = fmaf(h2[k], __ldg(W.w3t + (size_t)k * W.ld_action + lane), z);
    t = tanhf(z);
    const float mu = W.max_action * t, sg = expf(__ldg(W.sigma + lane));
    var = sg * sg;
    df = act[(int64_t)slot * nA + lane] - mu;
    lp = -(df * df) / (2.0f * var) - logf(sg) - CIRS_LOG_SQRT_2PI;
    en = 0.5f + CIRS_LOG_SQRT_2PI + logf(sg);   // Normal.entropy = 0.5 + 0.5 log(2 pi) + log(std)
  }
  const float logp = warp_sum(lp), ent = warp_sum(en);
  const float inv_n = 1.0f / (float)n_global;
  float A = adv[slot];
  if (cfg.norm_adv) {
    const double cnt = adv_stat[0], mean = adv_stat[1] / cnt;
    const double v2 = (adv_stat[2] - cnt * mean * mean) / (cnt - 1.0);
    A = (float)(((double)A - mean) / sqrt(fmax(v2, 0.0)));
  }
  const float ratio = expf(logp - logp_old[slot]);
  const float lo = 1.0f - cfg.eps_clip, hi = 1.0f + cfg.eps_clip;
  const float surr1 = ratio * A, surr2 = fminf(fmaxf(ratio, lo), hi) * A;
  const float clip_i = -fminf(surr1, surr2);
  const float g1 = surr1 < surr2 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
  const float g2 = surr2 < surr1 ? 1.f : (surr1 == surr2 ? 0.5f : 0.f);
  const bool in_clip = ratio >= lo && ratio <= hi;
  const float coef = -(g1 * A + (in_clip ? g2 * A : 0.f)) * ratio * inv_n;   // d loss / d logp_r
  // d logp / d mu_c = (a - mu) / var;  d mu / d z = max_action (1 - tanh^2);  d logp / d sigma_c = (a-mu)^2 / var - 1;
  // d (-ent_coef * mean entropy) / d sigma_c = -ent_coef / n
  float gz = 0.f, gs = 0.f;
  if (lane < nA) {
    gz = coef * (df / var) * W.max_action * (1.0f - t * t);
    gs = coef * ((df * df) / var - 1.0f) - cfg.ent_coef * inv_n;
  }
  dz[(int64_t)r * 32 + lane] = gz;
  dsg[(int64_t)r * 32 + lane] = gs;
  if (lane != 0) return;
  const float v = ws.value[r], vo = v_old[slot], R = returns[slot];
  float vf_i, dvf;
  if (cfg.value_clip) {
    const float dvc = v - vo;
    const float v_clip = vo + fminf(fmaxf(dvc, -cfg.eps_clip), cfg.eps_clip);
    const float vf1 = (R - v) * (R - v), vf2 = (R - v_clip) * (R - v_clip);
    vf_i = fmaxf(vf1, vf2);
    const float w1 = vf1 > vf2 ? 1.f : (vf1 == vf2 ? 0.5f : 0.f);
    const float w2 = vf2 > vf1 ? 1.f : (vf1 == vf2 ? 0.5f : 0.f);
    const bool in_v = dvc >= -cfg.eps_clip && dvc <= cfg.eps_clip;
    dvf = w1 * (-2.f * (R - v)) + (in_v ? w2 * (-2.f * (R - v_clip)) : 0.f);
  } else {
    vf_i = (R - v) * (R - v);
    dvf = -2.f * (R - v);
  }
  ws.dv[r] = cfg.vf_coef * dvf * inv_n;
  ws.terms[4 * r] = clip_i;
  ws.terms[4 * r + 1] = vf_i;
  ws.terms[4 * r + 2] = ent;
}

// out[c] = sum_r m[r][32 + c]-style column sums of an [n, 32] matrix (single CTA, fixed order: deterministic)
__global__ void __launch_bounds__(256) colsum32_kernel(int n, const float* __restrict__ m, int n_out, float* out) {
  __shared__ float sh[8][33];
  const int c = threadIdx.x & 31, part = threadIdx.x >> 5;
  float s = 0.f;
  for (int r = part; r < n; r += 8) s += m[(int64_t)r * 32 + c];
  sh[part][c] = s;
  __syncthreads();
  if (part == 0 && c < n_out) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += sh[i][c];
    out[c] = v;
  }
}

// ---- deterministic reduction of the per-row loss terms -> losses[4] = {loss, clip, vf, ent} / n_global
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(int n, int n_global, cirs_ppo_config cfg, const float* __restrict__ terms, float* losses,
                   const float* __restrict__ ent_part = nullptr, int n_split = 0) {
  // ent_part (tensor-core path): the rows' entropies still are per-split partials of pass B2, merged here
  __shared__ double sh[3][32];
  double a = 0, b = 0, c = 0;
  for (int r = threadIdx.x; r < n; r += 1024) {
    a += terms[4 * r]; b += terms[4 * r + 1];
    if (ent_part) {
      float e = 0.f;
      const float* ep = ent_part + (int64_t)r * n_split;
      int s = 0;
      for (; s + 8 <= n_split; s += 8) {   // eight loads in flight; the additions keep their left-to-right order
        const float v0 = ep[s], v1 = ep[s + 1], v2 = ep[s + 2], v3 = ep[s + 3], v4 = ep[s + 4], v5 = ep[s + 5],
                    v6 = ep[s + 6], v7 = ep[s + 7];
        e += v0; e += v1; e += v2; e += v3; e += v4; e += v5; e += v6; e += v7;
      }
      for (; s < n_split; ++s) e += ep[s];
      c += e;
    } else {
      c += terms[4 * r + 2];
    }
  }
  a = warp_sum_d(a); b = warp_sum_d(b); c = warp_sum_d(c);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = c = 0;
    for (int i = 0; i < 32; ++i) { a += sh[0][i]; b += sh[1][i]; c += sh[2][i]; }
    a /= n_global; b /= n_global; c /= n_global;
    losses[0] = (float)(a + cfg.vf_coef * b - cfg.ent_coef * c);   // ppo.py:211-212
    losses[1] = (float)a; losses[2] = (float)b; losses[3] = (float)c;
  }
}

// d loss / d logits[r][c], rebuilt on the fly:  coef_r (1[c == a_r] - p_rc)  +  ent_coef/n * p (log clamp(p) + inr + G_r)
struct DlCore {
  const float* logits; int64_t ld;
  const float *rowm, *rinvz, *coef, *rowG;
  const int32_t* acta;
  float ecn;  // ent_coef / n_global
  __device__ __forceinline__ float at(int r, int c) const {
    const float p = expf(__ldg(logits + (int64_t)r * ld + c) - __ldg(rowm + r)) * __ldg(rinvz + r);
    float v = __ldg(coef + r) * ((c == __ldg(acta + r) ? 1.f : 0.f) - p);
    if (ecn != 0.f) {
      const bool inr = p >= CATEGORICAL_EPS && p <= 1.0f - CATEGORICAL_EPS;
      const float pc = fminf(fmaxf(p, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
      v += ecn * p * (logf(pc) + (inr ? 1.f : 0.f) + __ldg(rowG + r));
    }
    return v;
  }
};
struct DlA {  // A(m = row, k = column): contiguous along k
  static constexpr bool INNER_IS_K = true;
  DlCore d;
  __device__ __forceinline__ float operator()(int m, int k) const { return d.at(m, k); }
};
struct DlB {  // B(k = row, n = column): contiguous along n
  static constexpr bool INNER_IS_K = false;
  DlCore d;
  __device__ __forceinline__ float operator()(int k, int n) const { return d.at(k, n); }
};

__global__ void __launch_bounds__(1024)
adv_stats_kernel(const int32_t* __restrict__ mb_off, const int32_t* __restrict__ idx, const float* __restrict__ adv,
                 double* __restrict__ stats, int n_rep_stride = 0) {
  // blockIdx.y = repeat: its permutation starts n_rep_stride entries further, its statistics 3 * gridDim.x doubles further
  __shared__ double sh[2][32];
  const int j = blockIdx.x, b = mb_off[j], e = mb_off[j + 1];
  idx += (size_t)blockIdx.y * n_rep_stride;
  stats += (size_t)blockIdx.y * gridDim.x * 3;
  double s = 0, ss = 0;
  for (int i = b + threadIdx.x; i < e; i += 1024) {
    const double a = adv[idx[i]];
    s += a; ss += a * a;
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = ss = 0;
    for (int i = 0; i < 32; ++i) { s += sh[0][i]; ss += sh[1][i]; }
    stats[3 * j] = (double)(e - b); stats[3 * j + 1] = s; stats[3 * j + 2] = ss;
  }
}

int split_for(int tiles, int K, int bk) {
  int s = (2 * 148 + tiles - 1) / tiles;
  const int max_s = (K + 4 * bk - 1) / (4 * bk);
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

}  // namespace

// cirs_policy_eval on the tensor cores (dispatched from actor.cu): trunk -> pass F -> merge.
namespace cirs_head_tc {
int64_t policy_eval_tc_workspace_bytes(int64_t n) {
  return (int64_t)sizeof(float) * (3 * align64(n * HID) / 1 + 2 * align64(n * MAX_SPLIT) + 2 * align64(n) +
                                   align64(head_tc_h2_image_floats(n))) + 256;
}
int64_t policy_eval_tc_image_bytes(int64_t ldA) {
  return (int64_t)sizeof(float) * align64(head_tc_image_floats(ldA < 128 ? 128 : ldA));
}
int policy_eval_tc(const cirs_policy_weights* w, int32_t n, const int32_t* row_idx, const float* obs,
                   const int32_t* act, float* value, float* logp, void* workspace, cudaStream_t st,
                   const int32_t* n_dev) {
  float* p = reinterpret_cast<float*>(workspace);
  auto take = [&](int64_t cnt) { float* r = p; p += align64(cnt); return r; };
  float *h1 = take((int64_t)n * HID), *h2 = take((int64_t)n * HID), *vtmp = take(n);
  float *pm = take((int64_t)n * MAX_SPLIT), *ps = take((int64_t)n * MAX_SPLIT), *la = take(n);
  float* himg = take(head_tc_h2_image_floats(n));
  float* img = take(head_tc_image_floats(w->ld_action));
  // trunk, h2 images and W3 images in one launch (head_tc_front); a value-only evaluation needs the trunk alone
  int rc = head_tc_front(w, n, row_idx, obs, h1, h2, vtmp, act ? himg : nullptr, act ? img : nullptr, st, n_dev);
  if (rc) return rc;
  int n_split = 0;
  if (act) {   // log-probs of the stored actions
    n_split = plan_split_f(n, w->n_action);
    HeadTc H{h2, n, w->w3t, w->ld_action, w->b3, w->n_action, img, himg};
    rc = head_tc_stats(H, row_idx, act, n_split, pm, ps, la, st, n_dev);
    if (rc) return rc;
  }
  CIRS_LAUNCH(eval_merge_kernel, (n + 255) / 256, 256, 0, st, n, n_split, row_idx, pm, ps, la, vtmp, value,
              act ? logp : nullptr, n_dev);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
}  // namespace cirs_head_tc

extern "C" int64_t cirs_ppo_workspace_bytes(int32_t n_rows, int32_t n_action) {
  const int64_t n = n_rows > 0 ? n_rows : 1, ldA = ((int64_t)n_action + 127) & ~127LL;
  int64_t cnt = 5 * align64(n * HID) + align64(n * ldA) + 7 * align64(n) + align64(4 * n) +
                3 * align64(n * TC_SPLIT) + align64(n) + align64(n * TC_SPLIT * HID) +
                align64(cirs_head_tc::head_tc_image_floats(ldA < 128 ? 128 : ldA)) +
                align64(cirs_head_tc::head_tc_h2_image_floats(n));
  return cnt * (int64_t)sizeof(float) + 256;
}

extern "C" int cirs_adv_stats(int32_t n_mb, const int32_t* mb_off, const int32_t* idx, const float* adv,
                              double* stats, void* stream) {
  if (n_mb < 0 || !mb_off || !idx || !adv || !stats) {
    cirs_set_error("cirs_adv_stats: null argument");
    return CIRS_ERR_ARG;
  }
  if (n_mb == 0) return CIRS_OK;
  CIRS_LAUNCH(adv_stats_kernel, n_mb, 1024, 0, (cudaStream_t)stream, mb_off, idx, adv, stats);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_ppo_minibatch(const cirs_policy_weights* w, const cirs_policy_weights* grads,
                                  const cirs_ppo_config* cfg, int32_t n, int32_t n_global, const int32_t* idx,
                                  const float* obs, const void* act_v, const float* adv, const float* returns,
                                  const float* v_old, const float* logp_old, const double* adv_stat, float* d_obs,
                                  float* losses, void* workspace, void* stream) {
  const int32_t* act = reinterpret_cast<const int32_t*>(act_v);
  if (!w || !grads || !cfg || !idx || !obs || !act || !adv || !returns || !v_old || !logp_old || !losses ||
      !workspace || n < 0 || n_global < n || (cfg->norm_adv && !adv_stat)) {
    cirs_set_error("cirs_ppo_minibatch: bad argument");
    return CIRS_ERR_ARG;
  }
  if (w->dim_state > 32 || !grads->flat) {
    cirs_set_error("cirs_ppo_minibatch: dim_state > 32 or grads->flat missing");
    return CIRS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nA = w->n_action, S = w->dim_state;
  const int64_t ldA = w->ld_action;
  cudaMemsetAsync(grads->flat, 0, sizeof(float) * grads->n_flat, st);
  if (n == 0) {
    cudaMemsetAsync(losses, 0, 4 * sizeof(float), st);
    return CIRS_OK;
  }
  const bool gauss = w->sigma != nullptr;
  if (gauss && (nA > 32 || !grads->sigma)) {
    cirs_set_error("cirs_ppo_minibatch: continuous actor needs n_action <= 32 and grads->sigma");
    return CIRS_ERR_ARG;
  }
  Workspace ws = carve(workspace, n, gauss ? 64 : ldA);
  bool tc = false;
  int tc_split = 0;

  // ---- forward
  const bool tc_path = !gauss && cfg->ent_coef == 0.f && cirs_head_tc::head_tc_enabled(n, nA, ldA);
  if (tc_path) {   // trunk + h2 images + W3 images (the weights changed in the last Adam step) in one launch
    int rc = cirs_head_tc::head_tc_front(w, n, idx, obs, ws.h1, ws.h2, ws.value, ws.h2img, ws.w3img, st);
    if (rc) return rc;
  } else {
    CIRS_LAUNCH(trunk_fwd_kernel, (n + 31) / 32, 256, 0, st, *w, n, idx, obs, ws.h1, ws.h2, ws.value);
    CIRS_CHECK_LAUNCH();
  }
  if (gauss) {
    float* dz = ws.logits;                      // [n, 32] d loss / d z
    float* dsg = ws.logits + (int64_t)n * 32;   // [n, 32] d loss / d sigma_param, per row
    CIRS_LAUNCH(gauss_row_kernel, (n + 7) / 8, 256, 0, st, *w, n, *cfg, n_global, idx,
                reinterpret_cast<const float*>(act_v), adv, returns, v_old, logp_old, adv_stat, ws, dz, dsg);
    CIRS_CHECK_LAUNCH();
    CIRS_LAUNCH(loss_reduce_kernel, 1, 1024, 0, st, n, n_global, *cfg, ws.terms, losses);
    CIRS_CHECK_LAUNCH();
    // dW3t[k][c] = sum_r h2[r][k] dz[r][c], db3[c] = sum_r dz[r][c]
    launch_gemm<64, 64, 16, 4>(ColMajorA{ws.h2, HID, nullptr}, RowMajorB{dz, 32, nullptr}, AtomicEp{grads->w3t, ldA},
                               HID, nA, n, split_for(1, n, 16), grads->b3, st, "gauss_dW3_gemm");
    CIRS_CHECK_LAUNCH();
    CIRS_LAUNCH(colsum32_kernel, 1, 256, 0, st, n, dsg, nA, grads->sigma);
    CIRS_CHECK_LAUNCH();
    // dh2[r][k] = sum_c dz[r][c] W3t[k][c]
    launch_gemm<64, 64, 16, 4>(RowMajorA{dz, 32, nullptr}, ColMajorB{w->w3t, ldA},
                               StoreEp{ws.dh2, HID, nullptr, 0, nullptr, nullptr, 0}, n, HID, nA, 1, nullptr, st,
                               "gauss_dh2_gemm");
    CIRS_CHECK_LAUNCH();
  } else if (tc_path) {
    // ---- actor head on the tensor cores: logits are recomputed per pass and never stored (head_tc.cu)
    tc = true;
    tc_split = cirs_head_tc::plan_split(n, nA);
    cirs_head_tc::HeadTc H{ws.h2, n, w->w3t, ldA, w->b3, nA, ws.w3img, ws.h2img};
    const int f_split = cirs_head_tc::plan_split_f(n, nA);
    int rc = cirs_head_tc::head_tc_stats(H, idx, act, f_split, ws.pm, ws.ps, ws.la, st);
    if (rc) return rc;
    CIRS_LAUNCH(row_loss_tc_kernel, (n + 127) / 128, 128, 0, st, n, f_split, *cfg, n_global, idx, act, adv, returns,
                v_old, logp_old, adv_stat, ws);
    CIRS_CHECK_LAUNCH();
    rc = cirs_head_tc::head_tc_dh2(H, ws.rowm, ws.rinvz, ws.coef, ws.acta, tc_split, ws.dh2_part, ws.ent_part, st);
    if (rc) return rc;
    CIRS_LAUNCH(loss_reduce_kernel, 1, 1024, 0, st, n, n_global, *cfg, ws.terms, losses, ws.ent_part, tc_split);
    CIRS_CHECK_LAUNCH();
    rc = cirs_head_tc::head_tc_dw3(H, ws.rowm, ws.rinvz, ws.coef, ws.acta, grads->w3t, grads->b3, st);
    if (rc) return rc;
  } else {
  launch_gemm<64, 128, 16, 8>(RowMajorA{ws.h2, HID, nullptr}, RowMajorB{w->w3t, ldA, nullptr},
                              StoreEp{ws.logits, ldA, w->b3, 0, nullptr, nullptr, 0}, n, nA, HID, 1, nullptr, st, "head_logits_gemm");
  CIRS_CHECK_LAUNCH();
  CIRS_LAUNCH(row_loss_kernel, n, 256, 0, st, nA, ldA, *cfg, n_global, idx, act, adv, returns, v_old, logp_old, adv_stat,
                                     ws);
  CIRS_CHECK_LAUNCH();
  CIRS_LAUNCH(loss_reduce_kernel, 1, 1024, 0, st, n, n_global, *cfg, ws.terms, losses);
  CIRS_CHECK_LAUNCH();

  // ---- backward through the actor head
  DlCore dl{ws.logits, ldA, ws.rowm, ws.rinvz, ws.coef, ws.rowG, ws.acta, cfg->ent_coef / (float)n_global};
  // dW3t[k][c] = sum_r h2[r][k] dl[r][c];  db3[c] = sum_r dl[r][c]      (M = 64, N = nA, K = n; split over rows)
  launch_gemm<64, 128, 16, 8>(ColMajorA{ws.h2, HID, nullptr}, DlB{dl}, AtomicEp{grads->w3t, ldA}, HID, nA, n,
                              split_for((nA + 127) / 128, n, 16), grads->b3, st, "head_dW3_gemm");
  CIRS_CHECK_LAUNCH();
  // dh2[r][k] = sum_c dl[r][c] W3t[k][c]                                  (M = n, N = 64, K = nA; split over columns)
  cudaMemsetAsync(ws.dh2, 0, sizeof(float) * (size_t)n * HID, st);
  launch_gemm<64, 64, 16, 4>(DlA{dl}, ColMajorB{w->w3t, ldA}, AtomicEp{ws.dh2, HID}, n, HID, nA,
                             split_for((n + 63) / 64, nA, 16), nullptr, st, "head_dh2_gemm");
  CIRS_CHECK_LAUNCH();
  }
  // ---- critic head + trunk: one fused kernel
  CIRS_LAUNCH(trunk_bwd_kernel, (n + 15) / 16, 256, 0, st, *w, *grads, n, idx, obs, tc ? ws.dh2_part : ws.dh2,
              tc ? tc_split : 0, ws.dv, ws.h1, ws.h2, d_obs);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

// The whole learn() loop of one update (core/policy/ppo.py:173-233): the advantage statistics of every minibatch of
// every repeat (they depend on the permutations only), then per repeat and minibatch forward / loss / backward /
// [gradient all-reduce] / clip / Adam -- issued back to back from C so that the host pays one call instead of ~20
// launches' worth of interpreter overhead per minibatch.  With a communicator (comm.cu) this is the data-parallel
// loop of SURVEY 8e: every rank holds its own chunk of each global minibatch; the statistics are summed over ranks
// once, the flat gradient once per minibatch, everything stream-ordered.
int cirs_comm_allreduce_impl(void* comm, void* buf, int64_t count, int dtype, cudaStream_t st);

extern "C" int cirs_ppo_learn(const cirs_policy_weights* w, const cirs_policy_weights* grads, float* exp_avg,
                              float* exp_avg_sq, const cirs_ppo_config* cfg, int32_t n_repeat, int32_t n_mb,
                              const int32_t* mb_off_h, const int32_t* mb_off, const int32_t* slots, const float* obs,
                              const void* act, const float* adv, const float* returns, const float* v_old,
                              const float* logp_old, double* adv_stats, float* d_obs, int64_t d_obs_floats,
                              float* losses, int32_t* opt_state, double* opt_scratch, void* workspace, void* comm,
                              const int32_t* n_global_h, int32_t n_stats_tail, void* stream) {
  if (!w || !grads || !exp_avg || !exp_avg_sq || !cfg || !mb_off_h || !mb_off || !slots || !adv_stats || !losses ||
      !opt_state || !opt_scratch || n_repeat < 0 || n_mb < 0 || (comm && !n_global_h)) {
    cirs_set_error("cirs_ppo_learn: bad argument");
    return CIRS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t n = mb_off_h[n_mb];
  if (n_mb > 0 && n_repeat > 0) {   // every repeat's minibatch statistics in one launch
    if (!adv) {
      cirs_set_error("cirs_ppo_learn: bad argument");
      return CIRS_ERR_ARG;
    }
    CIRS_LAUNCH(adv_stats_kernel, dim3(n_mb, n_repeat), 1024, 0, st, mb_off, slots, adv, adv_stats, n);
    CIRS_CHECK_LAUNCH();
  }
  if (comm) {
    // n_stats_tail more doubles behind the statistics ride on the same collective (the raw return moments)
    int rc = cirs_comm_allreduce_impl(comm, adv_stats, (int64_t)n_repeat * n_mb * 3 + (n_stats_tail > 0 ? n_stats_tail : 0),
                                      1, st);
    if (rc) return rc;
  }
  for (int r = 0; r < n_repeat; ++r) {
    const int32_t* sl = slots + (int64_t)r * n;
    double* stats = adv_stats + (int64_t)r * n_mb * 3;
    if (d_obs) cudaMemsetAsync(d_obs, 0, sizeof(float) * d_obs_floats, st);   // optim_state.zero_grad(), ppo.py:174
    for (int j = 0; j < n_mb; ++j) {
      const int b = mb_off_h[j], cnt = mb_off_h[j + 1] - b;
      int rc = cirs_ppo_minibatch(w, grads, cfg, cnt, n_global_h ? n_global_h[j] : cnt, sl + b, obs, act, adv, returns,
                                  v_old, logp_old, stats + 3 * j, d_obs, losses + 4 * ((int64_t)r * n_mb + j), workspace,
                                  stream);
      if (rc) return rc;
      if (comm) {   // ONE collective per minibatch: the flat actor / critic gradient (2.8 MB at 10728 items)
        rc = cirs_comm_allreduce_impl(comm, grads->flat, grads->n_flat, 0, st);
        if (rc) return rc;
      }
      rc = cirs_clip_adam(w->flat, grads->flat, exp_avg, exp_avg_sq, w->n_flat, w->n_trunk, cfg, opt_state,
                          opt_scratch, stream);
      if (rc) return rc;
    }
  }
  return CIRS_OK;
}
