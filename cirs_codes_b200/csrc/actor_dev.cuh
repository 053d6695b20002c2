// Device code of K3 (actor head + sampler), shared by the stand-alone kernels (actor.cu) and the persistent rollout
// kernel (rollout.cu).  See actor.cu for the description.
#pragma once
#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace cirs_actor {


constexpr int BM = 64, BN = 128, TM = 8, NT = 256, HID = CIRS_HIDDEN;
constexpr int LDH = BM + 4;   // k-major activations: hT[k][row]
constexpr int LDB = BN + 4;
constexpr int MAX_S = 32;     // dim_state <= 32
constexpr size_t SMEM_BYTES = sizeof(float) * (2 * HID * LDH + HID * LDB + BM * (MAX_S + 1));

struct Partial {
  float m, z, best_s, best_l;
  int best_i;
};

enum { MODE_SAMPLE = 0, MODE_ARGMAX = 1, MODE_EVAL = 2 };

struct HeadArgs {
  cirs_policy_weights W;
  int n_rows;
  const int32_t* gather;  // row k -> environment slot / buffer slot (NULL: identity)
  int state_by_k;         // state row = k (compact) or gathered id
  int out_by_k;           // outputs indexed by k or gathered id
  const uint8_t* active;  // indexed by gathered id
  const float* state;
  int64_t state_stride;
  const float* noise_q;
  uint64_t seed, offset;
  unsigned long long* rng_counter;  // optional device counter added to `offset` and bumped once per call
  int mode;
  const uint32_t* seen;
  const int32_t* act_in;  // MODE_EVAL: action per output index (may be NULL -> value only)
  const float* h2_in;     // optional precomputed trunk output [id][64] (actor_trunk_warp): skips trunk and critic
  int tiles_per_split, n_split;
  Partial* part;          // [n_split, n_rows]
  float* value;
  int icdf;               // MODE_SAMPLE without the per-element race: the partials carry (m, z) only and the caller
                          // samples by inverse CDF (actor_combine_icdf_warp; the persistent rollout)
};

__device__ __forceinline__ void merge_ms(float& m, float& z, float m2, float z2) {
  const float M = fmaxf(m, m2);
  if (M == -INFINITY) return;
  z = z * __expf(m - M) + z2 * __expf(m2 - M);
  m = M;
}
__device__ __forceinline__ void merge_best(float& s, float& l, int& i, float s2, float l2, int i2) {
  if (s2 > s || (s2 == s && i2 < i)) { s = s2; l = l2; i = i2; }
}

// One (64-row tile, catalogue split) work item; blockDim.x == NT.  Every thread of the CTA must call it.
__device__ __forceinline__ void actor_head_body(const HeadArgs& P, int bx, int by, float* smem_dyn) {
  float(*h1T)[LDH] = reinterpret_cast<float(*)[LDH]>(smem_dyn);                       // [HID][LDH]
  float(*h2T)[LDH] = reinterpret_cast<float(*)[LDH]>(smem_dyn + HID * LDH);           // [HID][LDH]
  float(*Bs)[LDB] = reinterpret_cast<float(*)[LDB]>(smem_dyn + 2 * HID * LDH);        // [HID][LDB]
  float(*s_in)[MAX_S + 1] = reinterpret_cast<float(*)[MAX_S + 1]>(smem_dyn + 2 * HID * LDH + HID * LDB);
  __shared__ int s_id[BM];   // gathered id per row, -1 when the row is skipped
  __shared__ int s_any;

  const int tid = threadIdx.x;
  const int r0 = bx * BM, split = by;
  const int S = P.W.dim_state;
  if (tid == 0) s_any = 0;
  __syncthreads();
  if (tid < BM) {
    const int k = r0 + tid;
    int id = -1;
    if (k < P.n_rows) {
      id = P.gather ? P.gather[k] : k;
      if (P.active && !P.active[id]) id = -1;
    }
    s_id[tid] = id;
    if (id >= 0) s_any = 1;
  }
  __syncthreads();
  if (!s_any) return;

  if (P.h2_in) {
    // trunk already evaluated per row (persistent rollout): stage h2 k-major
    for (int i = tid; i < BM * HID; i += NT) {
      const int r = i / HID, k = i % HID;
      const int id = s_id[r];
      h2T[k][r] = id >= 0 ? P.h2_in[(size_t)id * HID + k] : 0.f;
    }
    __syncthreads();
  } else {
    // ---- trunk: h1 = relu(W1 s + b1), h2 = relu(W2 h1 + b2)   (common.py:87-92)
    for (int i = tid; i < BM * S; i += NT) {
      const int r = i / S, c = i % S;
      const int id = s_id[r];
      float v = 0.f;
      if (id >= 0) v = P.state[(int64_t)(P.state_by_k ? (r0 + r) : id) * P.state_stride + c];
      s_in[r][c] = v;
    }
    __syncthreads();
    {
      const int row = tid % BM, cg = tid / BM;  // 4 column groups of 16
      float acc[16];
  #pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __ldg(P.W.b1 + cg * 16 + j);
#pragma unroll 4
      for (int k = 0; k < S; ++k) {
        const float x = s_in[row][k];
        const float4* w = reinterpret_cast<const float4*>(P.W.w1t + (size_t)k * HID + cg * 16);
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = __ldg(w + j);
          acc[4 * j] = fmaf(x, t.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(x, t.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(x, t.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(x, t.w, acc[4 * j + 3]);
        }
      }
  #pragma unroll
      for (int j = 0; j < 16; ++j) h1T[cg * 16 + j][row] = fmaxf(acc[j], 0.f);
      __syncthreads();
  #pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __ldg(P.W.b2 + cg * 16 + j);
#pragma unroll 4
      for (int k = 0; k < HID; ++k) {
        const float x = h1T[k][row];
        const float4* w = reinterpret_cast<const float4*>(P.W.w2t + (size_t)k * HID + cg * 16);
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = __ldg(w + j);
          acc[4 * j] = fmaf(x, t.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(x, t.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(x, t.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(x, t.w, acc[4 * j + 3]);
        }
      }
  #pragma unroll
      for (int j = 0; j < 16; ++j) h2T[cg * 16 + j][row] = fmaxf(acc[j], 0.f);
      __syncthreads();
    }
    // ---- critic: V = wv . h2 + bv   (discrete.py:109-114); written once (split 0)
    if (split == 0 && P.value && tid < BM && s_id[tid] >= 0) {
      float v = __ldg(P.W.bv);
      for (int k = 0; k < HID; ++k) v = fmaf(h2T[k][tid], __ldg(P.W.wv + k), v);
      P.value[P.out_by_k ? (r0 + tid) : s_id[tid]] = v;
    }
  }
  if (P.mode == MODE_EVAL && P.act_in == nullptr) return;  // value only

  // ---- logits tiles + online softmax / race epilogue
  const int tx = tid & 31, ty = tid >> 5;  // a warp = one group of 8 rows; lane = 4 columns
  float m[TM], z[TM], bs[TM], bl[TM];
  int bi[TM], a_given[TM];
  int rid[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    m[i] = -INFINITY; z[i] = 0.f; bs[i] = -INFINITY; bl[i] = 0.f; bi[i] = 0x7fffffff;
    rid[i] = s_id[ty * TM + i];
    a_given[i] = -1;
    if (P.mode == MODE_EVAL && rid[i] >= 0) a_given[i] = P.act_in[P.out_by_k ? (r0 + ty * TM + i) : rid[i]];
  }
  const int nA = P.W.n_action, ldA = P.W.ld_action;
  const int n_tiles = (nA + BN - 1) / BN;
  const int t_beg = split * P.tiles_per_split, t_end = min(n_tiles, t_beg + P.tiles_per_split);
  const int seen_words = (nA + 31) >> 5;
  const uint64_t offset = P.offset + (P.rng_counter ? (uint64_t)*P.rng_counter : 0ull);

  for (int t = t_beg; t < t_end; ++t) {
    const int n0 = t * BN;
    // stage W3t[0..64)[n0 .. n0+128): ld_action is a multiple of 128, so the tile is always in bounds
    for (int i = tid; i < HID * (BN / 4); i += NT) {
      const int k = i / (BN / 4), c4 = i % (BN / 4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(P.W.w3t + (size_t)k * ldA + n0) + c4);
      *reinterpret_cast<float4*>(&Bs[k][c4 * 4]) = v;
    }
    __syncthreads();
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll 8
    for (int k = 0; k < HID; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&h2T[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&h2T[k][ty * TM + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
    // epilogue for columns n0 + tx*4 .. +3
    const int nb = n0 + tx * 4;
    if (nb < nA) {
      const float4 bias = __ldg(reinterpret_cast<const float4*>(P.W.b3 + nb));
      const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        if (rid[i] < 0) continue;
        float l[4];
        uint32_t seen_bits = 0u;
        if (P.seen) seen_bits = P.seen[(size_t)rid[i] * seen_words + (nb >> 5)] >> (nb & 31);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (P.mode == MODE_SAMPLE && !P.icdf) {
          if (P.noise_q) {
            const float* q = P.noise_q + (size_t)(r0 + ty * TM + i) * nA + nb;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = (nb + j < nA) ? -logf(q[j]) : 0.f;
          } else {
            const uint4 rnd = philox4x32(make_uint4((uint32_t)rid[i], (uint32_t)(nb >> 2), (uint32_t)offset,
                                                    (uint32_t)(offset >> 32)),
                                         make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
            g[0] = -logf(-logf(u01(rnd.x)) + 1e-30f);  // Gumbel = -log(q), q = -log(u) ~ Exp(1)
            g[1] = -logf(-logf(u01(rnd.y)) + 1e-30f);
            g[2] = -logf(-logf(u01(rnd.z)) + 1e-30f);
            g[3] = -logf(-logf(u01(rnd.w)) + 1e-30f);
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = (nb + j < nA) && !((seen_bits >> j) & 1u);
          l[j] = ok ? acc[i][j] + bb[j] : -INFINITY;
          mx = fmaxf(mx, l[j]);
        }
        if (mx == -INFINITY) continue;
        const float M = fmaxf(m[i], mx);
        float zz = z[i] * __expf(m[i] - M);
#pragma unroll
        for (int j = 0; j < 4; ++j) zz += __expf(l[j] - M);
        z[i] = zz;
        m[i] = M;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (l[j] == -INFINITY) continue;
          float sc;
          if (P.mode == MODE_EVAL) sc = (nb + j == a_given[i]) ? 0.f : -INFINITY;
          else sc = l[j] + g[j];
          if (sc > bs[i]) { bs[i] = sc; bl[i] = l[j]; bi[i] = nb + j; }
        }
      }
    }
  }
  // ---- merge the 32 lanes of each warp (same 8 rows, different columns)
#pragma unroll
  for (int i = 0; i < TM; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(FULL_MASK, m[i], o), z2 = __shfl_xor_sync(FULL_MASK, z[i], o);
      const float s2 = __shfl_xor_sync(FULL_MASK, bs[i], o), l2 = __shfl_xor_sync(FULL_MASK, bl[i], o);
      const int i2 = __shfl_xor_sync(FULL_MASK, bi[i], o);
      merge_ms(m[i], z[i], m2, z2);
      merge_best(bs[i], bl[i], bi[i], s2, l2, i2);
    }
    if (tx == 0 && rid[i] >= 0) {
      Partial p;
      p.m = m[i]; p.z = z[i]; p.best_s = bs[i]; p.best_l = bl[i]; p.best_i = bi[i];
      P.part[(size_t)split * P.n_rows + r0 + ty * TM + i] = p;
    }
  }
}

// merge the catalogue splits; Categorical.log_prob = log(clamp(p_a / sum p, eps, 1 - eps))  
// returns the sampled action of row k (or -1 when the row is skipped)
__device__ __forceinline__ int actor_combine_row(const HeadArgs& P, int k, int32_t* __restrict__ act,
                                                 float* __restrict__ logp) {
  const int id = P.gather ? P.gather[k] : k;
  if (P.active && !P.active[id]) return -1;
  float m = -INFINITY, z = 0.f, bs = -INFINITY, bl = 0.f;
  int bi = 0x7fffffff;
  for (int s = 0; s < P.n_split; ++s) {
    const Partial p = P.part[(size_t)s * P.n_rows + k];
    merge_ms(m, z, p.m, p.z);
    merge_best(bs, bl, bi, p.best_s, p.best_l, p.best_i);
  }
  const int o = P.out_by_k ? k : id;
  if (act) act[o] = bi;
  if (logp) {
    float pa = expf(bl - m) / z;
    pa = fminf(fmaxf(pa, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
    logp[o] = (bs == -INFINITY) ? logf(CATEGORICAL_EPS) : logf(pa);
  }
  return bi;
}


// Trunk + critic of ONE row by one warp, in exactly the operation order of the tile version above (bias first, k
// ascending, fmaf), so both produce bit-identical h2.  s: the row's state (dim_state floats); sh: >= 160 floats of
// this warp's shared memory; h2_out: 64 floats (global).
__device__ __forceinline__ void actor_trunk_warp(const cirs_policy_weights& W, const float* __restrict__ s, int lane,
                                                 float* sh, float* __restrict__ h2_out, float* __restrict__ value_out) {
  const int S = W.dim_state;
  float* sx = sh;         // [32]
  float* h1 = sh + 32;    // [64]
  float* h2 = sh + 96;    // [64]
  if (lane < S) sx[lane] = s[lane];
  __syncwarp();
  float a0 = __ldg(W.b1 + lane), a1 = __ldg(W.b1 + lane + 32);
#pragma unroll 4
  for (int k = 0; k < S; ++k) {
    const float x = sx[k];
    a0 = fmaf(x, __ldg(W.w1t + (size_t)k * HID + lane), a0);
    a1 = fmaf(x, __ldg(W.w1t + (size_t)k * HID + lane + 32), a1);
  }
  h1[lane] = fmaxf(a0, 0.f);
  h1[lane + 32] = fmaxf(a1, 0.f);
  __syncwarp();
  a0 = __ldg(W.b2 + lane); a1 = __ldg(W.b2 + lane + 32);
#pragma unroll 16
  for (int k = 0; k < HID; ++k) {
    const float x = h1[k];
    a0 = fmaf(x, __ldg(W.w2t + (size_t)k * HID + lane), a0);
    a1 = fmaf(x, __ldg(W.w2t + (size_t)k * HID + lane + 32), a1);
  }
  a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
  h2[lane] = a0; h2[lane + 32] = a1;
  h2_out[lane] = a0; h2_out[lane + 32] = a1;
  __syncwarp();
  if (lane == 0 && value_out) {
    float v = __ldg(W.bv);
    for (int k = 0; k < HID; ++k) v = fmaf(h2[k], __ldg(W.wv + k), v);
    *value_out = v;
  }
  __syncwarp();
}

// warp-parallel merge of the catalogue splits of row k (all 32 lanes call it; every lane returns the action)
__device__ __forceinline__ int actor_combine_warp(const HeadArgs& P, int k, int lane, int32_t* __restrict__ act,
                                                  float* __restrict__ logp) {
  const int id = P.gather ? P.gather[k] : k;
  float m = -INFINITY, z = 0.f, bs = -INFINITY, bl = 0.f;
  int bi = 0x7fffffff;
  if (P.n_split <= 160) {   // all of the lane's partials in flight at once (one L2 round trip instead of five)
    Partial pp[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int s = lane + 32 * i;
      if (s < P.n_split) pp[i] = P.part[(size_t)s * P.n_rows + k];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (lane + 32 * i < P.n_split) {
        merge_ms(m, z, pp[i].m, pp[i].z);
        merge_best(bs, bl, bi, pp[i].best_s, pp[i].best_l, pp[i].best_i);
      }
    }
  } else {
    for (int s = lane; s < P.n_split; s += 32) {
      const Partial p = P.part[(size_t)s * P.n_rows + k];
      merge_ms(m, z, p.m, p.z);
      merge_best(bs, bl, bi, p.best_s, p.best_l, p.best_i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(FULL_MASK, m, o), z2 = __shfl_xor_sync(FULL_MASK, z, o);
    const float s2 = __shfl_xor_sync(FULL_MASK, bs, o), l2 = __shfl_xor_sync(FULL_MASK, bl, o);
    const int i2 = __shfl_xor_sync(FULL_MASK, bi, o);
    merge_ms(m, z, m2, z2);
    merge_best(bs, bl, bi, s2, l2, i2);
  }
  if (lane == 0) {
    const int o = P.out_by_k ? k : id;
    if (act) act[o] = bi;
    if (logp) {
      float pa = expf(bl - m) / z;
      pa = fminf(fmaxf(pa, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
      logp[o] = (bs == -INFINITY) ? logf(CATEGORICAL_EPS) : logf(pa);
    }
  }
  return bi;
}

// Categorical.sample by INVERSE CDF over the catalogue-split partials (the persistent rollout's sampler).
// The exponential race needs one Exp(1) draw per (row, item): 5.5 M Philox + log evaluations per turn at 512 rows, which
// was 60 % of the head phase.  Inverse CDF needs ONE uniform per row: with M = max_s m_s and Z = sum_s z_s e^{m_s - M}
// (the softmax denominator) draw target = u Z, walk the splits in catalogue order to the one that contains the target,
// recompute that split's `width` logits from h2 (64 x width FMAs by one warp, W3 from L2) and walk its columns.  The
// result is distributed exactly as softmax(logits) (up to float rounding of the partial sums; a target that lands
// beyond the recomputed columns by rounding takes the last unmasked one).  Masked items (seen) are excluded in both
// passes.  All 32 lanes call it; every lane returns the action.  h2: the row's trunk output (64 floats, global).
// (__noinline__: called once per row and turn; keeps its ~100 registers of load batches out of the persistent kernel's
// own allocation, which sits at the 255-register limit)
template <int MAXC>   // columns per lane of the recomputed split: split_width <= 32 * MAXC
__device__ __noinline__ int actor_combine_icdf_warp(const HeadArgs& P, int k, int lane, int32_t* __restrict__ act,
                                                       float* __restrict__ logp, const float* h2, int split_width,
                                                       uint64_t offset) {
  const int id = P.gather ? P.gather[k] : k;
  const int nA = P.W.n_action;
  constexpr int MAXS = 8;   // splits per lane: n_split <= 256
  float ms[MAXS], zs[MAXS];
  float M = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXS; ++i) {
    const int s = lane + 32 * i;
    ms[i] = -INFINITY; zs[i] = 0.f;
    if (s < P.n_split) {
      const Partial p = P.part[(size_t)s * P.n_rows + k];
      ms[i] = p.m; zs[i] = p.z;
      M = fmaxf(M, p.m);
    }
  }
  M = warp_max(M);
  float Z = 0.f;
#pragma unroll
  for (int i = 0; i < MAXS; ++i) {
    zs[i] = (ms[i] == -INFINITY) ? 0.f : zs[i] * __expf(ms[i] - M);
    Z += zs[i];
  }
  Z = warp_sum(Z);
  // one uniform in (0, 1) per (row, turn)
  const uint4 rnd = philox4x32(make_uint4((uint32_t)id, 0x1cdfu, (uint32_t)offset, (uint32_t)(offset >> 32)),
                               make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
  const float u = ((rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float target = u * Z;
  // split containing the target: inclusive prefix sums in split order (lane-strided ownership: round i covers 32 splits)
  float carry = 0.f, before = 0.f;
  int s_star = -1;
#pragma unroll
  for (int i = 0; i < MAXS; ++i) {
    if (32 * i >= P.n_split) break;
    float v = zs[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(FULL_MASK, v, o);
      if (lane >= o) v += t;
    }
    const float incl = carry + v;
    const unsigned hit = __ballot_sync(FULL_MASK, s_star < 0 && zs[i] > 0.f && incl >= target);
    if (s_star < 0 && hit) {
      const int l = __ffs(hit) - 1;
      s_star = 32 * i + l;
      before = __shfl_sync(FULL_MASK, incl - zs[i], l);
    }
    carry = __shfl_sync(FULL_MASK, incl, 31);
  }
  if (s_star < 0) {   // rounding: the target is a hair above the total -> last split with mass
    for (int i = MAXS - 1; i >= 0 && s_star < 0; --i) {
      const unsigned any = __ballot_sync(FULL_MASK, zs[i] > 0.f);
      if (any) { const int l = 31 - __clz(any); s_star = 32 * i + l; before = carry - __shfl_sync(FULL_MASK, zs[i], l); }
    }
    if (s_star < 0) s_star = 0;
  }
  const float resid = target - before;
  // recompute the split's logits: lane owns columns c0 + lane + 32 j
  const int c0 = s_star * split_width;
  float e[MAXC], lg[MAXC];
  const int seen_words = (nA + 31) >> 5;
#pragma unroll
  for (int j = 0; j < MAXC; ++j) { e[j] = 0.f; lg[j] = -INFINITY; }
  {
    float acc[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j) acc[j] = 0.f;
    const int ldA = P.W.ld_action;
    const float h_lo = h2[lane], h_hi = h2[lane + 32];   // the row's trunk output: two values per lane, shuffled out below
    bool okc[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j) okc[j] = lane + 32 * j < split_width && c0 + lane + 32 * j < ldA;
    const float* wbase = P.W.w3t + c0 + lane;
    constexpr int KB = MAXC <= 3 ? 16 : 4;   // KB x MAXC loads in flight per batch (W3 is L2 resident)
#pragma unroll
    for (int k0 = 0; k0 < HID; k0 += KB) {
      float wv[KB][MAXC];
#pragma unroll
      for (int u = 0; u < KB; ++u)
#pragma unroll
        for (int j = 0; j < MAXC; ++j) wv[u][j] = okc[j] ? __ldg(wbase + (size_t)(k0 + u) * ldA + 32 * j) : 0.f;
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        const float hv = __shfl_sync(FULL_MASK, (k0 + u) < 32 ? h_lo : h_hi, (k0 + u) & 31);
#pragma unroll
        for (int j = 0; j < MAXC; ++j) acc[j] = fmaf(hv, wv[u][j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int c = c0 + lane + 32 * j;
      if (lane + 32 * j < split_width && c < nA) {
        bool ok = true;
        if (P.seen) ok = !((P.seen[(size_t)id * seen_words + (c >> 5)] >> (c & 31)) & 1u);
        if (ok) { lg[j] = acc[j] + __ldg(P.W.b3 + c); e[j] = __expf(lg[j] - M); }
      }
    }
  }
  // column containing the residual target, in column order (round j covers columns c0 + 32 j .. + 31)
  int a = -1;
  float la = 0.f;
  carry = 0.f;
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    if (32 * j >= split_width) break;
    float v = e[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(FULL_MASK, v, o);
      if (lane >= o) v += t;
    }
    const float incl = carry + v;
    const unsigned hit = __ballot_sync(FULL_MASK, a < 0 && e[j] > 0.f && incl >= resid);
    if (a < 0 && hit) {
      const int l = __ffs(hit) - 1;
      a = c0 + 32 * j + l;
      la = __shfl_sync(FULL_MASK, lg[j], l);
    }
    carry = __shfl_sync(FULL_MASK, incl, 31);
  }
  if (a < 0) {   // rounding (FFMA vs 3xTF32 sums): the last unmasked column of the split
    for (int j = MAXC - 1; j >= 0 && a < 0; --j) {
      const unsigned any = __ballot_sync(FULL_MASK, e[j] > 0.f);
      if (any) { const int l = 31 - __clz(any); a = c0 + 32 * j + l; la = __shfl_sync(FULL_MASK, lg[j], l); }
    }
    if (a < 0) { a = min(c0, nA - 1); la = M; }
  }
  if (lane == 0) {
    const int o = P.out_by_k ? k : id;
    if (act) act[o] = a;
    if (logp) {
      float pa = expf(la - M) / Z;
      pa = fminf(fmaxf(pa, CATEGORICAL_EPS), 1.0f - CATEGORICAL_EPS);
      logp[o] = logf(pa);
    }
  }
  return a;
}

inline int pick_split(int n_rows, int n_action, int target_ctas = 2 * 148) {
  const int row_tiles = (n_rows + BM - 1) / BM, n_tiles = (n_action + BN - 1) / BN;
  int want = (target_ctas + row_tiles - 1) / row_tiles;  // aim for ~2 CTAs per SM
  if (want < 1) want = 1;
  if (want > n_tiles) want = n_tiles;
  return want;
}

// fills tiles_per_split / n_split; returns false on unsupported shapes
inline bool plan_head(HeadArgs& P, int target_ctas = 2 * 148) {
  const int nA = P.W.n_action;
  if (P.W.dim_state > MAX_S || P.W.ld_action % BN != 0 || P.W.ld_action < nA) return false;
  const int n_tiles = (nA + BN - 1) / BN;
  const int n_split = pick_split(P.n_rows, nA, target_ctas);
  P.tiles_per_split = (n_tiles + n_split - 1) / n_split;
  P.n_split = (n_tiles + P.tiles_per_split - 1) / P.tiles_per_split;
  return true;
}
}  // namespace cirs_actor
