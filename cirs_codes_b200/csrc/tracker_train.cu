// K6: training pass of the StateTracker -- one full-sequence causal forward over every environment's token
// sequence and the backward pass, given d loss / d obs for every stored observation.
//
// Replaces autograd through the observations held in the replay buffer (core/policy/ppo.py:215
// loss.backward(retain_graph=True) reaching core/state_tracker.py:170-250 through tianshou/data/batch.py:256-258).
// The reference re-encodes the whole prefix at every step, so its graph holds T separate encoder passes per
// environment; because the mask is causal, position p of ONE pass over the full sequence equals the last
// position of the prefix-p pass, and the gradient of sum_p <d_obs[p], s_p> is the same.
//
// Layout: token (e, p) lives in row e*L + p -- the replay buffer's slot index -- of every [M, .] activation
// matrix (M = n_env * L).  Rows with p >= ep_len[e] are padding: their token is zero, nothing valid attends to
// them (causality) and their upstream gradient is zero, so they contribute nothing.
// Linear layers, their input gradients and their weight gradients run on the FP32 tile GEMM (gemm.cuh); attention,
// LayerNorm, the reward gate and the embedding scatter are small dedicated kernels.  All reductions that feed a
// parameter gradient use float atomics at CTA granularity (order-dependent in the last bits only).
#include <stdlib.h>

#include "gemm.cuh"
#include "tracker_fused.cuh"
#include "../../include/cirs_b200.h"

namespace {
using namespace cirs;

__host__ __device__ inline int64_t al(int64_t x) { return (x + 63) & ~(int64_t)63; }
inline int up32(int v) { return (v + 31) & ~31; }

// Row <-> (environment, position) mapping.  Padded: row == buffer slot e*L + p.  Compact (tok_slot given): row i holds
// the token of buffer slot tok_slot[i]; slots are env-major sorted, so environment e owns rows env_off[e] ..
// env_off[e] + ep_len[e) and only valid tokens exist.
struct RowMap {
  int L;
  const int32_t* tok_slot;
  const int32_t* env_off;
  __device__ __forceinline__ int slot(int row) const { return tok_slot ? tok_slot[row] : row; }
  __device__ __forceinline__ size_t base(int e) const { return env_off ? (size_t)env_off[e] : (size_t)e * L; }
};

struct LayerBufs {
  float *qkv, *o, *r1, *st1, *x1, *h, *r2, *st2, *x2;
};
struct Bufs {
  float *u, *in, *g, *x0, *tok0, *da, *db, *dqkv, *dh, *dtmp, *din, *du, *dtok0;
  LayerBufs layer[CIRS_MAX_LAYERS];
  int64_t total;
};

Bufs carve(float* base, int64_t B, int64_t M, int d, int dhid, int nl, int d_user_in) {
  Bufs b;
  int64_t off = 0;
  auto take = [&](int64_t n) { float* r = base ? base + off : nullptr; off += al(n); return r; };
  b.u = take(B * d_user_in); b.in = take(M * (1 + d)); b.g = take(M * d); b.x0 = take(M * d); b.tok0 = take(B * d);
  b.da = take(M * d); b.db = take(M * d); b.dqkv = take(M * 3 * d); b.dh = take(M * dhid); b.dtmp = take(M * d);
  b.din = take(M * (1 + d)); b.du = take(B * d_user_in); b.dtok0 = take(B * d);
  for (int l = 0; l < nl; ++l) {
    LayerBufs& y = b.layer[l];
    y.qkv = take(M * 3 * d); y.o = take(M * d); y.r1 = take(M * d); y.st1 = take(M * 2); y.x1 = take(M * d);
    y.h = take(M * dhid); y.r2 = take(M * d); y.st2 = take(M * 2); y.x2 = take(M * d);
  }
  b.total = off;
  return b;
}

// ---- gather the token inputs: U[e] = user embedding / dense user;  IN[row] = [rew ; item embedding] for p >= 1
__global__ void gather_inputs_kernel(cirs_tracker_weights W, int B, RowMap map, const int32_t* __restrict__ users,
                                     const int32_t* __restrict__ act, const float* __restrict__ rew,
                                     const int32_t* __restrict__ ep_len, const float* __restrict__ dense_user,
                                     const float* __restrict__ dense_item, float* __restrict__ U,
                                     float* __restrict__ IN) {
  const int row = blockIdx.x, L = map.L, slot = map.slot(row), e = slot / L, p = slot % L, d = W.d, tid = threadIdx.x;
  const int n = ep_len[e];
  if (p == 0) {
    const int du = W.d_user_in;
    const float* src = W.emb_user ? W.emb_user + (size_t)users[e] * d : dense_user + (size_t)e * du;
    for (int c = tid; c < du; c += blockDim.x) U[(size_t)e * du + c] = src[c];
  }
  float* out = IN + (size_t)row * (1 + d);
  if (p >= 1 && p < n) {
    const size_t prev = (size_t)e * L + p - 1;
    const float* src = W.emb_item ? W.emb_item + (size_t)act[prev] * d : dense_item + prev * d;
    if (tid == 0) out[0] = rew[prev];
    for (int c = tid; c < d; c += blockDim.x) out[1 + c] = src[c];
  } else {
    for (int c = tid; c < 1 + d; c += blockDim.x) out[c] = 0.f;
  }
}

// ---- tokens: x0 = sqrt(d) * tok + PE[p];  tok = ffn_user(u) at p = 0, sigmoid(Z) * a afterwards (Z -> G in place)
__global__ void token_fwd_kernel(cirs_tracker_weights W, RowMap map, const int32_t* __restrict__ ep_len,
                                 const float* __restrict__ IN, const float* __restrict__ tok0, float* __restrict__ G,
                                 float* __restrict__ X0) {
  const int row = blockIdx.x, L = map.L, slot = map.slot(row), e = slot / L, p = slot % L, d = W.d;
  const bool valid = p < ep_len[e];
  const float sq = sqrtf((float)d);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const size_t i = (size_t)row * d + c;
    const float pe = p < W.max_len ? W.pe[(size_t)p * d + c] : 0.f;  // rows beyond max_len are always padding
    float tok = 0.f, g = 0.f;
    if (valid) {
      if (p == 0) tok = tok0[(size_t)e * d + c];
      else {
        g = 1.f / (1.f + expf(-G[i]));
        tok = g * IN[(size_t)row * (1 + d) + 1 + c];
      }
    }
    G[i] = g;
    X0[i] = tok * sq + pe;
  }
}

// d tok = sqrt(d) * dX0;  p = 0 -> dTOK0[e];  p >= 1 -> dZ = dtok * a * g (1 - g),  DA = dtok * g
__global__ void token_bwd_kernel(cirs_tracker_weights W, RowMap map, const int32_t* __restrict__ ep_len,
                                 const float* __restrict__ IN, const float* __restrict__ G,
                                 const float* __restrict__ dX0, float* __restrict__ dTOK0, float* __restrict__ dZ,
                                 float* __restrict__ DA) {
  const int row = blockIdx.x, L = map.L, slot = map.slot(row), e = slot / L, p = slot % L, d = W.d;
  const bool valid = p < ep_len[e];
  const float sq = sqrtf((float)d);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const size_t i = (size_t)row * d + c;
    const float dt = valid ? dX0[i] * sq : 0.f;
    float dz = 0.f, da = 0.f;
    if (p == 0) dTOK0[(size_t)e * d + c] = dt;
    else if (valid) {
      const float g = G[i], a = IN[(size_t)row * (1 + d) + 1 + c];
      dz = dt * a * g * (1.f - g);
      da = dt * g;
    }
    dZ[i] = dz;
    DA[i] = da;
  }
}

// embedding gradients (dense nn.Embedding grads, scatter-add)
__global__ void emb_scatter_kernel(cirs_tracker_weights W, cirs_tracker_weights Gr, int B, RowMap map,
                                   const int32_t* __restrict__ users, const int32_t* __restrict__ act,
                                   const int32_t* __restrict__ ep_len, const float* __restrict__ dIN,
                                   const float* __restrict__ DA, const float* __restrict__ dU) {
  const int row = blockIdx.x, L = map.L, slot = map.slot(row), e = slot / L, p = slot % L, d = W.d;
  if (p == 0) {
    if (Gr.emb_user)
      for (int c = threadIdx.x; c < d; c += blockDim.x)
        atomicAdd(Gr.emb_user + (size_t)users[e] * d + c, dU[(size_t)e * d + c]);
  } else if (p < ep_len[e] && Gr.emb_item) {
    const int item = act[(size_t)e * L + p - 1];
    for (int c = threadIdx.x; c < d; c += blockDim.x)
      atomicAdd(Gr.emb_item + (size_t)item * d + c, dIN[(size_t)row * (1 + d) + 1 + c] + DA[(size_t)row * d + c]);
  }
}

// ---- LayerNorm (eps 1e-5, biased variance): one warp per row
__global__ void __launch_bounds__(256)
ln_fwd_kernel(int M, int d, const float* __restrict__ R, const float* __restrict__ w, const float* __restrict__ b,
              float* __restrict__ X, float* __restrict__ ST) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* r = R + (size_t)row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += r[c];
  const float mu = warp_sum(s) / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) { const float v = r[c] - mu; q += v * v; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / d + 1e-5f);
  for (int c = lane; c < d; c += 32) X[(size_t)row * d + c] = (r[c] - mu) * rstd * __ldg(w + c) + __ldg(b + c);
  if (lane == 0) { ST[2 * row] = mu; ST[2 * row + 1] = rstd; }
}

// dR = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dY * w;  gw += dY * xhat, gb += dY
constexpr int LN_MAXC = 8;  // d <= 256
__global__ void __launch_bounds__(256)
ln_bwd_kernel(int M, int d, const float* __restrict__ dY, const float* __restrict__ R, const float* __restrict__ ST,
              const float* __restrict__ w, float* __restrict__ dR, float* gw, float* gb) {
  __shared__ float sw[8][256], sb[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[LN_MAXC], ab[LN_MAXC];
#pragma unroll
  for (int k = 0; k < LN_MAXC; ++k) aw[k] = ab[k] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const float mu = ST[2 * row], rstd = ST[2 * row + 1];
    const float* r = R + (size_t)row * d;
    const float* dy = dY + (size_t)row * d;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        const float xh = (r[c] - mu) * rstd, dyc = dy[c], dxh = dyc * __ldg(w + c);
        s1 += dxh; s2 += dxh * xh;
        aw[k] += dyc * xh; ab[k] += dyc;
      }
    }
    s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        const float xh = (r[c] - mu) * rstd, dxh = dy[c] * __ldg(w + c);
        dR[(size_t)row * d + c] = rstd * (dxh - s1 - xh * s2);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LN_MAXC; ++k) { sw[warp][lane + 32 * k] = aw[k]; sb[warp][lane + 32 * k] = ab[k]; }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; ++k) { a += sw[k][c]; b += sb[k][c]; }
    atomicAdd(gw + c, a);
    atomicAdd(gb + c, b);
  }
}

// ---- causal multi-head self-attention, one CTA per (environment, head)
struct AttnSmem {
  float *q, *k, *v, *dO, *mx, *iz, *dd, *pw, *sw;
};
__device__ __forceinline__ AttnSmem attn_carve(float* s, int Lmax, int dh, int nwarp) {
  AttnSmem a;
  const int ld = dh + 1;
  a.q = s; a.k = a.q + Lmax * ld; a.v = a.k + Lmax * ld; a.dO = a.v + Lmax * ld;
  a.mx = a.dO + Lmax * ld; a.iz = a.mx + Lmax; a.dd = a.iz + Lmax;
  a.pw = a.dd + Lmax; a.sw = a.pw + nwarp * Lmax;
  return a;
}
inline size_t attn_smem_bytes(int Lmax, int dh, int nwarp) {
  return sizeof(float) * ((size_t)4 * Lmax * (dh + 1) + 3 * Lmax + 2 * nwarp * Lmax);
}
constexpr int ATT_WARPS = 4;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(RowMap map, int d, int nhead, const int32_t* __restrict__ ep_len, const float* __restrict__ QKV,
                float* __restrict__ O) {
  extern __shared__ float sm[];
  const int e = blockIdx.x, h = blockIdx.y, dh = d / nhead, ld = dh + 1, L = map.L;
  const int n = min(ep_len[e], L), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r0 = map.base(e);
  const int n_rows = map.env_off ? n : L;   // rows this environment owns (padding rows exist only when padded)
  AttnSmem S = attn_carve(sm, L, dh, ATT_WARPS);
  for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
    const int p = i / dh, c = i % dh;
    const float* src = QKV + (r0 + p) * 3 * d + h * dh + c;
    S.q[p * ld + c] = src[0]; S.k[p * ld + c] = src[d]; S.v[p * ld + c] = src[2 * d];
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  float* pw = S.pw + warp * L;
  for (int i = warp; i < n_rows; i += ATT_WARPS) {
    float* out = O + (r0 + i) * d + h * dh;
    if (i >= n) {
      for (int c = lane; c < dh; c += 32) out[c] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      float s = 0.f;
      for (int c = 0; c < dh; ++c) s = fmaf(S.q[i * ld + c] * scale, S.k[j * ld + c], s);
      pw[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j <= i; j += 32) { const float ex = expf(pw[j] - mx); pw[j] = ex; z += ex; }
    z = warp_sum(z);
    __syncwarp();
    const float inv = 1.0f / z;
    for (int c = lane; c < dh; c += 32) {
      float a = 0.f;
      for (int j = 0; j <= i; ++j) a = fmaf(pw[j] * inv, S.v[j * ld + c], a);
      out[c] = a;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_kernel(RowMap map, int d, int nhead, const int32_t* __restrict__ ep_len, const float* __restrict__ QKV,
                const float* __restrict__ dOg, float* __restrict__ dQKV) {
  extern __shared__ float sm[];
  const int e = blockIdx.x, h = blockIdx.y, dh = d / nhead, ld = dh + 1, L = map.L;
  const int n = min(ep_len[e], L), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t r0 = map.base(e);
  const int n_rows = map.env_off ? n : L;   // rows this environment owns (padding rows exist only when padded)
  AttnSmem S = attn_carve(sm, L, dh, ATT_WARPS);
  for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
    const int p = i / dh, c = i % dh;
    const size_t row = r0 + p;
    const float* src = QKV + row * 3 * d + h * dh + c;
    S.q[p * ld + c] = src[0]; S.k[p * ld + c] = src[d]; S.v[p * ld + c] = src[2 * d];
    S.dO[p * ld + c] = dOg[row * d + h * dh + c];
  }
  // padding rows: zero gradient
  for (int i = threadIdx.x; i < (n_rows - n) * dh; i += blockDim.x) {
    const int p = n + i / dh, c = i % dh;
    float* dst = dQKV + (r0 + p) * 3 * d + h * dh + c;
    dst[0] = 0.f; dst[d] = 0.f; dst[2 * d] = 0.f;
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  float* pw = S.pw + warp * L;
  float* sw = S.sw + warp * L;
  // phase 1: per query i -- row statistics, D_i = sum_j p_ij dP_ij, dq_i
  for (int i = warp; i < n; i += ATT_WARPS) {
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      float s = 0.f;
      for (int c = 0; c < dh; ++c) s = fmaf(S.q[i * ld + c] * scale, S.k[j * ld + c], s);
      pw[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j <= i; j += 32) { const float ex = expf(pw[j] - mx); pw[j] = ex; z += ex; }
    z = warp_sum(z);
    const float inv = 1.0f / z;
    float dsum = 0.f;
    for (int j = lane; j <= i; j += 32) {
      float dp = 0.f;
      for (int c = 0; c < dh; ++c) dp = fmaf(S.dO[i * ld + c], S.v[j * ld + c], dp);
      const float p = pw[j] * inv;
      pw[j] = p;
      sw[j] = dp;
      dsum = fmaf(p, dp, dsum);
    }
    dsum = warp_sum(dsum);
    for (int j = lane; j <= i; j += 32) sw[j] = pw[j] * (sw[j] - dsum);   // dS_ij
    if (lane == 0) { S.mx[i] = mx; S.iz[i] = inv; S.dd[i] = dsum; }
    __syncwarp();
    float* dq = dQKV + (r0 + i) * 3 * d + h * dh;
    for (int c = lane; c < dh; c += 32) {
      float a = 0.f;
      for (int j = 0; j <= i; ++j) a = fmaf(sw[j], S.k[j * ld + c], a);
      dq[c] = a * scale;
    }
    __syncwarp();
  }
  __syncthreads();
  // phase 2: per key j -- dk_j = scale * sum_{i>=j} dS_ij q_i,  dv_j = sum_{i>=j} p_ij dO_i
  for (int j = warp; j < n; j += ATT_WARPS) {
    for (int i = j + lane; i < n; i += 32) {
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < dh; ++c) {
        s = fmaf(S.q[i * ld + c] * scale, S.k[j * ld + c], s);
        dp = fmaf(S.dO[i * ld + c], S.v[j * ld + c], dp);
      }
      const float p = expf(s - S.mx[i]) * S.iz[i];
      pw[i] = p;
      sw[i] = p * (dp - S.dd[i]);
    }
    __syncwarp();
    float* dk = dQKV + (r0 + j) * 3 * d + d + h * dh;
    for (int c = lane; c < dh; c += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = j; i < n; ++i) {
        ak = fmaf(sw[i], S.q[i * ld + c], ak);
        av = fmaf(pw[i], S.dO[i * ld + c], av);
      }
      dk[c] = ak * scale;
      dk[d + c] = av;
    }
    __syncwarp();
  }
}

int splitk(int tiles, int K) {
  int s = (2 * 148 + tiles - 1) / tiles;
  const int max_s = (K + 63) / 64;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

// ---- one-shot linear kernel for the tracker's skinny GEMMs (K <= 128): Y[M, N] = ep(X[M, K] . Wop[K, N]).
// The tile GEMM walks K in steps with a load -> sync -> compute round trip per step; for these shapes (a few thousand
// rows, K and N of 32 .. 128) that chain IS the run time.  Here a CTA loads its 32 rows of X and the whole K x 128
// weight block with every load in flight at once, synchronises once and computes from shared memory.
//   B_TRANS = false: Wop[k][n] = W[k * ldw + n]   (forward: Wt k-major)
//   B_TRANS = true : Wop[k][n] = W[n * ldw + k]   (input gradient: dX = dY . Wt^T)
// Epilogue as StoreEp: + bias, relu, * [mask > 0], + residual.
constexpr int SK_BM = 32, SK_BN = 128, SK_MAXK = 128, SK_LDX = SK_BM + 4, SK_LDW = SK_BN + 4;
constexpr size_t SK_SMEM = sizeof(float) * (SK_MAXK * SK_LDX + SK_MAXK * SK_LDW);

template <bool B_TRANS>
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                     const float* __restrict__ bias, float* __restrict__ Y, int ldy, int M, int N, int K, int relu,
                     const float* __restrict__ mask, int ldm, const float* __restrict__ res, int ldr) {
  extern __shared__ __align__(16) float sk_smem[];
  float* Xs = sk_smem;                       // [K][SK_LDX]  k-major: Xs[k][m]
  float* Ws = sk_smem + SK_MAXK * SK_LDX;    // [K][SK_LDW]  Ws[k][n]
  const int tid = threadIdx.x, m0 = blockIdx.x * SK_BM, n0 = blockIdx.y * SK_BN;
  const int nc = min(SK_BN, N - n0);         // columns of this CTA
  const int k4 = K >> 2;
  for (int i = tid; i < SK_BM * k4; i += 256) {     // X tile, transposed into Xs[k][m]
    const int m = i / k4, c = i % k4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + m < M) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + m) * ldx) + c);
    Xs[(4 * c) * SK_LDX + m] = v.x; Xs[(4 * c + 1) * SK_LDX + m] = v.y;
    Xs[(4 * c + 2) * SK_LDX + m] = v.z; Xs[(4 * c + 3) * SK_LDX + m] = v.w;
  }
  if (!B_TRANS) {
    const int n4 = (nc + 3) >> 2;
    for (int i = tid; i < K * n4; i += 256) {
      const int k = i / n4, c = i % n4;
      *reinterpret_cast<float4*>(Ws + k * SK_LDW + 4 * c) =
          __ldg(reinterpret_cast<const float4*>(W + (size_t)k * ldw + n0) + c);   // ldw is padded to 32: in bounds
    }
  } else {
    for (int i = tid; i < nc * k4; i += 256) {
      const int n = i / k4, c = i % k4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + n) * ldw) + c);
      Ws[(4 * c) * SK_LDW + n] = v.x; Ws[(4 * c + 1) * SK_LDW + n] = v.y;
      Ws[(4 * c + 2) * SK_LDW + n] = v.z; Ws[(4 * c + 3) * SK_LDW + n] = v.w;
    }
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;    // 4 rows (ty) x 4 columns (tx) per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  if (4 * tx < nc) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(Xs + k * SK_LDX + 4 * ty);
      const float4 b = *reinterpret_cast<const float4*>(Ws + k * SK_LDW + 4 * tx);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + 4 * ty + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + 4 * tx + j;
        if (n >= N) continue;
        float v = acc[i][j];
        if (bias) v += __ldg(bias + n);
        if (relu) v = fmaxf(v, 0.f);
        if (mask && !(mask[(size_t)m * ldm + n] > 0.f)) v = 0.f;
        if (res) v += res[(size_t)m * ldr + n];
        Y[(size_t)m * ldy + n] = v;
      }
    }
  }
}

// true when the one-shot kernel handles the shape (16-byte aligned rows, K a multiple of 4 and <= 128)
inline bool skinny_ok(const void* X, int ldx, const void* W, int ldw, int K) {
  return K >= 4 && K <= SK_MAXK && (K & 3) == 0 && (ldx & 3) == 0 && (ldw & 3) == 0 &&
         ((uintptr_t)X & 15) == 0 && ((uintptr_t)W & 15) == 0;
}
template <bool B_TRANS>
void launch_skinny(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M, int N,
                   int K, int relu, const float* mask, int ldm, const float* res, int ldr, cudaStream_t st,
                   const char* tag) {
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(skinny_linear_kernel<B_TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM);
    once = true;
  }
  dim3 grid((M + SK_BM - 1) / SK_BM, (N + SK_BN - 1) / SK_BN);
  const bool prof = cirs_profile_begin(tag, st);
  skinny_linear_kernel<B_TRANS><<<grid, 256, SK_SMEM, st>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K, relu, mask, ldm, res,
                                                             ldr);
  cirs_note_launch();
  if (prof) cirs_profile_end(st);
}

// Y[M,N] = X[M,K] Wt[K][ldw] + b (+ relu) (+ res)
void linear_fwd(const float* X, int ldx, const float* Wt, int ldw, const float* b, float* Y, int ldy, int M, int N,
                int K, int relu, const float* res, int ldr, cudaStream_t st) {
  if (skinny_ok(X, ldx, Wt, ldw, K)) {
    launch_skinny<false>(X, ldx, Wt, ldw, b, Y, ldy, M, N, K, relu, nullptr, 0, res, ldr, st, "tracker_linear_fwd_gemm");
    return;
  }
  launch_gemm<64, 64, 32, 4>(RowMajorA{X, ldx, nullptr}, RowMajorB{Wt, ldw, nullptr},
                             StoreEp{Y, ldy, b, relu, nullptr, nullptr, 0, res, ldr}, M, N, K, 1, nullptr, st, "tracker_linear_fwd_gemm");
}
// dX[M,K] = dY[M,N] Wt^T (* mask) (+ res)
void linear_bwd_x(const float* dY, int ldy, const float* Wt, int ldw, float* dX, int ldx, int M, int N, int K,
                  const float* mask, int ldm, const float* res, int ldr, cudaStream_t st) {
  // dX[m][kk] = sum_n dY[m][n] Wt[kk][n]: contraction over N (the layer's outputs), K_in outputs per row
  if (skinny_ok(dY, ldy, Wt, ldw, N)) {
    launch_skinny<true>(dY, ldy, Wt, ldw, nullptr, dX, ldx, M, K, N, 0, mask, ldm, res, ldr, st, "tracker_linear_dx_gemm");
    return;
  }
  launch_gemm<64, 64, 32, 4>(RowMajorA{dY, ldy, nullptr}, ColMajorB{Wt, ldw},
                             StoreEp{dX, ldx, nullptr, 0, nullptr, mask, ldm, res, ldr}, M, K, N, 1, nullptr, st, "tracker_linear_dx_gemm");
}
// gWt[K][ldw] += X^T dY,  gb[N] += colsum(dY)
void linear_bwd_w(const float* X, int ldx, const float* dY, int ldy, float* gWt, int ldw, float* gb, int M, int N,
                  int K, cudaStream_t st) {
  const int tiles = ((K + 63) / 64) * ((N + 63) / 64);
  launch_gemm<64, 64, 32, 4>(ColMajorA{X, ldx, nullptr}, RowMajorB{dY, ldy, nullptr}, AtomicEp{gWt, ldw}, K, N, M,
                             splitk(tiles, M), gb, st, "tracker_linear_dw_gemm");
}

// The weight-gradient GEMM of a layer (X^T dY) and the GEMM that propagates dY to the layer's input are independent:
// both only read dY.  These small GEMMs fill a fraction of the GPU each (25-50 CTAs), so the weight-gradient one is
// forked onto a side stream and joined again before dY's buffer can be reused.
struct SideStream {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok() {
    if (!side) {
      if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) { side = nullptr; return false; }
      cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    }
    return true;
  }
};
SideStream g_side;

template <class FW, class FX>
void fork_join(cudaStream_t st, FW dw, FX dx) {
  if (!g_side.ok()) { dw(st); dx(st); return; }
  cudaEventRecord(g_side.fork, st);
  cudaStreamWaitEvent(g_side.side, g_side.fork, 0);
  dw(g_side.side);
  dx(st);
  cudaEventRecord(g_side.join, g_side.side);
  cudaStreamWaitEvent(st, g_side.join, 0);
}

}  // namespace

// ---- fused path (tracker_fused.cuh): chunk kernel + grouped weight-gradient launch
static int g_fused_mode = -1;   // -1 default (on unless CIRS_K6_UNFUSED=1), 0 off, 1 on
static bool fused_enabled() {
  if (g_fused_mode >= 0) return g_fused_mode != 0;
  const char* e = getenv("CIRS_K6_UNFUSED");
  return !(e && e[0] == '1');
}
extern "C" void cirs_tracker_train_fused_enable(int on) { g_fused_mode = on < 0 ? -1 : (on ? 1 : 0); }

struct FusedPlan { bool ok, res; int TM, ldx, ldb, q; size_t smem; };
static FusedPlan fused_plan(const cirs_tracker_weights& W, int max_ep_len, int M = 0, int B = 0) {
  FusedPlan P{};
  const int d = W.d;
  int wide = 3 * d;
  if (W.d_hid > wide) wide = W.d_hid;
  if (1 + W.d_user_in > wide) wide = 1 + W.d_user_in;
  if (1 + d > wide) wide = 1 + d;
  if (W.dim_state > wide) wide = W.dim_state;
  P.ldx = cirs_k6::up4(d) + 4;
  P.ldb = cirs_k6::up4(wide) + 4;
  // measured on B200: the chunk kernel beats the layer-by-layer launches for d <= 64 (configs[1]: 0.30 vs 0.56 ms,
  // configs[2]: 1.03 vs 1.40 ms); at d = 128 its 32-row chunks lose (2.0 vs 1.25 ms), so that shape keeps the launches
  if (d > 64) return P;
  // d <= 32: every Linear weight resident in shared memory (~120 KB), 32-row chunks -- no weight traffic and no
  // barriers inside the stage products; the per-chunk latency drops ~4x against the streaming variant
  if (d <= 32 && max_ep_len <= 32) {
    const size_t smem = cirs_k6::chunk_smem_bytes(32, d, W.nhead, P.ldx, P.ldb, cirs_k6::resident_floats(W));
    if (smem <= 224 * 1024 && !getenv("CIRS_K6_STREAM")) {
      P.ok = true; P.res = true; P.TM = 32; P.smem = smem; P.q = 32 - max_ep_len + 1;
      // 16-row chunks (about half the latency per chunk: the stage products scale with the rows) when the greedy plan
      // still fits ONE wave of the 148 SMs: a chunk then holds ~16 - (mean episode length) / 2 rows
      const char* force = getenv("CIRS_K6_TM");
      const double mean_len = B > 0 ? (double)M / B : max_ep_len;
      const bool fits16 = max_ep_len <= 16 && M > 0 && (double)M / (16.0 - 0.5 * mean_len) <= 140.0;
      if ((force ? atoi(force) == 16 && max_ep_len <= 16 : fits16)) {
        P.TM = 16; P.q = 16 - max_ep_len + 1;
        P.smem = cirs_k6::chunk_smem_bytes(16, d, W.nhead, P.ldx, P.ldb, cirs_k6::resident_floats(W));
      }
      return P;
    }
  }
  for (int TM : {64, 32}) {
    const size_t smem = cirs_k6::chunk_smem_bytes(TM, d, W.nhead, P.ldx, P.ldb);
    // the probabilities of a chunk's attention live in the weight stage: [longest episode][TM * nhead] floats
    if (smem <= 224 * 1024 && max_ep_len <= TM &&
        (size_t)max_ep_len * TM * W.nhead <= (size_t)cirs_k6::WS_K * cirs_k6::WS_LD) {
      P.ok = true; P.TM = TM; P.smem = smem; P.q = TM - max_ep_len + 1;
      return P;
    }
  }
  return P;
}

static int fused_train(const cirs_tracker_weights& W, const cirs_tracker_weights& G, int B, int L, int M,
                       const int32_t* users, const int32_t* act, const float* rew, const int32_t* ep_len,
                       const float* dense_user, const float* dense_item, const int32_t* tok_slot,
                       const int32_t* env_off, const float* d_obs, float* obs_check, float* workspace,
                       const FusedPlan& P, int phase, cudaStream_t st) {
  using namespace cirs_k6;
  const int d = W.d, dhid = W.d_hid, S = W.dim_state, nl = W.nlayers, dui = W.d_user_in;
  cirs_k6::Args A{};
  A.W = W; A.G = G;
  A.S = cirs_k6::carve(workspace, M, d, dhid, nl, dui);
  A.n_env = B; A.L = L; A.M = M; A.q = P.q;
  A.users = users; A.act = act; A.ep_len = ep_len; A.tok_slot = tok_slot; A.env_off = env_off;
  A.rew = rew; A.dense_user = dense_user; A.dense_item = dense_item; A.d_obs = d_obs; A.obs_check = obs_check;
  A.ldx = P.ldx; A.ldb = P.ldb;
  A.phase = phase;
  const int n_chunks = (M + P.q - 1) / P.q;   // the quantum rule's count: an upper bound of the greedy plan's
  const int grid = n_chunks < 148 ? n_chunks : 148;
  if (P.res && B <= cirs_k6::PLAN_MAX_ENV) {   // greedy chunk plan behind the workspace's activations
    int32_t* plan = reinterpret_cast<int32_t*>(workspace + A.S.total);
    CIRS_LAUNCH(chunk_plan_kernel, 1, 1024, 0, st, B, env_off, P.TM, plan);
    CIRS_CHECK_LAUNCH();
    A.chunk_e0 = plan;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(tracker_chunk_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(tracker_chunk_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(tracker_chunk_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(tracker_chunk_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    attr_set = true;
  }
  {
    const bool prof = cirs_profile_begin("tracker_chunk_kernel", st);
    if (P.res && P.TM == 16) tracker_chunk_kernel<16, true><<<grid, NT, P.smem, st>>>(A);
    else if (P.res) tracker_chunk_kernel<32, true><<<grid, NT, P.smem, st>>>(A);
    else if (P.TM == 64) tracker_chunk_kernel<64, false><<<grid, NT, P.smem, st>>>(A);
    else tracker_chunk_kernel<32, false><<<grid, NT, P.smem, st>>>(A);
    cirs_note_launch();
    if (prof) cirs_profile_end(st);
  }
  CIRS_CHECK_LAUNCH();
  if (!d_obs || !(phase & 2)) return CIRS_OK;
  // ---- every Linear's weight / bias gradient in one grouped split-K launch
  const int ldd = up32(d), ld3 = up32(3 * d), ldh = up32(dhid), lds = up32(S);
  DwArgs D{};
  int np = 0, cta = 0;
  auto add = [&](const float* x, int ldx, const float* dy, int ldy, const int32_t* rows, float* gw, int ldw, float* gb,
                 int k_in, int n_out) {
    DwProblem& p = D.p[np++];
    p.x = x; p.ldx = ldx; p.dy = dy; p.ldy = ldy; p.dy_rows = rows; p.gw = gw; p.ldw = ldw; p.gb = gb;
    p.k_in = k_in; p.n_out = n_out; p.tiles_n = (n_out + 63) / 64; p.tiles_k = (k_in + 63) / 64; p.cta0 = cta;
    cta += p.tiles_n * p.tiles_k;   // per split; scaled below
  };
  const float* x_last = nl ? A.S.layer[nl - 1].x2 : A.S.x0;
  add(x_last, d, d_obs, S, tok_slot, G.dec_wt, lds, G.dec_b, d, S);
  for (int l = 0; l < nl; ++l) {
    const LayerSave& y = A.S.layer[l];
    const cirs_encoder_layer& Gy = G.layer[l];
    const float* xin = l == 0 ? A.S.x0 : A.S.layer[l - 1].x2;
    add(y.h, dhid, y.dr2, d, nullptr, Gy.l2_wt, ldd, Gy.l2_b, dhid, d);
    add(y.x1, d, y.dh, dhid, nullptr, Gy.l1_wt, ldh, Gy.l1_b, d, dhid);
    add(y.o, d, y.dr1, d, nullptr, Gy.out_wt, ldd, Gy.out_b, d, d);
    add(xin, d, y.dqkv, 3 * d, nullptr, Gy.in_wt, ld3, Gy.in_b, d, 3 * d);
  }
  add(A.S.in, 1 + d, A.S.dz, d, nullptr, G.gate_wt, ldd, G.gate_b, 1 + d, d);
  add(A.S.u, dui, A.S.dtok0, d, nullptr, G.user_wt, ldd, G.user_b, dui, d);
  const int tiles_total = cta;
  int splits = (2 * 148 + tiles_total - 1) / tiles_total;
  const int max_s = (M + 63) / 64;
  if (splits > max_s) splits = max_s;
  if (splits < 1) splits = 1;
  int kps = (M + splits - 1) / splits;
  kps = ((kps + 31) / 32) * 32;
  splits = (M + kps - 1) / kps;
  for (int i = 0; i < np; ++i) D.p[i].cta0 *= splits;
  D.n_prob = np; D.M = M; D.splits = splits; D.k_per_split = kps;
  CIRS_LAUNCH(tracker_dw_grouped_kernel, tiles_total * splits, 256, 0, st, D);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int64_t cirs_tracker_train_workspace_bytes(const cirs_tracker_weights* w, int32_t n_env,
                                                      int64_t n_rows) {
  if (!w) return 0;
  const Bufs b = carve(nullptr, n_env, n_rows, w->d, w->d_hid, w->nlayers, w->d_user_in);
  const cirs_k6::Save s = cirs_k6::carve(nullptr, n_rows, w->d, w->d_hid, w->nlayers, w->d_user_in);
  const int64_t t = b.total > s.total ? b.total : s.total;
  return t * (int64_t)sizeof(float) + 256 + ((int64_t)n_env + 64) * (int64_t)sizeof(int32_t);   // + the greedy chunk plan
}

extern "C" int cirs_tracker_train(const cirs_tracker_weights* w, const cirs_tracker_weights* grads, int32_t n_env,
                                  int32_t traj_len, const int32_t* users, const int32_t* traj_act,
                                  const float* traj_rew, const int32_t* ep_len, const float* dense_user,
                                  const float* dense_item, int32_t n_tok, const int32_t* tok_slot,
                                  const int32_t* env_off, int32_t max_ep_len, const float* d_obs, float* obs_check,
                                  void* workspace, int64_t workspace_bytes, int32_t phase, void* stream) {
  if (phase == 0) phase = 3;
  if (phase < 1 || phase > 3) {
    cirs_set_error("cirs_tracker_train: phase must be 0 / 3 (forward + backward), 1 (forward only) or 2 (backward only)");
    return CIRS_ERR_ARG;
  }
  if (!w || !grads || !traj_rew || !ep_len || !workspace || n_env < 0 || traj_len < 1) {
    cirs_set_error("cirs_tracker_train: null argument");
    return CIRS_ERR_ARG;
  }
  if ((w->emb_user && !users) || (!w->emb_user && !dense_user) || (w->emb_item && !traj_act) ||
      (!w->emb_item && !dense_item)) {
    cirs_set_error("cirs_tracker_train: token inputs missing (ids for embedding tables, dense features otherwise)");
    return CIRS_ERR_ARG;
  }
  if (w->d % w->nhead != 0 || w->d > 32 * LN_MAXC || w->nlayers > CIRS_MAX_LAYERS || w->d_item_in != w->d ||
      (w->emb_user && w->d_user_in != w->d)) {
    cirs_set_error("cirs_tracker_train: unsupported shape");
    return CIRS_ERR_ARG;
  }
  if ((tok_slot == nullptr) != (env_off == nullptr) || n_tok < 0) {
    cirs_set_error("cirs_tracker_train: tok_slot and env_off must be given together");
    return CIRS_ERR_ARG;
  }
  const int64_t n_rows = tok_slot ? (int64_t)n_tok : (int64_t)n_env * traj_len;
  if (cirs_tracker_train_workspace_bytes(w, n_env, n_rows) > workspace_bytes) {
    cirs_set_error("cirs_tracker_train: workspace too small");
    return CIRS_ERR_ARG;
  }
  if (n_env == 0 || n_rows == 0) return CIRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const RowMap map{traj_len, tok_slot, env_off};
  const int B = n_env, L = traj_len, M = (int)n_rows, d = w->d, dhid = w->d_hid, S = w->dim_state, nl = w->nlayers;
  const int ldd = up32(d), ld3 = up32(3 * d), ldh = up32(dhid), lds = up32(S), dui = w->d_user_in;
  const int nh = w->nhead, dh = d / nh;
  if (tok_slot && fused_enabled()) {
    const FusedPlan P = fused_plan(*w, max_ep_len > 0 && max_ep_len <= L ? max_ep_len : L, M, B);
    // one CTA carries a chunk through ~100 dependent stages: a latency design that wins while the chunks fit a few
    // waves of the 148 SMs (Kuaishou: 30-200 chunks); with thousands of chunks (VirtualTaobao at 2048 x 50 tokens) the
    // throughput-oriented layer-by-layer launches are faster (measured 5.0 vs 9.7 ms)
    if (P.ok && (M + P.q - 1) / P.q <= 4 * 148)
      return fused_train(*w, *grads, B, L, M, users, traj_act, traj_rew, ep_len, dense_user, dense_item, tok_slot,
                         env_off, d_obs, obs_check, reinterpret_cast<float*>(workspace), P, phase, st);
  }
  // layer-by-layer launches: the early forward-only call is a no-op, the later call runs the whole pass
  if (phase == 1 && !obs_check) return CIRS_OK;
  Bufs b = carve(reinterpret_cast<float*>(workspace), B, M, d, dhid, nl, dui);
  const size_t att_smem = attn_smem_bytes(L, dh, ATT_WARPS);
  if (att_smem > 200 * 1024) {
    cirs_set_error("cirs_tracker_train: sequence too long for the attention kernel's shared memory");
    return CIRS_ERR_ARG;
  }
  if (att_smem > 48 * 1024) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem);
    cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem);
  }
  const cirs_tracker_weights& W = *w;
  const cirs_tracker_weights& G = *grads;

  // ================= forward
  CIRS_LAUNCH(gather_inputs_kernel, M, 64, 0, st, W, B, map, users, traj_act, traj_rew, ep_len, dense_user, dense_item, b.u, b.in);
  CIRS_CHECK_LAUNCH();
  linear_fwd(b.u, dui, W.user_wt, ldd, W.user_b, b.tok0, d, B, d, dui, 0, nullptr, 0, st);           // ffn_user
  linear_fwd(b.in, 1 + d, W.gate_wt, ldd, W.gate_b, b.g, d, M, d, 1 + d, 0, nullptr, 0, st);          // fnn_gate (pre-act)
  CIRS_LAUNCH(token_fwd_kernel, M, 64, 0, st, W, map, ep_len, b.in, b.tok0, b.g, b.x0);
  CIRS_CHECK_LAUNCH();
  const float* x = b.x0;
  for (int l = 0; l < nl; ++l) {
    const cirs_encoder_layer& Y = W.layer[l];
    LayerBufs& y = b.layer[l];
    linear_fwd(x, d, Y.in_wt, ld3, Y.in_b, y.qkv, 3 * d, M, 3 * d, d, 0, nullptr, 0, st);
    CIRS_LAUNCH(attn_fwd_kernel, dim3(B, nh), ATT_WARPS * 32, att_smem, st, map, d, nh, ep_len, y.qkv, y.o);
    CIRS_CHECK_LAUNCH();
    linear_fwd(y.o, d, Y.out_wt, ldd, Y.out_b, y.r1, d, M, d, d, 0, x, d, st);                         // r1 = x + attn
    CIRS_LAUNCH(ln_fwd_kernel, (M + 7) / 8, 256, 0, st, M, d, y.r1, Y.n1_w, Y.n1_b, y.x1, y.st1);
    CIRS_CHECK_LAUNCH();
    linear_fwd(y.x1, d, Y.l1_wt, ldh, Y.l1_b, y.h, dhid, M, dhid, d, 1, nullptr, 0, st);
    linear_fwd(y.h, dhid, Y.l2_wt, ldd, Y.l2_b, y.r2, d, M, d, dhid, 0, y.x1, d, st);                  // r2 = x1 + ffn
    CIRS_LAUNCH(ln_fwd_kernel, (M + 7) / 8, 256, 0, st, M, d, y.r2, Y.n2_w, Y.n2_b, y.x2, y.st2);
    CIRS_CHECK_LAUNCH();
    x = y.x2;
  }
  if (obs_check)   // decoded states, scattered to their buffer slots
    launch_gemm<64, 64, 32, 4>(RowMajorA{x, d, nullptr}, RowMajorB{W.dec_wt, lds, nullptr},
                               StoreEp{obs_check, S, W.dec_b, 0, tok_slot, nullptr, 0, nullptr, 0}, M, S, d, 1, nullptr,
                               st, "tracker_linear_fwd_gemm");
  CIRS_CHECK_LAUNCH();
  if (!d_obs) return CIRS_OK;

  // ================= backward
  const int ln_grid = min((M + 7) / 8, 148 * 4);
  // decoder: d_obs rows are gathered from their buffer slots
  launch_gemm<64, 64, 32, 4>(ColMajorA{x, d, nullptr}, RowMajorB{d_obs, S, tok_slot}, AtomicEp{G.dec_wt, lds}, d, S, M,
                             splitk(1, M), G.dec_b, st, "tracker_linear_dw_gemm");
  launch_gemm<64, 64, 32, 4>(RowMajorA{d_obs, S, tok_slot}, ColMajorB{W.dec_wt, lds},
                             StoreEp{b.da, d, nullptr, 0, nullptr, nullptr, 0, nullptr, 0}, M, d, S, 1, nullptr, st,
                             "tracker_linear_dx_gemm");                                               // da = dX_last
  for (int l = nl - 1; l >= 0; --l) {
    const cirs_encoder_layer& Y = W.layer[l];
    const cirs_encoder_layer& Gy = G.layer[l];
    LayerBufs& y = b.layer[l];
    const float* xin = l == 0 ? b.x0 : b.layer[l - 1].x2;
    CIRS_LAUNCH(ln_bwd_kernel, ln_grid, 256, 0, st, M, d, b.da, y.r2, y.st2, Y.n2_w, b.db, Gy.n2_w, Gy.n2_b);  // db = dR2
    CIRS_CHECK_LAUNCH();
    fork_join(st, [&](cudaStream_t s2) { linear_bwd_w(y.h, dhid, b.db, d, Gy.l2_wt, ldd, Gy.l2_b, M, d, dhid, s2); },
              [&](cudaStream_t s1) { linear_bwd_x(b.db, d, Y.l2_wt, ldd, b.dh, dhid, M, d, dhid, y.h, dhid, nullptr, 0, s1); });  // dh (relu mask)
    fork_join(st, [&](cudaStream_t s2) { linear_bwd_w(y.x1, d, b.dh, dhid, Gy.l1_wt, ldh, Gy.l1_b, M, dhid, d, s2); },
              [&](cudaStream_t s1) { linear_bwd_x(b.dh, dhid, Y.l1_wt, ldh, b.da, d, M, dhid, d, nullptr, 0, b.db, d, s1); });    // da = dX1
    CIRS_LAUNCH(ln_bwd_kernel, ln_grid, 256, 0, st, M, d, b.da, y.r1, y.st1, Y.n1_w, b.db, Gy.n1_w, Gy.n1_b);  // db = dR1
    CIRS_CHECK_LAUNCH();
    fork_join(st, [&](cudaStream_t s2) { linear_bwd_w(y.o, d, b.db, d, Gy.out_wt, ldd, Gy.out_b, M, d, d, s2); },
              [&](cudaStream_t s1) { linear_bwd_x(b.db, d, Y.out_wt, ldd, b.dtmp, d, M, d, d, nullptr, 0, nullptr, 0, s1); });    // dtmp = dO
    CIRS_LAUNCH(attn_bwd_kernel, dim3(B, nh), ATT_WARPS * 32, att_smem, st, map, d, nh, ep_len, y.qkv, b.dtmp, b.dqkv);
    CIRS_CHECK_LAUNCH();
    fork_join(st, [&](cudaStream_t s2) { linear_bwd_w(xin, d, b.dqkv, 3 * d, Gy.in_wt, ld3, Gy.in_b, M, 3 * d, d, s2); },
              [&](cudaStream_t s1) { linear_bwd_x(b.dqkv, 3 * d, Y.in_wt, ld3, b.da, d, M, 3 * d, d, nullptr, 0, b.db, d, s1); });  // da = dX_in
  }
  // tokens: dZ -> dtmp, direct item gradient -> db
  CIRS_LAUNCH(token_bwd_kernel, M, 64, 0, st, W, map, ep_len, b.in, b.g, b.da, b.dtok0, b.dtmp, b.db);
  CIRS_CHECK_LAUNCH();
  linear_bwd_w(b.in, 1 + d, b.dtmp, d, G.gate_wt, ldd, G.gate_b, M, d, 1 + d, st);
  linear_bwd_w(b.u, dui, b.dtok0, d, G.user_wt, ldd, G.user_b, B, d, dui, st);
  if (G.emb_item || G.emb_user) {
    linear_bwd_x(b.dtmp, d, W.gate_wt, ldd, b.din, 1 + d, M, d, 1 + d, nullptr, 0, nullptr, 0, st);
    linear_bwd_x(b.dtok0, d, W.user_wt, ldd, b.du, dui, B, d, dui, nullptr, 0, nullptr, 0, st);
    CIRS_LAUNCH(emb_scatter_kernel, M, 64, 0, st, W, G, B, map, users, traj_act, ep_len, b.din, b.db, b.du);
  }
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
