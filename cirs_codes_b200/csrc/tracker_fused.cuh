// K6 fused: the tracker's whole training pass -- token construction, nlayers x (in-proj, causal attention, out-proj,
// LayerNorm, FFN, LayerNorm), decoder, and the complete backward to the tokens and embedding rows -- for a CHUNK of
// whole environments inside ONE CTA, activations staged through shared memory; plus ONE grouped launch for all the
// weight-gradient contractions.  Two launches replace ~50 dependent ones (skinny GEMMs, attention, LayerNorm, token
// kernels) whose launch + drain latency was the run time at these sizes (a few thousand tokens, d = 32 .. 128).
//
// Chunking: tokens are stored compactly env-major (row i = i-th stored transition, environment e owns rows
// env_off[e] .. env_off[e+1]).  Chunk c takes every environment whose FIRST row lies in [c q, (c+1) q), q = TM - (longest
// episode) + 1, so a chunk never exceeds TM rows and no environment is split: sequences never interact, hence no
// inter-CTA synchronisation.  Everything a later stage needs again (layer inputs, qkv, attention output, pre-LayerNorm
// sums and statistics, FFN hidden) goes to the workspace in global memory (L2 resident: a few MB) because the backward
// pass and the weight-gradient launch read it; the chain of stages itself runs out of shared memory.
//
// Parameter gradients: LayerNorm weights / biases and embedding rows by atomics from the chunk kernel; every Linear's
// weight and bias gradient (gW = X^T dY over ALL tokens) by the grouped split-K tile GEMM (gemm.cuh core).
#pragma once
#include "gemm.cuh"
#include "../../include/cirs_b200.h"

namespace cirs_k6 {
using namespace cirs;

constexpr int NT = 256;
constexpr int WS_K = 128, WS_N = 128, WS_LD = WS_N + 4;   // weight stage: [128 k][128 n] block of an operand

__host__ __device__ inline int up4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline int64_t al64(int64_t x) { return (x + 63) & ~(int64_t)63; }

struct LayerSave {           // per layer, [M, .] row-major in the workspace
  float *qkv, *o, *r1, *st1, *x1, *h, *r2, *st2, *x2;   // forward
  float *dqkv, *dr1, *dh, *dr2;                          // backward (operands of the weight-gradient GEMMs)
};
struct Save {
  float *u, *in, *g, *x0, *dz, *dtok0;    // u [M, dui] (rows of position 0, else 0), in [M, 1+d], g / x0 / dz / dtok0 [M, d]
  LayerSave layer[CIRS_MAX_LAYERS];
  int64_t total;
};
inline Save carve(float* base, int64_t M, int d, int dhid, int nl, int dui) {
  Save s;
  int64_t off = 0;
  auto take = [&](int64_t n) { float* r = base ? base + off : nullptr; off += al64(n); return r; };
  s.u = take(M * dui); s.in = take(M * (1 + d)); s.g = take(M * d); s.x0 = take(M * d); s.dz = take(M * d);
  s.dtok0 = take(M * d);
  for (int l = 0; l < nl; ++l) {
    LayerSave& y = s.layer[l];
    y.qkv = take(M * 3 * d); y.o = take(M * d); y.r1 = take(M * d); y.st1 = take(M * 2); y.x1 = take(M * d);
    y.h = take(M * dhid); y.r2 = take(M * d); y.st2 = take(M * 2); y.x2 = take(M * d);
    y.dqkv = take(M * 3 * d); y.dr1 = take(M * d); y.dh = take(M * dhid); y.dr2 = take(M * d);
  }
  s.total = off;
  return s;
}

struct Args {
  cirs_tracker_weights W, G;
  Save S;
  int n_env, L, M, q;                 // q = first-row quantum of a chunk
  const int32_t *users, *act, *ep_len, *tok_slot, *env_off;
  const float *rew, *dense_user, *dense_item, *d_obs;
  float* obs_check;
  int ldx, ldb;                       // shared tile strides: [TM][ldx] (width d), [TM][ldb] (width max(3d, dhid, 1+dui))
  const int32_t* chunk_e0;            // greedy plan (chunk_plan_kernel): chunk c = environments chunk_e0[c] .. chunk_e0[c+1]; [0] = count
  int phase;                          // 3 = forward + backward in one launch, 1 = forward only (activations -> workspace),
                                      // 2 = backward only (after a phase-1 launch on the same workspace)
};

// ---- Y[m x N] = X[m x K] . Wop[K x N] out of shared memory; W streamed through the stage in 128 x 128 blocks.
//   TRANS = false: Wop[k][n] = W[k * ldw + n]  (forward: Wt is k-major)    TRANS = true: Wop[k][n] = W[n * ldw + k]
// Thread (ty = tid / 32, tx = tid % 32) owns rows ty * RPT .. + RPT and columns 4 tx .. + 3 of each 128-column block.
// ep(row, col, value) is called for row < m, col < N.  X rows >= m and columns >= K must be finite (tiles are zeroed).
template <int TM, bool TRANS, class Ep>
__device__ __forceinline__ void tile_linear(const float* Xs, int ldx, const float* __restrict__ W, int ldw, int m,
                                            int N, int K, float* stage, Ep ep) {
  constexpr int RPT = TM / 8;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  for (int n0 = 0; n0 < N; n0 += WS_N) {
    const int nb = min(WS_N, N - n0);
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int k0 = 0; k0 < K; k0 += WS_K) {
      const int kb = min(WS_K, K - k0), kb4 = up4(kb);
      __syncthreads();   // the stage (and, on the first pass, the X tile) is free / complete
      if (!TRANS) {
        const int nb4 = up4(nb);
        for (int i = tid; i < kb4 * (nb4 >> 2); i += NT) {
          const int k = i / (nb4 >> 2), c = (i % (nb4 >> 2)) << 2;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < kb) {   // ldw is a multiple of 32 floats and n0 of 128: aligned, and in bounds up to the padded width
            const float* src = W + (size_t)(k0 + k) * ldw + n0 + c;
            if (n0 + c + 3 < ldw) v = __ldg(reinterpret_cast<const float4*>(src));
            else { v.x = __ldg(src); if (n0 + c + 1 < ldw) v.y = __ldg(src + 1); if (n0 + c + 2 < ldw) v.z = __ldg(src + 2); }
          }
          *reinterpret_cast<float4*>(stage + k * WS_LD + c) = v;
        }
      } else {
        for (int i = tid; i < nb * kb4; i += NT) {
          const int n = i / kb4, k = i % kb4;
          stage[k * WS_LD + n] = k < kb ? __ldg(W + (size_t)(n0 + n) * ldw + k0 + k) : 0.f;
        }
      }
      __syncthreads();
      if (4 * tx < nb) {
        const float* xr = Xs + (size_t)(ty * RPT) * ldx + k0;
#pragma unroll 2
        for (int k = 0; k < kb4; k += 4) {
          float4 b[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(stage + (k + j) * WS_LD + 4 * tx);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(xr + (size_t)i * ldx + k);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][0] = fmaf(av[j], b[j].x, acc[i][0]);
              acc[i][1] = fmaf(av[j], b[j].y, acc[i][1]);
              acc[i][2] = fmaf(av[j], b[j].z, acc[i][2]);
              acc[i][3] = fmaf(av[j], b[j].w, acc[i][3]);
            }
          }
        }
      }
    }
    if (4 * tx < nb) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = ty * RPT + i;
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + 4 * tx + j;
          if (n < N) ep(r, n, acc[i][j]);
        }
      }
    }
  }
  __syncthreads();   // the outputs are visible to the whole CTA
}

// ---- the same product with the weight matrix RESIDENT in shared memory (rows k-major like the global layout, row
// stride ldw = padded width + 4 floats, so both orientations read conflict-free 16-byte words):
//   TRANS = false: thread owns columns 4 tx .. + 3 of each 128-column block  (Wop[k][n] = Ws[k * ldw + n])
//   TRANS = true : thread owns the n-indices tx + 32 j, j < 4, and reads Ws[n * ldw + k .. k + 3] as one float4: lane
//                  stride = one row = 4 (mod 32) floats                     (Wop[k][n] = Ws[n * ldw + k])
// No weight traffic, no barrier inside: one barrier in front (the X tile is complete) and one behind.
template <int TM, bool TRANS, class Ep>
__device__ __forceinline__ void tile_linear_res(const float* Xs, int ldx, const float* Ws, int ldw, int m, int N, int K,
                                                Ep ep) {
  constexpr int RPT = TM / 8;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int K4 = up4(K);
  const float* xr = Xs + (size_t)(ty * RPT) * ldx;
  __syncthreads();
  for (int n0 = 0; n0 < N; n0 += 128) {
    float acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    if (!TRANS) {
      if (n0 + 4 * tx < N) {
        const float* wp = Ws + n0 + 4 * tx;
#pragma unroll 2
        for (int k = 0; k < K4; k += 4) {
          float4 b[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(wp + (size_t)(k + j) * ldw);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(xr + (size_t)i * ldx + k);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][0] = fmaf(av[j], b[j].x, acc[i][0]);
              acc[i][1] = fmaf(av[j], b[j].y, acc[i][1]);
              acc[i][2] = fmaf(av[j], b[j].z, acc[i][2]);
              acc[i][3] = fmaf(av[j], b[j].w, acc[i][3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = ty * RPT + i;
          if (r >= m) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + 4 * tx + j < N) ep(r, n0 + 4 * tx + j, acc[i][j]);
        }
      }
    } else {
      const float* wp[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ok[j] = n0 + tx + 32 * j < N;
        wp[j] = Ws + (size_t)(ok[j] ? n0 + tx + 32 * j : 0) * ldw;
      }
      if (ok[0]) {
#pragma unroll 2
        for (int k = 0; k < K4; k += 4) {
          float4 w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(wp[j] + k);
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(xr + (size_t)i * ldx + k);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[i][j] = fmaf(a.x, w[j].x, fmaf(a.y, w[j].y, fmaf(a.z, w[j].z, fmaf(a.w, w[j].w, acc[i][j]))));
          }
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = ty * RPT + i;
          if (r >= m) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ok[j]) ep(r, n0 + tx + 32 * j, acc[i][j]);
        }
      }
    }
  }
  __syncthreads();
}

// the tracker's Linear weights: id -> (global pointer, its row stride, rows, resident copy, the copy's row stride)
enum { M_USER = 0, M_GATE = 1, M_LAYER0 = 2, M_PER_LAYER = 4 };   // per layer: in, out, l1, l2; then the decoder
constexpr int MAX_MATS = 2 + 4 * CIRS_MAX_LAYERS + 1;
struct Mat { const float* g; const float* s; int ldw, rows, lds; };
__device__ __forceinline__ int m_in(int l) { return M_LAYER0 + M_PER_LAYER * l; }
__device__ __forceinline__ int m_out(int l) { return M_LAYER0 + M_PER_LAYER * l + 1; }
__device__ __forceinline__ int m_l1(int l) { return M_LAYER0 + M_PER_LAYER * l + 2; }
__device__ __forceinline__ int m_l2(int l) { return M_LAYER0 + M_PER_LAYER * l + 3; }

// floats of shared memory the resident copies need (rows rounded up to 4, row stride = padded width + 4)
__host__ __device__ inline int resident_floats(const cirs_tracker_weights& W) {
  const int d = W.d, ldd = (d + 31) & ~31, ld3 = (3 * d + 31) & ~31, ldh = (W.d_hid + 31) & ~31,
            lds = (W.dim_state + 31) & ~31;
  int n = up4(W.d_user_in) * (ldd + 4) + up4(1 + W.d_item_in) * (ldd + 4) + up4(d) * (lds + 4);
  n += W.nlayers * (up4(d) * (ld3 + 4) + up4(d) * (ldd + 4) + up4(d) * (ldh + 4) + up4(W.d_hid) * (ldd + 4));
  return n;
}

template <int TM, bool TRANS, bool RES, class Ep>
__device__ __forceinline__ void lin(const float* Xs, int ldx, const Mat& M, int m, int N, int K, float* stage, Ep ep) {
  if (RES) tile_linear_res<TM, TRANS>(Xs, ldx, M.s, M.lds, m, N, K, ep);
  else tile_linear<TM, TRANS>(Xs, ldx, M.g, M.ldw, m, N, K, stage, ep);
}

// one warp per row: X = LayerNorm(R) * w + b (biased variance, eps 1e-5); statistics to ST (global)
__device__ __forceinline__ void ln_fwd_rows(const float* Rs, float* Xs, int ldx, int m, int d, const float* __restrict__ w,
                                            const float* __restrict__ b, float* __restrict__ xg, float* __restrict__ st,
                                            size_t row0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < m; r += NT / 32) {
    const float* rr = Rs + (size_t)r * ldx;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += rr[c];
    const float mu = warp_sum(s) / d;
    float q = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = rr[c] - mu; q += v * v; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / d + 1e-5f);
    for (int c = lane; c < d; c += 32) {
      const float v = (rr[c] - mu) * rstd * __ldg(w + c) + __ldg(b + c);
      Xs[(size_t)r * ldx + c] = v;
      xg[(row0 + r) * d + c] = v;
    }
    if (lane == 0) { st[2 * (row0 + r)] = mu; st[2 * (row0 + r) + 1] = rstd; }
  }
  __syncthreads();
}

// dR = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)), dxhat = dY w;  gw += sum dY xhat, gb += sum dY (atomics per CTA).
// dYs, Rs (the saved pre-LayerNorm sums, loaded by the caller) -> dRs (may alias dYs); acc = 2 * d floats of scratch.
__device__ __forceinline__ void ln_bwd_rows(const float* dYs, const float* Rs, float* dRs, int ldx, int m, int d,
                                            const float* __restrict__ st, size_t row0, const float* __restrict__ w,
                                            float* gw, float* gb, float* acc, float* __restrict__ drg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < 2 * d; c += NT) acc[c] = 0.f;
  __syncthreads();
  for (int r = warp; r < m; r += NT / 32) {
    const float mu = st[2 * (row0 + r)], rstd = st[2 * (row0 + r) + 1];
    const float* rr = Rs + (size_t)r * ldx;
    const float* dy = dYs + (size_t)r * ldx;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float xh = (rr[c] - mu) * rstd, dyc = dy[c], dxh = dyc * __ldg(w + c);
      s1 += dxh; s2 += dxh * xh;
      atomicAdd(acc + c, dyc * xh);
      atomicAdd(acc + d + c, dyc);
    }
    s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
    for (int c = lane; c < d; c += 32) {
      const float xh = (rr[c] - mu) * rstd, dxh = dy[c] * __ldg(w + c);
      const float v = rstd * (dxh - s1 - xh * s2);
      dRs[(size_t)r * ldx + c] = v;
      drg[(row0 + r) * d + c] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += NT) {
    atomicAdd(gw + c, acc[c]);
    atomicAdd(gb + c, acc[d + c]);
  }
  __syncthreads();
}

// rows of a [M, width] global matrix -> shared tile (zero beyond m / width)
template <int TM>
__device__ __forceinline__ void load_rows(float* Ts, int ld, const float* __restrict__ g, size_t row0, int m, int width,
                                          int gld) {
  for (int i = threadIdx.x; i < TM * ld; i += NT) {
    const int r = i / ld, c = i % ld;
    Ts[i] = (r < m && c < width) ? g[(row0 + r) * gld + c] : 0.f;
  }
  __syncthreads();
}

// ---- causal multi-head attention of a chunk, thread per (row, head) task.  QKV tile [TM][ldb] = [q | k | v].
// P (>= Lmax * TM * nh floats) receives the normalised probabilities, P[(j - s0) * ntask + task].
template <int TM>
__device__ __forceinline__ void attn_forward(const float* QKV, int ldb, float* Os, int ldx, float* P,
                                             const int* rstart, int m, int d, int nh, int dh, float scale,
                                             float* __restrict__ og, size_t row0) {
  const int ntask = m * nh;
  for (int task = threadIdx.x; task < ntask; task += NT) {
    const int r = task / nh, h = task % nh, s0 = rstart[r], n = r - s0 + 1;
    const float* q = QKV + (size_t)r * ldb + h * dh;
    float mx = -INFINITY;
    for (int jj = 0; jj < n; ++jj) {
      const float* k = QKV + (size_t)(s0 + jj) * ldb + d + h * dh;
      float sc = 0.f;
      for (int c = 0; c < dh; ++c) sc = fmaf(q[c] * scale, k[c], sc);
      P[jj * ntask + task] = sc;
      mx = fmaxf(mx, sc);
    }
    float z = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const float ex = expf(P[jj * ntask + task] - mx);
      P[jj * ntask + task] = ex;
      z += ex;
    }
    const float inv = 1.0f / z;
    for (int jj = 0; jj < n; ++jj) P[jj * ntask + task] *= inv;
    for (int c0 = 0; c0 < dh; c0 += 8) {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      for (int jj = 0; jj < n; ++jj) {
        const float p = P[jj * ntask + task];
        const float* v = QKV + (size_t)(s0 + jj) * ldb + 2 * d + h * dh + c0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < dh) acc[u] = fmaf(p, v[u], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < dh) {
          Os[(size_t)r * ldx + h * dh + c0 + u] = acc[u];
          og[(row0 + r) * d + h * dh + c0 + u] = acc[u];
        }
    }
  }
  __syncthreads();
}

// Backward.  dOs [TM][ldx] = d loss / d (attention output); writes dQKV tile [TM][ldb] (zeroed by the caller).
//   phase 1, task (query i, head): P_ij (recomputed as in the forward), D_i = sum_j P_ij dP_ij with dP_ij = <dO_i, v_j>,
//            dq_i = scale sum_j P_ij (dP_ij - D_i) k_j
//   phase 2, task (key j, head): dk_j = scale sum_{i >= j} P_ij (dP_ij - D_i) q_i,  dv_j = sum_{i >= j} P_ij dO_i
// dP is recomputed per block of 8 output columns instead of being stored (one dot product of dh terms).
template <int TM>
__device__ __forceinline__ void attn_backward(const float* QKV, float* dQKV, int ldb, const float* dOs, int ldx, float* P,
                                              float* dsum_s, const int* rstart, const int* renv,
                                              const int32_t* __restrict__ env_off, int m, int d, int nh, int dh,
                                              float scale) {
  const int ntask = m * nh;
  for (int task = threadIdx.x; task < ntask; task += NT) {
    const int r = task / nh, h = task % nh, s0 = rstart[r], n = r - s0 + 1;
    const float* q = QKV + (size_t)r * ldb + h * dh;
    const float* dO = dOs + (size_t)r * ldx + h * dh;
    float mx = -INFINITY;
    for (int jj = 0; jj < n; ++jj) {
      const float* k = QKV + (size_t)(s0 + jj) * ldb + d + h * dh;
      float sc = 0.f;
      for (int c = 0; c < dh; ++c) sc = fmaf(q[c] * scale, k[c], sc);
      P[jj * ntask + task] = sc;
      mx = fmaxf(mx, sc);
    }
    float z = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const float ex = expf(P[jj * ntask + task] - mx);
      P[jj * ntask + task] = ex;
      z += ex;
    }
    const float inv = 1.0f / z;
    float dsum = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const float p = P[jj * ntask + task] * inv;
      P[jj * ntask + task] = p;
      const float* v = QKV + (size_t)(s0 + jj) * ldb + 2 * d + h * dh;
      float dp = 0.f;
      for (int c = 0; c < dh; ++c) dp = fmaf(dO[c], v[c], dp);
      dsum = fmaf(p, dp, dsum);
    }
    dsum_s[task] = dsum;
    for (int c0 = 0; c0 < dh; c0 += 8) {
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      for (int jj = 0; jj < n; ++jj) {
        const float p = P[jj * ntask + task];
        const float* v = QKV + (size_t)(s0 + jj) * ldb + 2 * d + h * dh;
        const float* k = QKV + (size_t)(s0 + jj) * ldb + d + h * dh + c0;
        float dp = 0.f;
        for (int c = 0; c < dh; ++c) dp = fmaf(dO[c], v[c], dp);
        const float ds = p * (dp - dsum);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < dh) acc[u] = fmaf(ds, k[u], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < dh) dQKV[(size_t)r * ldb + h * dh + c0 + u] = acc[u] * scale;
    }
  }
  __syncthreads();
  for (int task = threadIdx.x; task < ntask; task += NT) {
    const int j = task / nh, h = task % nh, s0 = rstart[j];
    const int e = renv[j];
    const int end = s0 + (env_off[e + 1] - env_off[e]) - 1;   // last chunk-local row of this environment
    const int jj = j - s0;
    const float* k = QKV + (size_t)j * ldb + d + h * dh;
    const float* v = QKV + (size_t)j * ldb + 2 * d + h * dh;
    for (int c0 = 0; c0 < dh; c0 += 8) {
      float ak[8], av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) ak[u] = av[u] = 0.f;
      for (int i = j; i <= end; ++i) {
        const int ti = i * nh + h;
        const float p = P[jj * ntask + ti];
        const float* dO = dOs + (size_t)i * ldx + h * dh;
        const float* q = QKV + (size_t)i * ldb + h * dh;
        float dp = 0.f;
        for (int c = 0; c < dh; ++c) dp = fmaf(dO[c], v[c], dp);
        const float ds = p * (dp - dsum_s[ti]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < dh) {
            ak[u] = fmaf(ds, q[c0 + u], ak[u]);
            av[u] = fmaf(p, dO[c0 + u], av[u]);
          }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < dh) {
          dQKV[(size_t)j * ldb + d + h * dh + c0 + u] = ak[u] * scale;
          dQKV[(size_t)j * ldb + 2 * d + h * dh + c0 + u] = av[u];
        }
    }
  }
  __syncthreads();
}

// RES: every Linear weight of the tracker is copied into shared memory once per CTA (d <= 32: ~120 KB) and all the
// stage products run from there (tile_linear_res); otherwise weight blocks stream through the stage (tile_linear).
template <int TM, bool RES>
__global__ void __launch_bounds__(NT, 1) tracker_chunk_kernel(const Args A) {
  extern __shared__ __align__(16) float sm[];
  const cirs_tracker_weights& W = A.W;
  const cirs_tracker_weights& G = A.G;
  const int d = W.d, nh = W.nhead, dh = d / nh, dhid = W.d_hid, S = W.dim_state, nl = W.nlayers, dui = W.d_user_in;
  const int ldd = (d + 31) & ~31, ld3 = (3 * d + 31) & ~31, ldh = (dhid + 31) & ~31, lds = (S + 31) & ~31;
  const int ldx = A.ldx, ldb = A.ldb, L = A.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared memory: three [TM][ldx] tiles, two [TM][ldb] tiles, the weight stage, per-row bookkeeping, small scratch
  float* xs = sm;
  float* ys = xs + TM * ldx;
  float* zs = ys + TM * ldx;
  float* big = zs + TM * ldx;
  float* big2 = big + TM * ldb;
  float* stage = big2 + TM * ldb;                // weight stage (streaming) / attention probabilities
  float* stat = stage + (RES ? TM * TM * nh : WS_K * WS_LD);   // [3][TM][nh] attention row statistics (backward)
  float* pw = stat + 3 * TM * nh;                // [8 warps][2][TM] probability / dS scratch
  float* lnacc = pw + 8 * 2 * TM;                // [2 d]
  int* rpos = reinterpret_cast<int*>(lnacc + 2 * ((d + 3) & ~3));   // [TM] position of the row inside its episode
  int* renv = rpos + TM;                         // [TM] environment of the row
  int* rstart = renv + TM;                       // [TM] chunk-local first row of the row's environment
  __shared__ Mat mats[MAX_MATS];
  if (tid == 0) {
    float* wres = reinterpret_cast<float*>(rstart + TM);
    wres = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wres) + 15) & ~(uintptr_t)15);
    auto put = [&](int id, const float* g, int ldw, int rows) {
      Mat& M = mats[id];
      M.g = g; M.ldw = ldw; M.rows = rows; M.lds = ldw + 4; M.s = wres;
      if (RES) wres += up4(rows) * (ldw + 4);
    };
    put(M_USER, W.user_wt, ldd, dui);
    put(M_GATE, W.gate_wt, ldd, 1 + d);
    for (int l = 0; l < nl; ++l) {
      put(m_in(l), W.layer[l].in_wt, ld3, d);
      put(m_out(l), W.layer[l].out_wt, ldd, d);
      put(m_l1(l), W.layer[l].l1_wt, ldh, d);
      put(m_l2(l), W.layer[l].l2_wt, ldd, dhid);
    }
    put(M_LAYER0 + M_PER_LAYER * nl, W.dec_wt, lds, d);
  }
  __syncthreads();
  const int M_DEC = M_LAYER0 + M_PER_LAYER * nl;
  if (RES) {   // resident copies: [up4(rows)][ldw + 4], zero beyond the matrix
    for (int id = 0; id <= M_DEC; ++id) {
      const Mat M = mats[id];
      float* dst = const_cast<float*>(M.s);
      const int r4 = up4(M.rows), w4 = M.ldw >> 2;
      for (int i = tid; i < r4 * (w4 + 1); i += NT) {
        const int r = i / (w4 + 1), c = (i % (w4 + 1)) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < M.rows && c < M.ldw) v = __ldg(reinterpret_cast<const float4*>(M.g + (size_t)r * M.ldw + c));
        *reinterpret_cast<float4*>(dst + (size_t)r * M.lds + c) = v;
      }
    }
    __syncthreads();
  }

  const float sq = sqrtf((float)d), scale = 1.0f / sqrtf((float)dh);
  const int n_chunks = A.chunk_e0 ? A.chunk_e0[0] : (A.M + A.q - 1) / A.q;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    // ---- the chunk's environments: the greedy plan's range, or (no plan) those whose first row lies in
    // [chunk q, (chunk + 1) q)  (binary search over env_off)
    __shared__ int s_e0, s_e1;
    if (tid < 2) {
      int lo = 0;
      if (A.chunk_e0) {
        lo = A.chunk_e0[1 + chunk + tid];
      } else {
        const int target = (chunk + tid) * A.q;
        int hi = A.n_env;            // first e with env_off[e] >= target
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (A.env_off[mid] < target) lo = mid + 1; else hi = mid;
        }
      }
      // environments with no stored transition own no row: skip them at the lower end
      if (tid == 0) s_e0 = lo; else s_e1 = lo;
    }
    __syncthreads();
    const int e0 = s_e0, e1 = s_e1;
    if (e0 >= e1) { __syncthreads(); continue; }
    const size_t row0 = (size_t)A.env_off[e0];
    const int m = A.env_off[e1] - A.env_off[e0];
    if (m <= 0 || m > TM) { __syncthreads(); continue; }   // m > TM is excluded by the host's choice of q
    for (int r = tid; r < TM; r += NT) {
      int e = -1, p = 0, st = 0;
      if (r < m) {
        const int slot = A.tok_slot[row0 + r];
        e = slot / L; p = slot % L;
        st = A.env_off[e] - (int)row0;
      }
      renv[r] = e; rpos[r] = p; rstart[r] = st;
    }
    for (int i = tid; i < 3 * TM * ldx; i += NT) xs[i] = 0.f;
    for (int i = tid; i < 2 * TM * ldb; i += NT) big[i] = 0.f;
    __syncthreads();

    // ================================================= forward
    if (A.phase & 1) {
    // ---- tokens.  user rows: big = u (position 0), tok = ffn_user(u);  action rows: big = [rew ; a], tok = sigmoid(gate) a
    for (int i = tid; i < m * dui; i += NT) {
      const int r = i / dui, c = i % dui;
      float v = 0.f;
      if (rpos[r] == 0) {
        const int e = renv[r];
        v = W.emb_user ? __ldg(W.emb_user + (size_t)A.users[e] * d + c) : A.dense_user[(size_t)e * dui + c];
      }
      big[(size_t)r * ldb + c] = v;
      A.S.u[(row0 + r) * dui + c] = v;
    }
    lin<TM, false, RES>(big, ldb, mats[M_USER], m, d, dui, stage, [&](int r, int c, float v) {
      if (rpos[r] == 0) zs[(size_t)r * ldx + c] = v + __ldg(W.user_b + c);
    });
    for (int i = tid; i < TM * ldb; i += NT) big[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < m * (1 + d); i += NT) {
      const int r = i / (1 + d), c = i % (1 + d);
      float v = 0.f;
      if (rpos[r] >= 1) {
        const size_t prev = (size_t)renv[r] * L + rpos[r] - 1;
        if (c == 0) v = A.rew[prev];
        else v = W.emb_item ? __ldg(W.emb_item + (size_t)A.act[prev] * d + c - 1) : A.dense_item[prev * d + c - 1];
      }
      big[(size_t)r * ldb + c] = v;
      A.S.in[(row0 + r) * (1 + d) + c] = v;
    }
    lin<TM, false, RES>(big, ldb, mats[M_GATE], m, d, 1 + d, stage, [&](int r, int c, float v) {
      float g = 0.f;
      if (rpos[r] >= 1) {
        g = 1.f / (1.f + expf(-(v + __ldg(W.gate_b + c))));
        zs[(size_t)r * ldx + c] = g * big[(size_t)r * ldb + 1 + c];
      }
      A.S.g[(row0 + r) * d + c] = g;
    });
    for (int i = tid; i < m * d; i += NT) {
      const int r = i / d, c = i % d;
      const float v = zs[(size_t)r * ldx + c] * sq + __ldg(W.pe + (size_t)rpos[r] * d + c);
      xs[(size_t)r * ldx + c] = v;
      A.S.x0[(row0 + r) * d + c] = v;
    }
    __syncthreads();

    for (int l = 0; l < nl; ++l) {
      const cirs_encoder_layer& Y = W.layer[l];
      const LayerSave& y = A.S.layer[l];
      // ---- qkv = x Win + b
      lin<TM, false, RES>(xs, ldx, mats[m_in(l)], m, 3 * d, d, stage, [&](int r, int c, float v) {
        v += __ldg(Y.in_b + c);
        big[(size_t)r * ldb + c] = v;
        y.qkv[(row0 + r) * 3 * d + c] = v;
      });
      // ---- causal attention inside each environment: one THREAD per (row, head) -- episodes are a handful of tokens
      // long, so a warp per task would idle 29 of its lanes.  Probabilities go to P[key offset][task] (the weight
      // stage is free between two linears), conflict-free because consecutive threads are consecutive tasks.
      attn_forward<TM>(big, ldb, ys, ldx, stage, rstart, m, d, nh, dh, scale, y.o, row0);
      // ---- r1 = x + o Wout + b;  x1 = LN1(r1)
      lin<TM, false, RES>(ys, ldx, mats[m_out(l)], m, d, d, stage, [&](int r, int c, float v) {
        v += __ldg(Y.out_b + c) + xs[(size_t)r * ldx + c];
        zs[(size_t)r * ldx + c] = v;
        y.r1[(row0 + r) * d + c] = v;
      });
      ln_fwd_rows(zs, ys, ldx, m, d, Y.n1_w, Y.n1_b, y.x1, y.st1, row0);
      // ---- h = relu(x1 W1 + b1);  r2 = x1 + h W2 + b2;  x2 = LN2(r2)
      lin<TM, false, RES>(ys, ldx, mats[m_l1(l)], m, dhid, d, stage, [&](int r, int c, float v) {
        v = fmaxf(v + __ldg(Y.l1_b + c), 0.f);
        big[(size_t)r * ldb + c] = v;
        y.h[(row0 + r) * dhid + c] = v;
      });
      lin<TM, false, RES>(big, ldb, mats[m_l2(l)], m, d, dhid, stage, [&](int r, int c, float v) {
        v += __ldg(Y.l2_b + c) + ys[(size_t)r * ldx + c];
        zs[(size_t)r * ldx + c] = v;
        y.r2[(row0 + r) * d + c] = v;
      });
      ln_fwd_rows(zs, xs, ldx, m, d, Y.n2_w, Y.n2_b, y.x2, y.st2, row0);
    }
    if (A.obs_check)   // decoded states at their buffer slots (tests)
      lin<TM, false, RES>(xs, ldx, mats[M_DEC], m, S, d, stage, [&](int r, int c, float v) {
        A.obs_check[(size_t)A.tok_slot[row0 + r] * S + c] = v + __ldg(W.dec_b + c);
      });
    }   // forward
    if (!A.d_obs || !(A.phase & 2)) { __syncthreads(); continue; }

    // ================================================= backward
    // ---- decoder: da = d_obs Wdec^T  (the weight gradient comes from the grouped launch)
    for (int i = tid; i < TM * ldb; i += NT) {
      const int r = i / ldb, c = i % ldb;
      big[i] = (r < m && c < S) ? A.d_obs[(size_t)A.tok_slot[row0 + r] * S + c] : 0.f;
    }
    lin<TM, true, RES>(big, ldb, mats[M_DEC], m, d, S, stage,
                          [&](int r, int c, float v) { xs[(size_t)r * ldx + c] = v; });   // xs = da
    for (int l = nl - 1; l >= 0; --l) {
      const cirs_encoder_layer& Y = W.layer[l];
      const cirs_encoder_layer& Gy = G.layer[l];
      const LayerSave& y = A.S.layer[l];
      // ---- LN2: zs = dR2
      load_rows<TM>(ys, ldx, y.r2, row0, m, d, d);
      ln_bwd_rows(xs, ys, zs, ldx, m, d, y.st2, row0, Y.n2_w, Gy.n2_w, Gy.n2_b, lnacc, y.dr2);
      // ---- dh = (dR2 W2^T) [h > 0]  -> big;   dx1 = dh W1^T + dR2 -> ys
      for (int i = tid; i < TM * ldb; i += NT) big[i] = 0.f;
      lin<TM, true, RES>(zs, ldx, mats[m_l2(l)], m, dhid, d, stage, [&](int r, int c, float v) {
        if (!(y.h[(row0 + r) * dhid + c] > 0.f)) v = 0.f;
        big[(size_t)r * ldb + c] = v;
        y.dh[(row0 + r) * dhid + c] = v;
      });
      lin<TM, true, RES>(big, ldb, mats[m_l1(l)], m, d, dhid, stage, [&](int r, int c, float v) {
        ys[(size_t)r * ldx + c] = v + zs[(size_t)r * ldx + c];
      });
      // ---- LN1: zs = dR1
      load_rows<TM>(xs, ldx, y.r1, row0, m, d, d);
      ln_bwd_rows(ys, xs, zs, ldx, m, d, y.st1, row0, Y.n1_w, Gy.n1_w, Gy.n1_b, lnacc, y.dr1);
      // ---- dO = dR1 Wout^T -> ys
      lin<TM, true, RES>(zs, ldx, mats[m_out(l)], m, d, d, stage,
                            [&](int r, int c, float v) { ys[(size_t)r * ldx + c] = v; });
      // ---- attention backward: qkv -> big, dqkv -> big2
      load_rows<TM>(big, ldb, y.qkv, row0, m, 3 * d, 3 * d);
      for (int i = tid; i < TM * ldb; i += NT) big2[i] = 0.f;
      __syncthreads();
      attn_backward<TM>(big, big2, ldb, ys, ldx, stage, stat, rstart, renv, A.env_off, m, d, nh, dh, scale);
      for (int i = tid; i < m * 3 * d; i += NT) {
        const int r = i / (3 * d), c = i % (3 * d);
        y.dqkv[(row0 + r) * 3 * d + c] = big2[(size_t)r * ldb + c];
      }
      // ---- dx_in = dqkv Win^T + dR1 -> xs (da of the layer below)
      lin<TM, true, RES>(big2, ldb, mats[m_in(l)], m, d, 3 * d, stage, [&](int r, int c, float v) {
        xs[(size_t)r * ldx + c] = v + zs[(size_t)r * ldx + c];
      });
    }
    // ---- tokens: dtok = sqrt(d) dx0.  position 0 -> dtok0 (ffn_user);  afterwards dz = dtok a g (1 - g), da = dtok g
    for (int i = tid; i < TM * ldx; i += NT) { ys[i] = 0.f; zs[i] = 0.f; }
    __syncthreads();
    for (int i = tid; i < m * d; i += NT) {
      const int r = i / d, c = i % d;
      const float dt = xs[(size_t)r * ldx + c] * sq;
      float dz = 0.f, da = 0.f, d0 = 0.f;
      if (rpos[r] == 0) d0 = dt;
      else {
        const float g = A.S.g[(row0 + r) * d + c], a = A.S.in[(row0 + r) * (1 + d) + 1 + c];
        dz = dt * a * g * (1.f - g);
        da = dt * g;
      }
      ys[(size_t)r * ldx + c] = dz;       // gate pre-activation gradient
      zs[(size_t)r * ldx + c] = da;       // direct item-embedding gradient
      xs[(size_t)r * ldx + c] = d0;       // user-token gradient
      A.S.dz[(row0 + r) * d + c] = dz;
      A.S.dtok0[(row0 + r) * d + c] = d0;
    }
    __syncthreads();
    if (G.emb_item)
      lin<TM, true, RES>(ys, ldx, mats[M_GATE], m, 1 + d, d, stage, [&](int r, int c, float v) {
        if (c >= 1 && rpos[r] >= 1) {
          const int item = A.act[(size_t)renv[r] * L + rpos[r] - 1];
          atomicAdd(G.emb_item + (size_t)item * d + c - 1, v + zs[(size_t)r * ldx + c - 1]);
        }
      });
    if (G.emb_user)
      lin<TM, true, RES>(xs, ldx, mats[M_USER], m, dui, d, stage, [&](int r, int c, float v) {
        if (rpos[r] == 0) atomicAdd(G.emb_user + (size_t)A.users[renv[r]] * d + c, v);
      });
    __syncthreads();
  }
}

// Greedy chunk plan: whole environments are packed in order into chunks of at most `cap` token rows (the optimum for a
// contiguous partition; the first-row-quantum rule above fills a chunk to ~(cap - longest episode / 2) rows only).
// plan[0] = number of chunks, plan[1 + c] = first environment of chunk c, plan[1 + n_chunks] = n_env.  One CTA, offsets
// in shared memory (n_env <= PLAN_MAX_ENV): every thread finds by binary search where a chunk that STARTS at its
// environment would end (jump pointer), then one thread follows the pointers from environment 0 -- ~n_chunks dependent
// shared-memory loads instead of a scan over all environments.
constexpr int PLAN_MAX_ENV = 6144;   // 36 KB of static shared memory
__global__ void __launch_bounds__(1024) chunk_plan_kernel(int n_env, const int32_t* __restrict__ env_off, int cap,
                                                           int32_t* __restrict__ plan) {
  __shared__ int s_off[PLAN_MAX_ENV + 1];
  __shared__ uint16_t s_next[PLAN_MAX_ENV];
  for (int i = threadIdx.x; i <= n_env; i += blockDim.x) s_off[i] = env_off[i];
  __syncthreads();
  for (int e = threadIdx.x; e < n_env; e += blockDim.x) {
    const int limit = s_off[e] + cap;
    int lo = e + 1, hi = n_env;          // largest j in [e + 1, n_env] with s_off[j] <= limit (j = e + 1 always fits)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= limit) lo = mid; else hi = mid - 1;
    }
    s_next[e] = (uint16_t)lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nc = 0, e = 0;
    while (e < n_env) {
      plan[1 + nc] = e;
      ++nc;
      e = s_next[e];
    }
    plan[0] = nc;
    plan[1 + nc] = n_env;
  }
}

inline size_t chunk_smem_bytes(int TM, int d, int nh, int ldx, int ldb, int resident_fl = -1) {
  const size_t stage = resident_fl >= 0 ? (size_t)TM * TM * nh : (size_t)WS_K * WS_LD;
  return sizeof(float) * ((size_t)3 * TM * ldx + 2 * TM * ldb + stage + 3 * TM * nh + 8 * 2 * TM +
                          2 * ((d + 3) & ~3) + (resident_fl >= 0 ? resident_fl + 4 : 0)) + sizeof(int) * 3 * TM + 64;
}

// ---- grouped weight-gradient launch: problem i computes gW_i[K_i][ldw_i] += X_i^T dY_i over all M rows (+ bias
// gradient = column sums of dY_i), as split-K tiles of the FP32 tile GEMM; one launch for every Linear of the tracker.
constexpr int MAX_PROB = 4 * CIRS_MAX_LAYERS + 3;
struct DwProblem {
  const float* x; int ldx;            // [M, k_in]
  const float* dy; int ldy;           // [M, n_out]
  const int32_t* dy_rows;             // optional row gather of dY (d_obs lives at buffer slots)
  float* gw; int ldw; float* gb;
  int k_in, n_out;
  int cta0, tiles_n, tiles_k;         // CTAs [cta0, cta0 + tiles_n * tiles_k * splits) belong to this problem
};
struct DwArgs {
  DwProblem p[MAX_PROB];
  int n_prob, M, splits, k_per_split;
};

__global__ void __launch_bounds__(256) tracker_dw_grouped_kernel(const DwArgs A) {
  int pi = 0;
  while (pi + 1 < A.n_prob && (int)blockIdx.x >= A.p[pi + 1].cta0) ++pi;
  const DwProblem& P = A.p[pi];
  const int local = blockIdx.x - P.cta0;
  const int tiles = P.tiles_n * P.tiles_k;
  const int split = local / tiles, t = local % tiles;
  gemm_tile<64, 64, 32, 4>(ColMajorA{P.x, P.ldx, nullptr}, RowMajorB{P.dy, P.ldy, P.dy_rows}, AtomicEp{P.gw, P.ldw},
                           P.k_in, P.n_out, A.M, A.k_per_split, P.gb, t % P.tiles_n, t / P.tiles_n, split);
}

}  // namespace cirs_k6
