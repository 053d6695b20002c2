// All-pairs DeepFM user-model inference -> KuaishouEnv's normed_mat.
// Replaces KuaishouEnv.compute_normed_reward (environments/KuaishouRec/env/kuaishouEnv.py:113-145): for every user
// the reference builds X = [user, photo_id, feat0..3, photo_duration] for all items, runs
// UserModel_Pairwise._deepfm (core/user_model_pairwise.py:98-132) and min-max normalises the U x I table.
//
// B200 design.  The model is  y(u, i) = linear + FM + w_last . relu(W2 relu(W1 x(u, i) + b1) + b2) + bias  with
// x(u, i) = [e_user(u) | e_item(i) | e_feat(i).. | dense(i)].  Everything that depends on ONE side only is hoisted
// out of the U x I loop (um_prep_users / um_prep_items):
//     W1 x + b1         = P[u] + Q[i]                 (P: user block of W1, Q: item blocks + dense + b1)
//     FM(u, i)          = <e_user(u), T[i]> + fm_i    (T = sum of the item-side embeddings)
//     linear            = lin_u[u] + lin_i[i]
// which leaves, per pair, h1 = relu(P[u] + Q[i]) (64 adds), the 64 x 64 contraction W2 h1 -- 97 % of the remaining
// flops -- and two short dot products.  The contraction runs on the tcgen05 tensor cores with 3xTF32 split precision
// (tc_dev.cuh; the 1e-5 parity bar rules out a single TF32 pass): a CTA owns one 128-item tile (its Q / T rows live in
// registers), walks a range of users, builds the 128 x 64 h1 tile as (hi, lo) K-major operand tiles in shared memory,
// issues 24 M128 x N64 x K8 MMAs against the resident (hi, lo) image of W2 into TMEM, and its epilogue applies
// b2 / ReLU / w_last straight out of TMEM -- h1 and h2 never exist in HBM; the only HBM traffic is the 4-byte result.
// Two CTAs per SM alternate (one stages / runs its epilogue while the other's MMAs are in flight).
// A plain FP32-FFMA kernel of the same factorisation is the second CUDA path (cirs_user_model_tc_enable(0), and the
// cross-check of the tensor-core path at full size).  A last pass normalises in place with float64 arithmetic like
// the reference (kuaishouEnv.py:139-143).
#include "common.cuh"
#include "tc_dev.cuh"
#include "../../include/cirs_b200.h"
#include <limits.h>
#include <stdlib.h>

namespace cirs_um {
using namespace cirs_tc;

constexpr int HID = CIRS_HIDDEN, TM = 128, NT = 256;
constexpr int KB = HID + 8;                   // contraction depth incl. the bias k-step: h1 | 1 0 0 0 0 0 0 0
constexpr uint32_t B_BYTES = HID * KB * 4;    // operand tile of [W2 | b2 | 0..]: 64 rows x 72 columns, 18 KB
constexpr uint32_t B_LBO = HID * 16, B_STEP = 2 * HID * 16;
constexpr uint32_t SBO = 128;
constexpr uint32_t IDESC = idesc_tf32(TM, HID, 0, 0);
// TMEM columns of one CTA (256 allocated; two CTAs per SM own all 512): accumulator | h1 hi (+ bias k-step) | h1 lo
constexpr uint32_t T_D = 0, T_AHI = 64, T_ALO = 64 + KB, T_COLS = 256;
// dynamic shared memory: the (hi, lo) image of W2 + w_last + half-row partials; padded to 100 KB so that at most two
// CTAs share an SM whatever the register allocation (a third CTA would block in tcgen05.alloc)
constexpr size_t TC_SMEM_USED = 2 * B_BYTES + HID * 4 + TM * 4;
constexpr size_t TC_SMEM = 100 * 1024;
static_assert(TC_SMEM_USED <= TC_SMEM, "smem");

__device__ int g_um_timeout = 0;
__device__ unsigned long long g_um_phase[8];   // TIMING instantiation: cycles of CTA 0 / thread 0 per phase (profiling aid)

// order-preserving float <-> int map (an involution) so that atomicMin / atomicMax on ints order floats
__device__ __forceinline__ int f2ord(float f) {
  const int o = __float_as_int(f);
  return o >= 0 ? o : o ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7FFFFFFF); }

struct Ws {   // carved out of the caller's workspace
  float *P, *Q, *T, *ci, *eu, *lu, *w2img, *b2w;
  int* minmax;
};
__host__ __device__ inline int64_t up4(int64_t x) { return (x + 3) & ~(int64_t)3; }
static int64_t ws_floats(int64_t U, int64_t I, int de) {
  return up4(U * HID) + up4(I * HID) + up4(I * de) + up4(I) + up4(U * de) + up4(U) + 2 * (B_BYTES / 4) + 2 * HID + 4;
}
static Ws carve(void* ws, int64_t U, int64_t I, int de) {
  Ws w;
  float* p = reinterpret_cast<float*>(ws);
  w.P = p; p += up4(U * HID);
  w.Q = p; p += up4(I * HID);
  w.T = p; p += up4(I * de);
  w.ci = p; p += up4(I);
  w.eu = p; p += up4(U * de);
  w.lu = p; p += up4(U);
  w.w2img = p; p += 2 * (B_BYTES / 4);
  w.b2w = p; p += 2 * HID;
  w.minmax = reinterpret_cast<int*>(p);
  return w;
}

// ---------------------------------------------------------------------------------------------- one-sided parts
// block = 64 threads = one user; thread j owns hidden unit j.  W1 is the reference's [64][in_dim] matrix.
__global__ void __launch_bounds__(HID)
um_prep_users_kernel(cirs_user_model m, const int32_t* __restrict__ user_ids, Ws w) {
  __shared__ float se[64];
  const int u = blockIdx.x, j = threadIdx.x, de = m.emb_dim;
  const int64_t uid = user_ids[u];
  if (j < de) {
    const float e = m.emb_user[uid * de + j];
    se[j] = e;
    w.eu[(int64_t)u * de + j] = e;
  }
  if (j == 0) w.lu[u] = m.lin_user[uid];
  __syncthreads();
  const int in_dim = de * (2 + m.n_feat) + m.n_dense;
  const float* w1 = m.w1 + (int64_t)j * in_dim;
  float acc = 0.f;
  for (int c = 0; c < de; ++c) acc = fmaf(w1[c], se[c], acc);
  w.P[(int64_t)u * HID + j] = acc;
}

// block = 64 threads = one item: Q (hidden pre-activation of the item side, bias included), T (sum of the item-side
// embeddings), ci (item-side FM term + linear logit + output bias).  Also builds the (hi, lo) operand image of W2 and
// the interleaved (b2, w_last) table (block 0) and resets the min / max cells.
__global__ void __launch_bounds__(HID)
um_prep_items_kernel(cirs_user_model m, const int32_t* __restrict__ item_ids, const int32_t* __restrict__ item_feat,
                     const float* __restrict__ item_dense, Ws w) {
  __shared__ float sx[64 * 9 + 16];   // item-side input vector: (1 + n_feat) * de + n_dense <= 592
  __shared__ float sred[2];
  const int i = blockIdx.x, j = threadIdx.x, de = m.emb_dim, nf = m.n_feat, nd = m.n_dense;
  const int64_t pid = item_ids[i];
  const int n_sp = (1 + nf) * de;
  for (int c = j; c < n_sp; c += HID) {
    const int f = c / de, k = c - f * de;
    sx[c] = f == 0 ? m.emb_item[pid * de + k] : m.emb_feat[(int64_t)item_feat[(int64_t)i * nf + f - 1] * de + k];
  }
  for (int c = j; c < nd; c += HID) sx[n_sp + c] = item_dense[(int64_t)i * nd + c];
  __syncthreads();
  const int in_dim = de * (2 + nf) + nd;
  const float* w1 = m.w1 + (int64_t)j * in_dim + de;   // skip the user block
  float acc = m.b1[j];
  for (int c = 0; c < n_sp + nd; ++c) acc = fmaf(w1[c], sx[c], acc);
  w.Q[(int64_t)i * HID + j] = acc;
  float part = 0.f;
  if (j < de) {
    float t = 0.f, s2 = 0.f;
    for (int f = 0; f <= nf; ++f) {
      const float e = sx[f * de + j];
      t += e;
      s2 = fmaf(e, e, s2);
    }
    w.T[(int64_t)i * de + j] = t;
    part = 0.5f * (t * t - s2);
  }
  part = warp_sum(part);
  if ((j & 31) == 0) sred[j >> 5] = part;
  __syncthreads();
  if (j == 0) {
    float c = sred[0] + sred[1] + m.lin_item[pid] + m.out_bias;
    for (int f = 0; f < nf; ++f) c += m.lin_feat[item_feat[(int64_t)i * nf + f]];
    for (int d = 0; d < nd; ++d) c = fmaf(m.lin_dense[d], sx[n_sp + d], c);
    w.ci[i] = c;
  }
  if (i == 0) {
    // W2 [out j][in k] row-major is already K-contiguous: B operand tile row = j, column = k
    char* hi = reinterpret_cast<char*>(w.w2img);
    char* lo = hi + B_BYTES;
    for (int c4 = 0; c4 < HID / 4; ++c4) tile_store_split(hi, lo, HID, j, c4, ld4(m.w2 + j * HID + 4 * c4));
    // bias k-step: column 64 = b2[j] (the A operand carries a constant 1 there), columns 65..71 zero
    tile_store_split(hi, lo, HID, j, HID / 4, make_float4(m.b2[j], 0.f, 0.f, 0.f));
    tile_store_split(hi, lo, HID, j, HID / 4 + 1, make_float4(0.f, 0.f, 0.f, 0.f));
    w.b2w[2 * j] = m.b2[j];
    w.b2w[2 * j + 1] = m.w_last[j];
    if (j == 0) { w.minmax[0] = INT_MAX; w.minmax[1] = INT_MIN; }
  }
}

struct PairArgs {
  const float *P, *Q, *T, *ci, *eu, *lu, *w2img, *b2w;
  float* out;
  int* minmax;
  int U, I, n_itile, n_uchunk, users_per_chunk;
};

__device__ __forceinline__ void publish_minmax(float vmin, float vmax, int* minmax) {
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(FULL_MASK, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(FULL_MASK, vmax, o));
  }
  if ((threadIdx.x & 31) == 0 && vmin <= vmax) {
    atomicMin(minmax, f2ord(vmin));
    atomicMax(minmax + 1, f2ord(vmax));
  }
}

// ---------------------------------------------------------------------------------------------- tensor-core path
// 256 threads: thread t owns item row t % 128 (= TMEM lane) and half t / 128 of the 64 hidden columns.
// The h1 tile never touches shared memory: each thread writes its 32 (hi, lo) values straight into TMEM
// (tcgen05.st) and the MMAs take A from TMEM, B = [W2 | b2] from shared memory -- an M128 x N64 x K8 MMA with both
// operands in shared memory reads 6 KB per instruction and is bound by that (measured 57 cycles); with A in TMEM it
// reads 2 KB.  The bias rides in a ninth k-step (A column 64 == 1), so the epilogue is relu + one FMA per element.
__device__ __forceinline__ void um_issue(uint32_t tb, uint32_t b_hi, uint32_t b_lo) {
  // One descriptor per operand tile; k-step j is the same descriptor with the start-address field advanced by
  // j * B_STEP / 16 (shared addresses are < 256 KB, so the 14-bit field never carries): one add per MMA operand
  // instead of rebuilding the descriptor (shift, mask, or) -- the issuing lane's instruction stream is what paces
  // the batch.
  const uint64_t dh = smem_desc(b_hi, B_LBO, SBO), dl = smem_desc(b_lo, B_LBO, SBO);
  constexpr uint64_t KSTEP = B_STEP >> 4;
  // bias k-step first (overwrites the accumulator), then the 8 k-steps of the contraction, small terms first
  mma_tf32_ts(tb + T_D, tb + T_AHI + HID, dl + (HID / 8) * KSTEP, IDESC, 0u);
  mma_tf32_ts(tb + T_D, tb + T_AHI + HID, dh + (HID / 8) * KSTEP, IDESC, 1u);
#pragma unroll
  for (int j = 0; j < HID / 8; ++j) {
    mma_tf32_ts(tb + T_D, tb + T_ALO + 8 * j, dh + j * KSTEP, IDESC, 1u);
    mma_tf32_ts(tb + T_D, tb + T_AHI + 8 * j, dl + j * KSTEP, IDESC, 1u);
    mma_tf32_ts(tb + T_D, tb + T_AHI + 8 * j, dh + j * KSTEP, IDESC, 1u);
  }
}

// round-to-nearest (ties away) TF32 of a non-negative finite float in two integer instructions (cvt.rna.tf32.f32
// compiles to four: add, mask and a NaN / infinity select)
__device__ __forceinline__ float tf32_rna_pos(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// one lane of a converged warp (elect.sync): code under it is provably executed by a single thread, which lets the
// compiler keep the MMA descriptors in uniform registers.  Issued under `if (tid == 0)` every tcgen05.mma cost ~8 SASS
// instructions (R2UR + a divergence waterfall) and 72 cycles: 1860 of a tile's 3350 cycles went into issuing.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int DE, bool TIMING>
__global__ void __launch_bounds__(NT, 2) um_pairs_tc_kernel(PairArgs a) {
  extern __shared__ __align__(1024) char smem[];
  char* b_hi = smem;
  char* b_lo = b_hi + B_BYTES;
  float* swl = reinterpret_cast<float*>(b_lo + B_BYTES);   // w_last[64]
  float* part = swl + HID;                                 // half 1's partial result per row
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const int warp_u = __shfl_sync(FULL_MASK, warp, 0);   // the same value, known to be warp-uniform
  const uint32_t lane_base = (warp & 3) * 32;
  constexpr int DH = DE / 2;   // FM dimensions per half
  const bool timing = TIMING && blockIdx.x == 0 && tid == 0;
  unsigned tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 32-bit cycle sums: one CTA's launch is ~10^7 cycles
  if (warp == 0) tmem_alloc(&tmem_base, T_COLS);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int k = tid; k < (int)(2 * B_BYTES / 16); k += NT)
    reinterpret_cast<float4*>(b_hi)[k] = __ldg(reinterpret_cast<const float4*>(a.w2img) + k);
  if (tid < HID) swl[tid] = a.b2w[2 * tid + 1];
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (half == 0) {   // constant A columns of the bias k-step
    const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    tmem_st8(tmem_addr(tb + T_AHI + HID, lane_base, 0), one);
    tmem_st_wait();
  }
  float vmin = INFINITY, vmax = -INFINITY;
  uint32_t ph = 0;
  const int n_chunks = a.n_itile * a.n_uchunk;
  bool dead = false;
  for (int c = blockIdx.x; c < n_chunks && !dead; c += gridDim.x) {
    const int it = c % a.n_itile, uc = c / a.n_itile;
    const int i = it * TM + row;
    const bool iv = i < a.I;
    float4 q[8];
    float t[DH];
    float ci = 0.f;
    if (iv) {
      const float4* qs = reinterpret_cast<const float4*>(a.Q + (int64_t)i * HID + half * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = __ldg(qs + k);
#pragma unroll
      for (int k = 0; k < DH; ++k) t[k] = __ldg(a.T + (int64_t)i * DE + half * DH + k);
      if (half == 0) ci = __ldg(a.ci + i);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < DH; ++k) t[k] = 0.f;
    }
    const int u0 = uc * a.users_per_chunk, u1 = min(a.U, u0 + a.users_per_chunk);
    float4 p[8];
    if (u0 < u1) {
      const float4* ps = reinterpret_cast<const float4*>(a.P + (int64_t)u0 * HID + half * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) p[k] = __ldg(ps + k);
    }
    for (int u = u0; u < u1; ++u) {
      unsigned tk0 = 0;
      if (timing) tk0 = (unsigned)clock();
      // h1 = relu(P[u] + Q[i]) -> (hi, lo) A operand in TMEM: lane = row, column = hidden index
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float hi[16], lo[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 pp = p[4 * g + k], qq = q[4 * g + k];
          const float x0 = fmaxf(pp.x + qq.x, 0.f), x1 = fmaxf(pp.y + qq.y, 0.f), x2 = fmaxf(pp.z + qq.z, 0.f),
                      x3 = fmaxf(pp.w + qq.w, 0.f);
          hi[4 * k] = tf32_rna_pos(x0); hi[4 * k + 1] = tf32_rna_pos(x1); hi[4 * k + 2] = tf32_rna_pos(x2);
          hi[4 * k + 3] = tf32_rna_pos(x3);
          // x - hi is exact in FP32 and has <= 13 significant bits; the MMA truncates it to TF32 (2^-22 of x)
          lo[4 * k] = x0 - hi[4 * k]; lo[4 * k + 1] = x1 - hi[4 * k + 1]; lo[4 * k + 2] = x2 - hi[4 * k + 2];
          lo[4 * k + 3] = x3 - hi[4 * k + 3];
        }
        tmem_st16(tmem_addr(tb + T_AHI, lane_base, half * 32 + g * 16), hi);
        tmem_st16(tmem_addr(tb + T_ALO, lane_base, half * 32 + g * 16), lo);
      }
      tmem_st_wait();
      fence_before_sync();
      if (timing) { const unsigned t = (unsigned)clock(); tacc[0] += t - tk0; tk0 = t; }
      __syncthreads();
      fence_after_sync();
      if (timing) { const unsigned t = (unsigned)clock(); tacc[1] += t - tk0; tk0 = t; }
      if (warp_u == 0) {
        if (elect_one()) {
          um_issue(tb, smem_u32(b_hi), smem_u32(b_lo));
          mma_commit(&bar);
        }
        __syncwarp();
      }
      if (timing) { const unsigned t = (unsigned)clock(); tacc[2] += t - tk0; tk0 = t; }
      // behind the MMAs: this user's FM / linear terms, the next user's P
      float fm = half == 0 ? ci + __ldg(a.lu + u) : 0.f;
      {
        const float4* es = reinterpret_cast<const float4*>(a.eu + (int64_t)u * DE + half * DH);
#pragma unroll
        for (int k = 0; k < DH / 4; ++k) {
          const float4 e = __ldg(es + k);
          fm = fmaf(e.x, t[4 * k], fm);
          fm = fmaf(e.y, t[4 * k + 1], fm);
          fm = fmaf(e.z, t[4 * k + 2], fm);
          fm = fmaf(e.w, t[4 * k + 3], fm);
        }
      }
      if (u + 1 < u1) {
        const float4* ps = reinterpret_cast<const float4*>(a.P + (int64_t)(u + 1) * HID + half * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = __ldg(ps + k);
      }
      if (timing) { const unsigned t = (unsigned)clock(); tacc[3] += t - tk0; tk0 = t; }
      const bool ok = mbar_wait(&bar, ph);
      ph ^= 1;
      fence_after_sync();
      if (timing) { const unsigned t = (unsigned)clock(); tacc[4] += t - tk0; tk0 = t; }
      float v[32];
      tmem_ld32(tmem_addr(tb + T_D, lane_base, half * 32), v);
      float acc0 = fm, acc1 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 w = *reinterpret_cast<const float4*>(swl + half * 32 + k);   // broadcast
        acc0 = fmaf(fmaxf(v[k], 0.f), w.x, acc0);
        acc1 = fmaf(fmaxf(v[k + 1], 0.f), w.y, acc1);
        acc0 = fmaf(fmaxf(v[k + 2], 0.f), w.z, acc0);
        acc1 = fmaf(fmaxf(v[k + 3], 0.f), w.w, acc1);
      }
      const float y = acc0 + acc1;
      if (half == 1) part[row] = y;
      fence_before_sync();
      if (timing) { const unsigned t = (unsigned)clock(); tacc[5] += t - tk0; tk0 = t; }
      // accumulator and operand columns are free again; half 1's partial is visible.  A wait that gave up (never
      // expected) ends the kernel for the whole CTA instead of spinning once per tile.
      if (__syncthreads_or(!ok)) {
        if (tid == 0) g_um_timeout = 1;
        dead = true;
        break;
      }
      if (timing) { const unsigned t = (unsigned)clock(); tacc[6] += t - tk0; tk0 = t; }
      if (half == 0 && iv) {
        const float r = y + part[row];
        a.out[(int64_t)u * a.I + i] = r;
        vmin = fminf(vmin, r);
        vmax = fmaxf(vmax, r);
      }
      if (timing) tacc[7] += 1;
    }
  }
  if (timing)
    for (int k = 0; k < 8; ++k) g_um_phase[k] = tacc[k];
  publish_minmax(vmin, vmax, a.minmax);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, T_COLS);
}

// ---------------------------------------------------------------------------------------------- FFMA path
// grid (item tiles, user chunks), 128 threads: thread = item, h1 in registers, W2 / b2 / w_last in shared memory
// (warp-uniform addresses: broadcast reads).
constexpr int FF_UCHUNK = 32;
__global__ void __launch_bounds__(TM) um_pairs_ffma_kernel(PairArgs a, const float* __restrict__ w2,
                                                           const float* __restrict__ b2,
                                                           const float* __restrict__ w_last, int de) {
  __shared__ __align__(16) float sw2[HID * HID];
  __shared__ float sb2[HID], swl[HID];
  __shared__ __align__(16) float sp[HID];
  __shared__ float seu[64];
  const int tid = threadIdx.x;
  for (int k = tid; k < HID * HID; k += TM) sw2[k] = w2[k];
  if (tid < HID) { sb2[tid] = b2[tid]; swl[tid] = w_last[tid]; }
  const int i = blockIdx.x * TM + tid;
  const bool iv = i < a.I;
  float q[HID];
#pragma unroll
  for (int k = 0; k < HID; ++k) q[k] = iv ? a.Q[(int64_t)i * HID + k] : 0.f;
  const float ci = iv ? a.ci[i] : 0.f;
  float vmin = INFINITY, vmax = -INFINITY;
  const int u0 = blockIdx.y * FF_UCHUNK, u1 = min(a.U, u0 + FF_UCHUNK);
  for (int u = u0; u < u1; ++u) {
    __syncthreads();
    if (tid < HID) sp[tid] = a.P[(int64_t)u * HID + tid];
    if (tid < de) seu[tid] = a.eu[(int64_t)u * de + tid];
    __syncthreads();
    float h1[HID];
#pragma unroll
    for (int k = 0; k < HID; ++k) h1[k] = fmaxf(sp[k] + q[k], 0.f);
    float y = ci + a.lu[u];
    if (iv)
      for (int k = 0; k < de; ++k) y = fmaf(seu[k], a.T[(int64_t)i * de + k], y);
    float d = 0.f;
#pragma unroll 2
    for (int j = 0; j < HID; ++j) {
      float acc = sb2[j];
      const float4* wr = reinterpret_cast<const float4*>(sw2 + j * HID);
#pragma unroll
      for (int k = 0; k < HID / 4; ++k) {
        const float4 w = wr[k];
        acc = fmaf(w.x, h1[4 * k], acc);
        acc = fmaf(w.y, h1[4 * k + 1], acc);
        acc = fmaf(w.z, h1[4 * k + 2], acc);
        acc = fmaf(w.w, h1[4 * k + 3], acc);
      }
      d = fmaf(fmaxf(acc, 0.f), swl[j], d);
    }
    y += d;
    if (iv) {
      a.out[(int64_t)u * a.I + i] = y;
      vmin = fminf(vmin, y);
      vmax = fmaxf(vmax, y);
    }
  }
  publish_minmax(vmin, vmax, a.minmax);
}

// ---------------------------------------------------------------------------------------------- normalise
// (pred - min) / (max - min) in float64 like the reference (kuaishouEnv.py:139-143), rounded once to float32
__global__ void __launch_bounds__(256) um_normalise_kernel(float* __restrict__ out, int64_t n, const int* __restrict__ minmax,
                                                           float* __restrict__ minmax_out) {
  const double mn = (double)ord2f(minmax[0]), mx = (double)ord2f(minmax[1]);
  const double den = mx - mn;
  if (minmax_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { minmax_out[0] = (float)mn; minmax_out[1] = (float)mx; }
  const int64_t n4 = n >> 2;
  float4* o4 = reinterpret_cast<float4*>(out);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    float4 v = o4[k];
    v.x = (float)(((double)v.x - mn) / den);
    v.y = (float)(((double)v.y - mn) / den);
    v.z = (float)(((double)v.z - mn) / den);
    v.w = (float)(((double)v.w - mn) / den);
    o4[k] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t k = (n4 << 2) + threadIdx.x;
    out[k] = (float)(((double)out[k] - mn) / den);
  }
}

static int g_um_tc = -1;   // -1: default (on unless CIRS_NO_TC=1), 0: FFMA, 1: tensor cores
static bool tc_on() {
  if (g_um_tc >= 0) return g_um_tc != 0;
  const char* e = getenv("CIRS_NO_TC");
  return !(e && e[0] == '1');
}

}  // namespace cirs_um

using namespace cirs_um;

extern "C" void cirs_user_model_tc_enable(int on) { g_um_tc = on; }

extern "C" int cirs_user_model_timeout(void) {
  int v = 0, zero = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&v, g_um_timeout, sizeof(int));
  cudaMemcpyToSymbol(g_um_timeout, &zero, sizeof(int));
  return v;
}

extern "C" int cirs_user_model_debug_phases(int64_t* out8_h) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out8_h, g_um_phase, 8 * sizeof(int64_t)) == cudaSuccess ? CIRS_OK : CIRS_ERR_CUDA;
}

extern "C" int64_t cirs_user_model_workspace_bytes(int32_t n_user, int32_t n_item, int32_t emb_dim) {
  return ws_floats(n_user, n_item, emb_dim) * 4;
}

extern "C" int cirs_user_model_predict_all(const cirs_user_model* m, int32_t n_user, const int32_t* user_ids,
                                           int32_t n_item, const int32_t* item_ids, const int32_t* item_feat,
                                           const float* item_dense, int32_t normalise, float* out, float* minmax,
                                           void* workspace, void* stream) {
  if (!m || !user_ids || !item_ids || !out || !workspace || n_user <= 0 || n_item <= 0) {
    cirs_set_error("cirs_user_model_predict_all: null / empty argument");
    return CIRS_ERR_ARG;
  }
  if (m->hidden != HID) {
    cirs_set_error("cirs_user_model_predict_all: dnn_hidden_units must be (64, 64)");
    return CIRS_ERR_ARG;
  }
  if (m->emb_dim < 1 || m->emb_dim > 64 || m->n_feat < 0 || m->n_feat > 8 || m->n_dense < 0 || m->n_dense > 16 ||
      (m->n_feat > 0 && !item_feat) || (m->n_dense > 0 && (!item_dense || !m->lin_dense))) {
    cirs_set_error("cirs_user_model_predict_all: unsupported feature layout (emb_dim <= 64, n_feat <= 8, n_dense <= 16)");
    return CIRS_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
    cirs_set_error("cirs_user_model_predict_all: out / workspace must be 16-byte aligned");
    return CIRS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int de = m->emb_dim;
  Ws w = carve(workspace, n_user, n_item, de);
  CIRS_LAUNCH(um_prep_items_kernel, n_item, HID, 0, st, *m, item_ids, item_feat, item_dense, w);
  CIRS_CHECK_LAUNCH();
  CIRS_LAUNCH(um_prep_users_kernel, n_user, HID, 0, st, *m, user_ids, w);
  CIRS_CHECK_LAUNCH();
  PairArgs a;
  a.P = w.P; a.Q = w.Q; a.T = w.T; a.ci = w.ci; a.eu = w.eu; a.lu = w.lu; a.w2img = w.w2img; a.b2w = w.b2w;
  a.out = out; a.minmax = w.minmax; a.U = n_user; a.I = n_item;
  a.n_itile = (n_item + TM - 1) / TM;
  const bool tc = tc_on() && (de == 8 || de == 16 || de == 32);
  if (tc) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid_max = 2 * sms;
    // user chunks: enough (item tile, user chunk) pairs for ~20 waves of the persistent grid, >= 32 users per chunk
    int upc = (int)(((int64_t)n_user * a.n_itile + 20 * (int64_t)grid_max - 1) / (20 * (int64_t)grid_max));
    if (upc < 32) upc = 32;
    if (upc > n_user) upc = n_user;
    a.users_per_chunk = upc;
    a.n_uchunk = (n_user + upc - 1) / upc;
    const int n_chunks = a.n_itile * a.n_uchunk;
    const int grid = n_chunks < grid_max ? n_chunks : grid_max;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(um_pairs_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
      cudaFuncSetAttribute(um_pairs_tc_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
      cudaFuncSetAttribute(um_pairs_tc_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
      cudaFuncSetAttribute(um_pairs_tc_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
      attr_done = true;
    }
    const char* fl = getenv("CIRS_UM_FLAGS");   // 4: per-phase cycle counters (cirs_user_model_debug_phases)
    const bool timing = fl && (atoi(fl) & 4) && de == 16;
    // the macro stringifies its first argument: keep the template commas inside parentheses
    if (timing) CIRS_LAUNCH((um_pairs_tc_kernel<16, true>), grid, NT, TC_SMEM, st, a);
    else if (de == 8) CIRS_LAUNCH((um_pairs_tc_kernel<8, false>), grid, NT, TC_SMEM, st, a);
    else if (de == 16) CIRS_LAUNCH((um_pairs_tc_kernel<16, false>), grid, NT, TC_SMEM, st, a);
    else CIRS_LAUNCH((um_pairs_tc_kernel<32, false>), grid, NT, TC_SMEM, st, a);
  } else {
    a.users_per_chunk = FF_UCHUNK;
    a.n_uchunk = (n_user + FF_UCHUNK - 1) / FF_UCHUNK;
    CIRS_LAUNCH(um_pairs_ffma_kernel, dim3(a.n_itile, a.n_uchunk), TM, 0, st, a, m->w2, m->b2, m->w_last, de);
  }
  CIRS_CHECK_LAUNCH();
  if (normalise) {
    CIRS_LAUNCH(um_normalise_kernel, 148 * 8, 256, 0, st, out, (int64_t)n_user * n_item, w.minmax, minmax);
    CIRS_CHECK_LAUNCH();
  } else if (minmax) {
    CIRS_LAUNCH(um_normalise_kernel, 1, 32, 0, st, out, (int64_t)0, w.minmax, minmax);
    CIRS_CHECK_LAUNCH();
  }
  return CIRS_OK;
}
