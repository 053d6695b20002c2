// Actor head over the item catalogue on tcgen05 tensor cores (see head_tc.cuh for the three passes).
// Replaces, for the PPO update (core/policy/ppo.py:181-220) and the policy evaluation of process_fn (ppo.py:96-109),
// the FP32-FFMA tile GEMMs  logits = h2 W3t + b3,  dW3t = h2^T dlogits,  dh2 = dlogits W3  and the [n, n_action]
// logits round trip through HBM that they needed.
//
// All MMAs are M128 x N64 x K8 kind::tf32 with K-major operands, issued by one thread; every FP32 operand is a (hi, lo)
// pair of TF32-exact tiles and each product is hi.hi + hi.lo + lo.hi (tc_dev.cuh).  W3 is split once per weight update
// into tile images in global memory (head_tc_pack) that the passes copy into shared memory; h2 is split while staging;
// d logits go back to TMEM as the A operand of the second MMA.
// Thread t owns TMEM lane t % 128 (warp % 4 selects the lane quarter, as tcgen05.ld requires) and the column group
// t / 128 of each accumulator: quarters with 512 threads (pass F on 128-column tiles, passes B2 / B3 on 64-column tiles;
// one CTA per SM) -- four resident warps per scheduler hide the epilogue's instruction latency; halves with 256 threads
// in the register-staged pass F (two CTAs per SM).
#include "common.cuh"
#include "tc_dev.cuh"
#include "head_tc.cuh"
#include "../../include/cirs_b200.h"
#include <stdlib.h>

namespace cirs_head_tc {
using namespace cirs_tc;

namespace {
constexpr int NT = 256, NTB = 512, TM = 128, TN = 64, HID = 64;   // threads: pass F / passes B2, B3
constexpr uint32_t A_BYTES = TM * HID * 4;   // 128-row operand tile (32 KB)
constexpr uint32_t B_BYTES = TN * HID * 4;   // 64-row operand tile (16 KB)
constexpr uint32_t A_LBO = TM * 16, A_STEP = 2 * TM * 16;   // K-major, R = 128
constexpr uint32_t B_LBO = TN * 16, B_STEP = 2 * TN * 16;   // K-major, R = 64
constexpr uint32_t SBO = 128;
constexpr uint32_t IDESC = idesc_tf32(TM, TN, 0, 0);
constexpr int KSTEPS = HID / 8;   // every contraction here has depth 64
#define MASKED (-1.0e30f)                         /* bias of the padding columns: exp(. - max) == 0 */
#define LOG_EPS (-15.942385152878742f)            /* log(CATEGORICAL_EPS) */
#define LOG_1M_EPS (-1.1920929665620963e-07f)     /* log(1 - CATEGORICAL_EPS) */

__device__ int g_tc_timeout = 0;   // set when an mbarrier wait gives up (never expected; checked by the tests)
// Phase counters (cycles summed over CTAs) of pass F [0, 16) and the bulk-copy fed pass B2 [16, 48); filled only
// by the PH instantiations (cirs_head_tc_debug_phases; scratch/head_phases.py).  One representative worker warp (warp 0)
// and the issuer warp stamp clock64() around their waits.
__device__ unsigned long long g_phase[64];
template <bool PH>
struct PhaseClock {   // compiled out entirely unless PH (the stamped instantiations are launched while g_phase_host is set)
  bool on;
  long long t;
  __device__ __forceinline__ void start(bool enable) {
    if constexpr (PH) { on = enable; if (on) t = clock64(); } else { on = false; }
  }
  __device__ __forceinline__ void lap(int slot) {
    if constexpr (PH) {
      if (on) {
        const long long n = clock64();
        atomicAdd(&g_phase[slot], (unsigned long long)(n - t));
        t = n;
      }
    }
  }
  __device__ __forceinline__ void count(int slot, int v) {
    if constexpr (PH) { if (on) atomicAdd(&g_phase[slot], (unsigned long long)v); }
  }
};

__device__ __forceinline__ void issue(uint32_t d, const char* a_hi, const char* a_lo, const char* b_hi, const char* b_lo,
                                      bool accumulate) {
  mma_3xtf32(d, smem_u32(a_hi), smem_u32(a_lo), A_STEP, A_LBO, SBO, smem_u32(b_hi), smem_u32(b_lo), B_STEP, B_LBO, SBO,
             IDESC, KSTEPS, accumulate);
}
// warp-collective forms for the issuer warp of the TMA-fed kernels (tc_dev.cuh elect_one)
__device__ __forceinline__ void w_issue(uint32_t d, const char* a_hi, const char* a_lo, const char* b_hi, const char* b_lo,
                                        bool accumulate) {
  if (elect_one()) issue(d, a_hi, a_lo, b_hi, b_lo, accumulate);
}
// Bounded wait: ~2^28 polls before the first thread gives up and raises the flag; every other waiter then leaves within
// 2^16 polls (a kernel whose pipeline is wedged ends in seconds instead of one full timeout per remaining wait).
__device__ __forceinline__ void wait_or_flag(uint64_t* bar, uint32_t parity) {
  for (int c = 0; c < 4096; ++c) {
    if (mbar_wait_n(bar, parity, 1u << 16)) return;
    if (*reinterpret_cast<volatile int*>(&g_tc_timeout)) return;
  }
  g_tc_timeout = 1;
}
// v[k] for a run-time k with v kept in REGISTERS.  The plain forms (v[k], or an unrolled `if (j == k) x = v[j]`) make
// the compiler index a local-memory copy of v: 8 x STL.128 per thread per tile in the hot loop of every pass-F kernel,
// which -- with the L1 carved down to nothing beside ~200 KB of shared memory -- went to L2 and paced the epilogue.
__device__ __forceinline__ float pick32(const float (&v)[32], int k) {
  float x = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.f32 %0, %3, %0, p;\n\t}" : "+f"(x) : "r"(k), "r"(j), "f"(v[j]));
  return x;
}

// ------------------------------------------------------------------------------------------------- pass F
constexpr size_t F_SMEM = 2 * A_BYTES + 2 * B_BYTES + 64 * 4 + 2 * NT * 4;

// ---- pre-split operand-tile images of W3 --------------------------------------------------------------------
// For every 64-column catalogue tile ct:   imgN[ct] = { hi, lo } of the tile laid out with tile row = column, tile
// column = hidden (B operand of the logits MMA);  imgK[ct] = { hi, lo } with tile row = hidden, tile column = column
// (B operand of the d h2 MMA).  For every 128-column tile: imgA = { hi, lo } with tile row = column (128), tile
// column = hidden (A operand of pass B3's transposed logits MMA).  All in the exact shared-memory byte layout.
constexpr int64_t IMG_B = 2 * (B_BYTES / 4);   // floats per (hi, lo) pair of a 64-row tile
constexpr int64_t IMG_A = 2 * (A_BYTES / 4);   // ... of a 128-row tile
__host__ __device__ inline int64_t img_n_off(int64_t ct) { return ct * IMG_B; }
__host__ __device__ inline int64_t img_k_off(int64_t n64, int64_t ct) { return n64 * IMG_B + ct * IMG_B; }
__host__ __device__ inline int64_t img_a_off(int64_t n64, int64_t c128) { return 2 * n64 * IMG_B + c128 * IMG_A; }
__host__ __device__ inline int64_t img_bias_off(int64_t n64) { return 2 * n64 * IMG_B + (n64 / 2) * IMG_A; }   // [ldA] bias, MASKED padding

// one 64-column catalogue tile ct by the whole CTA (any block size); s: [HID][TN + 1] floats of shared memory
__device__ __forceinline__ void pack_w3_tile(int ct, const float* __restrict__ w3t, int64_t ldA,
                                             const float* __restrict__ b3, int nA, float* __restrict__ img,
                                             float (*s)[TN + 1]) {
  const int tid = threadIdx.x, nthr = blockDim.x, c0 = ct * TN;
  const int64_t n64 = ldA / TN;
  if (tid < TN) img[img_bias_off(n64) + c0 + tid] = c0 + tid < nA ? __ldg(b3 + c0 + tid) : MASKED;
  for (int i = tid; i < HID * TN / 4; i += nthr) {
    const int k = i / (TN / 4), c4 = i % (TN / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(w3t + (size_t)k * ldA + c0) + c4);
    s[k][4 * c4] = v.x; s[k][4 * c4 + 1] = v.y; s[k][4 * c4 + 2] = v.z; s[k][4 * c4 + 3] = v.w;
  }
  __syncthreads();
  float4* n_hi = reinterpret_cast<float4*>(img + img_n_off(ct));
  float4* n_lo = n_hi + B_BYTES / 16;
  float4* k_hi = reinterpret_cast<float4*>(img + img_k_off(n64, ct));
  float4* k_lo = k_hi + B_BYTES / 16;
  float4* a_hi = reinterpret_cast<float4*>(img + img_a_off(n64, ct >> 1));
  float4* a_lo = a_hi + A_BYTES / 16;
  auto split4 = [](float4 v, float4& h, float4& l) {
    h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    l = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
  };
  for (int i = tid; i < HID * TN / 4; i += nthr) {
    // 16-byte chunk i of a 64-row tile: chunk column c4 = i / 64, tile row r = i % 64 (tile_chunk_off)
    const int c4 = i / TN, r = i % TN;
    float4 h, l;
    split4(make_float4(s[4 * c4][r], s[4 * c4 + 1][r], s[4 * c4 + 2][r], s[4 * c4 + 3][r]), h, l);   // row = column r
    n_hi[i] = h; n_lo[i] = l;
    const int ia = c4 * TM + (ct & 1) * TN + r;                                                     // 128-row tile
    a_hi[ia] = h; a_lo[ia] = l;
    split4(make_float4(s[r][4 * c4], s[r][4 * c4 + 1], s[r][4 * c4 + 2], s[r][4 * c4 + 3]), h, l);   // row = hidden r
    k_hi[i] = h; k_lo[i] = l;
  }
}
__global__ void __launch_bounds__(256)
head_tc_pack_kernel(const float* __restrict__ w3t, int64_t ldA, const float* __restrict__ b3, int nA,
                    float* __restrict__ img) {
  __shared__ float s[HID][TN + 1];   // s[k][col]
  pack_w3_tile(blockIdx.x, w3t, ldA, b3, nA, img, s);
}

// h2 [n, 64]: for every 64-row tile t  Hn[t] = { hi, lo } with tile row = row, tile column = hidden (B operand of pass
// B3's logits MMA), Ht[t] = { hi, lo } with tile row = hidden, tile column = row (B operand of the d W3 MMA); for every
// 128-row tile Ha = { hi, lo } with tile row = row (A operand of passes F / B2).  n64 = number of 64-row tiles (even).
__host__ __device__ inline int64_t himg_n_off(int64_t t) { return t * IMG_B; }
__host__ __device__ inline int64_t himg_t_off(int64_t n64, int64_t t) { return n64 * IMG_B + t * IMG_B; }
__host__ __device__ inline int64_t himg_a_off(int64_t n64, int64_t t128) { return 2 * n64 * IMG_B + t128 * IMG_A; }
__host__ __device__ inline int64_t h2_tiles64(int64_t n) { return 2 * ((n + TM - 1) / TM); }

__global__ void __launch_bounds__(256)
head_tc_pack_h2_kernel(const float* __restrict__ h2, int n, float* __restrict__ himg) {
  __shared__ float s[TN][HID + 1];   // s[row][hidden]
  const int t = blockIdx.x, tid = threadIdx.x, r0 = t * TN;
  const int64_t n64 = h2_tiles64(n);
  for (int i = tid; i < TN * HID / 4; i += 256) {
    const int r = i / (HID / 4), c4 = i % (HID / 4);
    const float4 v = r0 + r < n ? ld4(h2 + (size_t)(r0 + r) * HID + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    s[r][4 * c4] = v.x; s[r][4 * c4 + 1] = v.y; s[r][4 * c4 + 2] = v.z; s[r][4 * c4 + 3] = v.w;
  }
  __syncthreads();
  float4* n_hi = reinterpret_cast<float4*>(himg + himg_n_off(t));
  float4* n_lo = n_hi + B_BYTES / 16;
  float4* t_hi = reinterpret_cast<float4*>(himg + himg_t_off(n64, t));
  float4* t_lo = t_hi + B_BYTES / 16;
  float4* a_hi = reinterpret_cast<float4*>(himg + himg_a_off(n64, t >> 1));
  float4* a_lo = a_hi + A_BYTES / 16;
  auto split4 = [](float4 v, float4& h, float4& l) {
    h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    l = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
  };
  for (int i = tid; i < TN * HID / 4; i += 256) {
    const int c4 = i / TN, r = i % TN;
    float4 h, l;
    split4(make_float4(s[r][4 * c4], s[r][4 * c4 + 1], s[r][4 * c4 + 2], s[r][4 * c4 + 3]), h, l);   // row = row r
    n_hi[i] = h; n_lo[i] = l;
    const int ia = c4 * TM + (t & 1) * TN + r;
    a_hi[ia] = h; a_lo[ia] = l;
    split4(make_float4(s[4 * c4][r], s[4 * c4 + 1][r], s[4 * c4 + 2][r], s[4 * c4 + 3][r]), h, l);   // row = hidden r
    t_hi[i] = h; t_lo[i] = l;
  }
}

// ---- front end of a pass over n rows in ONE launch (replaces trunk_fwd_kernel + head_tc_pack_h2 + head_tc_pack):
//   CTAs [0, n_pack)   re-split W3 into its operand-tile images (pack_w3_tile) -- concurrently with
//   CTAs [n_pack, ..)  the policy trunk (Net 2 x Linear + ReLU, Critic head; core/policy/ppo.py:122-126 through
//                      tianshou Net / Critic) of one 64-row tile each, two 256-thread halves of 32 rows, which then
//                      write the tile's h2 straight out of shared memory as (hi, lo) operand-tile images.
// Accumulation order of the trunk (bias first, k ascending, fmaf) is the one of actor_trunk_warp / trunk_fwd_kernel:
// bit-identical h2.
constexpr int FRONT_NT = 512;
constexpr size_t FRONT_SMEM = sizeof(float) * (32 * HID + HID * HID + 2 * 32 * 33 + 2 * HID * 33);
__global__ void __launch_bounds__(FRONT_NT)
head_tc_front_kernel(cirs_policy_weights W, int n, const int32_t* __restrict__ idx, const float* __restrict__ obs,
                     float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ value,
                     float* __restrict__ himg, int n_pack, float* __restrict__ img,
                     const int32_t* __restrict__ n_dev) {
  extern __shared__ __align__(16) float fsm[];
  if ((int)blockIdx.x < n_pack) {
    pack_w3_tile(blockIdx.x, W.w3t, W.ld_action, W.b3, W.n_action, img, reinterpret_cast<float(*)[TN + 1]>(fsm));
    return;
  }
  // n sizes the grid and the image layout; with n_dev the rows that exist are min(n, *n_dev) (count still on the device)
  const int n_lay = n;
  if (n_dev) n = min(n, __ldg(n_dev));
  if ((int)(blockIdx.x - n_pack) * TN >= n) return;
  constexpr int R = 32;
  float* s_w1 = fsm;                     // W1t [dim_state <= 32][64]
  float* s_w2 = s_w1 + 32 * HID;         // W2t [64][64]
  const int tid = threadIdx.x, half = tid >> 8, t = tid & 255, tile = blockIdx.x - n_pack, S = W.dim_state;
  float(*s_in)[33] = reinterpret_cast<float(*)[33]>(s_w2 + HID * HID + half * 32 * 33);
  float(*hT)[R + 1] = reinterpret_cast<float(*)[R + 1]>(s_w2 + HID * HID + 2 * 32 * 33 + half * HID * 33);
  const int r0 = tile * TN + half * R;
  for (int i = tid; i < S * HID / 4; i += FRONT_NT) reinterpret_cast<float4*>(s_w1)[i] = __ldg(reinterpret_cast<const float4*>(W.w1t) + i);
  for (int i = tid; i < HID * HID / 4; i += FRONT_NT) reinterpret_cast<float4*>(s_w2)[i] = __ldg(reinterpret_cast<const float4*>(W.w2t) + i);
  for (int i = t; i < R * S; i += 256) {
    const int r = i / S, c = i % S;
    s_in[r][c] = (r0 + r < n) ? obs[(int64_t)(idx ? idx[r0 + r] : r0 + r) * S + c] : 0.f;
  }
  __syncthreads();
  const int row = t % R, og = t / R;   // og: outputs og*8 .. og*8+7 (uniform per warp)
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
  if (ok && h1) {
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
    acc[j] = ok ? fmaxf(acc[j], 0.f) : 0.f;   // rows beyond n are zero in the images
    hT[og * 8 + j][row] = acc[j];
  }
  if (ok) {
    float4* d = reinterpret_cast<float4*>(h2 + (int64_t)(r0 + row) * HID + og * 8);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  if (t < R && r0 + t < n) {
    float v = __ldg(W.bv);
    for (int k = 0; k < HID; ++k) v = fmaf(hT[k][t], __ldg(W.wv + k), v);
    value[r0 + t] = v;
  }
  if (!himg) return;
  // ---- the tile's h2 images (layout of head_tc_pack_h2_kernel); element (row r, hidden k) = hT of half r / 32
  float(*hA)[R + 1] = reinterpret_cast<float(*)[R + 1]>(s_w2 + HID * HID + 2 * 32 * 33);
  auto Hrk = [&](int r, int k) { return hA[(r >> 5) * HID + k][r & 31]; };
  const int64_t n64 = h2_tiles64(n_lay);
  float4* n_hi = reinterpret_cast<float4*>(himg + himg_n_off(tile));
  float4* n_lo = n_hi + B_BYTES / 16;
  float4* t_hi = reinterpret_cast<float4*>(himg + himg_t_off(n64, tile));
  float4* t_lo = t_hi + B_BYTES / 16;
  float4* a_hi = reinterpret_cast<float4*>(himg + himg_a_off(n64, tile >> 1));
  float4* a_lo = a_hi + A_BYTES / 16;
  auto split4 = [](float4 v, float4& h, float4& l) {
    h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    l = make_float4(tf32_hi(v.x - h.x), tf32_hi(v.y - h.y), tf32_hi(v.z - h.z), tf32_hi(v.w - h.w));
  };
  for (int i = tid; i < TN * HID / 4; i += FRONT_NT) {
    const int c4 = i / TN, r = i % TN;
    float4 h, l;
    split4(make_float4(Hrk(r, 4 * c4), Hrk(r, 4 * c4 + 1), Hrk(r, 4 * c4 + 2), Hrk(r, 4 * c4 + 3)), h, l);   // row = row r
    n_hi[i] = h; n_lo[i] = l;
    const int ia = c4 * TM + (tile & 1) * TN + r;
    a_hi[ia] = h; a_lo[ia] = l;
    split4(make_float4(Hrk(4 * c4, r), Hrk(4 * c4 + 1, r), Hrk(4 * c4 + 2, r), Hrk(4 * c4 + 3, r)), h, l);   // row = hidden r
    t_hi[i] = h; t_lo[i] = l;
  }
}

// (hi, lo) image pair of one operand tile: global -> registers (early) -> shared (late), plain 16-byte copies
template <int TILE_BYTES, int NTHREADS>
struct TileImg {
  static constexpr int STEPS = TILE_BYTES / 16 / NTHREADS;
  static_assert(TILE_BYTES / 16 % NTHREADS == 0, "tile / thread count");
  float4 h[STEPS], l[STEPS];
  __device__ __forceinline__ void load(int tid, const float* pair) {
    const float4* ph = reinterpret_cast<const float4*>(pair);
    const float4* pl = ph + TILE_BYTES / 16;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) { h[i] = __ldg(ph + tid + i * NTHREADS); l[i] = __ldg(pl + tid + i * NTHREADS); }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int tid) const {
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      reinterpret_cast<float4*>(hi)[tid + i * NTHREADS] = h[i];
      reinterpret_cast<float4*>(lo)[tid + i * NTHREADS] = l[i];
    }
  }
};

__global__ void __launch_bounds__(NT, 2)
head_tc_stats_kernel(HeadTc H, const int32_t* __restrict__ idx, const int32_t* __restrict__ act, int tiles_per_split,
                     int n_split, float* __restrict__ pm, float* __restrict__ ps, float* __restrict__ la) {
  extern __shared__ __align__(1024) char smem[];
  char* a_hi = smem;
  char* a_lo = a_hi + A_BYTES;
  char* b_hi = a_lo + A_BYTES;
  char* b_lo = b_hi + B_BYTES;
  float* sb3 = reinterpret_cast<float*>(b_lo + B_BYTES);
  float* sm = sb3 + 64;
  float* ss = sm + NT;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const int r0 = blockIdx.x * TM, split = blockIdx.y;
  const int n_tiles = (H.nA + TN - 1) / TN;
  const int ct0 = split * tiles_per_split, ct1 = min(n_tiles, ct0 + tiles_per_split);
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  {
    TileImg<A_BYTES, NT> ta;
    ta.load(tid, H.himg + himg_a_off(h2_tiles64(H.n), blockIdx.x));
    ta.store(a_hi, a_lo, tid);
  }
  TileImg<B_BYTES, NT> tb_;   // the NEXT catalogue tile of W3 (pre-split image), prefetched into registers
  float b3n = 0.f;
  auto prefetch = [&](int ct) {
    const int c0 = ct * TN;
    tb_.load(tid, H.img + img_n_off(ct));
    b3n = (tid < TN && c0 + tid < H.nA) ? __ldg(H.b3 + c0 + tid) : MASKED;
  };
  if (ct0 < ct1) prefetch(ct0);
  int a = -1;
  if (act != nullptr && r0 + row < H.n) a = act[idx ? idx[r0 + row] : r0 + row];
  float m = MASKED, s = 0.f, lav = 0.f;
  bool found = false;
  uint32_t ph = 0, tb = 0;
  for (int ct = ct0; ct < ct1; ++ct) {
    const int c0 = ct * TN;
    tb_.store(b_hi, b_lo, tid);
    if (tid < TN) sb3[tid] = b3n;
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    tb = tmem_base;
    if (tid == 0) {
      issue(tb, a_hi, a_lo, b_hi, b_lo, false);
      mma_commit(&bar);
    }
    if (ct + 1 < ct1) prefetch(ct + 1);     // global loads fly behind the MMA and the epilogue
    wait_or_flag(&bar, ph);
    ph ^= 1;
    fence_after_sync();
    float v[32];
    tmem_ld32(tmem_addr(tb, (warp & 3) * 32, half * 32), v);
    const int cb = c0 + half * 32;
    float mt = MASKED;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = v[j] + sb3[half * 32 + j];   // padding columns carry the MASKED bias
      v[j] = x;
      mt = fmaxf(mt, x);
    }
    if (a >= cb && a < cb + 32) {
      lav = pick32(v, a - cb);
      found = true;
    }
    if (mt > 0.5f * MASKED) {
      const float mn = fmaxf(m, mt);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += fast_exp(v[j] - mn);
      s = s * fast_exp(m - mn) + acc;
      m = mn;
    }
    fence_before_sync();
    __syncthreads();   // the B tile and the accumulator are free again
  }
  sm[tid] = m;
  ss[tid] = s;
  __syncthreads();
  if (half == 0 && r0 + row < H.n) {
    const float m1 = sm[tid + TM], s1 = ss[tid + TM];
    const float M = fmaxf(m, m1);
    const float S = s * fast_exp(m - M) + s1 * fast_exp(m1 - M);   // an empty half has s == 0
    pm[(size_t)(r0 + row) * n_split + split] = M;
    ps[(size_t)(r0 + row) * n_split + split] = S;
  }
  if (found) la[r0 + row] = lav;
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// D (+)= A[tmem hi/lo] . B[smem hi/lo], 3 TF32 products per k-step of 8
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, const char* b_hi, const char* b_lo,
                                         bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  const uint64_t bh0 = smem_desc(smem_u32(b_hi), B_LBO, SBO), bl0 = smem_desc(smem_u32(b_lo), B_LBO, SBO);
  constexpr uint64_t BS = B_STEP >> 4;
#pragma unroll
  for (int j = 0; j < KSTEPS; ++j) {
    mma_tf32_ts(d, a_lo + 8 * j, bh0 + j * BS, IDESC, acc);
    mma_tf32_ts(d, a_hi + 8 * j, bl0 + j * BS, IDESC, 1u);
    mma_tf32_ts(d, a_hi + 8 * j, bh0 + j * BS, IDESC, 1u);
    acc = 1u;
  }
}
// warp-collective form for the issuer warp of the TMA-fed kernels (tc_dev.cuh elect_one)
__device__ __forceinline__ void w_issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, const char* b_hi, const char* b_lo,
                                           bool accumulate) {
  if (elect_one()) issue_ts(d, a_hi, a_lo, b_hi, b_lo, accumulate);
}

// ------------------------------------------------------------------------------------------------- pass F, 128-column tiles
// Measured on the B200 (cirs_head_tc_debug_phases; profiles/r2c_head_phases.txt): an M128 N64 K8 kind::tf32 MMA with both
// operands in shared memory executes in ~58 cycles, not the tensor pipe's 32 -- it re-reads the 4 KB A slice for 2 KB of
// B, and shared memory feeds 128 B per clock.  With N = 128 an MMA reads 4 + 4 KB for twice the math.  This variant
// walks the catalogue in 128-column tiles: the B operand is the 128-row W3 image that pass B3 uses as its A operand
// (img_a_off), two 64 KB stages in shared memory, two 128-column accumulators in TMEM, and 16 epilogue warps (thread =
// TMEM lane x 32-column quarter).  One CTA per SM.
constexpr int WN = 128;
constexpr uint32_t IDESC_W = idesc_tf32(TM, WN, 0, 0);
constexpr size_t FW_SMEM = 2 * A_BYTES + 2 * 2 * A_BYTES + 4 * WN * 4 + 2 * NTB * 4;

template <bool PH>
__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_stats_wide_kernel(HeadTc H, const int32_t* __restrict__ idx, const int32_t* __restrict__ act,
                          int tiles_per_split, int n_split, float* __restrict__ pm, float* __restrict__ ps,
                          float* __restrict__ la, const int32_t* __restrict__ n_dev) {
  extern __shared__ __align__(1024) char smem[];
  // H.n sizes the grid and the h2 image layout; with n_dev the rows that exist are min(H.n, *n_dev)
  const int n_rows = n_dev ? min(H.n, __ldg(n_dev)) : H.n;
  if ((int)blockIdx.x * TM >= n_rows) return;
  char* a_hi = smem;
  char* a_lo = a_hi + A_BYTES;
  char* bring = a_lo + A_BYTES;                                            // 2 x { hi, lo } of a 128-column W3 tile
  float* sb3 = reinterpret_cast<float*>(bring + 2 * 2 * A_BYTES);          // ring of 4 x 128 bias values
  float* sm = sb3 + 4 * WN;
  float* ss = sm + NTB;
  __shared__ __align__(8) uint64_t tma_a, tma_b[2], mma[2], dfree[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, q = (tid >> 7) & 3;
  const bool worker = tid < NTB, issuer = __shfl_sync(FULL_MASK, warp, 0) == NTB / 32;
  const int r0 = blockIdx.x * TM, split = blockIdx.y;
  const int n_tiles = (H.nA + WN - 1) / WN;
  const int ct0 = split * tiles_per_split, T = min(n_tiles, ct0 + tiles_per_split) - ct0;
  const int64_t n64 = H.ldA / TN;
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (tid == 0) {
    mbar_init(&tma_a, 1); mbar_init(&tma_b[0], 1); mbar_init(&tma_b[1], 1);
    mbar_init(&mma[0], 1); mbar_init(&mma[1], 1);
    mbar_init(&dfree[0], NTB / 32); mbar_init(&dfree[1], NTB / 32);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (issuer && T > 0) {
    auto copy_tile = [&](int t) {   // tile t -> stage t & 1 (phase parity (t >> 1) & 1), bias -> ring slot t & 3
      const int ct = ct0 + t, b = t & 1;
      w_expect_tx(&tma_b[b], 2 * A_BYTES + WN * 4);
      w_bulk_g2s(bring + (size_t)b * 2 * A_BYTES, H.img + img_a_off(n64, ct), 2 * A_BYTES, &tma_b[b]);
      w_bulk_g2s(sb3 + WN * (t & 3), H.img + img_bias_off(n64) + (int64_t)ct * WN, WN * 4, &tma_b[b]);
    };
    w_expect_tx(&tma_a, 2 * A_BYTES);
    w_bulk_g2s(a_hi, H.himg + himg_a_off(h2_tiles64(H.n), blockIdx.x), 2 * A_BYTES, &tma_a);
    copy_tile(0);
    if (T > 1) copy_tile(1);
    wait_or_flag(&tma_a, 0);
    PhaseClock<PH> pc;
    pc.start(lane == 0);
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      wait_or_flag(&tma_b[b], (t >> 1) & 1);                          // tile t and its bias are in shared memory
      pc.lap(0);
      if (t >= 2) wait_or_flag(&dfree[b], ((t - 2) >> 1) & 1);        // accumulator b was read by epilogue(t-2)
      pc.lap(1);
      fence_after_sync();
      if (elect_one())
        mma_3xtf32(tb + (uint32_t)WN * b, smem_u32(a_hi), smem_u32(a_lo), A_STEP, A_LBO, SBO,
                   smem_u32(bring + (size_t)b * 2 * A_BYTES), smem_u32(bring + (size_t)b * 2 * A_BYTES + A_BYTES), A_STEP,
                   A_LBO, SBO, IDESC_W, KSTEPS, false);
      w_commit(&mma[b]);
      pc.lap(2);
      if (t >= 1 && t + 1 < T) {
        wait_or_flag(&mma[b ^ 1], ((t - 1) >> 1) & 1);                // MMA(t-1) released stage b ^ 1; bias slot (t+1) & 3
        copy_tile(t + 1);                                             // was tile t-3's
        pc.lap(3);
      }
    }
    pc.count(4, 2 * T);   // in 64-column units, comparable with the ring kernel's counters
  }
  int a = -1;
  if (worker && act != nullptr && r0 + row < n_rows) a = act[idx ? idx[r0 + row] : r0 + row];
  float m = MASKED, s = 0.f, lav = 0.f;
  bool found = false;
  if (worker) {
    PhaseClock<PH> pc;
    pc.start(tid == 0);
    for (int t = 0; t < T; ++t) {
      const int b = t & 1, cb = (ct0 + t) * WN + q * 32;
      wait_or_flag(&mma[b], (t >> 1) & 1);
      pc.lap(8);
      fence_after_sync();
      float v[32];
      tmem_ld32(tmem_addr(tb + (uint32_t)WN * b, (warp & 3) * 32, q * 32), v);
      pc.lap(9);
      float mt = MASKED;
      const float4* bias = reinterpret_cast<const float4*>(sb3 + WN * (t & 3) + q * 32);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 bb = bias[j4];   // padding columns carry the MASKED bias
        v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
      }
      {   // tree maximum (the serial chain of 32 dependent FMNMX was a third of this epilogue)
        float m8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m8[j] = fmaxf(fmaxf(v[j], v[j + 8]), fmaxf(v[j + 16], v[j + 24]));
        mt = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dfree[b]);
      pc.lap(10);
      if (a >= cb && a < cb + 32) {
        lav = pick32(v, a - cb);
        found = true;
      }
      if (mt > 0.5f * MASKED) {
        const float mn = fmaxf(m, mt);
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains
#pragma unroll
        for (int j = 0; j < 32; ++j) acc4[j & 3] += fast_exp(v[j] - mn);
        s = s * fast_exp(m - mn) + ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
        m = mn;
      }
      pc.lap(11);
    }
    sm[tid] = m;
    ss[tid] = s;
  }
  fence_before_sync();
  __syncthreads();
  if (worker && q == 0 && r0 + row < n_rows) {
    float M = m;
#pragma unroll
    for (int k = 1; k < 4; ++k) M = fmaxf(M, sm[tid + k * TM]);
    float S = s * fast_exp(m - M);
#pragma unroll
    for (int k = 1; k < 4; ++k) S += ss[tid + k * TM] * fast_exp(sm[tid + k * TM] - M);   // an empty quarter has s == 0
    pm[(size_t)(r0 + row) * n_split + split] = M;
    ps[(size_t)(r0 + row) * n_split + split] = S;
  }
  if (found) la[r0 + row] = lav;
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------- passes B2 / B3
// Software pipeline shared by both backward passes (tile t, buffers b = t & 1):
//     stage operands of tile t+1 -> MMA1(t+1) into the other logits accumulator   | overlaps
//     epilogue(t): TMEM logits -> d logits, written BACK TO TMEM as the (hi, lo) A operand  | MMA1(t+1) and MMA2(t-1)
//     MMA2(t): D2 += d logits[tmem] . B2[smem]
// d logits never touch shared memory (tcgen05.st), which frees the 64 KB that double-buffer the B operands.
// TMEM columns (512 allocated): logits accumulators [0,64) [64,128) | D2 [128,192) | d logits hi/lo, two buffers [192,448).
constexpr uint32_t T_D1 = 0, T_D2 = 128, T_DL = 192;
__device__ __forceinline__ uint32_t t_dl_hi(int b) { return T_DL + 128u * b; }
__device__ __forceinline__ uint32_t t_dl_lo(int b) { return T_DL + 128u * b + 64u; }

// 16 d-logit values of this thread's lane -> TMEM (hi by truncation, lo = x - hi: gradients need ~2^-22, tc_dev.cuh)
__device__ __forceinline__ void store_dl(uint32_t tb, int b, uint32_t lane_base, int col0, const float (&v)[16]) {
  float hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { hi[j] = tf32_trunc(v[j]); lo[j] = v[j] - hi[j]; }
  tmem_st16(tmem_addr(tb + t_dl_hi(b), lane_base, col0), hi);
  tmem_st16(tmem_addr(tb + t_dl_lo(b), lane_base, col0), lo);
  tmem_st_wait();
}

constexpr size_t B2_SMEM = 2 * A_BYTES + 6 * B_BYTES + 2 * 64 * 4 + NTB * 4;

__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_dh2_kernel(HeadTc H, const float* __restrict__ rowm, const float* __restrict__ rinvz,
                   const float* __restrict__ coef, const int32_t* __restrict__ acta, int tiles_per_split, int n_split,
                   float* __restrict__ dh2_part, float* __restrict__ ent_part) {
  extern __shared__ __align__(1024) char smem[];
  char* a_hi = smem;                         // h2 tile                            (A of MMA1)
  char* a_lo = a_hi + A_BYTES;
  char* bn = a_lo + A_BYTES;                 // 2 x { W3 tile r = column, c = hidden: hi, lo }   (B of MMA1)
  char* bk = bn + 4 * B_BYTES;               // 1 x { W3 tile r = hidden, c = column: hi, lo }   (B of MMA2)
  float* sb3 = reinterpret_cast<float*>(bk + 2 * B_BYTES);   // 2 x 64
  float* se = sb3 + 128;
  __shared__ __align__(8) uint64_t bar1[2], bar2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, qt = tid >> 7;   // qt: column quarter (16 columns)
  const bool worker = tid < NTB, issuer = tid == NTB;   // warp 16 only issues MMAs (the issue of a 24-MMA batch blocks
                                                        // for ~0.7 us: a worker that issued kept all others waiting)
  const int r0 = blockIdx.x * TM, split = blockIdx.y;
  const int n_tiles = (H.nA + TN - 1) / TN;
  const int ct0 = split * tiles_per_split, T = min(n_tiles, ct0 + tiles_per_split) - ct0;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar1[0], 1); mbar_init(&bar1[1], 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  if (worker) {
    TileImg<A_BYTES, NTB> ta;
    ta.load(tid, H.himg + himg_a_off(h2_tiles64(H.n), blockIdx.x));
    ta.store(a_hi, a_lo, tid);
  }
  // Register prefetch, one tile ahead of use: ``tn`` (transposed W3 tile, B of MMA1) holds tile t+1 while tile t is in
  // its epilogue; ``tk`` (natural W3 tile, B of MMA2) lags one tile behind it, because MMA2(t) is only issued after the
  // epilogue of tile t -- which lets its single shared buffer be rewritten late, after MMA2(t-1) has long finished.
  TileImg<B_BYTES, NTB> tn;
  TileImg<B_BYTES, NTB> tk;
  const int64_t n64 = H.ldA / TN;
  float b3n = 0.f;
  auto load_n = [&](int t) {
    if (!worker) return;
    const int c0 = (ct0 + t) * TN;
    tn.load(tid, H.img + img_n_off(ct0 + t));
    b3n = (tid < TN && c0 + tid < H.nA) ? __ldg(H.b3 + c0 + tid) : MASKED;
  };
  auto load_k = [&](int t) { if (worker) tk.load(tid, H.img + img_k_off(n64, ct0 + t)); };
  auto stage_n = [&](int b) {
    if (!worker) return;
    tn.store(bn + 2 * b * B_BYTES, bn + (2 * b + 1) * B_BYTES, tid);
    if (tid < TN) sb3[64 * b + tid] = b3n;
  };
  const bool live = worker && r0 + row < H.n;
  const float rm = live ? rowm[r0 + row] : 0.f, iz = live ? rinvz[r0 + row] : 0.f, cf = live ? coef[r0 + row] : 0.f;
  const int a = live ? acta[r0 + row] : -1;
  const float log_z = iz > 0.f ? -logf(iz) : 0.f;
  float ent = 0.f;
  uint32_t tb = 0;
  if (T > 0) {
    load_n(0);
    load_k(0);
    stage_n(0);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    tb = tmem_base;
    if (issuer) {
      issue(tb + T_D1, a_hi, a_lo, bn, bn + B_BYTES, false);
      mma_commit(&bar1[0]);
    }
    if (T > 1) load_n(1);
  }
  for (int t = 0; t < T; ++t) {
    const int b = t & 1, nb = b ^ 1;
    if (t + 1 < T) {   // logits MMA of the next tile runs behind this tile's epilogue
      stage_n(nb);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (issuer) {
        issue(tb + T_D1 + 64u * nb, a_hi, a_lo, bn + 2 * nb * B_BYTES, bn + (2 * nb + 1) * B_BYTES, false);
        mma_commit(&bar1[nb]);
      }
      if (t + 2 < T) load_n(t + 2);
    }
    if (worker) {
      wait_or_flag(&bar1[b], (t >> 1) & 1);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D1 + 64u * b, (warp & 3) * 32, qt * 16), v);
      const int cb = (ct0 + t) * TN + qt * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float xm = v[j] + sb3[64 * b + qt * 16 + j] - rm;   // logit - max (padding columns: -1e30)
        const float p = fast_exp(xm) * iz;
        // entropy of Categorical(probs): -sum p log(clamp(p, eps, 1 - eps)) with log clamp(p) = clamp(log p)
        const float lg = fminf(fmaxf(xm - log_z, LOG_EPS), LOG_1M_EPS);
        ent = fmaf(-p, lg, ent);
        v[j] = cf * ((cb + j == a ? 1.f : 0.f) - p);
      }
      store_dl(tb, b, (warp & 3) * 32, qt * 16, v);
      if (t >= 1) wait_or_flag(&bar2, (t - 1) & 1);   // MMA2(t-1), issued a whole epilogue ago, no longer reads bk
      tk.store(bk, bk + B_BYTES, tid);
      if (t + 1 < T) load_k(t + 1);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (issuer) {   // D2 += d logits . W3   (K = this tile's 64 columns); runs behind the next tile's staging + epilogue
      issue_ts(tb + T_D2, tb + t_dl_hi(b), tb + t_dl_lo(b), bk, bk + B_BYTES, t > 0);
      mma_commit(&bar2);
    }
  }
  if (T > 0 && worker) {
    wait_or_flag(&bar2, (T - 1) & 1);
    fence_after_sync();
    float v[16];
    tmem_ld16(tmem_addr(tb + T_D2, (warp & 3) * 32, qt * 16), v);
    if (live) {
      float* dst = dh2_part + ((size_t)split * H.n + r0 + row) * HID + qt * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  if (worker) se[tid] = ent;
  fence_before_sync();
  __syncthreads();
  if (qt == 0 && live)
    ent_part[(size_t)(r0 + row) * n_split + split] = (ent + se[tid + TM]) + (se[tid + 2 * TM] + se[tid + 3 * TM]);
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

constexpr size_t B3_SMEM = 2 * A_BYTES + 6 * B_BYTES + 2 * 4 * 64 * 4;

__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_dw3_kernel(HeadTc H, const float* __restrict__ rowm, const float* __restrict__ rinvz,
                   const float* __restrict__ coef, const int32_t* __restrict__ acta, int rows_per_split, int n_rsplit,
                   float* __restrict__ g_w3t, float* __restrict__ g_b3) {
  extern __shared__ __align__(1024) char smem[];
  char* wa_hi = smem;                        // W3 tile, r = column (128), c = hidden   (A of MMA1')
  char* wa_lo = wa_hi + A_BYTES;
  char* hb = wa_lo + A_BYTES;                // 2 x { h2 tile r = row (64), c = hidden: hi, lo }   (B of MMA1')
  char* ht = hb + 4 * B_BYTES;               // 1 x { h2 tile r = hidden, c = row: hi, lo }        (B of MMA3)
  float* stats = reinterpret_cast<float*>(ht + 2 * B_BYTES);   // 2 x { max[64], 1/Z[64], coef[64], action[64] }
  __shared__ __align__(8) uint64_t bar1[2], bar2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, cl = tid & 127, qt = tid >> 7;   // qt: quarter of 16 rows / hidden units
  const bool worker = tid < NTB, issuer = tid == NTB;   // warp 16 only issues MMAs
  const int c0 = blockIdx.x * TM, col = c0 + cl;
  const int rs0 = blockIdx.y * rows_per_split, rs1 = min(H.n, rs0 + rows_per_split);
  const int T = rs1 > rs0 ? (rs1 - rs0 + TN - 1) / TN : 0;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar1[0], 1); mbar_init(&bar1[1], 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  if (worker) {
    TileImg<A_BYTES, NTB> tw;
    tw.load(tid, H.img + img_a_off(H.ldA / TN, blockIdx.x));
    tw.store(wa_hi, wa_lo, tid);
  }
  // register prefetch as in pass B2: ``tv`` (natural h2 tile, B of MMA1') one tile ahead, ``tt`` (transposed, B of MMA3)
  // one tile behind it
  TileImg<B_BYTES, NTB> tv;
  TileImg<B_BYTES, NTB> tt;
  const int64_t h64 = h2_tiles64(H.n);
  float n_rm = 0.f, n_iz = 0.f, n_cf = 0.f;
  int n_ac = -1;
  auto load_v = [&](int t) {
    if (!worker) return;
    const int r0 = rs0 + t * TN;
    tv.load(tid, H.himg + himg_n_off(r0 / TN));
    const bool ok = tid < TN && r0 + tid < rs1;
    n_rm = ok ? rowm[r0 + tid] : 0.f;
    n_iz = ok ? rinvz[r0 + tid] : 0.f;
    n_cf = ok ? coef[r0 + tid] : 0.f;
    n_ac = ok ? acta[r0 + tid] : -1;
  };
  auto load_t = [&](int t) { if (worker) tt.load(tid, H.himg + himg_t_off(h64, rs0 / TN + t)); };
  auto stage_v = [&](int b) {
    if (!worker) return;
    tv.store(hb + 2 * b * B_BYTES, hb + (2 * b + 1) * B_BYTES, tid);
    if (tid < TN) {
      float* sp = stats + 256 * b;
      sp[tid] = n_rm; sp[64 + tid] = n_iz; sp[128 + tid] = n_cf; reinterpret_cast<int*>(sp)[192 + tid] = n_ac;
    }
  };
  const bool live = worker && col < H.nA;
  const float b3v = live ? __ldg(H.b3 + col) : MASKED;
  float db3 = 0.f;
  uint32_t tb = 0;
  if (T > 0) {
    load_v(0);
    load_t(0);
    stage_v(0);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    tb = tmem_base;
    if (issuer) {
      issue(tb + T_D1, wa_hi, wa_lo, hb, hb + B_BYTES, false);   // (logits tile)^T - b3: lane = column, 64 rows
      mma_commit(&bar1[0]);
    }
    if (T > 1) load_v(1);
  }
  for (int t = 0; t < T; ++t) {
    const int b = t & 1, nb = b ^ 1;
    if (t + 1 < T) {
      stage_v(nb);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (issuer) {
        issue(tb + T_D1 + 64u * nb, wa_hi, wa_lo, hb + 2 * nb * B_BYTES, hb + (2 * nb + 1) * B_BYTES, false);
        mma_commit(&bar1[nb]);
      }
      if (t + 2 < T) load_v(t + 2);
    }
    if (worker) {
      wait_or_flag(&bar1[b], (t >> 1) & 1);
      fence_after_sync();
      const float* sp = stats + 256 * b;
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D1 + 64u * b, (warp & 3) * 32, qt * 16), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int jj = qt * 16 + j;
        const float p = fast_exp(v[j] + b3v - sp[jj]) * sp[64 + jj];   // padding columns / rows: bias -1e30 or 1/Z = 0
        const float d = sp[128 + jj] * ((reinterpret_cast<const int*>(sp)[192 + jj] == col ? 1.f : 0.f) - p);
        db3 += d;
        v[j] = d;
      }
      store_dl(tb, b, (warp & 3) * 32, qt * 16, v);
      if (t >= 1) wait_or_flag(&bar2, (t - 1) & 1);   // MMA3(t-1) no longer reads ht
      tt.store(ht, ht + B_BYTES, tid);
      if (t + 1 < T) load_t(t + 1);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (issuer) {   // D3 += d logits^T . h2   (K = this tile's 64 rows)
      issue_ts(tb + T_D2, tb + t_dl_hi(b), tb + t_dl_lo(b), ht, ht + B_BYTES, t > 0);
      mma_commit(&bar2);
    }
  }
  if (T > 0 && worker) {
    wait_or_flag(&bar2, (T - 1) & 1);
    fence_after_sync();
    float v[16];
    tmem_ld16(tmem_addr(tb + T_D2, (warp & 3) * 32, qt * 16), v);
    if (live) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float* dst = g_w3t + (size_t)(qt * 16 + j) * H.ldA + col;
        if (n_rsplit > 1) atomicAdd(dst, v[j]); else *dst += v[j];
      }
      atomicAdd(g_b3 + col, db3);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------- passes B2 / B3, TMA-fed
// Same warp specialisation as pass F (TMA): warps 0-15 run epilogues only, lane 0 of warp 16 feeds the shared-memory
// operand buffers with cp.async.bulk and issues both MMAs of a tile.  The tensor pipe executes MMAs in issue order
// (MMA1(t+1) is issued before MMA2(t), MMA1(t+2) after it), so "MMA1(t+2) complete" implies "MMA2(t) complete": the
// workers may overwrite the d-logits operand buffer b of tile t when they start tile t+2 without a barrier of its own.
//   tma_n[b]  logits-MMA B tile (+ bias / row statistics) of tile t landed        issuer waits
//   tma_k     second-MMA B tile of tile t landed                                   issuer waits
//   mma1[b]   logits accumulator b holds tile t                                    workers wait
//   dlr[b]    (16 arrivals) d logits of tile t are in TMEM buffer b; accumulator b, bias / statistics buffer b consumed
//   mma2      second MMA of tile t complete (single B buffer free again)           issuer waits; workers at the end
// Both operand streams run TWO tiles ahead of the MMAs (double-buffered B tiles of both MMAs, bias in a ring of four):
// a 32 KB bulk copy takes ~1.5 us from L2, which the one-tile-ahead version paid twice per tile in the issuer's chain.
constexpr size_t B2T_SMEM = 2 * A_BYTES + 8 * B_BYTES + 4 * 64 * 4 + NTB * 4;

template <bool PH>
__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_dh2_tma_kernel(HeadTc H, const float* __restrict__ rowm, const float* __restrict__ rinvz,
                       const float* __restrict__ coef, const int32_t* __restrict__ acta, int tiles_per_split,
                       int n_split, float* __restrict__ dh2_part, float* __restrict__ ent_part) {
  extern __shared__ __align__(1024) char smem[];
  char* a_hi = smem;                         // h2 tile (hi, lo adjacent)          (A of MMA1)
  char* a_lo = a_hi + A_BYTES;
  char* bn = a_lo + A_BYTES;                 // 2 x { W3 tile r = column, c = hidden: hi, lo }   (B of MMA1)
  char* bk = bn + 4 * B_BYTES;               // 2 x { W3 tile r = hidden, c = column: hi, lo }   (B of MMA2)
  float* sb3 = reinterpret_cast<float*>(bk + 4 * B_BYTES);   // ring of 4 x 64 bias values (tile t in slot t & 3)
  float* se = sb3 + 256;
  __shared__ __align__(8) uint64_t tma_a, tma_n[2], tma_k[2], mma1[2], mma2, dlr[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, qt = (tid >> 7) & 3;
  const bool worker = tid < NTB, issuer = __shfl_sync(FULL_MASK, warp, 0) == NTB / 32;
  const int r0 = blockIdx.x * TM, split = blockIdx.y;
  const int n_tiles = (H.nA + TN - 1) / TN;
  const int ct0 = split * tiles_per_split, T = min(n_tiles, ct0 + tiles_per_split) - ct0;
  const int64_t n64 = H.ldA / TN;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) {
    mbar_init(&tma_a, 1); mbar_init(&tma_n[0], 1); mbar_init(&tma_n[1], 1); mbar_init(&tma_k[0], 1); mbar_init(&tma_k[1], 1);
    mbar_init(&mma1[0], 1); mbar_init(&mma1[1], 1); mbar_init(&mma2, 1);
    mbar_init(&dlr[0], NTB / 32); mbar_init(&dlr[1], NTB / 32);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (issuer && T > 0) {
    // tile t: logits-MMA B tile in bn[t & 1] (+ bias in ring slot t & 3) on tma_n[t & 1], second-MMA B tile in
    // bk[t & 1] on tma_k[t & 1]; each barrier's phase for tile t has parity (t >> 1) & 1
    auto copy_n = [&](int t) {
      const int ct = ct0 + t, b = t & 1;
      w_expect_tx(&tma_n[b], 2 * B_BYTES + 64 * 4);
      w_bulk_g2s(bn + 2 * b * B_BYTES, H.img + img_n_off(ct), 2 * B_BYTES, &tma_n[b]);
      w_bulk_g2s(sb3 + 64 * (t & 3), H.img + img_bias_off(n64) + (int64_t)ct * TN, 64 * 4, &tma_n[b]);
    };
    auto copy_k = [&](int t) {
      const int b = t & 1;
      w_expect_tx(&tma_k[b], 2 * B_BYTES);
      w_bulk_g2s(bk + 2 * b * B_BYTES, H.img + img_k_off(n64, ct0 + t), 2 * B_BYTES, &tma_k[b]);
    };
    w_expect_tx(&tma_a, 2 * A_BYTES);
    w_bulk_g2s(a_hi, H.himg + himg_a_off(h2_tiles64(H.n), blockIdx.x), 2 * A_BYTES, &tma_a);
    copy_n(0);
    if (T > 1) copy_n(1);
    copy_k(0);
    if (T > 1) copy_k(1);
    wait_or_flag(&tma_a, 0);
    wait_or_flag(&tma_n[0], 0);
    fence_after_sync();
    w_issue(tb + T_D1, a_hi, a_lo, bn, bn + B_BYTES, false);
    w_commit(&mma1[0]);
    PhaseClock<PH> pc;
    pc.start(lane == 0);
    pc.lap(16);                                                   // (start-up: first copies + MMA1(0) are not counted)
    for (int t = 0; t < T; ++t) {
      const int b = t & 1, nb = b ^ 1;
      if (t + 1 < T) {
        if (t >= 1) wait_or_flag(&dlr[nb], ((t - 1) >> 1) & 1);   // epilogue(t-1) consumed accumulator nb
        pc.lap(17);
        wait_or_flag(&tma_n[nb], ((t + 1) >> 1) & 1);             // copied one tile ago
        pc.lap(18);
        fence_after_sync();
        w_issue(tb + T_D1 + 64u * nb, a_hi, a_lo, bn + 2 * nb * B_BYTES, bn + (2 * nb + 1) * B_BYTES, false);
        w_commit(&mma1[nb]);
        pc.lap(19);
      }
      if (t + 2 < T) {
        wait_or_flag(&mma1[b], (t >> 1) & 1);                     // MMA1(t) no longer reads bn[b]; bias slot (t+2)&3 was
        copy_n(t + 2);                                            // tile t-2's, whose epilogue is long over
        pc.lap(20);
      }
      if (t >= 1 && t + 1 < T) {
        wait_or_flag(&mma2, (t - 1) & 1);                         // MMA2(t-1) no longer reads bk[nb]
        copy_k(t + 1);
        pc.lap(21);
      }
      wait_or_flag(&dlr[b], (t >> 1) & 1);                        // d logits of tile t are in TMEM
      pc.lap(22);
      wait_or_flag(&tma_k[b], (t >> 1) & 1);
      pc.lap(23);
      fence_after_sync();
      w_issue_ts(tb + T_D2, tb + t_dl_hi(b), tb + t_dl_lo(b), bk + 2 * b * B_BYTES, bk + (2 * b + 1) * B_BYTES, t > 0);
      w_commit(&mma2);
      pc.lap(24);
    }
    pc.count(25, T);
  }
  const bool live = worker && r0 + row < H.n;
  float ent = 0.f;
  if (worker) {
    const float rm = live ? rowm[r0 + row] : 0.f, iz = live ? rinvz[r0 + row] : 0.f, cf = live ? coef[r0 + row] : 0.f;
    const int a = live ? acta[r0 + row] : -1;
    const float log_z = iz > 0.f ? -logf(iz) : 0.f;
    PhaseClock<PH> pc;
    pc.start(tid == 0 || tid == NTB - 32);   // first and last worker warp
    const int ps = tid == 0 ? 32 : 40;
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      wait_or_flag(&mma1[b], (t >> 1) & 1);
      pc.lap(ps);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D1 + 64u * b, (warp & 3) * 32, qt * 16), v);
      pc.lap(ps + 1);
      const int cb = (ct0 + t) * TN + qt * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float xm = v[j] + sb3[64 * (t & 3) + qt * 16 + j] - rm;   // logit - max (padding columns: -1e30)
        const float p = fast_exp(xm) * iz;
        const float lg = fminf(fmaxf(xm - log_z, LOG_EPS), LOG_1M_EPS);
        ent = fmaf(-p, lg, ent);
        v[j] = cf * ((cb + j == a ? 1.f : 0.f) - p);
      }
      store_dl(tb, b, (warp & 3) * 32, qt * 16, v);
      pc.lap(ps + 2);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dlr[b]);
      // follow every phase of mma2 in order (a parity wait is only unambiguous one phase at a time); MMA2(t-1) was
      // issued a whole epilogue ago, so this does not stall
      if (t >= 1) wait_or_flag(&mma2, (t - 1) & 1);
      pc.lap(ps + 3);
    }
    if (T > 0) {
      wait_or_flag(&mma2, (T - 1) & 1);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D2, (warp & 3) * 32, qt * 16), v);
      if (live) {
        float* dst = dh2_part + ((size_t)split * H.n + r0 + row) * HID + qt * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    se[tid] = ent;
  }
  fence_before_sync();
  __syncthreads();
  if (qt == 0 && live)
    ent_part[(size_t)(r0 + row) * n_split + split] = (ent + se[tid + TM]) + (se[tid + 2 * TM] + se[tid + 3 * TM]);
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- pass B2 with ONE logits MMA batch per PAIR of 64-column tiles (experimental; CIRS_B2_WIDE=1) ------------------
// The logits of two neighbouring catalogue tiles come from M128 N128 K8 MMAs against the 128-row W3 image (the operand
// of pass F): 24 MMAs of ~74 cycles for 128 columns instead of 2 x 24 of ~51.  TMEM has no room for two 128-column
// logits accumulators AND two (hi, lo) d-logits operands, so the d-logits operand is single-buffered: the epilogue of
// tile q waits for MMA2(q - 1) right before it stores (the arithmetic in front of the store covers that MMA).  The
// 128-column B operand (64 KB) is single-staged in shared memory: the copy for pair p + 1 starts when MMA1(p) has
// completed and lands while the epilogue works on pair p.
//   TMEM: logits pair accumulators [0, 128) [128, 256) | D2 [256, 320) | d logits hi [320, 384) lo [384, 448)
//   tma_n     B operand + bias of pair p landed (phase parity p & 1)                  issuer waits
//   tma_k[s]  second-MMA B tile of tile q landed (s = q & 1, parity (q >> 1) & 1)       issuer waits
//   mma1[w]   pair accumulator w = p & 1 holds pair p (parity (p >> 1) & 1)            workers wait; issuer (B stage free)
//   d1r[w]    (16 arrivals) both halves of accumulator w are in registers             issuer waits before MMA1(p + 2)
//   dlr       (16 arrivals) d logits of tile q are in TMEM (parity q & 1)               issuer waits
//   mma2      second MMA of tile q complete (parity q & 1)                              workers (operand free), issuer (bk free)
constexpr uint32_t TW_D1 = 0, TW_D2 = 256, TW_DL = 320;
constexpr size_t B2W_SMEM = 2 * A_BYTES + 2 * A_BYTES + 4 * B_BYTES + 2 * WN * 4 + NTB * 4;

__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_dh2_wide_kernel(HeadTc H, const float* __restrict__ rowm, const float* __restrict__ rinvz,
                        const float* __restrict__ coef, const int32_t* __restrict__ acta, int pairs_per_split,
                        int n_split, float* __restrict__ dh2_part, float* __restrict__ ent_part) {
  extern __shared__ __align__(1024) char smem[];
  char* a_hi = smem;                         // h2 tile (hi, lo adjacent)                          (A of MMA1)
  char* a_lo = a_hi + A_BYTES;
  char* bn = a_lo + A_BYTES;                 // W3 pair tile r = column (128), c = hidden: hi, lo  (B of MMA1)
  char* bk = bn + 2 * A_BYTES;               // 2 x { W3 tile r = hidden, c = column (64): hi, lo } (B of MMA2)
  float* sb3 = reinterpret_cast<float*>(bk + 4 * B_BYTES);   // 2 x 128 bias values (pair p in slot p & 1)
  float* se = sb3 + 2 * WN;
  __shared__ __align__(8) uint64_t tma_a, tma_n, tma_k[2], mma1[2], d1r[2], dlr, mma2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, qt = (tid >> 7) & 3;
  const bool worker = tid < NTB, issuer = __shfl_sync(FULL_MASK, warp, 0) == NTB / 32;
  const int r0 = blockIdx.x * TM, split = blockIdx.y;
  const int n_pairs = (H.nA + WN - 1) / WN;
  const int cp0 = split * pairs_per_split, P = min(n_pairs, cp0 + pairs_per_split) - cp0, Q = 2 * P;
  const int64_t n64 = H.ldA / TN;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) {
    mbar_init(&tma_a, 1); mbar_init(&tma_n, 1); mbar_init(&tma_k[0], 1); mbar_init(&tma_k[1], 1);
    mbar_init(&mma1[0], 1); mbar_init(&mma1[1], 1); mbar_init(&mma2, 1);
    mbar_init(&d1r[0], NTB / 32); mbar_init(&d1r[1], NTB / 32); mbar_init(&dlr, NTB / 32);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (issuer && P > 0) {
    auto copy_n = [&](int p) {   // pair p: 64 KB operand image + 128 bias values
      w_expect_tx(&tma_n, 2 * A_BYTES + WN * 4);
      w_bulk_g2s(bn, H.img + img_a_off(n64, cp0 + p), 2 * A_BYTES, &tma_n);
      w_bulk_g2s(sb3 + WN * (p & 1), H.img + img_bias_off(n64) + (int64_t)(cp0 + p) * WN, WN * 4, &tma_n);
    };
    auto copy_k = [&](int q) {   // 64-column tile q of this CTA (catalogue tile 2 cp0 + q)
      const int s = q & 1;
      w_expect_tx(&tma_k[s], 2 * B_BYTES);
      w_bulk_g2s(bk + 2 * s * B_BYTES, H.img + img_k_off(n64, 2 * cp0 + q), 2 * B_BYTES, &tma_k[s]);
    };
    auto issue_mma1 = [&](int p) {
      const int w = p & 1;
      if (elect_one())
        mma_3xtf32(tb + TW_D1 + (uint32_t)WN * w, smem_u32(a_hi), smem_u32(a_lo), A_STEP, A_LBO, SBO, smem_u32(bn),
                   smem_u32(bn + A_BYTES), A_STEP, A_LBO, SBO, IDESC_W, KSTEPS, false);
      w_commit(&mma1[w]);
    };
    w_expect_tx(&tma_a, 2 * A_BYTES);
    w_bulk_g2s(a_hi, H.himg + himg_a_off(h2_tiles64(H.n), blockIdx.x), 2 * A_BYTES, &tma_a);
    copy_n(0);
    copy_k(0);
    copy_k(1);
    wait_or_flag(&tma_a, 0);
    wait_or_flag(&tma_n, 0);
    fence_after_sync();
    issue_mma1(0);
    for (int p = 0; p < P; ++p) {
      const int w = p & 1;
      if (p + 1 < P) {
        wait_or_flag(&mma1[w], (p >> 1) & 1);          // MMA1(p) complete: the B stage is free again (bias slot (p+1) & 1
        copy_n(p + 1);                                 // was pair p-1's: its last epilogue ended before MMA2(2p - 1))
      }
      for (int h = 0; h < 2; ++h) {
        const int q = 2 * p + h, s = q & 1;
        if (q >= 1 && q + 1 < Q) {
          wait_or_flag(&mma2, (q - 1) & 1);            // MMA2(q-1) complete: bk[(q+1) & 1] is free
          copy_k(q + 1);
        }
        wait_or_flag(&dlr, q & 1);                     // d logits of tile q are in TMEM
        wait_or_flag(&tma_k[s], (q >> 1) & 1);
        fence_after_sync();
        w_issue_ts(tb + TW_D2, tb + TW_DL, tb + TW_DL + 64u, bk + 2 * s * B_BYTES, bk + (2 * s + 1) * B_BYTES, q > 0);
        w_commit(&mma2);
        if (h == 0 && p + 1 < P) {                     // the next pair's logits, behind this pair's second half
          wait_or_flag(&tma_n, (p + 1) & 1);
          if (p >= 1) wait_or_flag(&d1r[w ^ 1], ((p - 1) >> 1) & 1);   // epilogue(p-1) has read accumulator w ^ 1
          fence_after_sync();
          issue_mma1(p + 1);
        }
      }
    }
  }
  const bool live = worker && r0 + row < H.n;
  float ent = 0.f;
  if (worker) {
    const float rm = live ? rowm[r0 + row] : 0.f, iz = live ? rinvz[r0 + row] : 0.f, cf = live ? coef[r0 + row] : 0.f;
    const int a = live ? acta[r0 + row] : -1;
    const float log_z = iz > 0.f ? -logf(iz) : 0.f;
    for (int q = 0; q < Q; ++q) {
      const int p = q >> 1, h = q & 1, w = p & 1;
      if (h == 0) wait_or_flag(&mma1[w], (p >> 1) & 1);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + TW_D1 + (uint32_t)WN * w + 64u * h, (warp & 3) * 32, qt * 16), v);
      if (h == 1) {   // both halves of accumulator w are in registers: the issuer may overwrite it (pair p + 2)
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d1r[w]);
      }
      const int cb = (2 * cp0 + q) * TN + qt * 16;
      const float* bias = sb3 + WN * (p & 1) + 64 * h + qt * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float xm = v[j] + bias[j] - rm;   // logit - max (padding columns: -1e30)
        const float pr = fast_exp(xm) * iz;
        const float lg = fminf(fmaxf(xm - log_z, LOG_EPS), LOG_1M_EPS);
        ent = fmaf(-pr, lg, ent);
        v[j] = cf * ((cb + j == a ? 1.f : 0.f) - pr);
      }
      if (q >= 1) wait_or_flag(&mma2, (q - 1) & 1);   // MMA2(q-1) has read the (single) d-logits operand
      {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { hi[j] = tf32_trunc(v[j]); lo[j] = v[j] - hi[j]; }
        tmem_st16(tmem_addr(tb + TW_DL, (warp & 3) * 32, qt * 16), hi);
        tmem_st16(tmem_addr(tb + TW_DL + 64u, (warp & 3) * 32, qt * 16), lo);
        tmem_st_wait();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dlr);
    }
    if (Q > 0) {
      wait_or_flag(&mma2, (Q - 1) & 1);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + TW_D2, (warp & 3) * 32, qt * 16), v);
      if (live) {
        float* dst = dh2_part + ((size_t)split * H.n + r0 + row) * HID + qt * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(dst + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    } else if (live) {   // a split beyond the catalogue's pairs: its partial is zero
      float* dst = dh2_part + ((size_t)split * H.n + r0 + row) * HID + qt * 16;
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(dst + 4 * k) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    se[tid] = ent;
  }
  fence_before_sync();
  __syncthreads();
  if (qt == 0 && live)
    ent_part[(size_t)(r0 + row) * n_split + split] = (ent + se[tid + TM]) + (se[tid + 2 * TM] + se[tid + 3 * TM]);
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// pass B3, TMA-fed.  Row statistics of a 64-row tile (max, 1/Z, coef, action: four contiguous 256-byte slices of the
// per-row arrays, which the caller pads with zeros / -1 up to a multiple of 64 rows) travel with the h2 tile.
constexpr size_t B3T_SMEM = 2 * A_BYTES + 8 * B_BYTES + 4 * 4 * 64 * 4;

__global__ void __launch_bounds__(NTB + 32, 1)
head_tc_dw3_tma_kernel(HeadTc H, const float* __restrict__ rowm, const float* __restrict__ rinvz,
                       const float* __restrict__ coef, const int32_t* __restrict__ acta, int rows_per_split,
                       int n_rsplit, float* __restrict__ g_w3t, float* __restrict__ g_b3) {
  extern __shared__ __align__(1024) char smem[];
  char* wa_hi = smem;                        // W3 tile, r = column (128), c = hidden (hi, lo adjacent)   (A of MMA1')
  char* wa_lo = wa_hi + A_BYTES;
  char* hb = wa_lo + A_BYTES;                // 2 x { h2 tile r = row (64), c = hidden: hi, lo }   (B of MMA1')
  char* ht = hb + 4 * B_BYTES;               // 2 x { h2 tile r = hidden, c = row: hi, lo }        (B of MMA3)
  float* stats = reinterpret_cast<float*>(ht + 4 * B_BYTES);   // ring of 4 x { max[64], 1/Z[64], coef[64], action[64] }
  __shared__ __align__(8) uint64_t tma_a, tma_n[2], tma_k[2], mma1[2], mma2, dlr[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cl = tid & 127, qt = (tid >> 7) & 3;
  const bool worker = tid < NTB, issuer = __shfl_sync(FULL_MASK, warp, 0) == NTB / 32;
  const int c0 = blockIdx.x * TM, col = c0 + cl;
  const int rs0 = blockIdx.y * rows_per_split, rs1 = min(H.n, rs0 + rows_per_split);
  const int T = rs1 > rs0 ? (rs1 - rs0 + TN - 1) / TN : 0;
  const int64_t h64 = h2_tiles64(H.n);
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) {
    mbar_init(&tma_a, 1); mbar_init(&tma_n[0], 1); mbar_init(&tma_n[1], 1); mbar_init(&tma_k[0], 1); mbar_init(&tma_k[1], 1);
    mbar_init(&mma1[0], 1); mbar_init(&mma1[1], 1); mbar_init(&mma2, 1);
    mbar_init(&dlr[0], NTB / 32); mbar_init(&dlr[1], NTB / 32);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (issuer && T > 0) {
    // same two-tiles-ahead operand streams as pass B2 (row statistics in ring slot t & 3)
    auto copy_n = [&](int t) {
      const int r0 = rs0 + t * TN, b = t & 1;
      float* sp = stats + 256 * (t & 3);
      w_expect_tx(&tma_n[b], 2 * B_BYTES + 4 * 64 * 4);
      w_bulk_g2s(hb + 2 * b * B_BYTES, H.himg + himg_n_off(r0 / TN), 2 * B_BYTES, &tma_n[b]);
      w_bulk_g2s(sp, rowm + r0, 64 * 4, &tma_n[b]);
      w_bulk_g2s(sp + 64, rinvz + r0, 64 * 4, &tma_n[b]);
      w_bulk_g2s(sp + 128, coef + r0, 64 * 4, &tma_n[b]);
      w_bulk_g2s(sp + 192, acta + r0, 64 * 4, &tma_n[b]);
    };
    auto copy_k = [&](int t) {
      const int b = t & 1;
      w_expect_tx(&tma_k[b], 2 * B_BYTES);
      w_bulk_g2s(ht + 2 * b * B_BYTES, H.himg + himg_t_off(h64, rs0 / TN + t), 2 * B_BYTES, &tma_k[b]);
    };
    w_expect_tx(&tma_a, 2 * A_BYTES);
    w_bulk_g2s(wa_hi, H.img + img_a_off(H.ldA / TN, blockIdx.x), 2 * A_BYTES, &tma_a);
    copy_n(0);
    if (T > 1) copy_n(1);
    copy_k(0);
    if (T > 1) copy_k(1);
    wait_or_flag(&tma_a, 0);
    wait_or_flag(&tma_n[0], 0);
    fence_after_sync();
    w_issue(tb + T_D1, wa_hi, wa_lo, hb, hb + B_BYTES, false);   // (logits tile)^T - b3: lane = column, 64 rows
    w_commit(&mma1[0]);
    for (int t = 0; t < T; ++t) {
      const int b = t & 1, nb = b ^ 1;
      if (t + 1 < T) {
        if (t >= 1) wait_or_flag(&dlr[nb], ((t - 1) >> 1) & 1);
        wait_or_flag(&tma_n[nb], ((t + 1) >> 1) & 1);
        fence_after_sync();
        w_issue(tb + T_D1 + 64u * nb, wa_hi, wa_lo, hb + 2 * nb * B_BYTES, hb + (2 * nb + 1) * B_BYTES, false);
        w_commit(&mma1[nb]);
      }
      if (t + 2 < T) {
        wait_or_flag(&mma1[b], (t >> 1) & 1);
        copy_n(t + 2);
      }
      if (t >= 1 && t + 1 < T) {
        wait_or_flag(&mma2, (t - 1) & 1);
        copy_k(t + 1);
      }
      wait_or_flag(&dlr[b], (t >> 1) & 1);
      wait_or_flag(&tma_k[b], (t >> 1) & 1);
      fence_after_sync();
      w_issue_ts(tb + T_D2, tb + t_dl_hi(b), tb + t_dl_lo(b), ht + 2 * b * B_BYTES, ht + (2 * b + 1) * B_BYTES,
                 t > 0);   // D3 += d logits^T . h2
      w_commit(&mma2);
    }
  }
  const bool live = worker && col < H.nA;
  if (worker) {
    const float b3v = live ? __ldg(H.b3 + col) : MASKED;
    float db3 = 0.f;
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      wait_or_flag(&mma1[b], (t >> 1) & 1);
      fence_after_sync();
      const float* sp = stats + 256 * (t & 3);
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D1 + 64u * b, (warp & 3) * 32, qt * 16), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int jj = qt * 16 + j;
        const float p = fast_exp(v[j] + b3v - sp[jj]) * sp[64 + jj];   // padding columns / rows: bias -1e30 or 1/Z = 0
        const float d = sp[128 + jj] * ((reinterpret_cast<const int*>(sp)[192 + jj] == col ? 1.f : 0.f) - p);
        db3 += d;
        v[j] = d;
      }
      store_dl(tb, b, (warp & 3) * 32, qt * 16, v);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dlr[b]);
      if (t >= 1) wait_or_flag(&mma2, (t - 1) & 1);   // follow every phase of mma2 in order
    }
    if (T > 0) {
      wait_or_flag(&mma2, (T - 1) & 1);
      fence_after_sync();
      float v[16];
      tmem_ld16(tmem_addr(tb + T_D2, (warp & 3) * 32, qt * 16), v);
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float* dst = g_w3t + (size_t)(qt * 16 + j) * H.ldA + col;
          if (n_rsplit > 1) atomicAdd(dst, v[j]); else *dst += v[j];
        }
        atomicAdd(g_b3 + col, db3);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <class K>
void set_smem(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
void tiles_for(int nA, int n_split, int* tiles_per_split) {
  const int n_tiles = (nA + TN - 1) / TN;
  *tiles_per_split = (n_tiles + n_split - 1) / n_split;
}
}  // namespace

int64_t head_tc_image_floats(int64_t ldA) {
  const int64_t n64 = ldA / TN;
  return 2 * n64 * IMG_B + (n64 / 2) * IMG_A + ldA;
}

int head_tc_pack(const float* w3t, int64_t ldA, const float* b3, int nA, float* img, cudaStream_t st) {
  CIRS_LAUNCH(head_tc_pack_kernel, (int)(ldA / TN), 256, 0, st, w3t, ldA, b3, nA, img);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

int64_t head_tc_h2_image_floats(int64_t n) {
  const int64_t n64 = h2_tiles64(n > 0 ? n : 1);
  return 2 * n64 * IMG_B + (n64 / 2) * IMG_A;
}

int head_tc_pack_h2(const float* h2, int n, float* himg, cudaStream_t st) {
  CIRS_LAUNCH(head_tc_pack_h2_kernel, (int)h2_tiles64(n), 256, 0, st, h2, n, himg);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

int head_tc_front(const cirs_policy_weights* w, int n, const int32_t* idx, const float* obs, float* h1, float* h2,
                  float* value, float* himg, float* img, cudaStream_t st, const int32_t* n_dev) {
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(head_tc_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM);
    once = true;
  }
  const int n_pack = img ? (int)(w->ld_action / TN) : 0;
  const int tiles = himg ? (int)h2_tiles64(n) : (n + TN - 1) / TN;
  CIRS_LAUNCH(head_tc_front_kernel, n_pack + tiles, FRONT_NT, FRONT_SMEM, st, *w, n, idx, obs, h1, h2, value, himg,
              n_pack, img, n_dev);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

static bool g_phase_host = false;   // launch the phase-stamped instantiations (cirs_head_tc_debug_phases)
static int g_tc_mode = -1;   // -1: environment default (CIRS_NO_TC), 0: off, 1: on (TMA-fed kernels), 2: on, register-staged
static bool tma_on() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("CIRS_NO_TMA");
    env = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return env == 1 && g_tc_mode != 2;
}
// Environment switches of the tile pipelines (A/B measurements; defaults are the fast paths)
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
// Catalogue splits of passes F / B2 (one CTA per SM): the split count that minimises  waves x (tiles per CTA + a fixed
// per-CTA cost of ~2 tile times: TMEM allocation, pipeline fill, the final partial store).  CIRS_TC_PLAN=0 restores the
// round-1 rule (about two CTAs per SM's worth of work items).
// tile = columns per tile (64, or 128 for the wide pass F); fixed = per-CTA fixed cost in tile times
static int plan_split_slots(int n, int nA, int slots, int tile = TN, int fixed = 2) {
  const int n_tiles = (nA + tile - 1) / tile, row_tiles = (n + TM - 1) / TM;
  static int plan = -1;
  if (plan < 0) plan = env_int("CIRS_TC_PLAN", 1);
  if (plan == 0 || row_tiles <= 0) {
    int want = (2 * 148 + row_tiles - 1) / (row_tiles > 0 ? row_tiles : 1);
    if (want > MAX_SPLIT) want = MAX_SPLIT;
    if (want > n_tiles) want = n_tiles;
    if (want < 1) want = 1;
    const int per = (n_tiles + want - 1) / want;
    return (n_tiles + per - 1) / per;
  }
  int best = 1, best_cost = 1 << 30;
  for (int want = 1; want <= MAX_SPLIT && want <= n_tiles; ++want) {
    const int per = (n_tiles + want - 1) / want, ns = (n_tiles + per - 1) / per;
    const int waves = (row_tiles * ns + slots - 1) / slots;
    const int cost = waves * (per + fixed);
    if (cost < best_cost || (cost == best_cost && ns > best)) { best_cost = cost; best = ns; }
  }
  return best;
}
int plan_split(int n, int nA) { return plan_split_slots(n, nA, 148); }
// pass F has its own split count: its partials (pm, ps) are merged separately from pass B2's (d h2, entropy)
int plan_split_f(int n, int nA) {
  if (tma_on()) return plan_split_slots(n, nA, 148, WN, 1);   // 128-column tiles, one CTA per SM
  return plan_split_slots(n, nA, 296);                        // register-staged kernel: two CTAs per SM
}

bool head_tc_enabled(int n, int nA, int64_t ldA) {
  if (g_tc_mode < 0) {
    const char* e = getenv("CIRS_NO_TC");
    g_tc_mode = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return g_tc_mode >= 1 && n > 0 && nA >= TN && (ldA % 128) == 0 && ldA >= nA;
}

int head_tc_stats(const HeadTc& H, const int32_t* idx, const int32_t* act, int n_split, float* pm, float* ps,
                  float* la, cudaStream_t st, const int32_t* n_dev) {
  static bool once = false;
  if (!once) {
    set_smem(head_tc_stats_kernel, F_SMEM);
    set_smem(head_tc_stats_wide_kernel<false>, FW_SMEM);
    set_smem(head_tc_stats_wide_kernel<true>, FW_SMEM);
    once = true;
  }
  int per;
  tiles_for(H.nA, n_split, &per);
  dim3 grid((H.n + TM - 1) / TM, n_split);
  if (tma_on()) {   // 128-column tiles, warp-specialised and bulk-copy fed (CIRS_NO_TMA=1 / mode 2: register-staged kernel)
    const int n_tiles = (H.nA + WN - 1) / WN, perw = (n_tiles + n_split - 1) / n_split;
    if (g_phase_host)
      CIRS_LAUNCH(head_tc_stats_wide_kernel<true>, grid, NTB + 32, FW_SMEM, st, H, idx, act, perw, n_split, pm, ps, la,
                  n_dev);
    else
      CIRS_LAUNCH(head_tc_stats_wide_kernel<false>, grid, NTB + 32, FW_SMEM, st, H, idx, act, perw, n_split, pm, ps, la,
                  n_dev);
    CIRS_CHECK_LAUNCH();
    return CIRS_OK;
  }
  if (n_dev) {
    cirs_set_error("head_tc_stats: a device-resident row count needs the bulk-copy fed pass F (CIRS_NO_TMA unset, mode 1)");
    return CIRS_ERR_ARG;
  }
  CIRS_LAUNCH(head_tc_stats_kernel, grid, NT, F_SMEM, st, H, idx, act, per, n_split, pm, ps, la);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

int head_tc_dh2(const HeadTc& H, const float* rowm, const float* rinvz, const float* coef, const int32_t* acta,
                int n_split, float* dh2_part, float* ent_part, cudaStream_t st) {
  static bool once = false;
  if (!once) {
    set_smem(head_tc_dh2_kernel, B2_SMEM);
    set_smem(head_tc_dh2_tma_kernel<false>, B2T_SMEM);
    set_smem(head_tc_dh2_wide_kernel, B2W_SMEM);
    set_smem(head_tc_dh2_tma_kernel<true>, B2T_SMEM);
    once = true;
  }
  int per;
  tiles_for(H.nA, n_split, &per);
  dim3 grid((H.n + TM - 1) / TM, n_split);
  static int b2_wide = -1;
  if (b2_wide < 0) b2_wide = env_int("CIRS_B2_WIDE", 0) ? 1 : 0;
  if (tma_on() && b2_wide && !g_phase_host) {   // one N = 128 logits MMA batch per pair of tiles (experimental)
    const int n_pairs = (H.nA + WN - 1) / WN, perp = (n_pairs + n_split - 1) / n_split;
    CIRS_LAUNCH(head_tc_dh2_wide_kernel, grid, NTB + 32, B2W_SMEM, st, H, rowm, rinvz, coef, acta, perp, n_split, dh2_part,
                ent_part);
    CIRS_CHECK_LAUNCH();
    return CIRS_OK;
  }
  if (tma_on()) {
    if (g_phase_host)
      CIRS_LAUNCH(head_tc_dh2_tma_kernel<true>, grid, NTB + 32, B2T_SMEM, st, H, rowm, rinvz, coef, acta, per, n_split,
                  dh2_part, ent_part);
    else
      CIRS_LAUNCH(head_tc_dh2_tma_kernel<false>, grid, NTB + 32, B2T_SMEM, st, H, rowm, rinvz, coef, acta, per, n_split,
                  dh2_part, ent_part);
    CIRS_CHECK_LAUNCH();
    return CIRS_OK;
  }
  CIRS_LAUNCH(head_tc_dh2_kernel, grid, NTB + 32, B2_SMEM, st, H, rowm, rinvz, coef, acta, per, n_split, dh2_part, ent_part);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

int head_tc_dw3(const HeadTc& H, const float* rowm, const float* rinvz, const float* coef, const int32_t* acta,
                float* g_w3t, float* g_b3, cudaStream_t st) {
  static bool once = false;
  if (!once) {
    set_smem(head_tc_dw3_kernel, B3_SMEM);
    set_smem(head_tc_dw3_tma_kernel, B3T_SMEM);
    once = true;
  }
  const int n_ct = (H.nA + TM - 1) / TM, row_tiles = (H.n + TN - 1) / TN;
  int n_rsplit = (3 * 148 + n_ct - 1) / n_ct;
  if (n_rsplit > row_tiles) n_rsplit = row_tiles;
  if (n_rsplit < 1) n_rsplit = 1;
  int tiles_per = (row_tiles + n_rsplit - 1) / n_rsplit;
  n_rsplit = (row_tiles + tiles_per - 1) / tiles_per;
  dim3 grid(n_ct, n_rsplit);
  if (tma_on()) {   // needs the per-row arrays padded with zeros / -1 up to a multiple of 64 rows (row_loss_tc_kernel)
    CIRS_LAUNCH(head_tc_dw3_tma_kernel, grid, NTB + 32, B3T_SMEM, st, H, rowm, rinvz, coef, acta, tiles_per * TN, n_rsplit,
                g_w3t, g_b3);
    CIRS_CHECK_LAUNCH();
    return CIRS_OK;
  }
  CIRS_LAUNCH(head_tc_dw3_kernel, grid, NTB + 32, B3_SMEM, st, H, rowm, rinvz, coef, acta, tiles_per * TN, n_rsplit, g_w3t,
              g_b3);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

}  // namespace cirs_head_tc

extern "C" void cirs_head_tc_enable(int on) { cirs_head_tc::g_tc_mode = on < 0 ? -1 : (on > 2 ? 1 : on); }

// debug: per-phase cycle counters of pass F (head_tc_stats_wide_kernel) and pass B2 (head_tc_dh2_tma_kernel), see g_phase.
// enable != 0 makes the launchers pick the stamped instantiations; out (64 values, host memory, may be null) receives
// the counters accumulated so far; reset != 0 clears them.  Synchronises the device.
extern "C" int cirs_head_tc_debug_phases(int32_t enable, int64_t* out64_h, int32_t reset) {
  cudaDeviceSynchronize();
  if (out64_h) cudaMemcpyFromSymbol(out64_h, cirs_head_tc::g_phase, 64 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(cirs_head_tc::g_phase, z, sizeof(z));
  }
  cirs_head_tc::g_phase_host = enable != 0;
  return cudaGetLastError() == cudaSuccess ? CIRS_OK : CIRS_ERR_CUDA;
}

// debug: 1 if any tensor-core kernel gave up waiting on an mbarrier since the last call (synchronises the device)
extern "C" int cirs_head_tc_timeout(void) {
  int v = 0, z = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&v, cirs_head_tc::g_tc_timeout, sizeof(int));
  if (v) cudaMemcpyToSymbol(cirs_head_tc::g_tc_timeout, &z, sizeof(int));
  return v;
}
// stream-ordered copy of the flag into PINNED host memory (no synchronisation): the product path folds it into the
// update's own read-back and raises when it is set (policy.py)
extern "C" int cirs_head_tc_timeout_peek(int32_t* out_pinned_h, void* stream) {
  if (!out_pinned_h) {
    cirs_set_error("cirs_head_tc_timeout_peek: null argument");
    return CIRS_ERR_ARG;
  }
  cudaError_t e = cudaMemcpyFromSymbolAsync(out_pinned_h, cirs_head_tc::g_tc_timeout, sizeof(int), 0,
                                            cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    cirs_set_error(cudaGetErrorString(e));
    return CIRS_ERR_CUDA;
  }
  return CIRS_OK;
}
