// Actor head + exponential-race sampler of the persistent rollout kernel on tcgen05 tensor cores (3xTF32).
//
// Decomposition: CTA s owns the catalogue slice [80 s, 80 s + 80) for the WHOLE rollout: its W3 columns are split
// into TF32 (hi, lo) operand tiles in shared memory once per launch.  Every turn the CTA walks the running
// environments in tiles of 128 rows: gathers their trunk outputs h2 (FP32, [B, 64]) into a K-major (hi, lo) tile,
// issues D[128 x 80] = h2 . W3_slice as 3 x 8 kind::tf32 MMAs into one of two TMEM accumulators, and while the
// tensor core works on the next row tile the 256 threads run the epilogue of the previous one straight out of TMEM:
// softmax partials (max, sum) and the slice's candidate action by the two-level sampler described at the epilogue
// (inverse CDF inside a 40-column unit, exponential race across units; Philox4x32-10 keyed by (environment, unit)).
// The draws differ from the FFMA path's per-element race (actor_dev.cuh) -- same distribution, different stream;
// MODE_ARGMAX is deterministic and identical on both paths.
// One Partial per (slice, row) goes to the workspace; actor_combine_warp merges the slices unchanged.
#pragma once
#include "actor_dev.cuh"
#include "tc_dev.cuh"

namespace cirs_actor_tc {
using namespace cirs_actor;
using namespace cirs_tc;

constexpr int SLICE = 80;                       // catalogue columns per CTA (10728 items -> 135 slices <= 148 SMs)
constexpr int ROWS = 128;                       // MMA M
constexpr uint32_t A_BYTES = ROWS * HID * 4;    // h2 tile (hi or lo)
constexpr uint32_t W_BYTES = SLICE * HID * 4;   // W3 slice (hi or lo)
constexpr uint32_t A_LBO = ROWS * 16, A_STEP = 2 * ROWS * 16;
constexpr uint32_t W_LBO = SLICE * 16, W_STEP = 2 * SLICE * 16;
constexpr uint32_t IDESC = idesc_tf32(ROWS, SLICE, 0, 0);
constexpr size_t TURN_BYTES = 2 * A_BYTES + 5 * NT * 4;              // rebuilt every turn (may alias phase-B scratch)
constexpr size_t KEEP_BYTES = 2 * W_BYTES + SLICE * 4 + 64;          // lives for the whole launch
constexpr int TMEM_COLS = 256;                                       // two accumulators at columns 0 and 128

struct TcSmem {
  char *a_hi, *a_lo;    // [TURN]  h2 row tile
  float* red;           // [TURN]  5 x NT floats: merge of the two column halves of a row
  char *w_hi, *w_lo;    // [KEEP]  W3 slice, tile row = catalogue column, tile column = hidden index
  float* b3;            // [KEEP]  bias of the slice
  uint64_t* bar;        // [KEEP]  2 mbarriers (one per accumulator)
  uint32_t* tmem;       // [KEEP]  TMEM base address
};
__device__ __forceinline__ TcSmem tc_carve(char* turn_region, char* keep_region) {
  TcSmem S;
  S.a_hi = turn_region; S.a_lo = S.a_hi + A_BYTES; S.red = reinterpret_cast<float*>(S.a_lo + A_BYTES);
  S.w_hi = keep_region; S.w_lo = S.w_hi + W_BYTES; S.b3 = reinterpret_cast<float*>(S.w_lo + W_BYTES);
  S.bar = reinterpret_cast<uint64_t*>(S.b3 + SLICE);
  S.tmem = reinterpret_cast<uint32_t*>(S.bar + 2);
  return S;
}

struct TcState { uint32_t use0, use1; };   // completed phases of the two mbarriers

// once per launch; every thread of the CTA calls it
__device__ __forceinline__ void tc_setup(const cirs_policy_weights& W, int slice, int n_slices, const TcSmem& S, int tid) {
  if (tid < 32) tmem_alloc(S.tmem, TMEM_COLS);
  if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
  if (slice < n_slices) {
    const int c0 = slice * SLICE;
    const int64_t ldA = W.ld_action;
    const float* w3t = W.w3t;
    TileT<SLICE, HID, NT> tw;
    tw.load(tid, [&](int r, int c) { return c0 + r < ldA ? __ldg(w3t + (size_t)c * ldA + c0 + r) : 0.f; });
    tw.store(S.w_hi, S.w_lo, tid);
    if (tid < SLICE) S.b3[tid] = c0 + tid < W.n_action ? __ldg(W.b3 + c0 + tid) : 0.f;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
}
__device__ __forceinline__ void tc_teardown(const TcSmem& S, int tid) {
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(*S.tmem, TMEM_COLS);
}

// The h2 rows of the running environments reach phase A as ready-made operand-tile IMAGES in global memory: phase B
// writes every row's trunk output, already split into TF32 (hi, lo), at its position in the NEXT turn's compact row
// list ([tile][hi | lo][A_BYTES], the shared-memory tile layout byte for byte), so staging a row tile is a straight
// 16-byte copy of the tile's first `rows` rows -- no gather, no split arithmetic, nothing for rows that are not there.
__device__ __forceinline__ void h2_image_store(char* img, int kn, int o, float v) {   // row kn of the list, hidden o
  char* tile = img + (size_t)(kn / ROWS) * (2 * A_BYTES);
  const uint32_t off = tile_chunk_off(ROWS, kn % ROWS, o >> 2) + (uint32_t)(o & 3) * 4u;
  const float h = tf32_hi(v);
  *reinterpret_cast<float*>(tile + off) = h;
  *reinterpret_cast<float*>(tile + A_BYTES + off) = tf32_hi(v - h);
}
struct RawTile {   // register prefetch of one tile image (behind the previous tile's MMA + epilogue)
  float4 h[ROWS * HID / 4 / NT], l[ROWS * HID / 4 / NT];
  __device__ __forceinline__ void load(int tid, const char* tile, int rows) {
#pragma unroll
    for (int i = 0; i < ROWS * HID / 4 / NT; ++i) {
      const int idx = tid + i * NT;
      if ((idx & (ROWS - 1)) < rows) {   // written during this launch by other CTAs: L2 loads (ld.global.cg)
        h[i] = __ldcg(reinterpret_cast<const float4*>(tile + (size_t)idx * 16));
        l[i] = __ldcg(reinterpret_cast<const float4*>(tile + A_BYTES + (size_t)idx * 16));
      }
    }
  }
  __device__ __forceinline__ void store(char* hi, char* lo, int tid, int rows) const {
#pragma unroll
    for (int i = 0; i < ROWS * HID / 4 / NT; ++i) {
      const int idx = tid + i * NT;
      if ((idx & (ROWS - 1)) < rows) {
        *reinterpret_cast<float4*>(hi + (size_t)idx * 16) = h[i];
        *reinterpret_cast<float4*>(lo + (size_t)idx * 16) = l[i];
      }
    }
  }
};

// One turn's actor-head partials for this CTA's slice.  P: n_rows / gather (compact list of running environments),
// part (n_split == number of slices), mode, seed, offset, rng_counter.  Every thread of the CTA calls it.
__device__ __forceinline__ void tc_head_turn(const HeadArgs& P, const char* h2_img, int slice, const TcSmem& S, int tid,
                                             TcState& st, int* timeout_flag, long long* tq = nullptr) {
  // tq (optional, one CTA): %globaltimer stamps {entry, h2 tile staged, first MMA done, epilogue done, partials written}
  auto stamp = [&](int i) {
    if (tq && tid == 0) { long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); tq[i] = t_; }
  };
  stamp(0);
  const int warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const bool warp0 = __shfl_sync(0xffffffffu, warp, 0) == 0;   // warp-uniform: the MMAs are issued under elect.sync (tc_dev.cuh)
  const int n_rt = (P.n_rows + ROWS - 1) / ROWS;
  const uint32_t tb = *S.tmem;
  const int nA = P.W.n_action;
  const int seen_words = (nA + 31) >> 5;
  const uint64_t offset = P.offset + (P.rng_counter ? (uint64_t)*P.rng_counter : 0ull);
  auto issue = [&](int rt) {   // called by all lanes of warp 0
    if (elect_one()) {
      mma_3xtf32(tb + 128u * (rt & 1), smem_u32(S.a_hi), smem_u32(S.a_lo), A_STEP, A_LBO, 128u, smem_u32(S.w_hi),
                 smem_u32(S.w_lo), W_STEP, W_LBO, 128u, IDESC, HID / 8, false);
      mma_commit(&S.bar[rt & 1]);
    }
    __syncwarp();
  };
  auto rows_of = [&](int rt) { return min(ROWS, P.n_rows - rt * ROWS); };
  {   // first tile: cp.async straight into the operand tile (nothing to hide it behind)
    const int rows = rows_of(0);
    for (int idx = tid; idx < ROWS * HID / 4; idx += NT)
      if ((idx & (ROWS - 1)) < rows) {
        cp_async16(S.a_hi + (size_t)idx * 16, h2_img + (size_t)idx * 16);
        cp_async16(S.a_lo + (size_t)idx * 16, h2_img + A_BYTES + (size_t)idx * 16);
      }
    cp_async_commit();
    cp_async_wait_all();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp0) issue(0);
  stamp(1);
  RawTile ta;
  if (n_rt > 1) ta.load(tid, h2_img + (size_t)1 * 2 * A_BYTES, rows_of(1));
  for (int rt = 0; rt < n_rt; ++rt) {
    const int b = rt & 1;
    uint32_t& use = b ? st.use1 : st.use0;
    if (!mbar_wait(&S.bar[b], use & 1u)) *timeout_flag = 1;
    ++use;
    fence_after_sync();
    if (rt == 0) stamp(2);
    if (rt + 1 < n_rt) {   // the h2 tile is free again: next row tile's MMA runs behind this tile's epilogue
      ta.store(S.a_hi, S.a_lo, tid, rows_of(rt + 1));
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (warp0) issue(rt + 1);
      if (rt + 2 < n_rt) ta.load(tid, h2_img + (size_t)(rt + 2) * 2 * A_BYTES, rows_of(rt + 2));
    }
    // ---- epilogue: thread = (row, column half of 40)
    const int k = rt * ROWS + row;
    const bool live = k < P.n_rows;
    const int rid = live ? (P.gather ? P.gather[k] : k) : -1;
    // Two-level sampler.  The thread's 40 columns form one UNIT of the catalogue.  Inside the unit the candidate is
    // drawn by inverse CDF (one uniform: P(j | unit) = e_j / z_unit); across the 270 units of a row the winner is the
    // exponential race over the unit masses (one Exp(1) draw q per unit: P(unit) = mass_unit / Z), carried in the log
    // domain as best_s = m + log z - log q, which is exactly what merge_best / actor_combine_warp compare.  Hence
    // P(j) = softmax(logits)_j as for Categorical.sample (core/policy/ppo.py:135), with 2 random numbers per unit instead
    // of one Philox draw + one log per ELEMENT (the per-element race was ~110 instructions per element, 8.2 us per row
    // tile; this is ~12).  MODE_ARGMAX: the unit's first maximum, best_s = its logit (ties -> lowest index in the merge).
    float m = -INFINITY, z = 0.f, bl = 0.f, bs = -INFINITY;
    int bi = 0x7fffffff;
    const int lbase = half * (SLICE / 2);          // column offset inside the slice
    const int cbase = slice * SLICE + lbase;       // catalogue column
    constexpr int NCH = SLICE / 16;                // chunks of 8 columns per thread
    constexpr int NC = SLICE / 2;                  // columns per thread
    uint32_t vr[NCH][8];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tmem_ld8_issue(tmem_addr(tb + 128u * b, (warp & 3) * 32, lbase + ch * 8), vr[ch]);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tmem_ld8_wait(vr[ch]);
    if (live) {
      float l[NC];
#pragma unroll
      for (int g = 0; g < NC / 4; ++g) {
        const int nb = cbase + 4 * g;
        // remove_recommended_ids (core/policy/utils.py:30-58): items already shown this episode leave the distribution
        uint32_t seen_bits = 0u;
        if (P.seen && nb < nA) seen_bits = P.seen[(size_t)rid * seen_words + (nb >> 5)] >> (nb & 31);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = nb + j < nA && !((seen_bits >> j) & 1u);
          l[4 * g + j] = ok ? __uint_as_float(vr[g >> 1][4 * (g & 1) + j]) + S.b3[lbase + 4 * g + j] : -INFINITY;
          m = fmaxf(m, l[4 * g + j]);
        }
      }
      if (m != -INFINITY) {
#pragma unroll
        for (int j = 0; j < NC; ++j) z += __expf(l[j] - m);
        if (P.mode == MODE_SAMPLE) {
          const uint4 rnd = philox4x32(make_uint4((uint32_t)rid, 0x40000000u + (uint32_t)(2 * slice + half),
                                                  (uint32_t)offset, (uint32_t)(offset >> 32)),
                                       make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
          const float q = fmaxf(-__logf(u01(rnd.x)), 1e-30f);            // Exp(1)
          const float target = (rnd.y >> 8) * (1.0f / 16777216.0f) * z;   // uniform in [0, z)
          float acc = 0.f, last_l = 0.f;
          int last_j = 0;
          bool found = false;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const float e = __expf(l[j] - m);   // the same values, in the same order, as the sum above
            acc += e;
            if (e > 0.f) { last_j = j; last_l = l[j]; }
            if (!found && e > 0.f && acc > target) { found = true; bi = j; bl = l[j]; }
          }
          if (!found) { bi = last_j; bl = last_l; }   // target rounded up to z
          bi += cbase;
          bs = m + __logf(z) - __logf(q);
        } else {
#pragma unroll
          for (int j = NC - 1; j >= 0; --j)
            if (l[j] == m) bi = cbase + j;
          bl = m;
          bs = m;
        }
      }
    }
    if (rt == 0) stamp(3);
    S.red[tid] = m; S.red[NT + tid] = z; S.red[2 * NT + tid] = bs; S.red[3 * NT + tid] = bl;
    S.red[4 * NT + tid] = __int_as_float(bi);
    fence_before_sync();
    __syncthreads();
    if (half == 0 && live) {
      const int o = tid + ROWS;
      merge_ms(m, z, S.red[o], S.red[NT + o]);
      merge_best(bs, bl, bi, S.red[2 * NT + o], S.red[3 * NT + o], __float_as_int(S.red[4 * NT + o]));
      Partial p;
      p.m = m; p.z = z; p.best_s = bs; p.best_l = bl; p.best_i = bi;
      P.part[(size_t)slice * P.n_rows + k] = p;
    }
    __syncthreads();   // red is reused by the next row tile
    if (rt == 0) stamp(4);
  }
}

}  // namespace cirs_actor_tc
