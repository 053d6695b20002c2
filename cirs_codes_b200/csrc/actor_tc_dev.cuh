// Actor head + exponential-race sampler of the persistent rollout kernel on tcgen05 tensor cores (3xTF32).
//
// Decomposition: CTA s owns the catalogue slice [80 s, 80 s + 80) for the WHOLE rollout: its W3 columns are split
// into TF32 (hi, lo) operand tiles in shared memory once per launch.  Every turn the CTA walks the running
// environments in tiles of 128 rows: gathers their trunk outputs h2 (FP32, [B, 64]) into a K-major (hi, lo) tile,
// issues D[128 x 80] = h2 . W3_slice as 3 x 8 kind::tf32 MMAs into one of two TMEM accumulators, and while the
// tensor core works on the next row tile the 256 threads run the epilogue of the previous one straight out of TMEM:
// online softmax (max, sum) and the running winner of the race  argmax_j logit_j + Gumbel_j  with the SAME Philox
// stream, keyed by (environment, column / 4), as the FFMA path (actor_dev.cuh) -- so both paths draw identical noise.
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

struct SrcRows {   // gathered h2 rows of one 128-row tile of the compact row list
  const float* h2; const int32_t* gather; int k0, n_rows;
  __device__ __forceinline__ float4 operator()(int r, int c4) const {
    const int k = k0 + r;
    if (k >= n_rows) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int id = gather ? gather[k] : k;
    return *reinterpret_cast<const float4*>(h2 + (size_t)id * HID + 4 * c4);   // written this launch: no __ldg
  }
};

// One turn's actor-head partials for this CTA's slice.  P: n_rows / gather (compact list of running environments),
// h2_in, part (n_split == number of slices), mode, seed, offset, rng_counter.  Every thread of the CTA calls it.
__device__ __forceinline__ void tc_head_turn(const HeadArgs& P, int slice, const TcSmem& S, int tid, TcState& st,
                                             int* timeout_flag, long long* tq = nullptr) {
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
  TileV<ROWS, HID, NT> ta;
  ta.load(tid, SrcRows{P.h2_in, P.gather, 0, P.n_rows});
  ta.store(S.a_hi, S.a_lo, tid);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp0) issue(0);
  stamp(1);
  if (n_rt > 1) ta.load(tid, SrcRows{P.h2_in, P.gather, ROWS, P.n_rows});
  for (int rt = 0; rt < n_rt; ++rt) {
    const int b = rt & 1;
    uint32_t& use = b ? st.use1 : st.use0;
    if (!mbar_wait(&S.bar[b], use & 1u)) *timeout_flag = 1;
    ++use;
    fence_after_sync();
    if (rt == 0) stamp(2);
    if (rt + 1 < n_rt) {   // the h2 tile is free again: next row tile's MMA runs behind this tile's epilogue
      ta.store(S.a_hi, S.a_lo, tid);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (warp0) issue(rt + 1);
      if (rt + 2 < n_rt) ta.load(tid, SrcRows{P.h2_in, P.gather, (rt + 2) * ROWS, P.n_rows});
    }
    // ---- epilogue: thread = (row, column half of 40)
    const int k = rt * ROWS + row;
    const bool live = k < P.n_rows;
    const int rid = live ? (P.gather ? P.gather[k] : k) : -1;
    // Race in the probability domain: candidate j beats the running winner iff  e_j / q_j > e_best / q_best  with
    // e = exp(l - m) relative to the running maximum (rescaled together with the softmax sum when m grows) and
    // q = -log(u) ~ Exp(1).  One fast log per element instead of the two of the Gumbel form; the winner's score
    // l - log(q) (log domain, what the cross-slice merge compares) is formed once per row tile.
    float m = -INFINITY, z = 0.f, e_best = 0.f, q_best = 1.f, bl = 0.f;
    int bi = 0x7fffffff;
    const int lbase = half * (SLICE / 2);          // column offset inside the slice
    const int cbase = slice * SLICE + lbase;       // catalogue column
    constexpr int NCH = SLICE / 16;                // chunks of 8 columns per thread
    uint32_t vr[2][8];
    tmem_ld8_issue(tmem_addr(tb + 128u * b, (warp & 3) * 32, lbase), vr[0]);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      tmem_ld8_wait(vr[ch & 1]);
      if (ch + 1 < NCH) tmem_ld8_issue(tmem_addr(tb + 128u * b, (warp & 3) * 32, lbase + (ch + 1) * 8), vr[(ch + 1) & 1]);
      if (!live) continue;
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        const int nb = cbase + ch * 8 + 4 * q4;
        if (nb >= nA) continue;
        float qn[4] = {1.f, 1.f, 1.f, 1.f};
        if (P.mode == MODE_SAMPLE && !P.icdf) {
          const uint4 rnd = philox4x32(make_uint4((uint32_t)rid, (uint32_t)(nb >> 2), (uint32_t)offset,
                                                  (uint32_t)(offset >> 32)),
                                       make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
          const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float u = u01(rr[j]), w = 1.0f - u;   // u in (0, 1], w exact
            const float ql = w * fmaf(w, fmaf(w, 0.33333334f, 0.5f), 1.0f);   // -log(1 - w) for small w
            qn[j] = fmaxf(w < 0.00390625f ? ql : -__logf(u), 1e-30f);
          }
        }
        // remove_recommended_ids (core/policy/utils.py:30-58): items already shown this episode leave the distribution
        uint32_t seen_bits = 0u;
        if (P.seen) seen_bits = P.seen[(size_t)rid * seen_words + (nb >> 5)] >> (nb & 31);
        float l[4], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = nb + j < nA && !((seen_bits >> j) & 1u);
          l[j] = ok ? __uint_as_float(vr[ch & 1][4 * q4 + j]) + S.b3[lbase + ch * 8 + 4 * q4 + j] : -INFINITY;
          mx = fmaxf(mx, l[j]);
        }
        if (mx == -INFINITY) continue;
        if (mx > m) {
          const float sc = __expf(m - mx);   // m = -inf at the start: exp(-inf) = 0
          z *= sc;
          e_best *= sc;
          m = mx;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float e = __expf(l[j] - m);   // 0 for masked columns
          z += e;
          if (e * q_best > qn[j] * e_best) { e_best = e; q_best = qn[j]; bl = l[j]; bi = nb + j; }
        }
      }
    }
    float bs = bi == 0x7fffffff ? -INFINITY : bl - logf(q_best);
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
