// Persistent rollout kernel: a whole Collector.collect(n_episode = B) in ONE cooperative launch.
//
// Replaces the reference's while-loop over turns (core/collector.py:219-320), i.e. per turn policy.forward
// (core/policy/ppo.py:111-163) -> env.step (simulated_env.py:111-168, kuaishouEnv.py:161-218) -> build_state
// (core/state_tracker.py:225-248) -> buffer.add (tianshou/data/buffer/manager.py:91-142), for all B environments.
//
// Why one kernel: a turn's work is tiny (~0.3 KB of algorithmic HBM traffic per env-step), so a
// launch-per-component rollout is bound by launch latency and dependent-kernel drain (measured: 2.1 ms of GPU time
// per collect as a 150-node CUDA graph, of which most is idle turns).  Here the grid stays resident; each turn is
// two phases separated by grid-wide barriers:
//   A  actor head: (64-row tile x catalogue split) work items, trunk + logits tiles + online softmax / race partials
//   B  per environment (one warp): merge the split partials -> action, log-prob; environment step; tracker token
//      against the K/V cache; trajectory written straight into the replay buffer's slots
// and the loop ends as soon as the device-side count of running environments reaches zero.
// Device code is shared with the stand-alone kernels (actor_dev.cuh, env_dev.cuh, tracker_dev.cuh), so both paths
// compute bit-identical results.
#include <cooperative_groups.h>

#include "actor_dev.cuh"
#include "actor_tc_dev.cuh"
#include "head_tc.cuh"
#include "env_dev.cuh"
#include "tracker_dev.cuh"
#include "tracker_cta_dev.cuh"
#include "tracker_fast_dev.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace cirs_actor;

__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct RolloutArgs {
  cirs_kuaishou_env E;
  cirs_tracker_weights T;
  HeadArgs H;               // policy weights, mode, seed / rng counter, partial workspace; state = cur_state
  int n_env, max_steps, force_length, traj_len;
  const int32_t* users;
  uint8_t* active;
  int32_t* act;             // [B] last action per slot
  float *logp, *value;      // [B]
  float* cur_state;         // [B, S]
  float *rew;               // [B]
  uint8_t* done;            // [B]
  float *traj_obs, *traj_obs_next;
  int32_t* traj_act;
  float* traj_rew;
  uint8_t* traj_done;
  int32_t* ep_len;
  float *kcache, *vcache;
  int* n_active;            // device counter of running environments
  int* count;               // [2] length of the compact list of turn t (index t & 1)
  int32_t* list;            // [2, B] environments still running at turn t
  float* h2;                // [B, 64] trunk output of the current state of every environment
  char* h2_img;             // tensor-core head: the same rows as TF32 (hi, lo) operand-tile images in list order
  const float* w_lo;        // tracker weights (everything but the embedding tables): start and float count
  int w_count;
  int smem_w_off;           // float offset of the staged weights inside dynamic shared memory
  long long* dbg;           // [1 + 3 * max_steps] turns played, then per turn {n_active, ns phase A, ns phase B}
  int row_floats;           // per-row scratch of the cooperative token step + 128 floats for the trunk
  // phase-B staging inside the per-turn shared region (behind the RB rows of scratch; float offsets, -1 = absent):
  int part_off;             // CTA_RB * 128 floats: partial sums of the split FFN product
  int trunk_w_off;          // the trunk's weights (w1t | b1 | w2t | b2 | wv | bv), re-fetched every turn with cp.async
  int kv_off, kv_cap;       // prefetch area for the rows' cached K / V positions and its capacity in floats
  // tensor-core head (actor_tc_dev.cuh): CTA s owns catalogue slice s
  int n_slices;             // ceil(n_action / 80) <= grid
  int tc_keep_off;          // float offset of the launch-lifetime shared-memory region (W3 slice, mbarriers)
  int* tc_timeout;          // set when an mbarrier wait gives up (never expected)
  // latency form of phase B (tracker_fast_dev.cuh; d = 32): transposed weights staged at smem_w_off, trunk image at
  // trunk_w_off (rebuilt by every CTA at launch, re-fetched per turn from trunk_img)
  cirs_tfast::Layout fast;
  float* trunk_img;         // [TR_FLOATS] global copy of the transposed trunk (written by CTA 0 at launch)
};

// Trunk + critic of R rows by the whole CTA, in the operation order of actor_trunk_warp / actor_head_body (bias first, k
// ascending, fmaf) so that h2 is bit-identical to the stand-alone kernels'.  The rows' states sit in the token step's
// scratch at sc[r * stride + state_off]; the last 128 floats of a row's scratch hold h1 | h2.
// layout of the trunk's weights when staged in shared memory (floats)
__host__ __device__ inline int trunk_w_floats(int S) { return S * HID + HID + HID * HID + HID + HID + 4; }

__device__ __forceinline__ void trunk_rows(const cirs_policy_weights& W, int R, float* sc, int stride, int state_off,
                                           const int* row_e, float* __restrict__ h2_out, float* __restrict__ value_out,
                                           const float* ws = nullptr, char* h2_img = nullptr,
                                           const int* row_kn = nullptr) {
  const int S = W.dim_state, tid = threadIdx.x;
  const int H1 = stride - 128, H2 = stride - 64;
  // ws: w1t [S][64] | b1 | w2t [64][64] | b2 | wv | bv in shared memory; otherwise the weights come from global memory
  const float* w1t = ws ? ws : W.w1t;
  const float* b1 = ws ? ws + S * HID : W.b1;
  const float* w2t = ws ? b1 + HID : W.w2t;
  const float* b2 = ws ? w2t + HID * HID : W.b2;
  const float* wv = ws ? b2 + HID : W.wv;
  const float* bv = ws ? wv + HID : W.bv;
  for (int idx = tid; idx < R * HID; idx += NT) {
    const int r = idx / HID, o = idx % HID;
    const float* s = sc + (size_t)r * stride + state_off;
    float a = b1[o];
    float wv1[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) wv1[k] = k < S ? w1t[(size_t)k * HID + o] : 0.f;   // every load in flight at once
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < S) a = fmaf(s[k], wv1[k], a);
    sc[(size_t)r * stride + H1 + o] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int idx = tid; idx < R * HID; idx += NT) {
    const int r = idx / HID, o = idx % HID;
    const float* h1 = sc + (size_t)r * stride + H1;
    float a = b2[o];
#pragma unroll
    for (int k0 = 0; k0 < HID; k0 += 32) {
      float w2[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) w2[k] = w2t[(size_t)(k0 + k) * HID + o];
#pragma unroll
      for (int k = 0; k < 32; ++k) a = fmaf(h1[k0 + k], w2[k], a);
    }
    a = fmaxf(a, 0.f);
    sc[(size_t)r * stride + H2 + o] = a;
    if (h2_img) {   // position row_kn[r] of the next turn's row list (< 0: the episode ended)
      if (row_kn[r] >= 0) cirs_actor_tc::h2_image_store(h2_img, row_kn[r], o, a);
    } else {
      h2_out[(size_t)row_e[r] * HID + o] = a;
    }
  }
  __syncthreads();
  if (tid < R && value_out) {
    const float* h2 = sc + (size_t)tid * stride + H2;
    float v = bv[0];
    for (int k = 0; k < HID; ++k) v = fmaf(h2[k], wv[k], v);
    value_out[row_e[tid]] = v;
  }
  __syncthreads();
}

// cp.async the trunk's weights into shared memory (issued at the start of phase B, consumed at its end)
__device__ __forceinline__ void prefetch_trunk(const cirs_policy_weights& W, float* ws) {
  const int S = W.dim_state, tid = threadIdx.x;
  float* b1 = ws + S * HID;
  float* w2t = b1 + HID;
  float* b2 = w2t + HID * HID;
  float* wv = b2 + HID;
  float* bv = wv + HID;
  for (int i = tid; i < S * HID / 4; i += NT) cp_async16(ws + 4 * i, W.w1t + 4 * i);
  for (int i = tid; i < HID * HID / 4; i += NT) cp_async16(w2t + 4 * i, W.w2t + 4 * i);
  if (tid < HID / 4) {
    cp_async16(b1 + 4 * tid, W.b1 + 4 * tid);
    cp_async16(b2 + 4 * tid, W.b2 + 4 * tid);
    cp_async16(wv + 4 * tid, W.wv + 4 * tid);
  }
  if (tid == 0) bv[0] = __ldg(W.bv);
}

// cp.async the cached K / V rows (positions 0 .. p-1, every layer) of the R rows into shared memory
__device__ __forceinline__ void prefetch_kv(const cirs_tracker_weights& W, int n_env, int R, const int* row_e, int p,
                                            const float* __restrict__ kcache, const float* __restrict__ vcache,
                                            float* kv_s, int kv_ld) {
  const int d = W.d, d4 = d >> 2, nl = W.nlayers;
  const int total = R * nl * 2 * p * d4;
  for (int i = threadIdx.x; i < total; i += NT) {
    int x = i;
    const int c4 = x % d4; x /= d4;
    const int j = x % p; x /= p;
    const int which = x & 1; x >>= 1;
    const int l = x % nl, r = x / nl;
    const float* src = (which ? vcache : kcache) + (((size_t)l * n_env + row_e[r]) * W.max_len + j) * d + 4 * c4;
    cp_async16(kv_s + ((size_t)((r * nl + l) * 2 + which) * p + j) * kv_ld + 4 * c4, src);
  }
}

// SMW: the tracker's weights are staged once in shared memory (they are re-read by every warp at every turn; from L2
// each token is ~55 dependent round trips of ~0.6 us, from shared memory ~20x less)
// TC: phase A on the tcgen05 tensor cores (3xTF32), each CTA keeping its slice of W3 in shared memory for the launch
// FAST: phase B in its latency form (tracker_fast_dev.cuh; implies SMW and TC)
template <bool SMW, bool TC, bool FAST = false>
__global__ void __launch_bounds__(NT, (SMW || TC) ? 1 : 2) rollout_kuaishou_kernel(const RolloutArgs A) {
  extern __shared__ __align__(128) float smem_dyn[];
  // The two structs that change during the launch (tracker weight pointers rebased into shared memory; the head's
  // per-turn row list / split plan) live ONCE per CTA in shared memory.  Writing to the kernel parameter instead makes
  // the compiler keep a private copy of the whole argument block in local memory for every thread (1.1 KB x 256),
  // which cannot stay in the small L1 left beside 226 KB of shared memory: every pointer fetch became an L2 round
  // trip and dominated the per-environment phase.
  __shared__ cirs_tracker_weights sT;
  __shared__ HeadArgs sH;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warps_per_cta = NT / 32;
  const int gwarp = blockIdx.x * warps_per_cta + warp, n_warps = gridDim.x * warps_per_cta;
  const int B = A.n_env;
  if (tid == 0) { sT = A.T; sH = A.H; }
  __syncthreads();
  if (FAST) {
    cirs_tfast::stage_tracker(A.T, A.fast, smem_dyn + A.smem_w_off);
    cirs_tfast::stage_trunk(A.H.W, smem_dyn + A.trunk_w_off);
    if (blockIdx.x == 0)
      for (int i = tid; i < cirs_tfast::TR_FLOATS; i += NT) A.trunk_img[i] = smem_dyn[A.trunk_w_off + i];
  } else if (SMW) {
    float* wsm = smem_dyn + A.smem_w_off;
    for (int i = tid; i < A.w_count / 4; i += NT)
      reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(A.w_lo) + i);
    if (tid == 0) {
      const ptrdiff_t shift = wsm - A.w_lo;   // rebase every weight pointer into the staged copy
      sT.user_wt += shift; sT.user_b += shift; sT.gate_wt += shift; sT.gate_b += shift;
      sT.dec_wt += shift; sT.dec_b += shift;
      for (int l = 0; l < sT.nlayers; ++l) {
        cirs_encoder_layer& Y = sT.layer[l];
        Y.in_wt += shift; Y.in_b += shift; Y.out_wt += shift; Y.out_b += shift; Y.l1_wt += shift; Y.l1_b += shift;
        Y.l2_wt += shift; Y.l2_b += shift; Y.n1_w += shift; Y.n1_b += shift; Y.n2_w += shift; Y.n2_b += shift;
      }
    }
    __syncthreads();
  }

  cirs_actor_tc::TcSmem TS{};
  cirs_actor_tc::TcState tst{0u, 0u};
  if (TC) {
    TS = cirs_actor_tc::tc_carve(reinterpret_cast<char*>(smem_dyn), reinterpret_cast<char*>(smem_dyn + A.tc_keep_off));
    cirs_actor_tc::tc_setup(A.H.W, blockIdx.x, A.n_slices, TS, tid);
  }

  // ---- reset + user token (position 0); every environment starts in the compact list of turn 0
  __shared__ int row_e[cirs_tracker::CTA_RB], row_a[cirs_tracker::CTA_RB], row_kn[cirs_tracker::CTA_RB];
  __shared__ float row_r[cirs_tracker::CTA_RB];
  constexpr int RB = cirs_tracker::CTA_RB;
  if (blockIdx.x == 0 && tid == 0) {
    *A.tc_timeout = 0;
    *A.n_active = B;
    A.count[0] = B;
    A.count[1] = 0;
  }
  for (int j0 = 0; (int)blockIdx.x + j0 * (int)gridDim.x < B; j0 += RB) {
    const int left = B - ((int)blockIdx.x + j0 * (int)gridDim.x);
    const int R = min(RB, (left + (int)gridDim.x - 1) / (int)gridDim.x);
    if (warp < R) {
      const int e = blockIdx.x + (j0 + warp) * gridDim.x;
      const int u = A.users[e];
      cirs_env::kuaishou_reset_warp(A.E, e, u, lane, A.active);
      if (lane == 0) {
        A.ep_len[e] = 0;
        A.list[e] = e;
        row_e[warp] = e; row_a[warp] = u; row_r[warp] = 0.f; row_kn[warp] = e;
      }
    }
    __syncthreads();
    if (FAST) {
      cirs_tfast::token_and_trunk(sT, A.fast, smem_dyn + A.smem_w_off, smem_dyn + A.trunk_w_off, B, R, 0, row_e, row_a,
                                  row_r, row_kn, A.kcache, A.vcache, smem_dyn, A.row_floats, A.cur_state, A.traj_len,
                                  A.traj_obs, A.traj_obs_next, nullptr, 0, A.value, [&](int r, int o, float v) {
                                    if (row_kn[r] >= 0) cirs_actor_tc::h2_image_store(A.h2_img, row_kn[r], o, v);
                                  }, false, [] {});
    } else {
      const int off = cirs_tracker::tracker_token_cta<SMW>(sT, B, R, 0, row_e, row_a, row_r, A.kcache, A.vcache, smem_dyn,
                                                           A.row_floats, A.cur_state, A.traj_len, A.traj_obs,
                                                           A.traj_obs_next, nullptr, nullptr, 0,
                                                           A.part_off >= 0 ? smem_dyn + A.part_off : nullptr);
      trunk_rows(A.H.W, R, smem_dyn, A.row_floats, off, row_e, A.h2, A.value, nullptr, TC ? A.h2_img : nullptr, row_kn);
    }
  }
  __threadfence();
  grid.sync();

  const int n_col_tiles = (A.H.W.n_action + BN - 1) / BN;
  for (int t = 0; t < A.max_steps; ++t) {
    // the environments still running, as a compact row list built during the previous turn (order is arbitrary:
    // every row's result is independent of its tile, and the Philox counter is keyed by the environment id)
    const int n_act = *reinterpret_cast<volatile int*>(A.count + (t & 1));
    if (n_act <= 0) break;
    const bool timer = blockIdx.x == 0 && tid == 0;
    long long t0 = 0, t1 = 0;
    if (timer) t0 = gtime_ns();
    const int row_tiles = (n_act + BM - 1) / BM;
    // the sampler's Philox offset of this turn (block 0 bumps the counter at the end of the turn, behind a grid barrier)
    const uint64_t turn_off = A.H.offset + (A.H.rng_counter ? (uint64_t)*A.H.rng_counter : 0ull);
    if (tid == 0) {   // this turn's head plan (shared by the CTA)
      sH.n_rows = n_act;
      sH.gather = A.list + (size_t)(t & 1) * B;
      if (TC) {
        sH.n_split = A.n_slices;
        sH.icdf = 0;   // the tensor-core epilogue samples inside the slice (two-level sampler): complete partials
      } else {
        int n_split = gridDim.x / row_tiles;
        n_split = n_split < 1 ? 1 : (n_split > n_col_tiles ? n_col_tiles : n_split);
        sH.tiles_per_split = (n_col_tiles + n_split - 1) / n_split;
        sH.n_split = (n_col_tiles + sH.tiles_per_split - 1) / sH.tiles_per_split;
        sH.icdf = sH.mode == MODE_SAMPLE && sH.tiles_per_split * BN <= 256 && sH.n_split <= 256 && row_tiles >= 10;
      }
    }
    __syncthreads();
    const HeadArgs& H = sH;
    // ---- phase A: actor head partials over the compact rows
    if (TC) {
      if ((int)blockIdx.x < A.n_slices)
        cirs_actor_tc::tc_head_turn(H, A.h2_img, blockIdx.x, TS, tid, tst, A.tc_timeout,
                                    blockIdx.x == 0 ? A.dbg + 1 + 3 * 512 + 32 : nullptr);
    } else {
      const int n_items = row_tiles * H.n_split;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        actor_head_body(H, w % row_tiles, w / row_tiles, smem_dyn);
        __syncthreads();
      }
    }
    __threadfence();
    grid.sync();
    if (timer) { t1 = gtime_ns(); A.dbg[1 + 3 * 512 + 32 + 5] = t1; A.dbg[1 + 3 * 512 + 32 + 6] = t0; }
    // ---- phase B: per running environment -- merge partials -> action, environment step (one warp per row), then the
    // tracker token and the trunk + critic of the new state for all of this CTA's rows together (tracker_cta_dev.cuh).
    // Row k of the compact list belongs to CTA k % gridDim: at most ceil(n_act / 148) rows per CTA, RB per pass.
    int32_t* list_next = A.list + (size_t)((t + 1) & 1) * B;
    for (int j0 = 0; (int)blockIdx.x + j0 * (int)gridDim.x < n_act; j0 += RB) {
      const int left = n_act - ((int)blockIdx.x + j0 * (int)gridDim.x);
      const int R = min(RB, (left + (int)gridDim.x - 1) / (int)gridDim.x);
      // stage stamps of CTA 0's first pass of the LAST turn played so far: dbg[1 + 3 * 512 ..) = {start, after combine +
      // env, 18 token stamps, after trunk}
      // (turn 0's stamps are kept separately at + 64: the turn with the most rows per CTA)
      long long* tq = (timer || (blockIdx.x == 0)) && j0 == 0 ? A.dbg + 1 + 3 * 512 + (t == 0 ? 64 : 0) : nullptr;
      if (tq && tid == 0) tq[0] = gtime_ns();
      // this turn's position is t + 1: the cached positions 0 .. t of the rows and the trunk's weights are fetched
      // into shared memory in the background (cp.async) while the rows' warps merge the head partials and step the
      // environment; the token step then never waits on L2 for them
      const int p = t + 1;
      const int kv_ld = A.T.d + 4;
      const bool kv_fit = A.kv_off >= 0 && (A.T.d & 3) == 0 &&
                          (int64_t)R * A.T.nlayers * 2 * p * kv_ld <= (int64_t)A.kv_cap;
      if (FAST) {
        if (j0 == 0)
          for (int i = tid; i < cirs_tfast::TR_FLOATS / 4; i += NT)
            cp_async16(smem_dyn + A.trunk_w_off + 4 * i, A.trunk_img + 4 * i);
      } else if (A.trunk_w_off >= 0 && j0 == 0) prefetch_trunk(A.H.W, smem_dyn + A.trunk_w_off);
      if (warp < R && lane == 0) row_e[warp] = H.gather[blockIdx.x + (j0 + warp) * gridDim.x];
      __syncthreads();
      if (kv_fit) prefetch_kv(A.T, B, R, row_e, p, A.kcache, A.vcache, smem_dyn + A.kv_off, kv_ld);
      cp_async_commit();
      int kn = -1, my_e = -1;   // lane 0 of a row's warp: the row's slot in the next turn's list (published late)
      if (warp < R) {
        const int k = blockIdx.x + (j0 + warp) * gridDim.x;
        const int e = row_e[warp];
        // the action-independent loads of the environment step go out before the partial merge
        const cirs_env::StepPre pre = cirs_env::kuaishou_step_pre(A.E, e, lane);
        int a;
        if (TC || !H.icdf) a = actor_combine_warp(H, k, lane, A.act, A.logp);
        else a = actor_combine_icdf_warp<8>(H, k, lane, A.act, A.logp, A.h2 + (size_t)e * HID,
                                            H.tiles_per_split * BN, turn_off);
        float4 emb = make_float4(0.f, 0.f, 0.f, 0.f);   // FAST: the action's embedding row, in flight behind the env step
        if (FAST && lane < 8) emb = __ldg(reinterpret_cast<const float4*>(A.T.emb_item + (size_t)a * 32) + lane);
        float r = 0.f;
        const bool d = cirs_env::kuaishou_step_warp(A.E, e, e, a, lane, A.active, A.rew, A.done, A.traj_len,
                                                    A.traj_act, A.traj_rew, A.traj_done, A.ep_len, A.force_length,
                                                    A.n_active, &pre, &r);
        if (FAST && lane < 8)
          *reinterpret_cast<float4*>(smem_dyn + (size_t)warp * A.row_floats + cirs_tfast::EMB_OFF + 4 * lane) = emb;
        __syncwarp();
        if (lane == 0) {
          if (!d) kn = atomicAdd(A.count + ((t + 1) & 1), 1);
          my_e = e;
          row_a[warp] = a; row_r[warp] = r;
          if (!FAST) {
            if (kn >= 0) list_next[kn] = e;
            row_kn[warp] = kn;
          }
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (tq && tid == 0) tq[1] = gtime_ns();
      if (FAST) {
        cirs_tfast::token_and_trunk(sT, A.fast, smem_dyn + A.smem_w_off, smem_dyn + A.trunk_w_off, B, R, p, row_e, row_a,
                                    row_r, row_kn, A.kcache, A.vcache, smem_dyn, A.row_floats, A.cur_state, A.traj_len,
                                    A.traj_obs, A.traj_obs_next, kv_fit ? smem_dyn + A.kv_off : nullptr, kv_ld, A.value,
                                    [&](int r, int o, float v) {
                                      if (row_kn[r] >= 0) cirs_actor_tc::h2_image_store(A.h2_img, row_kn[r], o, v);
                                    },
                                    true,
                                    [&] {   // the list slot's atomic has had the whole token step to return
                                      if (my_e >= 0) {
                                        row_kn[warp] = kn;
                                        if (kn >= 0) list_next[kn] = my_e;
                                      }
                                    },
                                    tq ? tq + 2 : nullptr);
      } else {
        const int off = cirs_tracker::tracker_token_cta<SMW>(sT, B, R, p, row_e, row_a, row_r, A.kcache, A.vcache,
                                                             smem_dyn, A.row_floats, A.cur_state, A.traj_len, A.traj_obs,
                                                             A.traj_obs_next, tq ? tq + 2 : nullptr,
                                                             kv_fit ? smem_dyn + A.kv_off : nullptr, kv_ld,
                                                             A.part_off >= 0 ? smem_dyn + A.part_off : nullptr);
        trunk_rows(A.H.W, R, smem_dyn, A.row_floats, off, row_e, A.h2, A.value,
                   A.trunk_w_off >= 0 ? smem_dyn + A.trunk_w_off : nullptr, TC ? A.h2_img : nullptr,
                   row_kn);   // consumed by the next turn's head phase
      }
      if (tq && tid == 0) tq[22] = gtime_ns();
    }
    if (blockIdx.x == 0 && tid == 0) {
      A.count[t & 1] = 0;   // consumed; it becomes the append counter of turn t + 1's phase B
      if (A.H.rng_counter) *A.H.rng_counter += 1ull;
    }
    __threadfence();
    grid.sync();
    if (timer) {
      A.dbg[0] = t + 1;
      A.dbg[1 + 3 * t] = n_act;
      A.dbg[2 + 3 * t] = t1 - t0;
      A.dbg[3 + 3 * t] = gtime_ns() - t1;
    }
  }
  if (TC) cirs_actor_tc::tc_teardown(TS, tid);
}

}  // namespace

// cooperative-launch capacity (resident CTAs) per (device, kernel variant, dynamic shared memory); thread-safe
#include <map>
#include <mutex>
#include <tuple>
static std::mutex g_occ_mu;
static std::map<std::tuple<int, int, size_t>, int> g_occ;
static int occupancy_cache_get(int dev, int variant, size_t smem) {
  std::lock_guard<std::mutex> lk(g_occ_mu);
  auto it = g_occ.find(std::make_tuple(dev, variant, smem));
  return it == g_occ.end() ? 0 : it->second;
}
static void occupancy_cache_put(int dev, int variant, size_t smem, int v) {
  std::lock_guard<std::mutex> lk(g_occ_mu);
  g_occ[std::make_tuple(dev, variant, smem)] = v;
}

// workspace: [counters 256 B][phase timers][list 2*B i32][h2 B*64 f32][head partials][h2 tile images]
constexpr int64_t DBG_BYTES = 8 * (1 + 6 * 512);   // phase timers of up to 512 turns
static int64_t partial_capacity(int32_t n_env, int32_t n_action) {
  const int64_t ffma = ((int64_t)n_env + 64) * 96 + 64 * 1024;
  const int64_t tc = (int64_t)((n_action + cirs_actor_tc::SLICE - 1) / cirs_actor_tc::SLICE + 1) * ((int64_t)n_env + 128);
  return ffma > tc ? ffma : tc;
}
static int64_t h2_image_bytes(int32_t n_env) {
  return (int64_t)((n_env + cirs_actor_tc::ROWS - 1) / cirs_actor_tc::ROWS) * 2 * cirs_actor_tc::A_BYTES;
}
extern "C" int64_t cirs_rollout_workspace_bytes(int32_t n_env, int32_t n_action) {
  return 256 + DBG_BYTES + (int64_t)sizeof(int32_t) * 2 * n_env + 64 + (int64_t)sizeof(float) * HID * n_env + 64 +
         (int64_t)sizeof(Partial) * partial_capacity(n_env, n_action) + 128 + h2_image_bytes(n_env) + 128 +
         (int64_t)sizeof(float) * cirs_tfast::TR_FLOATS;
}

extern "C" int cirs_rollout_kuaishou(const cirs_kuaishou_env* env, const cirs_tracker_weights* tw,
                                     const cirs_policy_weights* pw, const int32_t* users, uint8_t* active,
                                     int32_t* act, float* logp, float* value, float* cur_state, float* rew,
                                     uint8_t* done, int32_t traj_len, float* traj_obs, float* traj_obs_next,
                                     int32_t* traj_act, float* traj_rew, uint8_t* traj_done, int32_t* ep_len,
                                     float* kcache, float* vcache, int32_t kv_n_env, uint64_t seed,
                                     uint64_t* rng_counter, int32_t mode, int32_t max_steps, int32_t force_length,
                                     void* workspace, void* stream) {
  if (!env || !tw || !pw || !users || !active || !act || !logp || !value || !cur_state || !rew || !done || !ep_len ||
      !kcache || !vcache || !workspace || max_steps < 0 || (mode & 3) > 1) {
    cirs_set_error("cirs_rollout_kuaishou: null argument or bad mode");
    return CIRS_ERR_ARG;
  }
  if (!tw->emb_user || !tw->emb_item || tw->d % tw->nhead != 0 || tw->nlayers > CIRS_MAX_LAYERS ||
      tw->d_item_in != tw->d || tw->d_user_in != tw->d || max_steps + 1 > tw->max_len || max_steps > 512) {
    cirs_set_error("cirs_rollout_kuaishou: unsupported tracker shape (needs embedding tables, max_steps < max_len)");
    return CIRS_ERR_ARG;
  }
  if (kv_n_env != env->n_env) {
    cirs_set_error("cirs_rollout_kuaishou: the K/V caches are sized for a different number of environments "
                   "(kv_n_env != env->n_env): rebuild them (build_state(dim_batch, reset=True)) before the collect");
    return CIRS_ERR_ARG;
  }
  if (force_length > env->max_turn || max_steps > env->max_turn ||
      (traj_len > 0 && (max_steps > traj_len || force_length > traj_len))) {
    cirs_set_error("cirs_rollout_kuaishou: max_steps / force_length exceed env->max_turn or the trajectory length");
    return CIRS_ERR_ARG;
  }
  if (env->n_env <= 0) return CIRS_OK;
  // scratch of the cooperative token step: CTA_RB rows x (token buffers + 128 floats for the trunk's h1 | h2)
  const int row_floats = cirs_tracker::cta_row_floats(*tw) + 128;
  const size_t scratch_bytes = (size_t)row_floats * cirs_tracker::CTA_RB * sizeof(float);
  constexpr size_t SMEM_MAX = 224 * 1024;
  auto up128 = [](size_t x) { return (x + 127) & ~(size_t)127; };
  RolloutArgs A{};
  // tensor-core head: needs one catalogue slice per CTA (checked against the grid below) and plain sampling modes
  const int n_slices = (pw->n_action + cirs_actor_tc::SLICE - 1) / cirs_actor_tc::SLICE;
  const bool mask_seen = (mode & 4) != 0;   // remove_recommended_ids: mask the items in env->seen
  mode &= 3;
  if (mask_seen && !env->seen) {
    cirs_set_error("cirs_rollout_kuaishou: mode bit 2 (remove recommended ids) needs env->seen");
    return CIRS_ERR_ARG;
  }
  bool tc = cirs_head_tc::head_tc_enabled(env->n_env, pw->n_action, pw->ld_action) && (mode == 0 || mode == 1) &&
            pw->dim_state <= 32;
  // the tracker's weights staged in shared memory when they fit next to the head's buffers (d = 32: 113 KB)
  const float* w_lo = tw->user_wt;
  const int64_t w_count = tw->flat ? (tw->flat + tw->n_flat) - w_lo : 0;
  bool smw_ok = tw->flat != nullptr && w_count > 0 && (w_count % 4) == 0 && ((uintptr_t)w_lo % 16) == 0 &&
                tw->dec_b >= w_lo && tw->dec_b < w_lo + w_count && tw->gate_wt >= w_lo;
  for (int l = 0; smw_ok && l < tw->nlayers; ++l)
    smw_ok = tw->layer[l].in_wt >= w_lo && tw->layer[l].n2_b < w_lo + w_count;
  size_t smem = 0;
  bool smw = false, fast = false;
  int max_ctas = 0;
  const bool fast_ok = cirs_tfast::supported(*tw, *pw) && !getenv("CIRS_ROLLOUT_NO_FAST");
  const cirs_tfast::Layout fast_layout = cirs_tfast::layout(tw->nlayers);
  for (int attempt = 0; attempt < 2; ++attempt) {
    size_t head = tc ? up128(cirs_actor_tc::TURN_BYTES > scratch_bytes ? cirs_actor_tc::TURN_BYTES : scratch_bytes)
                     : up128(SMEM_BYTES > scratch_bytes ? SMEM_BYTES : scratch_bytes);
    size_t keep_off = head;
    if (tc) head += up128(cirs_actor_tc::KEEP_BYTES);
    if (head > SMEM_MAX) {
      if (tc) { tc = false; continue; }
      cirs_set_error("cirs_rollout_kuaishou: shared memory budget exceeded");
      return CIRS_ERR_ARG;
    }
    smw = smw_ok && head + (size_t)w_count * sizeof(float) <= SMEM_MAX;
    fast = tc && fast_ok && head + (size_t)fast_layout.total * sizeof(float) <= SMEM_MAX;
    smem = head;
    A.tc_keep_off = (int)(keep_off / sizeof(float));
    {   // phase-B staging behind the rows' scratch, inside the per-turn region [0, keep_off)
      size_t off = up128(scratch_bytes);
      const size_t part_bytes = (size_t)cirs_tracker::CTA_RB * 128 * sizeof(float);
      const size_t tw_bytes = up128((size_t)(fast ? cirs_tfast::TR_FLOATS : trunk_w_floats(pw->dim_state)) * sizeof(float));
      A.part_off = A.trunk_w_off = A.kv_off = -1;
      A.kv_cap = 0;
      if (off + part_bytes <= keep_off) { A.part_off = (int)(off / sizeof(float)); off += part_bytes; }
      if (off + tw_bytes <= keep_off && (pw->dim_state * HID) % 4 == 0) { A.trunk_w_off = (int)(off / sizeof(float)); off += tw_bytes; }
      if (off + 4096 <= keep_off) { A.kv_off = (int)(off / sizeof(float)); A.kv_cap = (int)((keep_off - off) / sizeof(float)); }
      if (fast && A.trunk_w_off < 0) fast = false;
    }
    if (fast) {
      smw = true;   // (the variant index and the occupancy cache treat it as a staged-weights kernel)
      A.w_lo = nullptr; A.w_count = 0; A.smem_w_off = (int)(smem / sizeof(float));
      A.fast = fast_layout;
      smem += (size_t)fast_layout.total * sizeof(float);
    } else if (smw) {
      A.w_lo = w_lo; A.w_count = (int)w_count; A.smem_w_off = (int)(smem / sizeof(float));
      smem += (size_t)w_count * sizeof(float);
    } else {
      A.w_lo = nullptr; A.w_count = 0; A.smem_w_off = 0;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    int mc = occupancy_cache_get(dev, (smw ? 1 : 0) + (tc ? 2 : 0) + (fast ? 4 : 0), smem);
    if (!mc) {
      int per_sm = 0, n_sm = 0;
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      const void* fn = fast ? (const void*)rollout_kuaishou_kernel<true, true, true>
                       : tc ? (smw ? (const void*)rollout_kuaishou_kernel<true, true> : (const void*)rollout_kuaishou_kernel<false, true>)
                            : (smw ? (const void*)rollout_kuaishou_kernel<true, false> : (const void*)rollout_kuaishou_kernel<false, false>);
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
      if (fast) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel<true, true, true>, NT, smem);
      else if (tc && smw) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel<true, true>, NT, smem);
      else if (tc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel<false, true>, NT, smem);
      else if (smw) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel<true, false>, NT, smem);
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_kuaishou_kernel<false, false>, NT, smem);
      if (per_sm < 1) {
        cirs_set_error("cirs_rollout_kuaishou: kernel does not fit on an SM");
        return CIRS_ERR_CUDA;
      }
      if (tc && per_sm > 1) per_sm = 1;   // TMEM: 256 columns per CTA, keep one CTA per SM
      mc = per_sm * n_sm;
      occupancy_cache_put(dev, (smw ? 1 : 0) + (tc ? 2 : 0) + (fast ? 4 : 0), smem, mc);
    }
    max_ctas = mc;
    if (tc && n_slices > max_ctas) { tc = false; continue; }   // catalogue too wide for one slice per CTA
    break;
  }
  A.n_slices = n_slices;
  A.E = *env; A.T = *tw;
  A.H.W = *pw; A.H.n_rows = env->n_env; A.H.gather = nullptr; A.H.state_by_k = 0; A.H.out_by_k = 0;
  A.H.active = active; A.H.state = cur_state; A.H.state_stride = tw->dim_state; A.H.noise_q = nullptr;
  A.H.seed = seed; A.H.offset = 1ull << 40; A.H.rng_counter = reinterpret_cast<unsigned long long*>(rng_counter);
  A.H.mode = mode; A.H.seen = mask_seen ? env->seen : nullptr; A.H.act_in = nullptr; A.H.value = value;
  int grid = max_ctas;
  if (!plan_head(A.H, grid)) {
    cirs_set_error("cirs_rollout_kuaishou: unsupported head shape");
    return CIRS_ERR_ARG;
  }
  // per turn: row_tiles * n_split <= grid work items of 64 rows each, n_split <= 84 column tiles
  char* wsp = reinterpret_cast<char*>(workspace);
  A.n_active = reinterpret_cast<int*>(wsp);
  A.count = reinterpret_cast<int*>(wsp + 64);
  A.dbg = reinterpret_cast<long long*>(wsp + 256);
  A.list = reinterpret_cast<int32_t*>(wsp + 256 + DBG_BYTES);
  A.h2 = reinterpret_cast<float*>(((uintptr_t)(A.list + 2 * (size_t)env->n_env) + 63) & ~(uintptr_t)63);
  A.H.h2_in = A.h2;
  A.H.part = reinterpret_cast<Partial*>(((uintptr_t)(A.h2 + (size_t)env->n_env * HID) + 63) & ~(uintptr_t)63);
  A.h2_img = reinterpret_cast<char*>(
      ((uintptr_t)(A.H.part + partial_capacity(env->n_env, pw->n_action)) + 127) & ~(uintptr_t)127);
  A.trunk_img = reinterpret_cast<float*>(((uintptr_t)(A.h2_img + h2_image_bytes(env->n_env)) + 127) & ~(uintptr_t)127);
  A.tc_timeout = reinterpret_cast<int*>(wsp + 128);
  if ((int64_t)(grid + 96) * BM > partial_capacity(env->n_env, pw->n_action)) {
    cirs_set_error("cirs_rollout_kuaishou: workspace too small for this grid");
    return CIRS_ERR_ARG;
  }
  A.n_env = env->n_env; A.max_steps = max_steps; A.force_length = force_length; A.traj_len = traj_len;
  A.users = users; A.active = active; A.act = act; A.logp = logp; A.value = value; A.cur_state = cur_state;
  A.rew = rew; A.done = done; A.traj_obs = traj_obs; A.traj_obs_next = traj_obs_next; A.traj_act = traj_act;
  A.traj_rew = traj_rew; A.traj_done = traj_done; A.ep_len = ep_len; A.kcache = kcache; A.vcache = vcache;
  A.row_floats = row_floats;
  void* params[] = {&A};
  const bool prof = cirs_profile_begin("rollout_kuaishou_kernel", (cudaStream_t)stream);
  void* fn = fast ? (void*)rollout_kuaishou_kernel<true, true, true>
             : tc ? (smw ? (void*)rollout_kuaishou_kernel<true, true> : (void*)rollout_kuaishou_kernel<false, true>)
                  : (smw ? (void*)rollout_kuaishou_kernel<true, false> : (void*)rollout_kuaishou_kernel<false, false>);
  cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NT), params, smem, (cudaStream_t)stream);
  cirs_note_launch();
  if (prof) cirs_profile_end((cudaStream_t)stream);
  if (err != cudaSuccess) {
    cirs_set_error(cudaGetErrorString(err));
    return CIRS_ERR_CUDA;
  }
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
