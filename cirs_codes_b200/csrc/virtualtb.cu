// The raw VirtualTaobao environment (environments/VirtualTaobao/virtualTB): its user generator and click model, one
// warp per environment -- the test environments of CIRS-RL-taobao.py:181-183 and the user distribution the training
// environments are reset from (SimulatedEnv.reset -> VirtualTB.reset -> UserModel.generate).
//
//   generate   model/UserModel.py:13-60   z ~ U(0,1)^128 -> Linear(128,128) -> LeakyReLU(0.01) -> Linear(128,88) ->
//              softmax over each of the 11 feature groups -> one multinomial draw per group -> one-hot x 11
//   click      model/ActionModel.py:6-23  [user 88, page 1, action 27] -> Linear(116,128) -> LeakyReLU -> Linear(128,256)
//              -> LeakyReLU -> Linear(256,21); a ~ multinomial(softmax(x[:11])), b ~ multinomial(softmax(x[11:]))
//   step       envs/virtualTB.py:74-100   done = Euclidean exit test over the last min(t, N-1) actions or t >= T-1;
//              reward = a (clicks on the page); cum_reward, total_turn; observation [action 27, a, b, total_turn]
// torch.multinomial(p, 1) is the exponential race argmax_j p_j / q_j, q ~ Exp(1): the draws q (and the
// generator's seeds z) can be supplied by the caller for parity runs, otherwise they come from Philox4x32-10.
#include "taobao_dev.cuh"

namespace {
using namespace cirs_taobao;
constexpr int WARPS = 4;
constexpr int NZ = 128, NH = 128, NA1 = 128, NA2 = 256, NCLK = 21, NA_IN = NU + 1 + NI;   // 116
__constant__ int GROUP_OFF[12] = {0, 8, 16, 27, 38, 49, 60, 62, 64, 67, 85, 88};          // UserModel.py:22-32

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.01f * v; }
__device__ __forceinline__ float exp1_draw(uint64_t seed, uint64_t off, int id, int c) {
  const uint4 r = philox4x32(make_uint4((uint32_t)id, (uint32_t)(c >> 2), (uint32_t)off, (uint32_t)(off >> 32)),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  return fmaxf(-logf(u01(rr[c & 3])), 1e-30f);
}

// y[o] = b[o] + sum_i Wt[i][ldo] x[i]  (k-major weights from global memory, x / y in shared memory)
__device__ __forceinline__ void dense(const float* __restrict__ Wt, const float* __restrict__ b, const float* x, int n_in,
                                      int n_out, int ldo, float* y, int lane, bool act) {
  for (int o = lane; o < n_out; o += 32) {
    float a = __ldg(b + o);
#pragma unroll 8
    for (int i = 0; i < n_in; ++i) a = fmaf(__ldg(Wt + (size_t)i * ldo + o), x[i], a);
    y[o] = act ? leaky(a) : a;
  }
  __syncwarp();
}

// argmax_j softmax(x)_j / q_j over x[lo, hi): the winner's index relative to lo.  All lanes return it.
__device__ __forceinline__ int race_group(const float* x, int lo, int hi, const float* q, uint64_t seed, uint64_t off,
                                          int id, int lane) {
  float mx = -INFINITY;
  for (int j = lo + lane; j < hi; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float z = 0.f;
  for (int j = lo + lane; j < hi; j += 32) z += expf(x[j] - mx);
  z = warp_sum(z);
  float best = -1.f;
  int bi = 0x7fffffff;
  for (int j = lo + lane; j < hi; j += 32) {
    const float p = expf(x[j] - mx) / z;
    const float qq = q ? q[j] : exp1_draw(seed, off, id, j);
    const float s = p / qq;
    if (s > best) { best = s; bi = j; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(FULL_MASK, best, o);
    const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  return bi - lo;
}

__global__ void __launch_bounds__(WARPS * 32)
virtualtb_generate_kernel(cirs_virtualtb_weights W, int n, const float* __restrict__ z, const float* __restrict__ q,
                          uint64_t seed, uint64_t offset, float* __restrict__ users) {
  __shared__ float sm[WARPS][NZ + NH + 96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS + warp;
  if (k >= n) return;
  float* sz = sm[warp];
  float* sh = sz + NZ;
  float* sx = sh + NH;
  for (int i = lane; i < NZ; i += 32) {
    float v;
    if (z) v = z[(size_t)k * NZ + i];
    else {   // torch.rand: uniform in [0, 1)
      const uint4 r = philox4x32(make_uint4((uint32_t)k, (uint32_t)(1000 + (i >> 2)), (uint32_t)offset,
                                            (uint32_t)(offset >> 32)), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
      v = (rr[i & 3] >> 8) * (1.0f / 16777216.0f);
    }
    sz[i] = v;
  }
  __syncwarp();
  dense(W.g1t, W.g1b, sz, NZ, NH, 128, sh, lane, true);
  dense(W.g2t, W.g2b, sh, NH, NU, 96, sx, lane, false);
  for (int i = lane; i < NU; i += 32) users[(size_t)k * NU + i] = 0.f;
  __syncwarp();
  for (int g = 0; g < 11; ++g) {
    const int win = race_group(sx, GROUP_OFF[g], GROUP_OFF[g + 1], q ? q + (size_t)k * NU : nullptr, seed, offset, k, lane);
    if (lane == 0) users[(size_t)k * NU + GROUP_OFF[g] + win] = 1.0f;
  }
}

// VirtualTB.step for one environment by one warp (E.um is unused: the raw environment has no reward model)
__global__ void __launch_bounds__(WARPS * 32)
virtualtb_step_kernel(cirs_taobao_env E, cirs_virtualtb_weights W, int n_rows, const int32_t* __restrict__ env_id,
                      const float* __restrict__ act, const float* __restrict__ q, uint64_t seed, uint64_t offset,
                      float* __restrict__ rew, uint8_t* __restrict__ done, int32_t* __restrict__ click,
                      int force_length) {
  __shared__ float sm[WARPS][128 + NA1 + NA2 + 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * WARPS + warp;
  if (k >= n_rows) return;
  const int e = env_id ? env_id[k] : k;
  const int T = E.max_turn, t = E.turn[e];
  float* sx = sm[warp];
  float* h1 = sx + 128;
  float* h2 = h1 + NA1;
  float* lo = h2 + NA2;
  float* hist = E.hist + (size_t)e * T * NI;
  const float a = lane < NI ? act[(size_t)k * NI + lane] : 0.f;
  bool leave = false;   // virtualTB.py:126-133 (float32 norm like numpy)
  for (int l = t - 1; l > max(-1, t - E.num_leave_compute); --l) {
    const float df = lane < NI ? __fsub_rn(a, hist[(size_t)l * NI + lane]) : 0.f;
    const float dist = sqrtf(warp_sum(__fmul_rn(df, df)));
    if ((double)dist <= E.leave_threshold) leave = true;
  }
  bool d = leave || (t >= T - 1);
  if (force_length > 0) d = (t + 1 >= force_length);
  if (t < T && lane < NI) hist[(size_t)t * NI + lane] = a;
  // click model input [user 88, page = total_turn, action 27]
  for (int i = lane; i < NU; i += 32) sx[i] = E.user[(size_t)e * NU + i];
  if (lane == 0) sx[NU] = (float)t;
  if (lane < NI) sx[NU + 1 + lane] = a;
  __syncwarp();
  dense(W.a1t, W.a1b, sx, NA_IN, NA1, 128, h1, lane, true);
  dense(W.a2t, W.a2b, h1, NA1, NA2, 256, h2, lane, true);
  dense(W.a3t, W.a3b, h2, NA2, NCLK, 32, lo, lane, false);
  const float* qq = q ? q + (size_t)k * NCLK : nullptr;
  const int ca = race_group(lo, 0, 11, qq, seed, offset, e, lane);
  const int cb = race_group(lo, 11, NCLK, qq, seed, offset, e, lane);
  if (lane == 0) {
    E.cum_rew[e] += (double)ca;
    E.prev_rew[e] = (double)ca;
    E.turn[e] = t + 1;
    rew[k] = (float)ca;
    done[k] = d ? 1 : 0;
    if (click) { click[2 * k] = ca; click[2 * k + 1] = cb; }
  }
}

bool bad(const cirs_virtualtb_weights* w, bool need_gen, bool need_act) {
  if (!w) return true;
  if (need_gen && (!w->g1t || !w->g1b || !w->g2t || !w->g2b)) return true;
  if (need_act && (!w->a1t || !w->a1b || !w->a2t || !w->a2b || !w->a3t || !w->a3b)) return true;
  return false;
}

}  // namespace

extern "C" int cirs_virtualtb_generate_users(const cirs_virtualtb_weights* w, int32_t n, const float* z, const float* q,
                                             uint64_t seed, uint64_t offset, float* users, void* stream) {
  if (bad(w, true, false) || n < 0 || !users) {
    cirs_set_error("cirs_virtualtb_generate_users: null argument / generator weights missing");
    return CIRS_ERR_ARG;
  }
  if (n == 0) return CIRS_OK;
  CIRS_LAUNCH(virtualtb_generate_kernel, (n + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream, *w, n, z, q, seed,
              offset, users);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}

extern "C" int cirs_virtualtb_step(const cirs_taobao_env* env, const cirs_virtualtb_weights* w, int32_t n_rows,
                                   const int32_t* env_id, const float* act, const float* q, uint64_t seed,
                                   uint64_t offset, float* rew, uint8_t* done, int32_t* click, int32_t force_length,
                                   void* stream) {
  if (!env || bad(w, false, true) || n_rows < 0 || !act || !rew || !done || !env->hist || !env->user || !env->turn ||
      !env->cum_rew || !env->prev_rew) {
    cirs_set_error("cirs_virtualtb_step: null argument / click-model weights missing");
    return CIRS_ERR_ARG;
  }
  if (force_length > env->max_turn) {
    cirs_set_error("cirs_virtualtb_step: force_length exceeds env->max_turn");
    return CIRS_ERR_ARG;
  }
  if (n_rows == 0) return CIRS_OK;
  CIRS_LAUNCH(virtualtb_step_kernel, (n_rows + WARPS - 1) / WARPS, WARPS * 32, 0, (cudaStream_t)stream, *env, *w, n_rows,
              env_id, act, q, seed, offset, rew, done, click, force_length);
  CIRS_CHECK_LAUNCH();
  return CIRS_OK;
}
