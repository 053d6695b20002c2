/* cirs_b200.h -- C ABI of libcirs_b200.so: the CIRS rollout + PPO-update hot path as sm_100a CUDA kernels.
 *
 * Drop-in boundary.  The reference is pure Python and has no FFI of its own; each entry
 * point below replaces one Python-level operator of the reference's hot path (file:line given per function,
 * paths relative to the reference root).  The reference-side binding a maintainer would add is a ctypes stub
 * (INTEGRATION.md); this repo's own binding is cirs_codes_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _h;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no entry point synchronises;
 *     re-entrant per device, one process per GPU; every entry point can be captured into a CUDA graph;
 *   - return 0 on success, non-zero on error; cirs_last_error() returns a message for the calling thread;
 *   - "rows" are the environments taking part in a call.  Row k refers to environment slot
 *     e = env_id ? env_id[k] : k, and is skipped when `active` is given and active[e] == 0;
 *   - a Linear(in -> out) is stored k-major: Wt[in][ldo], ldo = out rounded up to a multiple of 32 (padding
 *     columns are zero and have zero gradient); biases are padded to ldo as well.
 *     cirs_codes_b200/params.py converts from/to the reference's torch state_dict layout ([out][in]);
 *   - the weight structs hold non-const pointers because the same struct type describes the parameters, their
 *     gradients and the Adam moments (three parallel flat buffers with identical layout).
 */
#ifndef CIRS_B200_H
#define CIRS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CIRS_ABI_VERSION 9
#define CIRS_MAX_LAYERS 4
#define CIRS_HIDDEN 64 /* tianshou Net hidden_sizes=[64,64], CIRS-RL-kuaishou.py:88 */

const char* cirs_last_error(void);
int cirs_abi_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py reports it as gpu_launches) */
int64_t cirs_launch_count(void);
/* Optional per-kernel timing: while enabled every launch is bracketed by CUDA events on its own stream.
 * cirs_profile_report synchronises the device and writes "kernel_name count total_ms" lines into buf. */
void cirs_profile_enable(int on);
int cirs_profile_report(char* buf, int n);
/* Actor-head contractions of cirs_ppo_minibatch / cirs_policy_eval / cirs_rollout_kuaishou: 1 = tcgen05 tensor cores
 * with 3xTF32 split precision, warp-specialised kernels fed by cp.async.bulk (default; csrc/head_tc.cu), 2 = the same
 * with register-staged operands (also CIRS_NO_TMA=1), 0 = FP32 FFMA tile GEMMs (csrc/gemm.cuh), -1 = default
 * (environment variable CIRS_NO_TC=1 selects FFMA).  All are CUDA paths with the same results within the 1e-5 bar. */
void cirs_head_tc_enable(int on);
/* 1 if a tensor-core kernel gave up waiting on an mbarrier since the last call (synchronises; never expected). */
int cirs_head_tc_timeout(void);
/* Stream-ordered copy of that flag into PINNED host memory, without synchronising or clearing: the host path folds it
 * into the update's own read-back and raises when it is set. */
int cirs_head_tc_timeout_peek(int32_t* out_pinned_h, void* stream);

/* ------------------------------------------------------------------ KuaishouEnv / SimulatedEnv ---------- */
typedef struct {
  int32_t n_env;    /* B: environment slots */
  int32_t max_turn; /* T */
  int32_t num_leave_compute; /* N  (kuaishouEnv.py:38) */
  int32_t n_user, n_item;
  int32_t simulated; /* 1: SimulatedEnv reward (normed_mat, exposure); 0: raw KuaishouEnv reward mat[u,a] */
  int32_t version;   /* 1: r/(1+e)   2: r-e      (simulated_env.py:102-107) */
  float leave_threshold, tau, gamma_exposure, r_decay;
  /* read-only tables */
  const float* normed_mat;  /* [n_user, n_item]  simulated_env.py:100 */
  const float* mat;         /* [n_user, n_item]  kuaishouEnv.py:171 (may be NULL when simulated) */
  const uint32_t* cat_mask; /* [n_item] bit c set <=> category c (1..31) in list_feat_small[item] */
  const float* alpha_u;     /* [n_user] indexed by encoded user id, or NULL (simulated_env.py:157-164) */
  const float* beta_i;      /* [n_item] indexed by encoded item id, or NULL */
  const float* dist;        /* optional [n_item, n_item] df_dist_small; NULL -> 1/Jaccard(cat_mask) */
  /* per-environment state */
  int32_t* user;    /* [B] */
  int32_t* turn;    /* [B] total_turn */
  int32_t* hist;    /* [B, T] history_action / sequence_action */
  double* cum_rew;  /* [B] */
  uint32_t* seen;   /* optional [B, ceil(n_item/32)] bitset of items already recommended this episode */
} cirs_kuaishou_env;

/* reset(): kuaishouEnv.py:182-190 + simulated_env.py:59-72.  users[n] are injected by the caller (the
 * reference draws random.randint, kuaishouEnv.py:155-159).  Clears turn / history / cum_rew / seen and,
 * when given, sets active[e] = 1. */
int cirs_kuaishou_reset(const cirs_kuaishou_env* env, int32_t n_rows, const int32_t* env_id,
                        const int32_t* users, uint8_t* active, void* stream);

/* step(): simulated_env.py:111-168 + kuaishouEnv.py:161-218 + util.py:21-54 for n_rows environments.
 * act[n] -> rew[n] (f32), done[n] (u8); obs_next is the action itself (kuaishouEnv.py:147-153).
 * When `active` is given it is updated in place: active[e] &= !done (collector.py:303-311 drops finished
 * environments from the ready set).  Optional trajectory outputs (env-major, VectorReplayBuffer layout,
 * vecbuf.py:26-30; L = traj_len slots per environment): traj_act / traj_rew / traj_done [B, L] written at
 * [e, t]; ep_len[e] written on done.  force_length > 0 overrides done (collector.py:253-258). */
int cirs_kuaishou_step(const cirs_kuaishou_env* env, int32_t n_rows, const int32_t* env_id, uint8_t* active,
                       const int32_t* act, float* rew, uint8_t* done, int32_t traj_len, int32_t* traj_act,
                       float* traj_rew, uint8_t* traj_done, int32_t* ep_len, int32_t force_length,
                       void* stream);

/* ------------------------------------------------------------------ VirtualTaobao / SimulatedEnv -------- */
#define CIRS_TB_USER 88 /* one-hot x 11 user features, virtualTB.py:16 */
#define CIRS_TB_ITEM 27 /* continuous item / action features, virtualTB.py:17 */

/* UserModel_MMOE.forward (core/user_model_mmoe.py:144-220) for dense feature columns and one regression task: the
 * reward model SimulatedEnv evaluates inside every VirtualTaobao step (simulated_env.py:79-86).  Read only.
 *   y = x . lin_w  +  tower . (E(h) g(h))  + out_bias,  h = relu(W2 relu(W1 x + b1) + b2),
 *   E(h) = (We h + be) viewed [expert_dim][n_expert], g(h) = softmax(Wg h)            (core/layers.py:67-116) */
typedef struct {
  int32_t n_in;                 /* 118 = 88 user + [prev reward, 0, turn] + 27 item */
  int32_t h1, h2;               /* dnn_hidden_units, each <= 128 */
  int32_t n_expert, expert_dim; /* 4, 8: n_expert * expert_dim <= 64, n_expert <= 32 */
  const float* lin_w;           /* [n_in]  linear_model_task[0].weight */
  const float *w1t, *b1;        /* dnn.linears.0  Wt[n_in][ld(h1)] */
  const float *w2t, *b2;        /* dnn.linears.1  Wt[h1][ld(h2)] */
  const float *wet, *be;        /* mmoe_layer.expert_network  Wt[h2][ld(n_expert*expert_dim)], output o = dim*n_expert + expert */
  const float *wgt, *bg;        /* mmoe_layer.gating_networks[0]  Wt[h2][ld(n_expert)]; bg = zeros (the layer has no bias) */
  const float* tower;           /* [expert_dim]  tower_network[0].weight */
  float out_bias;               /* out[0].bias (PredictionLayer, deepctr layers/core.py:155-161) */
} cirs_mmoe_weights;

typedef struct {
  int32_t n_env;    /* B */
  int32_t max_turn; /* T */
  int32_t num_leave_compute; /* N (virtualTB.py:126-133: the last min(t, N-1) actions are compared) */
  int32_t version;  /* 1: r/(1+e)  2: r-e */
  int32_t map_action; /* 1: `act` is the policy's raw sample and the kernel applies policy.map_action first
                       * (clip to [-1,1], then low + (high-low)(a+1)/2 in float32; tianshou/policy/base.py:143-173) */
  float act_low, act_high; /* action_space.low / high (virtualTB.py:24: -1, 1) */
  double leave_threshold, tau, gamma_exposure; /* python floats in the reference */
  cirs_mmoe_weights um;
  /* per-environment state */
  float* user;      /* [B, 88] cur_user */
  int32_t* turn;    /* [B] */
  float* hist;      /* [B, T, 27] history_action (float32 values; the reference keeps them in a float64 array) */
  double* prev_rew; /* [B] self.reward, a feature of the next step (simulated_env.py:79) */
  double* cum_rew;  /* [B] */
} cirs_taobao_env;

/* reset(): virtualTB.py:102-113 + simulated_env.py:59-72 with injected users[n, 88] (the reference samples them
 * from its generator network, model/UserModel.py:40-60). */
int cirs_taobao_reset(const cirs_taobao_env* env, int32_t n_rows, const int32_t* env_id, const float* users,
                      uint8_t* active, void* stream);

/* step(): simulated_env.py:111-168 over virtualTB.py:74-100,126-133 for n_rows environments.
 *   act[n, 27]      action (raw policy sample when env->map_action, else already mapped)
 *   act_env[n, 27]  optional out: the action the environment used == obs_next[:, :27] (simulated_env.py:50)
 *   rew[n], done[n] as in cirs_kuaishou_step; obs_next = [act_env, rew, 0, turn + 1] is assembled by the caller
 * The real environment's click model and new-user draws (virtualTB.py:84-98) only consume torch RNG and are
 * discarded by SimulatedEnv (simulated_env.py:114,138); they are not computed.
 * Optional trajectory outputs: traj_act[B, L, 27] receives `act` as given (the buffer keeps the RAW action,
 * collector.py:246-250), traj_act_env[B, L, 27] the action the environment used (the tracker's token input when it
 * is trained), traj_rew / traj_done [B, L], ep_len[B]; force_length as in cirs_kuaishou_step. */
int cirs_taobao_step(const cirs_taobao_env* env, int32_t n_rows, const int32_t* env_id, uint8_t* active,
                     const float* act, float* act_env, float* rew, uint8_t* done, int32_t traj_len,
                     float* traj_act, float* traj_act_env, float* traj_rew, uint8_t* traj_done, int32_t* ep_len,
                     int32_t force_length, void* stream);

/* The raw VirtualTaobao environment: user generator and click model with the weights the reference ships
 * (virtualTB/data/generator_model.pt, action_model.pt), k-major like every Linear here.  Either half may be NULL when
 * only the other entry point is used. */
typedef struct {
  const float *g1t, *g1b; /* generator_model.0  Wt[128][128]   (model/UserModel.py:9-13) */
  const float *g2t, *g2b; /* generator_model.2  Wt[128][96]    (88 outputs) */
  const float *a1t, *a1b; /* ActionModel.model.0  Wt[116][128] (model/ActionModel.py:8-14) */
  const float *a2t, *a2b; /* ActionModel.model.2  Wt[128][256] */
  const float *a3t, *a3b; /* ActionModel.model.4  Wt[256][32]  (21 outputs: 11 click counts, 10 page actions) */
} cirs_virtualtb_weights;

/* UserModel.generate (model/UserModel.py:40-60): n users, one-hot x 11 feature groups -> users[n, 88].
 *   z [n, 128] the generator's seeds (torch.rand) or NULL -> Philox uniform [0, 1)
 *   q [n, 88]  Exp(1) race draws of the 11 multinomials (draw j belongs to feature j) or NULL -> Philox */
int cirs_virtualtb_generate_users(const cirs_virtualtb_weights* w, int32_t n, const float* z, const float* q,
                                  uint64_t seed, uint64_t offset, float* users, void* stream);

/* VirtualTB.step (envs/virtualTB.py:74-100) for n_rows environments of a cirs_taobao_env (its reward model `um` is
 * not used): act[n, 27] (already mapped by policy.map_action), reward = the click count a = ActionModel.predict(user,
 * total_turn, action)[0]; click[n, 2] (optional) = (a, b), the `lst_action` part of the next observation
 * [action 27, a, b, total_turn].  q [n, 21]: Exp(1) race draws of the two multinomials (11 + 10) or NULL -> Philox.
 * The new user drawn when an episode ends (virtualTB.py:96-98) is not generated: the Collector never steps a
 * finished environment again before its reset (core/collector.py:294-311). */
int cirs_virtualtb_step(const cirs_taobao_env* env, const cirs_virtualtb_weights* w, int32_t n_rows,
                        const int32_t* env_id, const float* act, const float* q, uint64_t seed, uint64_t offset,
                        float* rew, uint8_t* done, int32_t* click, int32_t force_length, void* stream);

/* ------------------------------------------------------------------ StateTracker ------------------------ */
typedef struct {
  float *in_wt, *in_b;     /* self_attn.in_proj  Wt[d][ld3d], b[3d] */
  float *out_wt, *out_b;   /* self_attn.out_proj Wt[d][ldd] */
  float *l1_wt, *l1_b;     /* linear1 Wt[d][ldh] */
  float *l2_wt, *l2_b;     /* linear2 Wt[d_hid][ldd] */
  float *n1_w, *n1_b, *n2_w, *n2_b; /* LayerNorm, eps 1e-5 */
} cirs_encoder_layer;

typedef struct {
  int32_t d, nhead, d_hid, nlayers, dim_state, max_len; /* max_len = MAX_TURN + 1 (state_tracker.py:144) */
  int32_t d_user_in;   /* input width of ffn_user: d (embedding) or 88 (VirtualTaobao dense) */
  int32_t d_item_in;   /* input width of the item part of fnn_gate: d, or 27 */
  int32_t n_user, n_item; /* embedding table rows (0 when dense) */
  float* emb_user; /* [n_user, d] or NULL when the user observation is dense (core/inputs.py:24-44) */
  float* emb_item; /* [n_item, d] or NULL */
  float *user_wt, *user_b; /* ffn_user  Wt[d_user_in][ldd] (state_tracker.py:146) */
  float *gate_wt, *gate_b; /* fnn_gate  Wt[1 + d_item_in][ldd], row 0 multiplies the reward (:150) */
  float* pe;               /* [max_len, d]  PositionalEncoding table (:255-279); not a parameter */
  cirs_encoder_layer layer[CIRS_MAX_LAYERS];
  float *dec_wt, *dec_b;   /* decoder Wt[d][ld_state] (:158) */
  float* flat;             /* base of the flat parameter buffer (everything above except pe) */
  int64_t n_flat;
} cirs_tracker_weights;

/* build_state(): state_tracker.py:188-250, dropout = 0, with a per-environment K/V cache instead of the
 * reference's whole-prefix recompute (exact by causality).
 *   pos[e] : sequence position to write (0 = user token, t >= 1 = action token of turn t-1); read per env slot
 *   expect_pos : >= 0 -> rows whose pos[e] differs are skipped (the fused rollout passes the current turn so that
 *            environments that finished earlier are skipped but those that finished THIS turn still get their
 *            last obs_next, as in collector.py:261-269); -1 -> no filter
 *   idx[n] : user id (pos 0) or item id (pos >= 1) when the corresponding embedding table is non-NULL
 *   dense[n, d_*_in] : dense user / item features otherwise
 *   rew[n] : reward of the transition (ignored at pos 0)
 * kcache / vcache: [nlayers, B, max_len, d].
 * Outputs (each optional): state_out[k * state_stride ..+dim_state) per row;  cur_state[e * dim_state ..] per
 * environment slot;  traj_obs[(e*traj_len + p) * dim_state ..] when p < traj_len and
 * traj_obs_next[(e*traj_len + p-1) * dim_state ..] when p >= 1 -- the replay buffer's obs / obs_next slots
 * (tianshou/data/buffer/base.py:238-275, env-major vecbuf.py:26-30). */
int cirs_tracker_step(const cirs_tracker_weights* w, int32_t n_env, int32_t n_rows, const int32_t* env_id,
                      const uint8_t* active, const int32_t* pos, int32_t expect_pos, const int32_t* idx,
                      const float* dense, const float* rew, float* kcache, float* vcache, float* state_out,
                      int64_t state_stride, float* cur_state, int32_t traj_len, float* traj_obs,
                      float* traj_obs_next, void* stream);

/* Training pass of the tracker (replaces autograd through the observations stored in the replay buffer,
 * core/policy/ppo.py:215 loss.backward(retain_graph=True) -> tianshou/data/batch.py:256-258):
 * one full-sequence causal forward over every environment's token sequence followed by the backward pass,
 * given d_obs = d loss / d obs for every stored observation (accumulated by cirs_ppo_minibatch over the last
 * repeat).  Sequences: environment e has n_tok[e] = ep_len[e] observation positions 0..ep_len[e]-1
 * (position 0 = user token, position p>=1 = action token of turn p-1, reward traj_rew[e, p-1]).
 *   users[B], traj_act[B, L], traj_rew[B, L], ep_len[B]   the rollout record (L = traj_len)
 *   d_obs[B*L, dim_state]                                  upstream gradient per buffer slot (e*L + p)
 *   grads                                                  same layout as w; ACCUMULATED into (zero it first)
 *   workspace / workspace_bytes                            cirs_tracker_train_workspace_bytes(...) bytes
 * dense_user[B, d_user_in] / dense_item[B, L, d_item_in] replace users / traj_act when the tables are NULL.
 * Compact mode (tok_slot != NULL): only the n_tok valid tokens are processed.  tok_slot[i] = buffer slot of compact
 * row i (env-major sorted, i.e. VectorReplayBuffer.sample_index(0)), env_off[e] = first compact row of environment e
 * (exclusive prefix sum of ep_len, B + 1 entries).  Without it every [B*L] slot is a row and padding is masked.
 * obs_check (optional, [B*L, dim_state]) receives the forward pass's decoded states at their buffer slots. */
int64_t cirs_tracker_train_workspace_bytes(const cirs_tracker_weights* w, int32_t n_env, int64_t n_rows);
int cirs_tracker_train(const cirs_tracker_weights* w, const cirs_tracker_weights* grads, int32_t n_env,
                       int32_t traj_len, const int32_t* users, const int32_t* traj_act, const float* traj_rew,
                       const int32_t* ep_len, const float* dense_user, const float* dense_item, int32_t n_tok,
                       const int32_t* tok_slot, const int32_t* env_off, int32_t max_ep_len, const float* d_obs,
                       float* obs_check, void* workspace, int64_t workspace_bytes, int32_t phase, void* stream);
/* phase: 0 or 3 = the whole pass; 1 = forward only (needs no d_obs: it can be issued on a side stream as soon as the
 * rollout has ended, beside the PPO minibatches); 2 = backward only, after a phase-1 call with the same arguments and
 * workspace and unchanged tracker weights.  The layer-by-layer path treats phase 1 as a no-op and runs everything in
 * phase 2.
 * Compact mode runs as TWO launches by default (csrc/tracker_fused.cuh): a chunk kernel that carries whole
 * environments through the forward and backward pass inside one CTA, and one grouped split-K launch for every
 * Linear's weight gradient.  max_ep_len (0 = unknown -> traj_len) is the longest stored episode: it sizes the chunks.
 * cirs_tracker_train_fused_enable(0) selects the layer-by-layer launches instead (-1 = default; CIRS_K6_UNFUSED=1);
 * both are CUDA paths with the same results within the parity bar (tests/test_gpu_tracker_train.py runs both). */
void cirs_tracker_train_fused_enable(int on);

/* ------------------------------------------------------------------ policy / value heads ---------------- */
typedef struct {
  int32_t dim_state, n_action;
  int32_t ld_action;  /* n_action rounded up to 128 */
  float *w1t, *b1; /* Net layer 0: Wt[dim_state][64]  (utils/net/common.py:87-92) */
  float *w2t, *b2; /* Net layer 1: Wt[64][64] */
  float *w3t, *b3; /* Actor.last: Wt[64][ld_action]  (utils/net/discrete.py:56-67) */
  float *wv, *bv;  /* Critic.last: [64], [1]          (discrete.py:109-114) */
  float* flat;     /* base of the flat buffer that holds all of the above, trunk first */
  int64_t n_flat;  /* floats in the flat buffer */
  int64_t n_trunk; /* leading floats that belong to the shared trunk (w1t, b1, w2t, b2) */
  /* continuous actor (tianshou ActorProb, utils/net/continuous.py:120-199; VirtualTaobao): w3t / b3 are the `mu`
   * layer Wt[64][ld_action] with n_action <= 32, sigma = ActorProb.sigma_param[n_action] (state independent,
   * std = exp(sigma)), mean = max_action * tanh(.).  sigma == NULL -> discrete actor over the catalogue. */
  float* sigma;
  float max_action;
} cirs_policy_weights;

/* policy.forward(): core/policy/ppo.py:111-163 for the discrete actor: softmax over the whole catalogue and
 * Categorical.sample(), i.e. the exponential race argmax_j p_j / q_j, q ~ Exp(1), fused with the
 * logits GEMM so the [n, n_action] probabilities are never written.
 *   state   : row k reads state + (env_id ? k : e) * state_stride   (compact rows with env_id, per-slot without)
 *   noise_q : [n_rows, n_action] Exp(1) draws supplied by the caller (parity tests) or NULL -> Philox(seed, offset)
 *   rng_counter : optional device uint64; the Philox offset becomes offset + *rng_counter and the call increments it,
 *             so that a captured CUDA graph draws fresh noise at every replay
 *   mode    : 0 sample, 1 argmax (deterministic_eval)
 *   seen    : optional [B, ceil(n_action/32)] bitset; set bits are removed from the distribution
 *             (remove_recommended_ids, core/policy/utils.py:30-58)
 * Outputs per row k: act (i32), logp = Categorical.log_prob(act) (f32), value = critic(s) (f32).
 * workspace: cirs_actor_workspace_bytes(n_rows, n_action) bytes of device scratch. */
int64_t cirs_actor_workspace_bytes(int32_t n_rows, int32_t n_action);
int cirs_actor_sample(const cirs_policy_weights* w, int32_t n_rows, const int32_t* env_id, const uint8_t* active,
                      const float* state, int64_t state_stride, const float* noise_q, uint64_t seed,
                      uint64_t offset, uint64_t* rng_counter, int32_t mode, const uint32_t* seen, int32_t* act,
                      float* logp, float* value, void* workspace, void* stream);

/* Critic / log-prob evaluation without sampling: A2CPolicy._compute_returns' critic(obs) calls
 * (tianshou/policy/modelfree/a2c.py:89-90) and PPOPolicy.process_fn's old log-prob (core/policy/ppo.py:104-108).
 * obs[n_rows, dim_state] (row r reads obs + (row_idx ? row_idx[r] : r) * dim_state);  act may be NULL (value
 * only).  value / logp are indexed like obs (by row_idx[r] when given). */
int cirs_policy_eval(const cirs_policy_weights* w, int32_t n_rows, const int32_t* row_idx, const float* obs,
                     const int32_t* act, float* value, float* logp, void* workspace, void* stream);
/* The same with the row count still on the DEVICE (n_dev, i32): n_cap sizes grids / layouts / the workspace and rows
 * >= min(n_cap, *n_dev) are skipped.  The host can queue process_fn's evaluations behind the rollout before it has read
 * the collect's transition count back (core/collector.py:147-367 returns before policy.update() starts in the reference;
 * here the two overlap).  Tensor-core head only; CIRS_ERR_ARG otherwise. */
int cirs_policy_eval_dev(const cirs_policy_weights* w, int32_t n_cap, const int32_t* n_dev, const int32_t* row_idx,
                         const float* obs, const int32_t* act, float* value, float* logp, void* workspace, void* stream);

/* policy.forward() for the continuous actor (core/policy/ppo.py:144-156 with dist_fn = Independent(Normal),
 * CIRS-RL-taobao.py:228-232): mu = max_action * tanh(W3 h + b3), std = exp(sigma); act = eps * std + mu
 * (torch.normal: multiply then add), eps ~ N(0,1) from noise_eps[n_rows, n_action] (parity runs) or Philox +
 * Box-Muller; mode 1 -> act = mu (deterministic_eval).  Outputs per row k: act[k, n_action] (RAW, unclipped: this is
 * what the buffer stores), logp[k] = sum_c Normal.log_prob, value[k]; mu_out[k, n_action] optional. */
int cirs_actorprob_sample(const cirs_policy_weights* w, int32_t n_rows, const int32_t* env_id,
                          const uint8_t* active, const float* state, int64_t state_stride, const float* noise_eps,
                          uint64_t seed, uint64_t offset, uint64_t* rng_counter, int32_t mode, float* act,
                          float* logp, float* value, float* mu_out, void* stream);
/* cirs_policy_eval for the continuous actor: act[., n_action] float, indexed like obs. */
int cirs_actorprob_eval(const cirs_policy_weights* w, int32_t n_rows, const int32_t* row_idx, const float* obs,
                        const float* act, float* value, float* logp, void* stream);

/* ------------------------------------------------------------------ fused rollout ----------------------- */
/* A whole Collector.collect(n_episode = B) (core/collector.py:147-367 with the fork's semantics: reset everything,
 * no reset on done, finished environments dropped) in ONE persistent cooperative kernel: reset + user token, then
 * per turn  actor head -> sample -> environment step -> tracker token -> replay-buffer slots, until every episode
 * has ended or max_steps turns were played.  All arrays are per environment slot ([B] or [B, .]); traj_* are the
 * replay buffer's env-major arrays (traj_len slots per environment); ep_len[e] = episode length.
 * rng_counter: device uint64, advanced once per turn (Philox offset of the sampler).  mode: 0 sample, 1 argmax;
 * + 4: remove_recommended_ids (core/policy/utils.py:7-58; the test collectors NX_0 / NX_x of core/collector_set.py:19) --
 * items in env->seen (maintained by the environment step) are masked out of the softmax and of the race.
 * Same device code as cirs_actor_sample / cirs_kuaishou_step / cirs_tracker_step (bit-identical results).
 * kv_n_env: environment capacity of kcache / vcache ([nlayers, kv_n_env, max_len, d]); must equal env->n_env (the
 * layer stride), anything else is rejected instead of writing past the caches.  force_length must not exceed
 * env->max_turn (the history has max_turn slots) and, like max_steps, traj_len when trajectories are recorded.
 * The workspace's bytes [256, 256 + 8 * (1 + 3 * 512)) hold int64 phase timers written by the kernel:
 * turns played, then per turn {running environments, ns in the actor-head phase, ns in the per-environment phase};
 * the int32 at byte 128 is set when a tensor-core mbarrier wait gave up (never expected; the host checks it after
 * the collect's read-back and raises). */
int64_t cirs_rollout_workspace_bytes(int32_t n_env, int32_t n_action);
int cirs_rollout_kuaishou(const cirs_kuaishou_env* env, const cirs_tracker_weights* tw, const cirs_policy_weights* pw,
                          const int32_t* users, uint8_t* active, int32_t* act, float* logp, float* value,
                          float* cur_state, float* rew, uint8_t* done, int32_t traj_len, float* traj_obs,
                          float* traj_obs_next, int32_t* traj_act, float* traj_rew, uint8_t* traj_done,
                          int32_t* ep_len, float* kcache, float* vcache, int32_t kv_n_env, uint64_t seed,
                          uint64_t* rng_counter, int32_t mode, int32_t max_steps, int32_t force_length,
                          void* workspace, void* stream);

/* A whole Collector.collect(n_episode = B) on SimulatedEnv(VirtualTB) in ONE kernel.  Nothing in a VirtualTaobao turn
 * couples environments (the action head is 27 wide, not a catalogue), so each warp plays its environment's entire
 * episode -- user token, then per turn actor -> sample -> map_action -> environment step (incl. the MMOE reward
 * model) -> tracker token -> replay-buffer slots -- without any grid-wide barrier.  users[B, 88]; traj_act[B, L, 27]
 * (raw actions), traj_act_env[B, L, 27] (mapped actions); other arrays as in cirs_rollout_kuaishou.  Philox offset of (environment e, turn t) is
 * *rng_counter + t; the kernel's last warp to finish adds max_steps to *rng_counter.  Same device code as
 * cirs_actorprob_sample / cirs_taobao_step / cirs_tracker_step (bit-identical results). */
int cirs_rollout_taobao(const cirs_taobao_env* env, const cirs_tracker_weights* tw, const cirs_policy_weights* pw,
                        const float* users, uint8_t* active, float* cur_state, int32_t traj_len, float* traj_obs,
                        float* traj_obs_next, float* traj_act, float* traj_act_env, float* traj_rew,
                        uint8_t* traj_done, int32_t* ep_len, float* kcache, float* vcache, int32_t kv_n_env,
                        uint64_t seed, uint64_t* rng_counter, int32_t mode, int32_t max_steps, int32_t force_length,
                        void* stream);

/* ------------------------------------------------------------------ returns (GAE) ----------------------- */
/* A2CPolicy._compute_returns (a2c.py:80-109) + BasePolicy.compute_episodic_return / _gae_return
 * (tianshou/policy/base.py:272-313, 380-396) + RunningMeanStd.update (utils/statistics.py:80-95), float64
 * arithmetic like the reference.  Buffer slots are env-major [B, L]; environment e holds n_slot[e] transitions.
 *   v_s, v_next : critic outputs (normalised space) per slot;  rew, done per slot
 *   ret_rms     : double[3] = {mean, var, count} on the device (reward_normalization = 1; read only here);
 *                 NULL -> no normalisation
 *   scratch     : double[2 * n_env] device scratch;  moments: double[3] = raw {sum, sumsq, count} of this call's
 *                 unnormalised returns (NULL -> not computed).  Ranks all-reduce `moments` before merging.
 *   out: returns[B*L] (normalised, f32), adv[B*L] (f32).  Slots t >= n_slot[e] are left untouched.
 * An episode still running at its last stored slot ends the scan there (unfinished_index, base.py:308-309).
 * cirs_rms_update merges the batch moments into ret_rms (RunningMeanStd.update, statistics.py:80-95). */
int cirs_compute_returns(int32_t n_env, int32_t traj_len, const int32_t* n_slot, const float* v_s,
                         const float* v_next, const float* rew, const uint8_t* done, double gamma,
                         double gae_lambda, const double* ret_rms, double* scratch, double* moments,
                         float* returns, float* adv, void* stream);
int cirs_rms_update(double* ret_rms, const double* moments, void* stream);

/* ------------------------------------------------------------------ PPO update -------------------------- */
typedef struct {
  float eps_clip, vf_coef, ent_coef, max_grad_norm; /* CIRS-RL-kuaishou.py:97-104 */
  int32_t value_clip, norm_adv;                     /* ppo.py:185-186, 200-205 */
  float lr, beta1, beta2, adam_eps;                 /* torch.optim.Adam defaults, lr 1e-3 */
} cirs_ppo_config;

/* Per-minibatch advantage statistics (ppo.py:185-186: mean and UNBIASED std over the minibatch), computed for
 * n_mb minibatches at once: minibatch j = slots idx[mb_off[j] .. mb_off[j+1]).  stats[j] = {count, sum, sumsq}
 * in float64 -- raw moments so that ranks can all-reduce them before use. */
int cirs_adv_stats(int32_t n_mb, const int32_t* mb_off, const int32_t* idx, const float* adv, double* stats,
                   void* stream);

/* One PPO minibatch, forward + loss + backward (core/policy/ppo.py:181-220):
 *   idx[n]  buffer slots of the minibatch;  obs[., dim_state], act, adv, returns, v_old, logp_old per slot
 *   adv_stat  double[3] {count, sum, sumsq} of this (global) minibatch;  n_global = number of rows the losses
 *             are averaged over (== n unless the minibatch is sharded over ranks)
 *   grads   gradient buffer (same layout as w), OVERWRITTEN with d loss / d params (local partial sums)
 *   d_obs   [., dim_state] d loss / d obs written at the minibatch's slots (the tracker's upstream gradient)
 *   losses  float[4] {loss, clip, vf, ent}: local partial sums already divided by n_global
 * act: int32 per slot (discrete actor) or float[., n_action] per slot (continuous actor, w->sigma != NULL: Gaussian
 *      log-prob / entropy, gradients into w3t / b3 / sigma; core/policy/ppo.py:181-220 with Independent(Normal)).
 * workspace: cirs_ppo_workspace_bytes(n_max, n_action). */
int64_t cirs_ppo_workspace_bytes(int32_t n_rows, int32_t n_action);
int cirs_ppo_minibatch(const cirs_policy_weights* w, const cirs_policy_weights* grads, const cirs_ppo_config* cfg,
                       int32_t n, int32_t n_global, const int32_t* idx, const float* obs, const void* act,
                       const float* adv, const float* returns, const float* v_old, const float* logp_old,
                       const double* adv_stat, float* d_obs, float* losses, void* workspace, void* stream);

/* clip_grad_norm_ + Adam over a flat parameter buffer (ppo.py:221-226; torch.optim.Adam single-tensor CPU
 * semantics).  The first n_dup elements belong to tensors that occur TWICE in the reference's parameter list
 * (the trunk shared by actor and critic, CIRS-RL-kuaishou.py:245-258): they count twice in the
 * norm, are scaled by coef^2, and receive two sequential Adam updates (step counter += 2).
 *   state   int32[2] on the device: {steps taken by ordinary tensors, steps taken by duplicated tensors}
 *   max_grad_norm <= 0 -> no clipping.   scratch: double[16] device scratch, ZERO-INITIALISED by the caller once and
 *   then left to this function (squared norm, per-step bias corrections, the fused kernel's alternating accumulators). */
int cirs_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t n_dup,
                   const cirs_ppo_config* cfg, int32_t* state, double* scratch, void* stream);

/* The whole learn() loop of one update (core/policy/ppo.py:173-233): n_repeat passes over n_mb minibatches;
 * slots[r * n + i] is the i-th buffer slot of repeat r's permutation (n = mb_off_h[n_mb]), minibatch j = entries
 * mb_off_h[j] .. mb_off_h[j+1].  mb_off_h is a HOST array, mb_off its device copy.
 * adv_stats: double[n_repeat * n_mb * 3] device scratch; losses: float[n_repeat * n_mb * 4] on the device.
 * d_obs (optional, d_obs_floats elements) is zeroed at the start of every repeat and receives d loss / d obs.
 * exp_avg / exp_avg_sq / opt_state / opt_scratch: Adam state as in cirs_clip_adam (n_dup = w->n_trunk).
 * Data parallel: comm != NULL (cirs_comm_create) -> this rank holds ITS chunk of every global minibatch;
 * the advantage moments are summed over ranks once, the flat gradient once per minibatch (stream-ordered
 * all-reduces between the minibatch kernels and clip + Adam); n_global_h[n_mb] (HOST) = rows of each global
 * minibatch, over which the losses are averaged.  comm == NULL: single process, n_global_h may be NULL.
 * n_stats_tail: that many further doubles stored right behind adv_stats are summed over ranks by the same collective
 * (this repo: the raw return moments of cirs_compute_returns, so that they cost no collective of their own). */
int cirs_ppo_learn(const cirs_policy_weights* w, const cirs_policy_weights* grads, float* exp_avg,
                   float* exp_avg_sq, const cirs_ppo_config* cfg, int32_t n_repeat, int32_t n_mb,
                   const int32_t* mb_off_h, const int32_t* mb_off, const int32_t* slots, const float* obs,
                   const void* act, const float* adv, const float* returns, const float* v_old,
                   const float* logp_old, double* adv_stats, float* d_obs, int64_t d_obs_floats, float* losses,
                   int32_t* opt_state, double* opt_scratch, void* workspace, void* comm, const int32_t* n_global_h,
                   int32_t n_stats_tail, void* stream);

/* Front end of the update on the device (no host round trip between the rollout and the update):
 * cirs_update_plan = VectorReplayBuffer.sample_index(0) (tianshou/data/buffer/manager.py:144-169) for buffers the
 * fused rollout filled: tok_slot[i] = i-th stored slot, env-major (e * traj_len + t, t < n_slot[e]); env_off[e] =
 * first compact row of environment e (n_env + 1 entries, env_off[n_env] = number of stored transitions).
 * tok_slot may be NULL.  cirs_gather_i32: dst[i] = src[idx[i]] -- the minibatch order indices[perm] of one repeat
 * (tianshou/data/batch.py:733-744). */
int cirs_update_plan(int32_t n_env, int32_t traj_len, const int32_t* n_slot, int32_t* tok_slot, int32_t* env_off,
                     void* stream);
int cirs_gather_i32(int32_t* dst, const int32_t* src, const int32_t* idx, int32_t n, void* stream);
/* Stream-ordered zero fill of a device buffer (optim_state.zero_grad(), core/policy/ppo.py:174, for the flat gradient buffers). */
int cirs_zero(void* ptr, int64_t bytes, void* stream);

/* Test-time coverage metrics of a collect on the device (evaluation.py:286-371 Callback_Coverage_Count):
 * over the n stored transitions act[idx[i]] (idx NULL -> act[i]): out3[0] = number of DISTINCT recommended items
 * (CV = out3[0] / n_item, CV_turn = out3[0] / n), out3[2] = sum_i item_weight[act_i] when item_weight is given (the
 * dominated-category rate ifeat_* = out3[2] / n with item_weight[j] = the value get_feat_dominate_dict's integer
 * arithmetic assigns to item j, evaluation.py:10-77; built once per callback on the host).  out3[1] is unused (0).
 * bits: uint32[ceil(n_item / 32)] scratch.  Stream-ordered. */
int cirs_coverage_count(int32_t n, const int32_t* idx, const int32_t* act, int32_t n_item, const int32_t* item_weight,
                        uint32_t* bits, int64_t* out3, void* stream);

/* ------------------------------------------------------------------ multi-GPU ---------------- */
/* One process per GPU, environments sharded over ranks, parameters replicated; the reference has no multi-GPU path.
 * A communicator wraps an NCCL communicator of the NCCL library already loaded in the process (resolved with dlopen:
 * no link-time dependency).  Rank 0 calls cirs_comm_unique_id (128 bytes, HOST) and sends the id to the other ranks
 * by any means (this repo: torch.distributed broadcast); every rank then calls cirs_comm_create with its CUDA device
 * current.  cirs_comm_allreduce: stream-ordered in-place SUM over ranks; dtype 0 = float32, 1 = float64, 2 = int32. */
int cirs_comm_unique_id(void* id128_h);
int cirs_comm_create(const void* id128_h, int32_t rank, int32_t world, void** comm_out);
int cirs_comm_destroy(void* comm);
int cirs_comm_allreduce(void* comm, void* buf, int64_t count, int32_t dtype, void* stream);
/* All-reduces issued between group_begin and group_end travel as one fused NCCL operation. */
int cirs_comm_group_begin(void* comm);
int cirs_comm_group_end(void* comm);

/* ------------------------------------------------------------------ user model -> normed_mat ------------ */
/* The DeepFM user model of stage 1 (UserModel_Pairwise, core/user_model_pairwise.py:36-132) with the feature columns
 * of CIRS-UserModel-kuaishou.py:115-123: sparse user_id, photo_id, n_feat slots sharing ONE "feat" embedding table
 * (padding row 0 = zeros), n_dense dense item features; dnn_hidden_units = (64, 64).  Tensors are in the reference's
 * torch layout (Linear weight = [out][in]); cirs_codes_b200/user_model.py fills the struct from a state_dict. */
typedef struct {
  const float* emb_user;  /* [vocab_user, emb_dim]   embedding_dict.user_id.weight */
  const float* emb_item;  /* [vocab_item, emb_dim]   embedding_dict.photo_id.weight */
  const float* emb_feat;  /* [vocab_feat, emb_dim]   embedding_dict.feat.weight */
  const float* lin_user;  /* [vocab_user]            linear.embedding_dict.user_id.weight (core/layers.py:20-73) */
  const float* lin_item;  /* [vocab_item] */
  const float* lin_feat;  /* [vocab_feat] */
  const float* lin_dense; /* [n_dense]               linear.weight */
  const float* w1;        /* [64][emb_dim * (2 + n_feat) + n_dense]   dnn.linears.0.weight; input order = user, item,
                             feat slots, dense (combined_dnn_input) */
  const float* b1;        /* [64] */
  const float* w2;        /* [64][64]                dnn.linears.1.weight */
  const float* b2;        /* [64] */
  const float* w_last;    /* [64]                    last.weight (no bias, user_model_pairwise.py:66) */
  float out_bias;         /* out.bias (PredictionLayer, task "regression": bias only) */
  int32_t emb_dim, n_feat, n_dense, hidden;
} cirs_user_model;

/* KuaishouEnv.compute_normed_reward (environments/KuaishouRec/env/kuaishouEnv.py:113-145): predict every user in
 * user_ids[n_user] (raw ids, lbe_user.classes_) on every item (item_ids[n_item] raw photo ids, item_feat[n_item,
 * n_feat], item_dense[n_item, n_dense]: the rows of df_photo_env) with UserModel_Pairwise.forward
 * (user_model_pairwise.py:98-132, 154-156), then, if normalise != 0, (pred - min) / (max - min) over the table.
 *   out     float[n_user * n_item] row-major, 16-byte aligned; minmax (optional) float[2] = {min, max} of the raw
 *           predictions.  Arithmetic is FP32 like the reference's torch path (normalisation in FP64 like numpy).
 *   workspace  cirs_user_model_workspace_bytes(n_user, n_item, emb_dim) bytes, 16-byte aligned.
 * The 64 x 64 hidden contraction runs on the tcgen05 tensor cores (3xTF32) for emb_dim in {8, 16, 32};
 * cirs_user_model_tc_enable(0) selects the FP32-FFMA kernel (-1 = default, CIRS_NO_TC=1 also selects FFMA). */
int64_t cirs_user_model_workspace_bytes(int32_t n_user, int32_t n_item, int32_t emb_dim);
int cirs_user_model_predict_all(const cirs_user_model* m, int32_t n_user, const int32_t* user_ids, int32_t n_item,
                                const int32_t* item_ids, const int32_t* item_feat, const float* item_dense,
                                int32_t normalise, float* out, float* minmax, void* workspace, void* stream);
void cirs_user_model_tc_enable(int on);
/* 1 if the tensor-core kernel gave up waiting on an mbarrier since the last call (synchronises; never expected). */
int cirs_user_model_timeout(void);
/* Profiling aid: with the environment variable CIRS_UM_FLAGS=4 thread 0 of CTA 0 of the tensor-core kernel accumulates
 * clock64() cycles per phase of its tile loop; out8_h (HOST int64[8]) = {stage, barrier, MMA issue, FM / prefetch,
 * MMA wait, epilogue, barrier, tiles} of the last launch.  Synchronises. */
int cirs_user_model_debug_phases(int64_t* out8_h);
/* Profiling aid for the tensor-core head passes (csrc/head_tc.cu): enable != 0 makes the issuer warp and the first / last
 * epilogue warp of every CTA of pass F and pass B2 (bulk-copy fed kernels) accumulate clock64() cycles per phase of
 * their tile loops into 64 device counters; out64_h (HOST int64[64], may be NULL) receives the counters accumulated so
 * far, reset != 0 clears them.  Layout: scratch/head_phases.py.  Synchronises. */
int cirs_head_tc_debug_phases(int32_t enable, int64_t* out64_h, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif
