// Actor head over the item catalogue on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split precision.
// Interface of head_tc.cu, used by ppo.cu (PPO minibatch) and actor.cu (policy evaluation).
//
// The [rows, n_action] logits are NEVER written to HBM: every pass recomputes its 128 x 64 logit tiles with
// tensor-core MMAs (the contraction depth is only 64) and consumes them from TMEM in the epilogue.
//   pass F  (head_tc_stats)  per row online-softmax partials (max, sum exp) per catalogue split + the logit of the
//                            taken action                      -> Categorical.log_prob, ratio, losses
//   pass B2 (head_tc_dh2)    d logits rebuilt in the epilogue, split hi/lo into a K-major operand tile, second MMA
//                            accumulates d h2 = d logits . W3 in TMEM over the CTA's catalogue tiles; entropy partials
//   pass B3 (head_tc_dw3)    transposed problem (lane = catalogue column): d logits^T tile -> MMA accumulates
//                            d W3 = d logits^T . h2 over the CTA's row tiles; d b3 column sums in registers
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cirs_b200.h"

namespace cirs_head_tc {

constexpr int MAX_SPLIT = 32;   // catalogue splits of passes F / B2 (partials per row)

struct HeadTc {
  const float* h2;      // [n, 64] row-major trunk output
  int n;                // rows
  const float* w3t;     // [64, ldA] k-major Actor.last weight (zero padded columns)
  int64_t ldA;
  const float* b3;      // [nA]
  int nA;
  // W3 pre-split into TF32 (hi, lo) operand-tile images (head_tc_pack): the passes copy them into shared memory
  // instead of splitting / transposing W3 again for every tile of every CTA
  const float* img;     // head_tc_image_floats(ldA) floats
  const float* himg;    // the same for h2 (head_tc_pack_h2): head_tc_h2_image_floats(n) floats
};

// floats of the image buffer for a catalogue padded to ldA columns
int64_t head_tc_image_floats(int64_t ldA);
// (re)build the images from w3t; call after every change of the weights (once per minibatch / evaluation)
int head_tc_pack(const float* w3t, int64_t ldA, const float* b3, int nA, float* img, cudaStream_t st);
// images of the trunk output h2 [n, 64] (rows beyond n are zero); rebuilt once per minibatch after the trunk forward
int64_t head_tc_h2_image_floats(int64_t n);
int head_tc_pack_h2(const float* h2, int n, float* himg, cudaStream_t st);

// Front end of a pass in one launch: policy trunk of n gathered observation rows (h1 optional, h2, critic value), the
// h2 images (himg, optional) written by the same CTAs, and -- concurrently, by further CTAs -- the W3 images (img,
// optional; pass it whenever the weights changed since they were last packed).
// n_dev (optional, DEVICE int32): the true row count when it is not known on the host yet -- n is then a host-side
// CAPACITY (grid and image layout are sized by n, rows >= min(n, *n_dev) are skipped).
int head_tc_front(const cirs_policy_weights* w, int n, const int32_t* idx, const float* obs, float* h1, float* h2,
                  float* value, float* himg, float* img, cudaStream_t st, const int32_t* n_dev = nullptr);

// number of catalogue splits used for n rows (<= MAX_SPLIT); partial arrays are [n, n_split]
int plan_split(int n, int nA);
// ... and of pass F alone (its partials pm / ps are merged separately from pass B2's)
int plan_split_f(int n, int nA);

// pass F.  act_of_row: action of row r = act[idx ? idx[r] : r] (may be NULL: no logit is picked).
// Outputs: pm, ps [n, n_split] partial (max, sum exp(l - max)); la[n] logit of the taken action.
int head_tc_stats(const HeadTc& H, const int32_t* idx, const int32_t* act, int n_split, float* pm, float* ps,
                  float* la, cudaStream_t st, const int32_t* n_dev = nullptr);   // n_dev: as for head_tc_front (wide kernel only)

// pass B2.  rowm / rinvz: softmax max and 1 / sum per row; coef: d loss / d logp per row; acta: taken action per row.
// Outputs: dh2_part [n_split, n, 64] (sum over splits = d loss / d h2 through the actor head);
//          ent_part [n, n_split] partial entropies  -sum_c p log clamp(p).
int head_tc_dh2(const HeadTc& H, const float* rowm, const float* rinvz, const float* coef, const int32_t* acta,
                int n_split, float* dh2_part, float* ent_part, cudaStream_t st);

// pass B3.  Accumulates (atomicAdd) into g_w3t [64, ldA] and g_b3 [nA]; both must be zeroed by the caller.
int head_tc_dw3(const HeadTc& H, const float* rowm, const float* rinvz, const float* coef, const int32_t* acta,
                float* g_w3t, float* g_b3, cudaStream_t st);

// cirs_policy_eval (values + log-probs of stored actions, process_fn) through trunk -> pass F -> merge (ppo.cu)
int64_t policy_eval_tc_workspace_bytes(int64_t n);
int64_t policy_eval_tc_image_bytes(int64_t ldA);

// true when the tensor-core path can be used for this shape (and CIRS_NO_TC is not set in the environment)
bool head_tc_enabled(int n, int nA, int64_t ldA);

}  // namespace cirs_head_tc
