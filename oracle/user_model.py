"""CPU oracle for the all-pairs user-model inference that produces KuaishouEnv's ``normed_mat``.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU arm, never by the product.

A numpy float32 restatement of the reference, row by row the way the reference evaluates it (one user at a time over
all items, nothing factorised):
  * ``deepfm_forward``   -- UserModel_Pairwise._deepfm (core/user_model_pairwise.py:98-132):
       linear logit      core/layers.py:47-73   (sum of the 1-d embeddings + dense . weight)
       FM cross term     DeepCTR-Torch/deepctr_torch/layers/interaction.py:26-34
       DNN               DeepCTR-Torch/deepctr_torch/layers/core.py:120-134 (Linear + ReLU, no BN, dropout 0)
       last / out        user_model_pairwise.py:66-67, core.py:155-161 (Linear(H, 1, bias=False) + scalar bias)
  * ``compute_normed_reward`` -- KuaishouEnv.compute_normed_reward (environments/KuaishouRec/env/kuaishouEnv.py:113-145):
       per user: X = [user, photo_id, feat0..3, photo_duration] as float32 rows, predict_mat (float64) row = forward(X),
       then (predict_mat - min) / (max - min).

Pinned: tests/golden/user_model_deepfm.npz holds the reference's own outputs (oracle/make_golden_user_model.py);
tests/test_oracle_golden.py::test_user_model_oracle_vs_golden compares.
``params`` is the reference's state_dict as numpy arrays (keys as in torch, e.g. "dnn.linears.0.weight").
"""
import numpy as np

F32 = np.float32


def deepfm_forward(params, user_id, item_ids, item_feat, item_dense):
    """Predictions of one user on n items -> float32 [n].  item_feat int [n, F], item_dense float [n, D]."""
    n = len(item_ids)
    e_user = params["embedding_dict.user_id.weight"].astype(F32)
    e_item = params["embedding_dict.photo_id.weight"].astype(F32)
    e_feat = params["embedding_dict.feat.weight"].astype(F32)
    # input_from_feature_columns (core/user_model.py:419-447): one [n, d] embedding per sparse column, in column order
    sparse = [np.broadcast_to(e_user[user_id], (n, e_user.shape[1])), e_item[item_ids]] + \
             [e_feat[item_feat[:, f]] for f in range(item_feat.shape[1])]
    dense = item_dense.astype(F32).reshape(n, -1)
    # linear logit (core/layers.py:47-73)
    l_user = params["linear.embedding_dict.user_id.weight"].astype(F32)[:, 0]
    l_item = params["linear.embedding_dict.photo_id.weight"].astype(F32)[:, 0]
    l_feat = params["linear.embedding_dict.feat.weight"].astype(F32)[:, 0]
    lin = np.full(n, l_user[user_id], F32) + l_item[item_ids]
    for f in range(item_feat.shape[1]):
        lin = lin + l_feat[item_feat[:, f]]
    lin = lin + dense @ params["linear.weight"].astype(F32)[:, 0]
    # FM (interaction.py:26-34) over the [n, fields, d] stack
    stack = np.stack(sparse, axis=1).astype(F32)
    sq_of_sum = stack.sum(axis=1) ** 2
    sum_of_sq = (stack * stack).sum(axis=1)
    fm = F32(0.5) * (sq_of_sum - sum_of_sq).sum(axis=1)
    # DNN on cat(sparse..., dense) (combined_dnn_input), then last + out bias
    x = np.concatenate([s.astype(F32) for s in sparse] + [dense], axis=1)
    k = 0
    while "dnn.linears.%d.weight" % k in params:
        w = params["dnn.linears.%d.weight" % k].astype(F32)
        b = params["dnn.linears.%d.bias" % k].astype(F32)
        x = np.maximum(x @ w.T + b, F32(0))
        k += 1
    dnn = x @ params["last.weight"].astype(F32)[0] + params["out.bias"].astype(F32).reshape(())
    return (lin + fm + dnn).astype(F32)


def predict_mat(params, users, items, item_feat, item_dense):
    """float64 [n_user, n_item] of float32 predictions (kuaishouEnv.py:131-137)."""
    out = np.zeros((len(users), len(items)), np.float64)
    for r, u in enumerate(users):
        out[r] = deepfm_forward(params, int(u), items, item_feat, item_dense)
    return out


def compute_normed_reward(params, users, items, item_feat, item_dense):
    pm = predict_mat(params, users, items, item_feat, item_dense)
    mn, mx = pm.min(), pm.max()
    return (pm - mn) / (mx - mn)


def synth_params(v_user, v_item, v_feat, dim=16, hidden=64, n_feat=4, n_dense=1, seed=2023, scale=1.0):
    """Random "trained-like" weights in the reference's state_dict layout (bench / full-size tests)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nrm = lambda *s, std: (rng.standard_normal(s) * std * scale).astype(F32)  # noqa: E731
    p = {"embedding_dict.user_id.weight": nrm(v_user, dim, std=0.3),
         "embedding_dict.photo_id.weight": nrm(v_item, dim, std=0.3),
         "embedding_dict.feat.weight": nrm(v_feat, dim, std=0.3),
         "linear.embedding_dict.user_id.weight": nrm(v_user, 1, std=0.3),
         "linear.embedding_dict.photo_id.weight": nrm(v_item, 1, std=0.3),
         "linear.embedding_dict.feat.weight": nrm(v_feat, 1, std=0.3),
         "linear.weight": nrm(n_dense, 1, std=0.01),
         "dnn.linears.0.weight": nrm(hidden, dim * (2 + n_feat) + n_dense, std=0.1),
         "dnn.linears.0.bias": nrm(hidden, std=0.1),
         "dnn.linears.1.weight": nrm(hidden, hidden, std=0.2),
         "dnn.linears.1.bias": nrm(hidden, std=0.1),
         "last.weight": nrm(1, hidden, std=0.25),
         "out.bias": nrm(1, 1, std=0.1)}
    p["embedding_dict.feat.weight"][0] = 0   # padding_idx = 0
    return p
