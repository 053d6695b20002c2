"""Synthetic KuaiRec-shaped / VirtualTaobao-shaped inputs.

Input generator shared by bench.py, the tests and the golden-vector script.  Pure numpy; every tensor is drawn
from ``numpy.random.Generator(PCG64(seed))`` so the GPU run, the CPU oracle and
the reference (when generating golden vectors) see identical tables.

The reference's real data files are not shipped (environments/KuaishouRec/data
holds only .gitkeep), so these tables stand in for:
  mat          <- KuaishouEnv.load_mat  (kuaishouEnv.py:61-80, watch_ratio clipped at 5)
  normed_mat   <- KuaishouEnv.compute_normed_reward (kuaishouEnv.py:113-145), values in [0,1]
  list_feat    <- item_categories.json feature_index (kuaishouEnv.py:84-96): 1..4 categories of 31 per item
  alpha_u/beta_i <- user-model ab_embedding_dict (CIRS-RL-kuaishou.py:157-163)
"""
import numpy as np

N_CAT = 31


def kuaishou_tables(n_user, n_item, seed=2023, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    mat = np.clip(rng.lognormal(-0.3, 0.6, size=(n_user, n_item)), 0, 5).astype(dtype)
    normed = rng.beta(8.0, 0.4, size=(n_user, n_item)).astype(dtype)
    # categories: Zipf-ish popularity over 31 categories, 1..4 distinct per item, values 1..31
    pop = 1.0 / np.arange(1, N_CAT + 1) ** 1.2
    pop /= pop.sum()
    n_cat = rng.integers(1, 5, size=n_item)
    cats = np.zeros((n_item, 4), dtype=np.int32)
    # vectorised draw: Gumbel top-k without replacement
    g = np.log(pop)[None, :] + rng.gumbel(size=(n_item, N_CAT))
    order = np.argsort(-g, axis=1)[:, :4] + 1
    for k in range(4):
        cats[:, k] = np.where(k < n_cat, order[:, k], 0)
    alpha = rng.normal(1.0, 0.05, size=n_user).astype(dtype)
    beta = rng.normal(1.0, 0.05, size=n_item).astype(dtype)
    return dict(mat=mat, normed_mat=normed, cats=cats, alpha_u=alpha, beta_i=beta)


def cats_to_list_feat(cats):
    """int32[I,4] zero padded -> list of python lists (the reference's list_feat)."""
    return [[int(c) for c in row if c > 0] for row in cats]


def cats_to_mask(cats):
    """int32[I,4] zero padded -> uint32[I] bitmask, bit c set for category c (1..31)."""
    m = np.zeros(cats.shape[0], dtype=np.uint32)
    for k in range(cats.shape[1]):
        c = cats[:, k].astype(np.uint32)
        m |= np.where(c > 0, np.uint32(1) << c, np.uint32(0)).astype(np.uint32)
    return m


def jaccard_distance_matrix(cats):
    """1 / Jaccard(categories), inf where disjoint (util.py:225-268). float64[I,I]."""
    m = cats_to_mask(cats)
    inter = np.bitwise_count(m[:, None] & m[None, :]).astype(np.float64)
    union = np.bitwise_count(m[:, None] | m[None, :]).astype(np.float64)
    with np.errstate(divide="ignore"):
        return 1.0 / (inter / union)


TAOBAO_GROUPS = (8, 8, 11, 11, 11, 11, 2, 2, 3, 18, 3)  # model/UserModel.py:22-32


def taobao_users(n, seed=2023):
    """one-hot x 11 user vectors f32[n,88], uniform category per group."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((n, 88), dtype=np.float32)
    off = 0
    for g in TAOBAO_GROUPS:
        k = rng.integers(0, g, size=n)
        out[np.arange(n), off + k] = 1.0
        off += g
    return out


def mmoe_state_dict(seed=2023, n_in=118, hidden=(64, 64), n_expert=4, expert_dim=8):
    """A synthetic UserModel_MMOE state_dict with the reference's parameter names and shapes for the VirtualTaobao
    columns (core/user_model_mmoe.py:80-98; CIRS-UserModel-taobao.py dnn_hidden_units (64, 64), 4 experts of width 8):
    the reward model SimulatedEnv evaluates inside every step.  Weights N(0, 0.15) so predictions land inside the
    clamp range [0, 10] with a useful spread; bias 4.0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = lambda *s: rng.normal(0.0, 0.15, size=s).astype(np.float32)  # noqa: E731
    h1, h2 = hidden
    return {
        "linear_model_task.0.weight": f(n_in, 1), "dnn.linears.0.weight": f(h1, n_in), "dnn.linears.0.bias": f(h1),
        "dnn.linears.1.weight": f(h2, h1), "dnn.linears.1.bias": f(h2),
        "mmoe_layer.expert_network.weight": f(n_expert * expert_dim, h2),
        "mmoe_layer.expert_network.bias": f(n_expert * expert_dim),
        "mmoe_layer.gating_networks.0.weight": f(n_expert, h2), "tower_network.0.weight": f(1, expert_dim),
        "out.0.bias": np.full((1, 1), 4.0, dtype=np.float32),
    }
