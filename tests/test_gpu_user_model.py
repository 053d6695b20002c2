"""All-pairs user-model inference (csrc/user_model.cu) through the C ABI (cirs_user_model_predict_all)
against the reference's recorded outputs (tests/golden/user_model_deepfm.npz: UserModel_Pairwise.forward and
KuaishouEnv.compute_normed_reward run by oracle/make_golden_user_model.py) and against the CPU oracle
(oracle/user_model.py).  Bars: raw predictions 1e-5 relative (absolute floor 1e-5 x the table's range: they are sums of
terms of that size); normalised table 1e-5 relative with absolute floor 2e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import user_model as oum

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "user_model_deepfm.npz")


@pytest.fixture(scope="module")
def um():
    from cirs_codes_b200 import user_model
    return user_model


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    return g, {k[3:]: g[k] for k in g.files if k.startswith("sd.")}


def _mode(tc):
    from cirs_codes_b200 import _lib
    _lib.load().cirs_user_model_tc_enable(tc)


def _no_timeout():
    from cirs_codes_b200 import _lib
    assert _lib.load().cirs_user_model_timeout() == 0


@pytest.mark.parametrize("tc", [1, 0])
def test_predictions_vs_reference_golden(um, gold, tc):
    g, P = gold
    _mode(tc)
    try:
        w = um.UserModelWeights(P)
        raw, mm = um.predict_all(w, g["users"], g["items"], g["item_feat"], g["item_dense"], normalise=False,
                                 return_minmax=True)
        raw, mm = raw.cpu().numpy(), mm.cpu().numpy()
        rng = float(g["raw_pred"].max() - g["raw_pred"].min())
        np.testing.assert_allclose(raw[g["raw_rows"]], g["raw_pred"], rtol=1e-5, atol=1e-5 * rng)
        assert mm[0] == raw.min() and mm[1] == raw.max()
        nm = um.predict_all(w, g["users"], g["items"], g["item_feat"], g["item_dense"]).cpu().numpy()
        np.testing.assert_allclose(nm, g["normed_mat"], rtol=1e-5, atol=2e-6)
        assert nm.min() == 0.0 and nm.max() == 1.0
        _no_timeout()
    finally:
        _mode(-1)


def test_compute_normed_reward_drop_in(um, gold):
    """Same call as KuaishouEnv.compute_normed_reward(user_model, lbe_user, lbe_photo, df_photo_env)."""
    import pandas as pd
    g, P = gold

    class Lbe:
        def __init__(self, c):
            self.classes_ = c
    v_item = P["embedding_dict.photo_id.weight"].shape[0]
    feats = np.zeros((v_item, 4), np.int64)
    dur = np.zeros(v_item)
    feats[g["items"]] = g["item_feat"]
    dur[g["items"]] = g["item_dense"][:, 0]
    df = pd.DataFrame({"feat0": feats[:, 0], "feat1": feats[:, 1], "feat2": feats[:, 2], "feat3": feats[:, 3],
                       "photo_duration": dur}, index=np.arange(v_item))
    sd = {k: torch.from_numpy(np.array(v)) for k, v in P.items()}
    nm = um.compute_normed_reward(sd, Lbe(g["users"]), Lbe(g["items"]), df)
    assert nm.dtype == np.float64 and nm.shape == g["normed_mat"].shape
    np.testing.assert_allclose(nm, g["normed_mat"], rtol=1e-5, atol=2e-6)
    dev = um.compute_normed_reward(sd, Lbe(g["users"]), Lbe(g["items"]), df, return_device=True)
    assert dev.is_cuda and dev.dtype == torch.float32


def _synth_inputs(n_user, n_item, v_feat, rng):
    feat = np.zeros((n_item, 4), np.int32)
    for i in range(n_item):
        k = int(rng.integers(1, 5))
        feat[i, :k] = rng.integers(1, v_feat, k)
    dense = rng.uniform(3, 60, (n_item, 1)).astype(np.float32)
    return feat, dense


@pytest.mark.parametrize("n_user,n_item,dim", [(1, 1, 16), (3, 129, 16), (70, 128, 8), (33, 257, 32), (5, 200, 12)])
def test_ragged_shapes_vs_oracle(um, n_user, n_item, dim):
    """single pair, partial item tiles, every tensor-core embedding width, and a width only the FFMA kernel takes"""
    rng = np.random.Generator(np.random.PCG64(n_user * 1000 + n_item))
    P = oum.synth_params(n_user + 5, n_item + 9, 32, dim=dim, seed=n_item)
    users = np.sort(rng.choice(n_user + 5, n_user, replace=False)).astype(np.int32)
    items = np.sort(rng.choice(n_item + 9, n_item, replace=False)).astype(np.int32)
    feat, dense = _synth_inputs(n_item, n_item, 32, rng)
    want = oum.predict_mat(P, users, items, feat, dense)
    w = um.UserModelWeights(P)
    got = um.predict_all(w, users, items, feat, dense, normalise=False).cpu().numpy()
    rngv = max(float(want.max() - want.min()), 1.0)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * rngv)
    _no_timeout()


def test_full_size_table_tc_vs_ffma_and_oracle(um):
    """BASELINE configs[1] shape: 7176 users x 10728 items.  Tensor-core table == FFMA table; sampled users == oracle;
    min-max properties of the normalised table."""
    U, I = 7176, 10728
    rng = np.random.Generator(np.random.PCG64(2023))
    P = oum.synth_params(U, I + 1, 32, seed=2023)
    users, items = np.arange(U, dtype=np.int32), np.arange(1, I + 1, dtype=np.int32)
    feat, dense = _synth_inputs(I, I, 32, rng)
    w = um.UserModelWeights(P)
    try:
        _mode(1)
        tc, mm = um.predict_all(w, users, items, feat, dense, normalise=False, return_minmax=True)
        _mode(0)
        ff, mm0 = um.predict_all(w, users, items, feat, dense, normalise=False, return_minmax=True)
        _mode(1)
        nm = um.predict_all(w, users, items, feat, dense)
    finally:
        _mode(-1)
    _no_timeout()
    span = float(mm[1] - mm[0])
    assert float((tc - ff).abs().max()) <= 1e-5 * span
    assert float((mm - mm0).abs().max()) <= 1e-5 * span
    assert float(tc.min()) == float(mm[0]) and float(tc.max()) == float(mm[1])
    assert float(nm.min()) == 0.0 and float(nm.max()) == 1.0
    rows = [0, 1, 97, 3587, 7175]
    want = oum.predict_mat(P, users[rows], items, feat, dense)
    np.testing.assert_allclose(tc[rows].cpu().numpy(), want, rtol=1e-5, atol=1e-5 * span)
    want_n = (want - float(mm[0])) / span
    np.testing.assert_allclose(nm[rows].cpu().numpy(), want_n, rtol=1e-5, atol=2e-6)


def test_argument_errors(um, gold):
    from cirs_codes_b200._lib import CirsError
    g, P = gold
    w = um.UserModelWeights(P)
    with pytest.raises(CirsError):
        um.predict_all(w, np.array([10 ** 6]), g["items"], g["item_feat"], g["item_dense"])
    with pytest.raises(CirsError):
        um.predict_all(w, g["users"], g["items"], g["item_feat"][:, :3], g["item_dense"])
    bad = dict(P)
    bad["dnn.linears.1.weight"] = np.zeros((32, 64), np.float32)
    with pytest.raises(CirsError):
        um.UserModelWeights(bad)
