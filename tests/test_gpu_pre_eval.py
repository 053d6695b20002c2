"""process_fn's kernels queued ahead of the collect's read-back (PPOPolicy.post_collect -> cirs_policy_eval_dev, the row
count still on the device) must give the same update as the in-order path (core/policy/ppo.py:96-109: process_fn after
the collect has returned)."""
import numpy as np
import pytest
import torch

import bench
from tests import goldutil as G

pytestmark = pytest.mark.gpu


def _run(pre_eval, iters=3):
    cfg = dict(bench.CONFIGS["small"])
    cfg.update(B=96, I=1500, batch_size=256)
    dev = torch.device("cuda", 0)
    tb = bench.tables(cfg)
    env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
    pol.pre_eval = pre_eval
    rng = np.random.default_rng(5)
    prng = np.random.default_rng(11)
    out, used = [], []
    for it in range(iters):
        col.collect(n_episode=cfg["B"], users=rng.integers(0, cfg["U"], size=cfg["B"]))
        used.append(getattr(pol, "_pre", None) is not None)
        n = len(buf)
        perms = [prng.permutation(n) for _ in range(cfg["repeat"])]
        res = pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"], perms=perms)
        out.append({k: np.asarray(res[k], dtype=np.float64) for k in ("loss", "loss/clip", "loss/vf", "loss/ent")})
    torch.cuda.synchronize()
    return out, pol.flat.cpu().numpy().copy(), trk.flat.cpu().numpy().copy(), pol.ret_rms.t.cpu().numpy().copy(), used


def test_pre_launched_process_fn_matches_in_order():
    a, pa, ta, ra, used_a = _run(True)
    b, pb, tb_, rb, used_b = _run(False)
    assert used_a[1:] == [True] * (len(used_a) - 1) and not used_a[0], used_a   # no capacity before the first update
    assert not any(used_b)
    for x, y in zip(a, b):
        for k in x:
            G.assert_close(x[k], y[k], 1e-5, 1e-6, what=k)
    G.assert_close(ra, rb, 1e-9, what="ret_rms")
    G.assert_close(pa, pb, 1e-5, 2e-5, what="policy parameters")
    G.assert_close(ta, tb_, 1e-5, 2e-5, what="tracker parameters")


def test_capacity_overflow_falls_back():
    """A collect whose transition count exceeds the queued capacity repeats the evaluations in order."""
    cfg = dict(bench.CONFIGS["small"])
    cfg.update(B=96, I=1500, batch_size=256)
    dev = torch.device("cuda", 0)
    tb = bench.tables(cfg)
    env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
    rng = np.random.default_rng(5)
    col.collect(n_episode=cfg["B"], users=rng.integers(0, cfg["U"], size=cfg["B"]))
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=1)
    users = rng.integers(0, cfg["U"], size=cfg["B"])
    col.collect(n_episode=cfg["B"], users=users)
    assert pol._pre is not None
    n = len(buf)
    pol._pre = (pol._pre[0], n - 1)            # pretend the capacity was one row short
    ref_flat = pol.flat.clone()
    res = pol.update(0, buf, batch_size=cfg["batch_size"], repeat=1, perms=[np.arange(n)])
    assert np.isfinite(np.asarray(res["loss"])).all() and not torch.equal(ref_flat, pol.flat)
