"""Helpers shared by the CPU (oracle vs golden) and GPU (CUDA vs oracle vs golden) parity tests."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KUAISHOU_CASES = ["kuaishou_N1", "kuaishou_N5", "kuaishou_v2"]
TAOBAO_CASES = ["taobao_N3"]
PARAM_ATOL = 1e-5  # 1 % of one Adam step (lr 1e-3); see tests/test_oracle_golden.py docstring


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def cfg(z):
    U, I, B, T, N, thr, d, nhead, batch_size, repeat, iters, seed = (int(x) for x in z["cfg"])
    tau, gamma_e, r_decay, ver, use_ab = (float(x) for x in z["cfg_f"])
    return dict(U=U, I=I, B=B, T=T, N=N, thr=thr, d=d, nhead=nhead, batch_size=batch_size, repeat=repeat,
                iters=iters, seed=seed, tau=tau, gamma_exposure=gamma_e, r_decay=r_decay,
                version="v1" if ver == 1.0 else "v2", use_ab=bool(use_ab))


def taobao_cfg(z):
    B, T, N, d, nhead, batch_size, repeat, iters, seed = (int(x) for x in z["cfg"])
    thr, tau, gamma_e, ver = (float(x) for x in z["cfg_f"])
    return dict(B=B, T=T, N=N, d=d, nhead=nhead, batch_size=batch_size, repeat=repeat, iters=iters, seed=seed, thr=thr,
                tau=tau, gamma_exposure=gamma_e, version="v1" if ver == 1.0 else "v2")


def turns(z, it):
    n = int(z[f"it{it}/n_turns"])
    keys = ["state", "q", "probs", "env_id", "obs_next_raw", "rew", "done", "state_next"]
    if f"it{it}/turn0/eps" in z.files:   # continuous actor (VirtualTaobao)
        keys = ["state", "eps", "mu", "sigma", "env_id", "obs_next_raw", "rew", "done", "state_next"]
    return [{k: z[f"it{it}/turn{t}/{k}"] for k in keys} for t in range(n)]


def perms(z, it, n):
    """The minibatch permutations the reference drew (np.random.seed(useed) before policy.update)."""
    c = taobao_cfg(z) if "usermodel_x" in z.files else cfg(z)
    st = np.random.get_state()
    np.random.seed(int(z[f"it{it}/upd/seed"]))
    out = [np.random.permutation(n) for _ in range(c["repeat"])]
    np.random.set_state(st)
    return out


def assert_close(a, b, rtol=1e-5, atol=0.0, what=""):
    """|a-b| <= rtol*max(|b|, scale) elementwise; ``atol`` is the absolute floor for quantities that are
    differences of O(1) numbers (e.g. the PPO clip loss, a mean of zero-mean advantages)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off; max err {err.max():.3e} (tol {tol[bad].min():.3e})"


def drop_key_bias(in_proj_bias):
    """in_proj_bias = [q | k | v]; return [q | v]."""
    d = in_proj_bias.shape[0] // 3
    return np.concatenate([in_proj_bias[:d], in_proj_bias[2 * d:]])
