"""Tensor-core actor head (csrc/head_tc.cu, tcgen05 + 3xTF32) against the FP32-FFMA path (csrc/gemm.cuh) and against
an FP64 autograd reference of the same minibatch loss (core/policy/ppo.py:181-220), all through the C ABI
(cirs_ppo_minibatch / cirs_policy_eval)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import goldutil as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from tests import gpu_harness
    return gpu_harness


def _minibatch(pol, n, seed, tc):
    """One cirs_ppo_minibatch on random rows; returns losses, the flat gradient and d_obs."""
    from cirs_codes_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev, S, A = pol.device, pol.dim_state, pol.n_action
    n_slots = n + 37
    obs = torch.randn(n_slots, S, generator=g).to(dev)
    act = torch.randint(0, A, (n_slots,), generator=g, dtype=torch.int32).to(dev)
    adv = torch.randn(n_slots, generator=g).to(dev)
    ret = torch.randn(n_slots, generator=g).to(dev)
    v_old = torch.randn(n_slots, generator=g).to(dev) * 0.1
    logp_old = (torch.randn(n_slots, generator=g) * 0.05 - float(np.log(A))).to(dev)
    idx = torch.randperm(n_slots, generator=g)[:n].to(torch.int32).to(dev)
    stat = torch.zeros(3, dtype=torch.float64, device=dev)
    a = adv[idx.long()].double()
    stat[0], stat[1], stat[2] = n, a.sum(), (a * a).sum()
    d_obs = torch.zeros(n_slots, S, device=dev)
    losses = torch.zeros(4, device=dev)
    ws = torch.empty(lib.cirs_ppo_workspace_bytes(n, A), dtype=torch.uint8, device=dev)
    lib.cirs_head_tc_enable(int(tc))   # 0 FFMA, 1 tcgen05 + TMA-fed warp-specialised kernels, 2 tcgen05 register-staged
    try:
        _lib.call("cirs_ppo_minibatch", C.byref(pol._w), C.byref(pol._g), C.byref(pol.cfg), n, n, _lib.ptr(idx),
                  _lib.ptr(obs), _lib.ptr(act), _lib.ptr(adv), _lib.ptr(ret), _lib.ptr(v_old), _lib.ptr(logp_old),
                  _lib.ptr(stat), _lib.ptr(d_obs), _lib.ptr(losses), _lib.ptr(ws), _lib.stream())
        torch.cuda.synchronize()
        assert lib.cirs_head_tc_timeout() == 0, "a tensor-core kernel timed out on an mbarrier"
    finally:
        lib.cirs_head_tc_enable(-1)
    inputs = dict(obs=obs, act=act, adv=adv, ret=ret, v_old=v_old, logp_old=logp_old, idx=idx)
    return losses.cpu().numpy(), pol.grad.clone(), d_obs.clone(), inputs


def _reference(pol, inp, n):
    """FP64 autograd of the same minibatch (the reference's arithmetic, ppo.py:181-212)."""
    sd = pol.layout.unpack(pol.flat)
    p = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    idx = inp["idx"].long()
    obs = inp["obs"][idx].double().requires_grad_(True)
    h = torch.relu(obs @ p["trunk.0.weight"].T + p["trunk.0.bias"])
    h = torch.relu(h @ p["trunk.2.weight"].T + p["trunk.2.bias"])
    logits = h @ p["actor.last.weight"].T + p["actor.last.bias"]
    value = h @ p["critic.last.weight"] + p["critic.last.bias"]
    probs = torch.softmax(logits, -1)
    eps = torch.finfo(torch.float32).eps
    lg = torch.log(probs.clamp(eps, 1 - eps))
    a = inp["act"][idx].long()
    logp = lg.gather(1, a[:, None]).flatten()
    adv = inp["adv"][idx].double()
    adv = (adv - adv.mean()) / adv.std()
    ratio = (logp - inp["logp_old"][idx].double()).exp()
    clip = -torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean()
    R, vo = inp["ret"][idx].double(), inp["v_old"][idx].double()
    vclip = vo + (value - vo).clamp(-0.2, 0.2)
    vf = torch.max((R - value) ** 2, (R - vclip) ** 2).mean()
    ent = -(probs * lg).sum(-1).mean()
    loss = clip + 0.25 * vf
    loss.backward()
    return dict(loss=loss.item(), clip=clip.item(), vf=vf.item(), ent=ent.item(), p=p, d_obs=obs.grad)


@pytest.mark.parametrize("n,I", [(200, 160), (300, 1000), (1573, 10728), (129, 4100)])
def test_tc_minibatch_matches_ffma_and_fp64(H, n, I):
    z, c = H.synthetic_case(I=I, seed=11)
    pol = H.make_policy(None, c, None)
    # random (not near-uniform) logits: scale the head so that probabilities spread over several orders of magnitude
    seg = pol.layout.segs["actor.last.weight"]
    pol.flat[seg.offset:seg.offset + seg.size].mul_(8.0)
    l_f, g_f, do_f, inp = _minibatch(pol, n, 3, tc=0)
    l_t, g_t, do_t, _ = _minibatch(pol, n, 3, tc=1)
    l_r, g_r, do_r, _ = _minibatch(pol, n, 3, tc=2)
    ref = _reference(pol, inp, n)
    for name, l in (("ffma", l_f), ("tc", l_t), ("tc-register-staged", l_r)):
        G.assert_close(l[0], ref["loss"], 1e-5, 1e-6, what=f"{name} loss")
        G.assert_close(l[1], ref["clip"], 1e-5, 1e-6, what=f"{name} clip")
        G.assert_close(l[2], ref["vf"], 1e-5, what=f"{name} vf")
        G.assert_close(l[3], ref["ent"], 1e-5, what=f"{name} entropy")
    # gradients: tensor-core path vs FP64 reference, tolerance relative to the largest entry of each tensor
    got = pol.layout.unpack(g_t)
    for k, pr in ref["p"].items():
        want, have = pr.grad.cpu().numpy(), got[k].double().numpy()
        scale = np.abs(want).max()
        assert np.abs(have - want).max() <= 2e-5 * scale + 1e-9, \
            f"{k}: max err {np.abs(have - want).max():.3e} vs scale {scale:.3e}"
    idx = inp["idx"].long()
    want = ref["d_obs"].cpu().numpy()
    have = do_t[idx].cpu().numpy()
    assert np.abs(have - want).max() <= 2e-5 * np.abs(want).max()
    # and the two CUDA paths agree with each other at the same level
    scale = g_f.abs().max().item()
    assert (g_f - g_t).abs().max().item() <= 2e-5 * scale
    assert (g_r - g_t).abs().max().item() <= 2e-5 * scale and (do_r - do_t).abs().max().item() <= 2e-5 * do_t.abs().max().item()


def test_tc_policy_eval_matches_ffma(H):
    from cirs_codes_b200 import _lib
    lib = _lib.load()
    z, c = H.synthetic_case(I=10728, seed=12)
    pol = H.make_policy(None, c, None)
    seg = pol.layout.segs["actor.last.weight"]
    pol.flat[seg.offset:seg.offset + seg.size].mul_(8.0)
    n = 777
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(n, 20, generator=g).cuda()
    act = torch.randint(0, 10728, (n,), generator=g, dtype=torch.int32).cuda()
    idx = torch.arange(n, dtype=torch.int32).cuda()
    out = {}
    for tc in (0, 1):
        lib.cirs_head_tc_enable(tc)
        v, lp = torch.zeros(n).cuda(), torch.zeros(n).cuda()
        ws = torch.empty(lib.cirs_actor_workspace_bytes(n, 10728), dtype=torch.uint8, device="cuda")
        _lib.call("cirs_policy_eval", C.byref(pol._w), n, _lib.ptr(idx), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(v),
                  _lib.ptr(lp), _lib.ptr(ws), _lib.stream())
        torch.cuda.synchronize()
        out[tc] = (v.cpu().numpy(), lp.cpu().numpy())
    lib.cirs_head_tc_enable(-1)
    assert lib.cirs_head_tc_timeout() == 0
    G.assert_close(out[1][0], out[0][0], 1e-6, 1e-7, what="value")
    G.assert_close(out[1][1], out[0][1], 1e-5, what="log-prob")


def test_b2_pair_kernel_variant_matches_ffma_and_fp64():
    """head_tc_dh2_wide_kernel (one N = 128 logits MMA batch per pair of tiles; CIRS_B2_WIDE=1, a switch the library reads
    once per process): the minibatch tests above -- tensor-core path against the FFMA path and FP64 autograd -- in a
    child process with the switch set."""
    import os
    import subprocess
    import sys
    if os.environ.get("CIRS_B2_WIDE") == "1":
        pytest.skip("this is the child process")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k",
                        "test_tc_minibatch_matches_ffma_and_fp64"], cwd=root, env=dict(os.environ, CIRS_B2_WIDE="1"),
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and " passed" in p.stdout, p.stdout[-3000:] + "\n" + p.stderr[-2000:]
