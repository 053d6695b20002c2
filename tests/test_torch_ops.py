"""torch.ops.cirs_b200.* (csrc/torch_ops.cpp, TORCH_LIBRARY): the C ABI registered as PyTorch operators.  CPU: the
library loads, the six operators of SURVEY 8b are registered with tensor-in / tensor-out schemas, and a call without CUDA
tensors fails loudly.  GPU: every operator gives bit-identical results to the ctypes call of the same C entry point."""
import ctypes as C

import numpy as np
import pytest
import torch

from cirs_codes_b200 import _lib, torch_ops

OPS = ("env_step_kuaishou", "tracker_step", "actor_sample", "gae", "ppo_minibatch", "adam_clip")


def _ops():
    import os
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(torch_ops.LIB):
        import __graft_entry__ as g
        g.build()
    return torch_ops.load()


def test_ops_are_registered_and_fail_loudly_without_cuda():
    ops = _ops()
    for name in OPS:
        schema = str(getattr(ops, name).default._schema)
        assert schema.startswith(f"cirs_b200::{name}("), schema
        assert "Tensor" in schema
    n_slot = torch.tensor([2, 1], dtype=torch.int32)
    x = torch.zeros(6)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.gae(n_slot, x, x, x, torch.zeros(6, dtype=torch.uint8), 0.95, 0.95, None, None)   # CPU tensors: no kernel


@pytest.mark.gpu
def test_gae_and_adam_clip_ops_match_c_abi():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, L = 37, 9
    n_slot = torch.randint(1, L + 1, (B,), device="cuda", generator=g).int()
    v_s, v_next, rew = (torch.randn(B * L, device="cuda", generator=g) for _ in range(3))
    done = (torch.rand(B * L, device="cuda", generator=g) < 0.2).to(torch.uint8)
    rms = torch.tensor([0.3, 1.7, 50.0], dtype=torch.float64, device="cuda")
    mom = torch.zeros(3, dtype=torch.float64, device="cuda")
    ret, adv = ops.gae(n_slot, v_s, v_next, rew, done, 0.95, 0.9, rms, mom)
    ret2, adv2 = torch.zeros_like(v_s), torch.zeros_like(v_s)
    mom2, scratch = torch.zeros_like(mom), torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    _lib.call("cirs_compute_returns", B, L, _lib.ptr(n_slot), _lib.ptr(v_s), _lib.ptr(v_next), _lib.ptr(rew),
              _lib.ptr(done), 0.95, 0.9, _lib.ptr(rms), _lib.ptr(scratch), _lib.ptr(mom2), _lib.ptr(ret2),
              _lib.ptr(adv2), _lib.stream())
    assert torch.equal(ret, ret2) and torch.equal(adv, adv2) and torch.equal(mom, mom2)

    n, n_dup = 5000, 1200
    cfg = _lib.PPOConfigStruct(0.2, 0.25, 0.0, 0.5, 1, 1, 1e-3, 0.9, 0.999, 1e-8)
    bufs = [torch.randn(n, device="cuda", generator=g) for _ in range(2)] + \
           [torch.rand(n, device="cuda", generator=g) * 1e-3 for _ in range(2)]
    a = [t.clone() for t in bufs] + [torch.zeros(2, dtype=torch.int32, device="cuda"),
                                     torch.zeros(16, dtype=torch.float64, device="cuda")]
    b = [t.clone() for t in bufs] + [torch.zeros(2, dtype=torch.int32, device="cuda"),
                                     torch.zeros(16, dtype=torch.float64, device="cuda")]
    for _ in range(2):
        ops.adam_clip(a[0], a[1], a[2], a[3], n_dup, torch_ops.handle(cfg), a[4], a[5])
        _lib.call("cirs_clip_adam", _lib.ptr(b[0]), _lib.ptr(b[1]), _lib.ptr(b[2]), _lib.ptr(b[3]), n, n_dup,
                  C.byref(cfg), _lib.ptr(b[4]), _lib.ptr(b[5]), _lib.stream())
    for x, y in zip(a[:5], b[:5]):
        assert torch.equal(x, y)
    assert not torch.equal(a[0], bufs[0])
    with pytest.raises(RuntimeError):
        ops.adam_clip(a[0], a[1][:10], a[2], a[3], n_dup, torch_ops.handle(cfg), a[4], a[5])


@pytest.mark.gpu
def test_env_tracker_actor_minibatch_ops_match_c_abi():
    from tests import gpu_harness as H
    ops = _ops()
    z, c = H.synthetic_case(U=64, I=300, B=16, T=10, N=3, thr=1)
    B = c["B"]
    rng = np.random.default_rng(4)
    users = torch.as_tensor(rng.integers(0, c["U"], B).astype(np.int32), device="cuda")
    # ---- environment step: two identical environments, one stepped through the operator
    e1, e2 = H.make_env(z, c), H.make_env(z, c)
    for e in (e1, e2):
        e.reset_device(users)
    for t in range(3):
        act = torch.as_tensor(rng.integers(0, c["I"], B).astype(np.int32), device="cuda")
        rew, done = ops.env_step_kuaishou(torch_ops.handle(e1._struct), act, None, e1.active, 0)
        rew2, done2 = torch.empty(B, device="cuda"), torch.empty(B, dtype=torch.uint8, device="cuda")
        e2.step_device(act, rew2, done2)
        assert torch.equal(rew, rew2) and torch.equal(done, done2) and torch.equal(e1.active, e2.active)
    # ---- tracker token: two trackers with the same weights and their own K/V caches
    t1, t2 = H.make_tracker(None, c), H.make_tracker(None, c)
    t2.load_state_dict(t1.state_dict())
    ids = torch.arange(B, dtype=torch.int32, device="cuda")
    for trk in (t1, t2):
        trk.build_state(dim_batch=B, reset=True)
    pos = torch.zeros(B, dtype=torch.int32, device="cuda")
    s1 = ops.tracker_step(torch_ops.handle(t1._w), B, t1.dim_state, ids, None, pos, -1, users, None, None, t1.kcache,
                          t1.vcache)
    s2 = torch.empty(B, t2.dim_state, device="cuda")
    t2.step_device(B, ids, None, pos, -1, users, None, None, state_out=s2)
    assert torch.equal(s1, s2) and torch.equal(t1.kcache, t2.kcache)
    # ---- actor: sample with the same Philox seed / offset
    pol = H.make_policy(None, c, None)
    ws = pol.actor_workspace(B)
    act, logp, value = ops.actor_sample(torch_ops.handle(pol._w), s1, ids, None, None, 11, 5, 0, ws)
    act2 = torch.empty(B, dtype=torch.int32, device="cuda")
    logp2, value2 = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    _lib.call("cirs_actor_sample", C.byref(pol._w), B, _lib.ptr(ids), None, _lib.ptr(s1), s1.shape[1], None, 11, 5,
              None, 0, None, _lib.ptr(act2), _lib.ptr(logp2), _lib.ptr(value2), _lib.ptr(ws), _lib.stream())
    assert torch.equal(act, act2) and torch.equal(logp, logp2) and torch.equal(value, value2)
    # ---- one PPO minibatch: losses and the accumulated gradients
    n = B
    idx = torch.arange(n, dtype=torch.int32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(9)
    adv, ret, v_old = (torch.randn(n, device="cuda", generator=g) for _ in range(3))
    stat = torch.tensor([float(n), float(adv.sum()), float((adv * adv).sum())], dtype=torch.float64, device="cuda")   # {count, sum, sum sq}
    pws = pol._ppo_ws(n)
    pol.grad.zero_()
    d1 = torch.zeros(n, pol.dim_state, device="cuda")
    losses = ops.ppo_minibatch(torch_ops.handle(pol._w), torch_ops.handle(pol._g), torch_ops.handle(pol.cfg), n, idx, s1,
                               act, adv, ret, v_old, logp, stat, d1, pws)
    g1 = pol.grad.clone()
    pol.grad.zero_()
    d2, losses2 = torch.zeros_like(d1), torch.zeros(4, device="cuda")
    _lib.call("cirs_ppo_minibatch", C.byref(pol._w), C.byref(pol._g), C.byref(pol.cfg), n, n, _lib.ptr(idx), _lib.ptr(s1),
              _lib.ptr(act), _lib.ptr(adv), _lib.ptr(ret), _lib.ptr(v_old), _lib.ptr(logp), _lib.ptr(stat), _lib.ptr(d2),
              _lib.ptr(losses2), _lib.ptr(pws), _lib.stream())
    torch.cuda.synchronize()
    assert torch.isfinite(losses).all() and float(g1.abs().sum()) > 0
    np.testing.assert_allclose(losses.cpu().numpy(), losses2.cpu().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g1.cpu().numpy(), pol.grad.cpu().numpy(), rtol=1e-5, atol=1e-7)   # atomics: order
    np.testing.assert_allclose(d1.cpu().numpy(), d2.cpu().numpy(), rtol=1e-5, atol=1e-7)
