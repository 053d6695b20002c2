"""GPU parity tests for the tracker's training pass (K6) and for complete reference iterations (collect + update,
tracker included) replayed on the CUDA path."""
import numpy as np
import pytest
import torch

from tests import goldutil as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from tests import gpu_harness
    return gpu_harness


def _replay(H, z, c, it, trk, pol):
    import cirs_codes_b200 as cb
    env = H.make_env(z, c)
    buf = cb.VectorReplayBuffer(c["B"] * (c["T"] + 2), c["B"])
    col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=False)
    gt = G.turns(z, it)
    res = col.collect(n_episode=c["B"], users=z[f"it{it}/users"], noise_fn=lambda t, n: gt[t]["q"])
    return buf, res


@pytest.fixture
def unfused():
    """Select the layer-by-layer launches of K6 for one test, then restore the default (fused chunk kernel)."""
    from cirs_codes_b200 import _lib
    _lib.load().cirs_tracker_train_fused_enable(0)
    yield
    _lib.load().cirs_tracker_train_fused_enable(-1)


@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_tracker_train_unfused_path_vs_autograd(H, name, unfused):
    """The second CUDA path of K6 (one launch per layer operation) against the same autograd reference."""
    test_tracker_train_forward_and_grads_vs_autograd(H, name, True)


@pytest.mark.parametrize("compact", [True, False])
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_tracker_train_forward_and_grads_vs_autograd(H, name, compact):
    """One full-sequence pass: (a) its decoded states equal the states the rollout stored; (b) the parameter
    gradients for a random upstream d_obs equal torch autograd through the oracle's encoder.  compact = True runs the
    fused chunk kernel + grouped weight-gradient launch (csrc/tracker_fused.cuh), False the padded layer-by-layer path."""
    from oracle import nets
    z = G.load(name)
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    pol = H.make_policy(z, c, None)
    buf, _ = _replay(H, z, c, 0, trk, pol)
    buf.sync_device()
    B, L, S = c["B"], buf.sub_size, 20
    lens = buf._lengths
    rng = np.random.default_rng(0)
    d_obs = np.zeros((B * L, S), dtype=np.float32)
    for e in range(B):
        d_obs[e * L:e * L + lens[e]] = rng.normal(size=(lens[e], S))
    d_dobs = torch.tensor(d_obs, device="cuda")
    check = torch.zeros(B * L, S, device="cuda")
    trk.zero_grad()
    trk.backward_from_buffer(buf, d_dobs, buf.d_users, obs_check=check, compact=compact)
    torch.cuda.synchronize()
    idx = buf.sample_index(0)
    it = torch.as_tensor(idx, device="cuda")
    G.assert_close(check[it].cpu().numpy(), buf.obs[it].cpu().numpy(), 1e-5, 1e-6, what="full-sequence forward")
    # autograd reference
    P = {k: v.clone().requires_grad_(k != "pos_encoder.pe") for k, v in nets.to_params(z, "init/tracker/").items()}
    users, acts, rews = z["it0/users"], buf.act.reshape(B, L), buf.rew.reshape(B, L)
    loss = 0.0
    for e in range(B):
        n = int(lens[e])
        toks = [nets.user_token(P, users=[users[e]])]
        if n > 1:
            toks.append(nets.action_token(P, rews[e, :n - 1], acts=acts[e, :n - 1]))
        X = torch.cat(toks, 0).unsqueeze(1)                                  # [n, 1, d]
        s = nets.encode(X, P, c["nhead"], all_positions=True)[:, 0]         # [n, S]
        loss = loss + (s * torch.tensor(d_obs[e * L:e * L + n])).sum()
    loss.backward()
    mine = trk.layout.unpack(trk.grad)
    for k, p in P.items():
        if k == "pos_encoder.pe":
            continue
        ref = p.grad.numpy()
        scale = float(np.abs(ref).max()) + 1e-12
        G.assert_close(mine[k].numpy(), ref, 1e-4, 2e-5 * scale, what=f"grad {k}")

@pytest.mark.parametrize("tm", ["16", "32"])
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_tracker_train_chunk_heights_vs_autograd(H, name, tm, monkeypatch):
    """The resident-weight chunk kernel with 16-row and with 32-row chunks (greedy chunk plan, chunk_plan_kernel): the
    same forward states and gradients against autograd.  (The library picks the height from the token count; the
    golden cases' episodes are short enough for both.)"""
    monkeypatch.setenv("CIRS_K6_TM", tm)   # (a request: an episode longer than 16 rows keeps 32-row chunks)
    test_tracker_train_forward_and_grads_vs_autograd(H, name, True)


@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_full_iterations_vs_golden(H, name):
    """Both recorded iterations of the reference run (collect -> update incl. the tracker's Adam step -> collect with
    the updated weights -> update): buffers, losses and all parameters must follow the reference."""
    z = G.load(name)
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    pol = H.make_policy(z, c, trk)
    assert pol.state_tracker is trk
    for it in range(c["iters"]):
        buf, res = _replay(H, z, c, it, trk, pol)
        idx = buf.sample_index(0)
        P = f"it{it}/"
        assert np.array_equal(buf._lengths, z[P + "buf/lengths"]), f"lengths it{it}"
        assert np.array_equal(buf.act[idx], z[P + "buf/act"]), f"actions it{it}"
        assert np.array_equal(buf.done[idx], z[P + "buf/done"])
        G.assert_close(buf.rew[idx], z[P + "buf/rew"], 1e-5, what="buf rew")
        dt = torch.as_tensor(idx, device="cuda")
        G.assert_close(buf.obs[dt].cpu().numpy(), z[P + "buf/obs"], 2e-5, 2e-6, what=f"buf obs it{it}")
        out = pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=G.perms(z, it, len(idx)))
        tol = 1e-5 if it == 0 else 5e-5   # iteration 1 runs on weights that already differ by float rounding
        G.assert_close(out["loss/clip"], z[P + "upd/loss_clip"], tol, tol, what=f"clip loss it{it}")
        G.assert_close(out["loss/vf"], z[P + "upd/loss_vf"], tol, what=f"vf loss it{it}")
        G.assert_close(out["loss/ent"], z[P + "upd/loss_ent"], tol, what=f"entropy it{it}")
        G.assert_close(pol.ret_rms.t.cpu().numpy(), z[P + "upd/ret_rms"], 1e-5, what="ret_rms")
        sd = pol.state_dict()
        for k in z.files:
            for net in ("actor", "critic"):
                pre = P + f"after/{net}/"
                if k.startswith(pre):
                    G.assert_close(sd[f"{net}." + k[len(pre):]].numpy(), z[k], 1e-5, 2 * G.PARAM_ATOL, what=k)
        tsd = trk.state_dict()
        for k in z.files:
            pre = P + "after/tracker/"
            if not k.startswith(pre):
                continue
            mine, ref = tsd[k[len(pre):]].numpy().reshape(z[k].shape), z[k]
            if k.endswith("in_proj_bias"):
                mine, ref = G.drop_key_bias(mine), G.drop_key_bias(ref)   # zero-gradient key bias, see CPU test
            G.assert_close(mine, ref, 1e-5, 2 * G.PARAM_ATOL, what=k)


def test_tracker_train_split_phases_match_single_pass(H):
    """cirs_tracker_train phase 1 (forward only, no upstream gradient needed) followed by phase 2 (backward only) on the
    same workspace == the single-launch pass: identical parameter gradients."""
    z = G.load("kuaishou_N5")
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    pol = H.make_policy(z, c, None)
    buf, _ = _replay(H, z, c, 0, trk, pol)
    buf.sync_device()
    B, L, S = c["B"], buf.sub_size, 20
    d_dobs = torch.tensor(np.random.default_rng(1).normal(size=(B * L, S)).astype(np.float32), device="cuda")
    trk.zero_grad()
    trk.backward_from_buffer(buf, d_dobs, buf.d_users)
    one = trk.grad.clone()
    trk.zero_grad()
    trk.forward_async(buf, buf.d_users)
    trk.backward_from_buffer(buf, d_dobs, buf.d_users, after_forward=True)
    torch.cuda.synchronize()
    # float atomics accumulate the chunks' contributions in a different order from launch to launch: rounding level
    scale = float(one.abs().max())
    G.assert_close(trk.grad.cpu().numpy(), one.cpu().numpy(), 1e-5, 1e-6 * scale, what="split-phase gradients")
