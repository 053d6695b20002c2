"""GPU parity tests of the VirtualTaobao path (SURVEY §8a E6 / E2-E3 Taobao / S1-S3 dense / P1 continuous): the CUDA
kernels through the C ABI against the golden vectors recorded from the reference (tests/golden/taobao_*.npz,
oracle/make_golden.py taobao) and against the CPU oracle.

Bars: done / episode lengths / ready sets exact; the action mapping (clip + float32 scaling) bit exact on identical
inputs; rewards, states, actions, values, log-probs, returns, losses <= 1e-5 relative (absolute floors as in
tests/test_oracle_golden.py)."""
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


def _dev(x, dt):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device="cuda")


# ------------------------------------------------------------------ K1': environment step incl. the reward model
@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_env_step_vs_golden(H, name):
    z = G.load(name)
    c = G.taobao_cfg(z)
    env = H.make_taobao_env(z, c)
    for it in range(c["iters"]):
        obs = env.reset(users=z[f"it{it}/users"])
        assert np.array_equal(obs, z[f"it{it}/reset_obs"])
        for t, ref in enumerate(G.turns(z, it)):
            obs_next, rew, done, info = env.step(ref["obs_next_raw"][:, :27], ref["env_id"])
            assert np.array_equal(done, ref["done"]), f"{name} it{it} turn {t}"
            G.assert_close(rew, ref["rew"], 1e-5, 1e-6, what=f"{name} it{it} turn {t} rew")
            assert np.array_equal(obs_next[:, :27], ref["obs_next_raw"][:, :27])
            assert np.array_equal(obs_next[:, 28:], ref["obs_next_raw"][:, 28:])      # [0, turn]
            G.assert_close(obs_next[:, 27], ref["obs_next_raw"][:, 27], 1e-5, 1e-6, what="obs reward column")


def test_env_map_action_in_kernel_is_bit_exact(H):
    """env->map_action = 1 (fused rollout): clip + float32 scaling inside the kernel == numpy's (base.py:164-172)."""
    from cirs_codes_b200 import _lib
    from oracle import nets
    z = G.load("taobao_N3")
    c = G.taobao_cfg(z)
    env = H.make_taobao_env(z, c, B=64)
    env.reset(users=np.tile(z["it0/users"], (11, 1))[:64])
    raw = np.random.default_rng(0).normal(0, 1.2, size=(64, 27)).astype(np.float32)
    d_raw, out = _dev(raw, torch.float32), torch.zeros(64, 27, device="cuda")
    rew, done = torch.zeros(64, device="cuda"), torch.zeros(64, dtype=torch.uint8, device="cuda")
    _lib.call("cirs_taobao_step", C.byref(env._struct_raw), 64, None, None, _lib.ptr(d_raw), _lib.ptr(out),
              _lib.ptr(rew), _lib.ptr(done), 0, None, None, None, None, None, 0, _lib.stream())
    want = nets.map_action(raw, z["action_low"], z["action_high"])
    assert np.array_equal(out.cpu().numpy(), want)


# ------------------------------------------------------------------ K2: tracker step with dense inputs (d = 27, 3 heads)
@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_tracker_step_vs_golden(H, name):
    z = G.load(name)
    c = G.taobao_cfg(z)
    trk = H.make_taobao_tracker(z, c)
    B = c["B"]
    trk.build_state(dim_batch=B, reset=True)
    s0 = trk.build_state(obs=z["it0/reset_obs"], env_id=np.arange(B))["obs"]
    G.assert_close(s0.cpu().numpy(), z["it0/s0"], 1e-5, 1e-6, what="s0")
    for t, ref in enumerate(G.turns(z, 0)):
        s = trk.build_state(obs_next=ref["obs_next_raw"], rew=ref["rew"], done=ref["done"], info={}, policy=None,
                            env_id=ref["env_id"])["obs_next"]
        G.assert_close(s.cpu().numpy(), ref["state_next"], 1e-5, 1e-6, what=f"{name} state_next turn {t}")


# ------------------------------------------------------------------ K3': continuous actor
@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_actorprob_sample_vs_golden(H, name):
    from oracle import nets
    z = G.load(name)
    c = G.taobao_cfg(z)
    pol = H.make_taobao_policy(z, c, None)
    R = nets.rl_params_continuous(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    for t, ref in enumerate(G.turns(z, 0)):
        out = pol.forward(_dev(ref["state"], torch.float32), noise_q=ref["eps"])
        mu, sigma = (x.cpu().numpy() for x in out.logits)
        G.assert_close(mu, ref["mu"], 1e-5, 1e-6, what=f"mu turn {t}")
        G.assert_close(sigma, ref["sigma"], 1e-5, what=f"sigma turn {t}")
        act = out.act.cpu().numpy()
        want = nets.sample_normal(torch.tensor(ref["mu"]), torch.tensor(ref["sigma"]), ref["eps"]).numpy()
        G.assert_close(act, want, 1e-5, 1e-6, what=f"act turn {t}")
        G.assert_close(pol.map_action(act), ref["obs_next_raw"][:, :27], 1e-5, 1e-6, what="mapped action")
        m, s = nets.actor_mu_sigma(R, torch.tensor(ref["state"]))
        want_lp = nets.normal_log_prob(m, s, act).detach().numpy()
        G.assert_close(out.logp.cpu().numpy(), want_lp, 1e-5, 1e-5, what="logp")
        want_v = nets.critic_value(R, torch.tensor(ref["state"])).detach().numpy()
        G.assert_close(out.value.cpu().numpy(), want_v, 1e-5, 1e-6, what="value")
    pol.eval()
    pol._deterministic_eval = True
    ref = G.turns(z, 0)[0]
    out = pol.forward(_dev(ref["state"], torch.float32))
    G.assert_close(out.act.cpu().numpy(), ref["mu"], 1e-5, 1e-6, what="deterministic act == mu")


def test_actorprob_philox_normal_moments(H):
    z = G.load("taobao_N3")
    c = G.taobao_cfg(z)
    pol = H.make_taobao_policy(z, c, None, seed=9)
    s = torch.tensor(G.turns(z, 0)[0]["state"][:1])
    n = 40000
    out = pol.forward(s.repeat(n, 1).cuda())
    mu, sigma = out.logits
    zed = ((out.act - mu) / sigma).cpu().numpy()
    assert abs(zed.mean()) < 0.01 and abs(zed.std() - 1.0) < 0.01
    assert abs(np.mean(zed ** 3)) < 0.03 and abs(np.mean(zed ** 4) - 3.0) < 0.1
    assert abs(np.corrcoef(zed[:, 0], zed[:, 1])[0, 1]) < 0.02


# ------------------------------------------------------------------ rollout driver + update
def _replay(H, z, c, it, trk, pol, fused=False):
    import cirs_codes_b200 as cb
    env = H.make_taobao_env(z, c)
    buf = cb.VectorReplayBuffer(c["B"] * (c["T"] + 2), c["B"])
    col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=fused)
    gt = G.turns(z, it)
    res = col.collect(n_episode=c["B"], users=z[f"it{it}/users"], noise_fn=lambda t, n: gt[t]["eps"])
    return buf, res


def _check_buffer(buf, res, z, it, tol=1e-5):
    P = f"it{it}/"
    idx = buf.sample_index(0)
    assert np.array_equal(buf._lengths, z[P + "buf/lengths"])
    assert np.array_equal(idx, z[P + "buf/index"])
    assert np.array_equal(buf.done[idx], z[P + "buf/done"])
    G.assert_close(buf.act[idx], z[P + "buf/act"], tol, 1e-6, what="buf act")
    G.assert_close(buf.rew[idx], z[P + "buf/rew"], tol, 1e-6, what="buf rew")
    dt = torch.as_tensor(idx, device="cuda")
    G.assert_close(buf.obs[dt].cpu().numpy(), z[P + "buf/obs"], 2 * tol, 2e-6, what="buf obs")
    G.assert_close(buf.obs_next[dt].cpu().numpy(), z[P + "buf/obs_next"], 2 * tol, 2e-6, what="buf obs_next")
    assert res["n/st"] == int(z[P + "res/n_st"]) and res["n/ep"] == int(z[P + "res/n_ep"])
    assert np.array_equal(res["lens"], z[P + "res/lens"])
    G.assert_close(res["rews"], z[P + "res/rews"], tol, 1e-6, what="episode rewards")
    return idx


@pytest.mark.parametrize("c_loop", [True, False])
@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_update_heads_vs_golden(H, name, c_loop):
    z = G.load(name)
    c = G.taobao_cfg(z)
    trk = H.make_taobao_tracker(z, c)
    pol = H.make_taobao_policy(z, c, None)
    pol.c_loop = c_loop
    buf, res = _replay(H, z, c, 0, trk, pol)
    idx = _check_buffer(buf, res, z, 0)
    out = pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=G.perms(z, 0, len(idx)))
    dt = torch.as_tensor(idx, device="cuda")
    for k, t in (("v_s", pol.v_s), ("returns", pol.returns), ("adv", pol.adv), ("logp_old", pol.logp_old)):
        G.assert_close(t[dt].cpu().numpy(), z[f"it0/upd/{k}"], 1e-5, 1e-5, what=k)
    G.assert_close(out["loss/clip"], z["it0/upd/loss_clip"], 1e-5, 1e-5, what="clip loss")
    G.assert_close(out["loss/vf"], z["it0/upd/loss_vf"], 1e-5, 1e-6, what="vf loss")
    G.assert_close(out["loss/ent"], z["it0/upd/loss_ent"], 1e-5, what="entropy")
    G.assert_close(out["loss"], z["it0/upd/loss"], 1e-5, 1e-5, what="loss")
    G.assert_close(pol.ret_rms.t.cpu().numpy(), z["it0/upd/ret_rms"], 1e-6, what="ret_rms")
    sd = pol.state_dict()
    n_checked = 0
    for k in z.files:
        for net in ("actor", "critic"):
            pre = f"it0/after/{net}/"
            if k.startswith(pre):
                G.assert_close(sd[f"{net}." + k[len(pre):]].numpy(), z[k], 1e-5, G.PARAM_ATOL, what=k)
                n_checked += 1
    assert n_checked == 13   # actor: sigma_param + 4 trunk + mu w/b; critic: 4 trunk + last w/b


def test_update_entropy_coef_vs_oracle(H):
    """ent_coef > 0 (gradient into sigma_param through the entropy), no value clip, no advantage normalisation."""
    from oracle import nets, ppo
    z = G.load("taobao_N3")
    c = G.taobao_cfg(z)
    trk = H.make_taobao_tracker(z, c)
    kw = dict(ent_coef=0.02, value_clip=0, advantage_normalization=0, max_grad_norm=None)
    pol = H.make_taobao_policy(z, c, None, **kw)
    buf, res = _replay(H, z, c, 0, trk, pol)
    idx = buf.sample_index(0)
    n = len(idx)
    perms = G.perms(z, 0, n)
    out = pol.update(0, buf, batch_size=c["batch_size"], repeat=2, perms=perms)
    dt = torch.as_tensor(idx, device="cuda")
    R = nets.rl_params_continuous(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    R = {k: v.clone() for k, v in R.items()}
    obs, obs_next = buf.obs[dt].cpu(), buf.obs_next[dt].cpu()
    act_t = torch.tensor(buf.act[idx])
    last = np.cumsum(buf._lengths) - 1
    unf = np.zeros(n, dtype=bool)
    unf[last] = ~buf.done[idx][last]
    rms = ppo.RunningMeanStd()
    with torch.no_grad():
        v_s, v_n = nets.critic_value(R, obs).numpy(), nets.critic_value(R, obs_next).numpy()
        lp = nets.normal_log_prob(*nets.actor_mu_sigma(R, obs), act_t).numpy()
    returns, adv = ppo.compute_returns(v_s, v_n, buf.rew[idx], buf.done[idx], unf, rms, 0.95, 0.95)
    want = {"loss": [], "loss/clip": [], "loss/vf": [], "loss/ent": []}
    opt = ppo.AdamDup()
    plist = ppo.rl_param_list(R)
    uniq = list({id(p): p for p in plist}.values())
    for p in uniq:
        p.requires_grad_(True)
    for perm in perms:
        for ch in ppo.split_indices(n, c["batch_size"], np.asarray(perm)):
            i_t = torch.as_tensor(ch, dtype=torch.long)
            mu, sigma = nets.actor_mu_sigma(R, obs[i_t])
            a = torch.as_tensor(adv)[i_t]
            ratio = (nets.normal_log_prob(mu, sigma, act_t[i_t]) - torch.as_tensor(lp)[i_t]).exp()
            clip_loss = -torch.min(ratio * a, ratio.clamp(0.8, 1.2) * a).mean()
            vf_loss = ((torch.as_tensor(returns)[i_t] - nets.critic_value(R, obs[i_t])) ** 2).mean()
            ent = nets.normal_entropy(sigma).mean()
            loss = clip_loss + 0.25 * vf_loss - 0.02 * ent
            for q in uniq:
                q.grad = None
            loss.backward()
            opt.step(plist, [q.grad for q in plist])
            for k, v in (("loss", loss), ("loss/clip", clip_loss), ("loss/vf", vf_loss), ("loss/ent", ent)):
                want[k].append(v.item())
    for k in want:
        G.assert_close(out[k], want[k], 1e-5, 1e-5, what=k)
    sd = pol.state_dict()
    mine = nets.rl_params_continuous({k[6:]: v for k, v in sd.items() if k.startswith("actor.")},
                                     {k[7:]: v for k, v in sd.items() if k.startswith("critic.")})
    for k in R:
        G.assert_close(mine[k].numpy(), R[k].detach().numpy(), 1e-5, G.PARAM_ATOL, what=f"param {k}")


@pytest.mark.parametrize("compact", [True, False])
def test_tracker_train_dense_vs_autograd(H, compact):
    """K6 with dense user / item inputs (d = 27, 3 heads): forward == stored states, gradients == autograd."""
    from oracle import nets
    z = G.load("taobao_N3")
    c = G.taobao_cfg(z)
    trk = H.make_taobao_tracker(z, c)
    pol = H.make_taobao_policy(z, c, None)
    buf, _ = _replay(H, z, c, 0, trk, pol)
    buf.sync_device()
    B, L, S = c["B"], buf.sub_size, 20
    lens = buf._lengths
    rng = np.random.default_rng(0)
    d_obs = np.zeros((B * L, S), dtype=np.float32)
    for e in range(B):
        d_obs[e * L:e * L + lens[e]] = rng.normal(size=(lens[e], S))
    check = torch.zeros(B * L, S, device="cuda")
    trk.zero_grad()
    trk.backward_from_buffer(buf, torch.tensor(d_obs, device="cuda"), None, obs_check=check, compact=compact)
    torch.cuda.synchronize()
    idx = buf.sample_index(0)
    it = torch.as_tensor(idx, device="cuda")
    G.assert_close(check[it].cpu().numpy(), buf.obs[it].cpu().numpy(), 1e-5, 1e-6, what="full-sequence forward")
    P = {k: v.clone().requires_grad_(k != "pos_encoder.pe") for k, v in nets.to_params(z, "init/tracker/").items()}
    users = z["it0/users"]
    act_env = buf.d_act_env.cpu().numpy().reshape(B, L, 27)
    rews = buf.rew.reshape(B, L)
    loss = 0.0
    for e in range(B):
        n = int(lens[e])
        toks = [nets.user_token(P, user_dense=users[e:e + 1])]
        if n > 1:
            toks.append(nets.action_token(P, rews[e, :n - 1], act_dense=act_env[e, :n - 1]))
        X = torch.cat(toks, 0).unsqueeze(1)
        s = nets.encode(X, P, c["nhead"], all_positions=True)[:, 0]
        loss = loss + (s * torch.tensor(d_obs[e * L:e * L + n])).sum()
    loss.backward()
    mine = trk.layout.unpack(trk.grad)
    for k, p in P.items():
        if k == "pos_encoder.pe":
            continue
        ref = p.grad.numpy()
        scale = float(np.abs(ref).max()) + 1e-12
        G.assert_close(mine[k].numpy(), ref, 1e-4, 2e-5 * scale, what=f"grad {k}")


@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_full_iterations_vs_golden(H, name):
    """Both recorded iterations of the reference's Taobao run: collect -> update (incl. the tracker) -> collect -> update."""
    z = G.load(name)
    c = G.taobao_cfg(z)
    trk = H.make_taobao_tracker(z, c)
    pol = H.make_taobao_policy(z, c, trk)
    assert pol.state_tracker is trk
    for it in range(c["iters"]):
        buf, res = _replay(H, z, c, it, trk, pol)
        tol = 1e-5 if it == 0 else 5e-5
        idx = _check_buffer(buf, res, z, it, tol)
        P = f"it{it}/"
        out = pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=G.perms(z, it, len(idx)))
        G.assert_close(out["loss/clip"], z[P + "upd/loss_clip"], tol, tol, what=f"clip loss it{it}")
        G.assert_close(out["loss/vf"], z[P + "upd/loss_vf"], tol, 1e-6, what=f"vf loss it{it}")
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
                mine, ref = G.drop_key_bias(mine), G.drop_key_bias(ref)
            G.assert_close(mine, ref, 1e-5, 2 * G.PARAM_ATOL, what=k)


def test_fused_rollout_matches_generic(H):
    """One-kernel rollout (one warp per environment, no grid barrier) == the generic loop over the stand-alone kernels
    (deterministic actions: act = mu), on more environments than one wave of warps; and the sampled fused rollout
    reproduces itself for the same Philox counter."""
    import cirs_codes_b200 as cb
    z = G.load("taobao_N3")
    c = dict(G.taobao_cfg(z), B=300, T=12, thr=-1.0)
    users = H.make_taobao_env(z, c, B=1, seed=3).draw_users(c["B"])
    outs, calibrated = [], False
    for fused in (False, True, False):
        if len(outs) == 1 and not calibrated:
            calibrated = True
            # first pass never leaves (thr < 0): pick the threshold as the median step-to-step action distance so that
            # episode lengths vary in the two compared passes
            ae = outs[0]["act_env_full"]
            dist = np.linalg.norm(ae[:, 1:] - ae[:, :-1], axis=-1)
            c["thr"] = float(np.median(dist))
            outs = []
        env = H.make_taobao_env(z, c)
        trk = H.make_taobao_tracker(z, dict(c), prefix=None)
        pol = H.make_taobao_policy(z, c, None, load=False, deterministic_eval=True)
        pol.eval()
        buf = cb.VectorReplayBuffer(c["B"] * c["T"], c["B"])
        col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=fused)
        assert col.fused == fused
        res = col.collect(n_episode=c["B"], users=users)
        idx = buf.sample_index(0)
        it = torch.as_tensor(idx, device="cuda")
        buf.sync_device()
        outs.append(dict(res=res, lens=buf._lengths.copy(), act=buf.act[idx].copy(), rew=buf.rew[idx].copy(),
                         done=buf.done[idx].copy(), obs=buf.obs[it].cpu().numpy(),
                         obs_next=buf.obs_next[it].cpu().numpy(), act_env=buf.d_act_env[it].cpu().numpy(),
                         act_env_full=buf.d_act_env.cpu().numpy().reshape(c["B"], buf.sub_size, 27)))
    a, b = outs
    assert np.array_equal(a["lens"], b["lens"]) and np.array_equal(a["done"], b["done"])
    assert a["lens"].min() >= 1 and a["lens"].max() <= c["T"] and len(np.unique(a["lens"])) > 1
    for k in ("act", "act_env", "rew", "obs", "obs_next"):
        G.assert_close(a[k], b[k], 1e-6, 1e-7, what=k)
    assert a["res"]["n/st"] == b["res"]["n/st"] and np.array_equal(a["res"]["lens"], b["res"]["lens"])
    G.assert_close(a["res"]["rews"], b["res"]["rews"], 1e-6, what="episode rewards")


# ------------------------------------------------------------------ E6: raw VirtualTB -- user generator, click model, step
def _vtb():
    z = G.load("taobao_usergen")
    sub = lambda p: {k[len(p):]: z[k] for k in z.files if k.startswith(p)}  # noqa: E731
    return z, sub("generator/"), sub("action/")


def test_virtualtb_user_generator_vs_golden():
    """UserModel.generate on the device (csrc/virtualtb.cu) with the reference's recorded seeds z and race noise q:
    the one-hot users are exactly the reference's; with its own Philox draws it still produces valid one-hot x 11
    users whose per-group frequencies follow the generator's softmax probabilities."""
    import cirs_codes_b200 as cb
    from cirs_codes_b200 import env as E, params
    from oracle import env as oenv
    z, gen, act = _vtb()
    packed = params.virtualtb_pack(generator_sd=E._cpu_sd(gen), device="cuda")
    got = E.generate_users(packed, len(z["gen/z"]), torch.device("cuda"), z=z["gen/z"], q=z["gen/q"]).cpu().numpy()
    assert np.array_equal(got, z["gen/user"])
    # the TaobaoVectorEnv draws its users through the same kernel when it is given the generator
    env = cb.TaobaoVectorEnv(4, H_um(), generator=gen, seed=3)
    u = env.draw_users(512)
    off = np.array(oenv.VirtualTBOracle.GROUPS)
    assert u.shape == (512, 88) and np.all(u.sum(1) == 11)
    assert all(np.all(u[:, lo:hi].sum(1) == 1) for lo, hi in zip(off[:-1], off[1:]))
    u2 = env.draw_users(512)
    assert not np.array_equal(u, u2)                       # a fresh Philox offset per call
    # distribution check: mean one-hot frequency vs mean softmax probability over many seeds (same z distribution)
    rng = np.random.default_rng(0)
    zz = rng.random((4096, 128), dtype=np.float32)
    o = oenv.VirtualTBOracle(gen, None)
    h = o._leaky(zz @ o.gen["0.weight"].T + o.gen["0.bias"])
    x = h @ o.gen["2.weight"].T + o.gen["2.bias"]
    want = np.concatenate([o._softmax(x[:, lo:hi]) for lo, hi in zip(off[:-1], off[1:])], axis=1).mean(0)
    mine = E.generate_users(packed, 4096, torch.device("cuda"), seed=11, offset=1, z=zz).cpu().numpy().mean(0)
    assert np.abs(mine - want).max() < 0.03, float(np.abs(mine - want).max())


def H_um():
    z = G.load("taobao_N3")
    return {k[len("usermodel/"):]: torch.tensor(np.asarray(z[k])) for k in z.files if k.startswith("usermodel/")}


def test_virtualtb_step_vs_golden_and_oracle():
    """VirtualTB.step on the device: clicks (a, b) of the reference's recorded click-model calls are exact; a
    teacher-forced episode against the oracle (exit test, turn counter, observation layout)."""
    import cirs_codes_b200 as cb
    from oracle import env as oenv
    z, gen, act = _vtb()
    n = len(z["click/act"])
    env = cb.VirtualTBVectorEnv(n, gen, act, max_turn=50, num_leave_compute=5, leave_threshold=3.0, seed=5)
    env.reset(users=z["click/user"])
    # the golden calls used arbitrary page numbers: set the environments' turn counters to them
    env.turn.copy_(torch.as_tensor(z["click/page"].reshape(-1).astype(np.int32), device="cuda"))
    obs, rew, done, info = env.step(z["click/act"], q=z["click/q"])
    assert np.array_equal(rew.astype(np.int64), z["click/result"][:, 0])
    live = ~done
    assert np.array_equal(obs[live, 27:29].astype(np.int64), z["click/result"][live])
    assert np.array_equal(obs[:, 29].astype(np.int64), z["click/page"].reshape(-1).astype(np.int64) + 1)
    # a short episode with the oracle's click model and its own noise: rewards exact, exit test on repeated actions
    o = oenv.VirtualTBOracle(gen, act)
    B = 8
    env = cb.VirtualTBVectorEnv(B, gen, act, max_turn=6, num_leave_compute=3, leave_threshold=1.0, seed=5)
    users = z["gen/user"][:B]
    first = env.reset(users=users)
    assert first.shape == (B, 91) and np.array_equal(first[:, :88], users) and np.all(first[:, 88:] == 0)
    rng = np.random.default_rng(2)
    prev = None
    for t in range(6):
        a = rng.uniform(-1, 1, size=(B, 27)).astype(np.float32)
        if t == 2:
            a[:4] = prev[:4] + 0.01                       # within the exit radius of the previous action -> leave
        q = rng.exponential(size=(B, 21)).astype(np.float32)
        obs, rew, done, info = env.step(a, q=q)
        want = o.click(users, np.full((B, 1), t, np.float32), a, q)
        assert np.array_equal(rew.astype(np.int64), want[:, 0])
        exp_done = np.zeros(B, bool)
        if t == 2:
            exp_done[:4] = True
        if t == 5:
            exp_done[:] = True                             # t >= max_turn - 1
        assert np.array_equal(done, exp_done), (t, done)
        assert np.all(obs[:, 29] == t + 1) and np.array_equal(obs[:, :27].astype(np.float32), a)
        prev = a
