"""GPU parity tests: the CUDA path (through the C ABI, via the host classes) against the golden vectors produced by
the reference itself (tests/golden/*.npz) and against the CPU oracle on the same seeded inputs.

Bars (north_star): actions / done / episode lengths bit-exact; rewards, states, values, log-probs, returns, losses
<= 1e-5 relative (absolute floors as documented in tests/test_oracle_golden.py)."""
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


# ------------------------------------------------------------------ K1: environment step
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
@pytest.mark.parametrize("use_dist", [False, True])
def test_env_step_vs_golden(H, name, use_dist):
    from cirs_codes_b200 import synth
    z = G.load(name)
    c = G.cfg(z)
    over = {"df_dist_small": synth.jaccard_distance_matrix(z["cats"])} if use_dist else {}
    env = H.make_env(z, c, **over)
    for it in range(c["iters"]):
        obs = env.reset(users=z[f"it{it}/users"])
        assert np.array_equal(obs[:, 0], z[f"it{it}/users"])
        for t, ref in enumerate(G.turns(z, it)):
            obs_next, rew, done, info = env.step(ref["obs_next_raw"], ref["env_id"])
            assert np.array_equal(obs_next, ref["obs_next_raw"])
            assert np.array_equal(done, ref["done"]), f"{name} it{it} turn {t}"
            G.assert_close(rew, ref["rew"], 1e-5, what=f"{name} it{it} turn {t} rew")


def test_env_raw_reward_and_seen(H):
    z = G.load("kuaishou_N1")
    c = G.cfg(z)
    env = H.make_env(z, c, simulated=False, track_seen=True)
    users = z["it0/users"]
    env.reset(users=users)
    ref = G.turns(z, 0)[0]
    act = ref["obs_next_raw"][:, 0]
    _, rew, done, info = env.step(ref["obs_next_raw"], ref["env_id"])
    G.assert_close(rew, z["mat"][users, act], 1e-6, what="raw reward mat[u,a]")       # kuaishouEnv.py:171
    G.assert_close(info["cum_reward"], z["mat"][users, act], 1e-6, what="cum_reward")
    seen = env.seen.cpu().numpy().view(np.uint32)
    for e, a in enumerate(act):
        assert seen[e, a >> 5] == np.uint32(1) << np.uint32(a & 31)


# ------------------------------------------------------------------ K2: tracker step
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_tracker_step_vs_golden(H, name):
    z = G.load(name)
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    B = c["B"]
    trk.build_state(dim_batch=B, reset=True)
    s0 = trk.build_state(obs=z["it0/users"].reshape(-1, 1), env_id=np.arange(B))["obs"]
    G.assert_close(s0.cpu().numpy(), z["it0/s0"], 1e-5, 1e-6, what="s0")
    for t, ref in enumerate(G.turns(z, 0)):
        s = trk.build_state(obs_next=ref["obs_next_raw"], rew=ref["rew"], done=ref["done"], info={}, policy=None,
                            env_id=ref["env_id"])["obs_next"]
        G.assert_close(s.cpu().numpy(), ref["state_next"], 1e-5, 1e-6, what=f"{name} state_next turn {t}")


# ------------------------------------------------------------------ K3: actor head + sampler
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_actor_sample_vs_golden(H, name):
    from oracle import nets
    z = G.load(name)
    c = G.cfg(z)
    pol = H.make_policy(z, c, None)
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    for t, ref in enumerate(G.turns(z, 0)):
        out = pol.forward(_dev(ref["state"], torch.float32), noise_q=ref["q"])
        act = out.act.cpu().numpy()
        assert np.array_equal(act, ref["obs_next_raw"][:, 0]), f"{name} turn {t}: sampled actions differ"
        want_logp = nets.log_prob(torch.tensor(ref["probs"]), act).numpy()
        G.assert_close(out.logp.cpu().numpy(), want_logp, 1e-5, 1e-6, what="logp")
        want_v = nets.critic_value(R, torch.tensor(ref["state"])).detach().numpy()
        G.assert_close(out.value.cpu().numpy(), want_v, 1e-5, 1e-6, what="value")
    # argmax mode == argmax of the reference's probabilities
    pol.eval()
    pol._deterministic_eval = True
    ref = G.turns(z, 0)[0]
    out = pol.forward(_dev(ref["state"], torch.float32))
    assert np.array_equal(out.act.cpu().numpy(), ref["probs"].argmax(-1))


def test_actor_philox_sampler_distribution(H):
    """In-kernel Philox race: empirical action frequencies follow the softmax probabilities (chi-square-ish bound)
    and are reproducible for a fixed (seed, call counter)."""
    from oracle import nets
    z = G.load("kuaishou_N1")
    c = G.cfg(z)
    pol = H.make_policy(z, c, None, seed=123)
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    s = torch.tensor(G.turns(z, 0)[0]["state"][:1])
    p = nets.actor_probs(R, s).detach().numpy()[0]
    n = 20000
    out = pol.forward(s.repeat(n, 1).cuda())
    cnt = np.bincount(out.act.cpu().numpy(), minlength=c["I"])
    err = np.abs(cnt / n - p)
    assert err.max() < 5 * np.sqrt(p.max() / n) + 1e-3, err.max()
    pol2 = H.make_policy(z, c, None, seed=123)
    out2 = pol2.forward(s.repeat(n, 1).cuda())
    assert torch.equal(out.act, out2.act)


def test_policy_eval_matches_oracle(H):
    from oracle import nets
    from cirs_codes_b200 import _lib
    z = G.load("kuaishou_N5")
    c = G.cfg(z)
    pol = H.make_policy(z, c, None)
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    obs = torch.tensor(z["it0/buf/obs"])
    act = z["it0/buf/act"].astype(np.int32)
    n = len(act)
    rows = np.random.default_rng(0).permutation(n).astype(np.int32)[: n - 3]
    d_obs, d_act, d_rows = obs.cuda(), _dev(act, torch.int32), _dev(rows, torch.int32)
    val = torch.full((n,), 7.0, device="cuda")
    lp = torch.full((n,), 7.0, device="cuda")
    _lib.call("cirs_policy_eval", C.byref(pol._w), len(rows), _lib.ptr(d_rows), _lib.ptr(d_obs), _lib.ptr(d_act),
              _lib.ptr(val), _lib.ptr(lp), _lib.ptr(pol._actor_ws(n)), _lib.stream())
    want_v = nets.critic_value(R, obs).detach().numpy()
    want_lp = nets.log_prob(nets.actor_probs(R, obs).detach(), act).numpy()
    G.assert_close(val.cpu().numpy()[rows], want_v[rows], 1e-5, 1e-6, what="eval value")
    G.assert_close(lp.cpu().numpy()[rows], want_lp[rows], 1e-5, 1e-6, what="eval logp")
    untouched = np.setdiff1d(np.arange(n), rows)
    assert np.all(val.cpu().numpy()[untouched] == 7.0)


# ------------------------------------------------------------------ K4: returns
def _returns(v_s, v_next, rew, done, lens, L, gamma, lam, rms=None):
    from cirs_codes_b200 import _lib
    B = len(lens)
    n = B * L
    ret, adv = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    scratch = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    mom = torch.zeros(3, dtype=torch.float64, device="cuda")
    # keep the device inputs alive until the call has been issued (a temporary would be freed and its block reused)
    d_len, d_vs, d_vn = _dev(lens, torch.int32), _dev(v_s, torch.float32), _dev(v_next, torch.float32)
    d_rew, d_done = _dev(rew, torch.float32), _dev(done, torch.uint8)
    _lib.call("cirs_compute_returns", B, L, _lib.ptr(d_len), _lib.ptr(d_vs), _lib.ptr(d_vn), _lib.ptr(d_rew),
              _lib.ptr(d_done), gamma, lam, _lib.ptr(rms), _lib.ptr(scratch), _lib.ptr(mom),
              _lib.ptr(ret), _lib.ptr(adv), _lib.stream())
    torch.cuda.synchronize()
    if rms is not None:
        _lib.call("cirs_rms_update", _lib.ptr(rms), _lib.ptr(mom), _lib.stream())
    return ret.cpu().numpy(), adv.cpu().numpy()


def test_gae_known_answers_tianshou():
    # tianshou/test/base/test_returns.py:58-72 (gamma .99, lambda .95, explicit values); one sub-buffer of 12 slots
    done = np.array([0, 0, 0, 1., 0, 0, 0, 1, 0, 0, 0, 1])
    rew = np.array([101, 102, 103., 200, 104, 105, 106, 201, 107, 108, 109, 202])
    v = np.array([2., 3., 4, -1, 5., 6., 7, -2, 8., 9., 10, -3])
    truth = [454.8344, 376.1143, 291.298, 200., 464.5610, 383.1085, 295.387, 201., 474.2876, 390.1027, 299.476, 202.]
    ret, _ = _returns(np.roll(v, 1), v, rew, done, [12], 12, 0.99, 0.95)
    assert np.allclose(ret, truth, rtol=1e-6)
    # test_returns.py:36-57 (gamma .1, lambda 1, zero values); second case ends unfinished inside the episode
    ret, _ = _returns(np.zeros(7), np.zeros(7), [7, 6, 1, 2, 3, 4, 5.], [0, 1, 0, 1, 0, 1, 0.], [7], 7, 0.1, 1.0)
    assert np.allclose(ret, [7.6, 6, 1.2, 2, 3.4, 4, 5])
    ret, _ = _returns(np.zeros(7), np.zeros(7), [7, 6, 1, 2, 3, 4, 5.], [0, 1, 0, 1, 0, 0, 1.], [7], 7, 0.1, 1.0)
    assert np.allclose(ret, [7.6, 6, 1.2, 2, 3.45, 4.5, 5])


@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_returns_vs_golden(name):
    """Feed the reference's own critic outputs (golden v_s) and rewards: returns / adv / ret_rms must match."""
    z = G.load(name)
    c = G.cfg(z)
    rms = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device="cuda")
    for it in range(c["iters"]):
        lens = z[f"it{it}/buf/lengths"]
        B, L = len(lens), int(lens.max())
        pos = np.concatenate([np.arange(l) + i * L for i, l in enumerate(lens)])
        def scatter(x):
            out = np.zeros(B * L, dtype=np.float64)
            out[pos] = x
            return out

        v_s = z[f"it{it}/upd/v_s"]
        done = z[f"it{it}/buf/done"]
        v_next = np.zeros_like(v_s)
        v_next[:-1] = v_s[1:]          # non-terminal obs_next is the next slot's obs; terminal ones are masked
        ret, adv = _returns(scatter(v_s), scatter(v_next), scatter(z[f"it{it}/buf/rew"]), scatter(done), lens, L,
                            0.95, 0.95, rms)
        G.assert_close(ret[pos], z[f"it{it}/upd/returns"], 1e-5, 1e-6, what="returns")
        G.assert_close(adv[pos], z[f"it{it}/upd/adv"], 1e-5, 1e-6, what="adv")
        G.assert_close(rms.cpu().numpy(), z[f"it{it}/upd/ret_rms"], 1e-6, what="ret_rms")


# ------------------------------------------------------------------ rollout driver + update
def _golden_collect(H, z, c, it, tracker, policy, fused=False):
    import cirs_codes_b200 as cb
    env = H.make_env(z, c)
    buf = cb.VectorReplayBuffer(c["B"] * (c["T"] + 2), c["B"])
    col = cb.Collector(policy, env, buf, preprocess_fn=tracker.build_state, fused=fused)
    gt = G.turns(z, it)
    res = col.collect(n_episode=c["B"], users=z[f"it{it}/users"], noise_fn=lambda t, n: gt[t]["q"])
    return col, buf, res


@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_generic_collect_vs_golden(H, name):
    """The reference's collect loop on the CUDA components with the reference's own race noise: the buffer content
    and the collect statistics must reproduce the reference run."""
    z = G.load(name)
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    pol = H.make_policy(z, c, None)
    col, buf, res = _golden_collect(H, z, c, 0, trk, pol)
    idx = buf.sample_index(0)
    assert np.array_equal(buf._lengths, z["it0/buf/lengths"])
    assert np.array_equal(idx, z["it0/buf/index"])
    assert np.array_equal(buf.act[idx], z["it0/buf/act"])
    assert np.array_equal(buf.done[idx], z["it0/buf/done"])
    G.assert_close(buf.rew[idx], z["it0/buf/rew"], 1e-5, what="buf rew")
    it = torch.as_tensor(idx, device="cuda")
    G.assert_close(buf.obs[it].cpu().numpy(), z["it0/buf/obs"], 1e-5, 1e-6, what="buf obs")
    G.assert_close(buf.obs_next[it].cpu().numpy(), z["it0/buf/obs_next"], 1e-5, 1e-6, what="buf obs_next")
    assert res["n/st"] == int(z["it0/res/n_st"]) and res["n/ep"] == int(z["it0/res/n_ep"])
    assert np.array_equal(res["lens"], z["it0/res/lens"])
    assert np.array_equal(res["idxs"], z["it0/res/idxs"])
    G.assert_close(res["rews"], z["it0/res/rews"], 1e-5, what="episode rewards")
    # buffer index arithmetic used by the reference's callbacks (evaluation.py:309-354)
    last = buf.last_index
    assert np.array_equal(buf.next(last), last) and len(buf.unfinished_index()) == 0
    first = np.arange(c["B"]) * buf.sub_size
    assert np.array_equal(buf.prev(first), first)
    if len(idx) > c["B"]:
        mid = idx[(~np.isin(idx, first))]
        assert np.array_equal(buf.prev(mid), mid - 1)


@pytest.mark.parametrize("c_loop", [True, False])
@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_update_heads_vs_golden_and_oracle(H, name, c_loop):
    """policy.update on the replayed rollout with the reference's minibatch permutations.  The tracker is frozen here
    (its step comes after all minibatches, core/policy/ppo.py:235, so iteration-0 losses and actor / critic weights
    do not depend on it); tests/test_gpu_tracker_train.py covers the tracker."""
    from oracle import nets, pipeline, ppo
    z = G.load(name)
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    pol = H.make_policy(z, c, None)
    pol.c_loop = c_loop     # one C call for the whole loop / the per-minibatch entry points the multi-GPU path uses
    col, buf, res = _golden_collect(H, z, c, 0, trk, pol)
    n = len(buf)
    perms = G.perms(z, 0, n)
    out = pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=perms)
    idx = torch.as_tensor(buf.sample_index(0), device="cuda")
    for k, t in (("v_s", pol.v_s), ("returns", pol.returns), ("adv", pol.adv), ("logp_old", pol.logp_old)):
        G.assert_close(t[idx].cpu().numpy(), z[f"it0/upd/{k}"], 1e-5, 1e-6, what=k)
    G.assert_close(out["loss/clip"], z["it0/upd/loss_clip"], 1e-5, 1e-5, what="clip loss")
    G.assert_close(out["loss/vf"], z["it0/upd/loss_vf"], 1e-5, what="vf loss")
    G.assert_close(out["loss/ent"], z["it0/upd/loss_ent"], 1e-5, what="entropy")
    G.assert_close(out["loss"], z["it0/upd/loss"], 1e-5, 1e-5, what="loss")
    G.assert_close(pol.ret_rms.t.cpu().numpy(), z["it0/upd/ret_rms"], 1e-6, what="ret_rms")
    sd = pol.state_dict()
    for k in z.files:
        for net in ("actor", "critic"):
            pre = f"it0/after/{net}/"
            if k.startswith(pre):
                G.assert_close(sd[f"{net}." + k[len(pre):]].numpy(), z[k], 1e-5, G.PARAM_ATOL, what=k)
    assert len(out["loss"]) == len(z["it0/upd/loss"])


def test_update_entropy_and_noclip_vs_oracle(H):
    """Non-default loss configuration (ent_coef > 0, no value clip, no advantage normalisation, no grad clipping)
    against the oracle's autograd."""
    from oracle import nets, pipeline, ppo
    z = G.load("kuaishou_N5")
    c = G.cfg(z)
    trk = H.make_tracker(z, c)
    kw = dict(ent_coef=0.01, value_clip=0, advantage_normalization=0, max_grad_norm=None)
    pol = H.make_policy(z, c, None, **kw)
    col, buf, res = _golden_collect(H, z, c, 0, trk, pol)
    n = len(buf)
    perms = G.perms(z, 0, n)
    out = pol.update(0, buf, batch_size=c["batch_size"], repeat=2, perms=perms)
    # oracle on the same buffer content
    idx = buf.sample_index(0)
    it = torch.as_tensor(idx, device="cuda")
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    R = {k: v.clone() for k, v in R.items()}
    traj = pipeline.Trajectory(buf.obs[it].cpu(), buf.obs_next[it].cpu(), buf.act[idx], buf.rew[idx], buf.done[idx],
                               buf._lengths)
    rms = ppo.RunningMeanStd()
    with torch.no_grad():
        v_s = nets.critic_value(R, traj.obs).numpy()
        v_n = nets.critic_value(R, traj.obs_next).numpy()
        lp = nets.log_prob(nets.actor_probs(R, traj.obs), traj.act).numpy()
    returns, adv = ppo.compute_returns(v_s, v_n, traj.rew, traj.done, traj.unfinished, rms, 0.95, 0.95)

    def learn_no_norm():
        # oracle ppo_learn normalises advantages unconditionally; emulate norm_adv=0 / value_clip=0 / no clipping here
        outl = {"loss": [], "loss/clip": [], "loss/vf": [], "loss/ent": []}
        opt = ppo.AdamDup()
        plist = ppo.rl_param_list(R)
        uniq = list({id(p): p for p in plist}.values())
        for p in uniq:
            p.requires_grad_(True)
        act_t = torch.as_tensor(traj.act, dtype=torch.long)
        for perm in perms:
            for ch in ppo.split_indices(n, c["batch_size"], np.asarray(perm)):
                i_t = torch.as_tensor(ch, dtype=torch.long)
                s = traj.obs[i_t]
                p = nets.actor_probs(R, s)
                a = torch.as_tensor(adv)[i_t]
                ratio = (nets.log_prob(p, act_t[i_t]) - torch.as_tensor(lp)[i_t]).exp()
                clip_loss = -torch.min(ratio * a, ratio.clamp(0.8, 1.2) * a).mean()
                vf_loss = ((torch.as_tensor(returns)[i_t] - nets.critic_value(R, s)) ** 2).mean()
                ent = nets.entropy(p).mean()
                loss = clip_loss + 0.25 * vf_loss - 0.01 * ent
                for q in uniq:
                    q.grad = None
                loss.backward()
                opt.step(plist, [q.grad for q in plist])
                for k, v in (("loss", loss), ("loss/clip", clip_loss), ("loss/vf", vf_loss), ("loss/ent", ent)):
                    outl[k].append(v.item())
        return outl

    want = learn_no_norm()
    for k in want:
        G.assert_close(out[k], want[k], 1e-5, 1e-5, what=k)
    sd = pol.state_dict()
    a_sd, c_sd = {k[6:]: v for k, v in sd.items() if k.startswith("actor.")}, \
        {k[7:]: v for k, v in sd.items() if k.startswith("critic.")}
    mine = nets.rl_params(a_sd, c_sd)
    for k in R:
        G.assert_close(mine[k].numpy(), R[k].detach().numpy(), 1e-5, G.PARAM_ATOL, what=f"param {k}")


def test_fused_collect_matches_generic(H):
    """Device-resident fused rollout (persistent kernel, CUDA graph, host-issued turns) == the generic loop (argmax
    actions so that all are deterministic)."""
    import cirs_codes_b200 as cb
    z, c = H.synthetic_case(U=64, I=300, B=48, T=12, N=3, thr=1, d=32)
    users = np.random.default_rng(1).integers(0, c["U"], size=c["B"])
    outs = []
    for fused, kw in ((True, dict(persistent=True)), (False, {}), (True, dict(persistent=False, use_graph=True)),
                      (True, dict(persistent=False, use_graph=False))):
        env = H.make_env(z, c)
        trk = H.make_tracker(None, c)
        with torch.no_grad():
            trk.flat.mul_(1.0)
        sd = trk.state_dict()
        g = torch.Generator().manual_seed(3)
        sd["embedding_dict.feat_user.weight"] = torch.randn(c["U"], c["d"], generator=g) * 0.1
        sd["embedding_dict.feat_item.weight"] = torch.randn(c["I"], c["d"], generator=g) * 0.1
        trk.load_state_dict(sd)
        pol = H.make_policy(None, c, None, deterministic_eval=True)
        pol.eval()
        buf = cb.VectorReplayBuffer(c["B"] * c["T"], c["B"])
        col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=fused, **kw)
        assert col.fused == fused
        res = col.collect(n_episode=c["B"], users=users)
        if fused:   # a second collect on the same objects (graph replay / re-launch) must reproduce the first
            res2 = col.collect(n_episode=c["B"], users=users)
            assert np.array_equal(res["lens"], res2["lens"]) and np.allclose(res["rews"], res2["rews"])
        idx = buf.sample_index(0)
        it = torch.as_tensor(idx, device="cuda")
        outs.append(dict(res=res, lens=buf._lengths.copy(), act=buf.act[idx].copy(), rew=buf.rew[idx].copy(),
                         done=buf.done[idx].copy(), obs=buf.obs[it].cpu().numpy(),
                         obs_next=buf.obs_next[it].cpu().numpy()))
    b = outs[1]
    for a in (outs[0], outs[2], outs[3]):
        assert np.array_equal(a["lens"], b["lens"]) and np.array_equal(a["act"], b["act"])
        assert np.array_equal(a["done"], b["done"])
        G.assert_close(a["rew"], b["rew"], 1e-6, what="rew")
        # states: FP32 rounding level -- the persistent kernel splits the tracker's matrix-vector products over warp
        # groups once few environments are left, which changes the summation order
        G.assert_close(a["obs"], b["obs"], 1e-6, 1e-6, what="obs")
        G.assert_close(a["obs_next"], b["obs_next"], 1e-6, 1e-6, what="obs_next")
        assert a["res"]["n/st"] == b["res"]["n/st"] and np.array_equal(a["res"]["lens"], b["res"]["lens"])
        G.assert_close(a["res"]["rews"], b["res"]["rews"], 1e-6, what="episode rewards")
        assert a["lens"].min() >= 1 and a["lens"].max() <= c["T"]


@pytest.mark.parametrize("force_length", [0, 6])
def test_fused_test_collectors_match_generic(H, force_length):
    """The test-time collectors of core/collector_set.py:19-33 (raw KuaishouEnv reward, remove_recommended_ids,
    force_length) on the fused path == the reference's generic loop (host-built seen bitset, core/policy/utils.py:7-27);
    argmax actions.  No item may repeat inside an episode."""
    import cirs_codes_b200 as cb
    z, c = H.synthetic_case(U=64, I=300, B=40, T=12, N=3, thr=1, d=32)
    users = np.random.default_rng(2).integers(0, c["U"], size=c["B"])
    outs = []
    for fused in (True, False):
        env = H.make_env(z, c, simulated=False)
        trk = H.make_tracker(None, c)
        sd = trk.state_dict()
        g = torch.Generator().manual_seed(3)
        sd["embedding_dict.feat_user.weight"] = torch.randn(c["U"], c["d"], generator=g) * 0.1
        sd["embedding_dict.feat_item.weight"] = torch.randn(c["I"], c["d"], generator=g) * 0.1
        trk.load_state_dict(sd)
        pol = H.make_policy(None, c, None, deterministic_eval=True)
        pol.eval()
        buf = cb.VectorReplayBuffer(c["B"] * c["T"], c["B"])
        col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=fused, remove_recommended_ids=True,
                           force_length=force_length)
        assert col.fused == fused
        res = col.collect(n_episode=c["B"], users=users)
        idx = buf.sample_index(0)
        outs.append(dict(res=res, lens=buf._lengths.copy(), act=buf.act[idx].copy(), rew=buf.rew[idx].copy(),
                         done=buf.done[idx].copy()))
    a, b = outs
    assert np.array_equal(a["lens"], b["lens"]) and np.array_equal(a["act"], b["act"])
    assert np.array_equal(a["done"], b["done"])
    G.assert_close(a["rew"], b["rew"], 1e-6, what="rew")
    if force_length:
        assert np.all(a["lens"] == force_length)
    off = np.concatenate([[0], np.cumsum(a["lens"])])
    for e in range(c["B"]):
        ep = a["act"][off[e]:off[e + 1]]
        assert len(np.unique(ep)) == len(ep), "an item was recommended twice in one episode"


def test_collector_set_keys(H):
    """CollectorSet.collect merges the three collectors' results with the reference's key prefixes."""
    import cirs_codes_b200 as cb
    z, c = H.synthetic_case(U=64, I=300, B=16, T=12, N=3, thr=1, d=32)
    trk = H.make_tracker(None, c)
    pol = H.make_policy(None, c, None)
    envs = {"FB": H.make_env(z, c, simulated=False), "NX_0": H.make_env(z, c, simulated=False),
            "NX_5": H.make_env(z, c, simulated=False)}
    cs = cb.CollectorSet(pol, envs, c["B"] * c["T"], c["B"], preprocess_fn=trk.build_state, force_length=5)
    res = cs.collect(n_episode=c["B"])
    for k in ("n/ep", "n/st", "rew", "len", "NX_0_rew", "NX_0_len", "NX_5_rew", "NX_5_len", "NX_5_lens"):
        assert k in res, k
    assert np.all(res["NX_5_lens"] == 5) and res["n/ep"] == c["B"]
    assert cs.collect_step == res["n/st"]


def test_trainer_end_to_end_and_checkpoint(H, tmp_path):
    """CIRS-RL-kuaishou.py:288-358 on the CUDA path: train collector + CollectorSet through onpolicy_trainer for two
    epochs, then the reference-layout checkpoint round trip."""
    import cirs_codes_b200 as cb
    z, c = H.synthetic_case(U=64, I=300, B=16, T=10, N=1, thr=0, d=32)
    trk = H.make_tracker(None, c)
    pol = H.make_policy(None, c, trk)
    train = cb.Collector(pol, H.make_env(z, c), cb.VectorReplayBuffer(c["B"] * c["T"], c["B"]),
                         preprocess_fn=trk.build_state)
    envs = {k: H.make_env(z, c, simulated=False) for k in ("FB", "NX_0", "NX_4")}
    test = cb.CollectorSet(pol, envs, c["B"] * c["T"], c["B"], preprocess_fn=trk.build_state, force_length=4)
    info = cb.onpolicy_trainer(pol, train, test, trk, max_epoch=2, step_per_epoch=60, repeat_per_collect=2,
                               episode_per_test=c["B"], batch_size=64, episode_per_collect=c["B"], verbose=False)
    assert info["train_step"] >= 120 and info["test_episode"] == 3 * c["B"] and np.isfinite(info["best_reward"])
    path = str(tmp_path / "ck.pt")
    cb.save_checkpoint(path, pol, trk)
    before = pol.flat.clone(), trk.flat.clone(), pol.exp_avg.clone()
    pol.flat.zero_(); trk.flat.zero_(); pol.exp_avg.zero_()
    ck = cb.load_checkpoint(path, pol, trk)
    assert set(ck) == {"policy", "optim_RL", "optim_state", "state_tracker", "ret_rms"}
    assert torch.equal(pol.flat, before[0]) and torch.equal(pol.exp_avg, before[2])
    # padding columns aside, the tracker's parameters survive the state_dict round trip
    sd = trk.state_dict()
    trk2 = H.make_tracker(None, c)
    trk2.load_state_dict(sd)
    assert torch.equal(trk2.flat, trk.flat)


def test_fused_collect_wide_catalogue_falls_back_to_ffma_head(H):
    """A catalogue wider than one 80-column slice per SM (148 x 80 = 11840 items) cannot use the tensor-core head
    phase of the persistent kernel; the launcher must fall back to the FFMA head phase and still match the generic loop."""
    import cirs_codes_b200 as cb
    z, c = H.synthetic_case(U=32, I=12100, B=8, T=5, N=2, thr=1, d=32)
    users = np.random.default_rng(5).integers(0, c["U"], size=c["B"])
    outs = []
    for fused in (True, False):
        env, trk = H.make_env(z, c), H.make_tracker(None, c)
        pol = H.make_policy(None, c, None, deterministic_eval=True)
        pol.eval()
        buf = cb.VectorReplayBuffer(c["B"] * c["T"], c["B"])
        col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, fused=fused)
        res = col.collect(n_episode=c["B"], users=users)
        idx = buf.sample_index(0)
        outs.append((buf._lengths.copy(), buf.act[idx].copy(), buf.rew[idx].copy(), res["n/st"]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    G.assert_close(outs[0][2], outs[1][2], 1e-6, what="rew")
    assert outs[0][3] == outs[1][3]


def test_checkpoint_is_reference_adam_format(H, tmp_path):
    """save_checkpoint writes torch.optim.Adam state_dicts in the reference's parameter order (CIRS-RL-kuaishou.py:
    340-358): they load into plain torch optimizers built the way the reference builds them (trunk listed twice), the
    moments equal the device-side Adam state, and a fresh policy restored from the file continues identically."""
    import cirs_codes_b200 as cb
    z = G.load("kuaishou_N5")
    c = G.cfg(z)

    def build():
        trk = H.make_tracker(z, c)
        pol = H.make_policy(z, c, trk)
        return trk, pol

    trk, pol = build()
    col, buf, res = _golden_collect(H, z, c, 0, trk, pol)
    perms = G.perms(z, 0, len(buf))
    pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=perms)
    path = str(tmp_path / "ck.pt")
    cb.save_checkpoint(path, pol, trk)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) >= {"policy", "optim_RL", "optim_state", "state_tracker"} and "ret_rms" not in ck["policy"]
    # (a) plain torch optimizers over reference-shaped modules accept the entries
    net = cb.Net(20, hidden_sizes=[64, 64])
    actor, critic = cb.Actor(net, c["I"]), cb.Critic(net)
    opt_rl = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=1e-3)
    opt_rl.load_state_dict(ck["optim_RL"])
    n_mb = len(G.perms(z, 0, len(buf))) * len(pol_split(len(buf), c["batch_size"]))
    trunk_p = next(iter(net.parameters()))
    assert float(opt_rl.state[trunk_p]["step"]) == 2 * n_mb            # the shared trunk steps twice per minibatch
    last_p = actor.last.model[0].weight
    assert float(opt_rl.state[last_p]["step"]) == n_mb
    m = pol.layout.unpack(pol.exp_avg)
    assert torch.equal(opt_rl.state[last_p]["exp_avg"], m["actor.last.weight"])
    assert torch.equal(opt_rl.state[trunk_p]["exp_avg"], m["trunk.0.weight"])
    ref_like = [torch.nn.Parameter(v.clone()) for v in trk.layout.unpack(trk.flat).values()]
    opt_tr = torch.optim.Adam(ref_like, lr=1e-3)
    opt_tr.load_state_dict(ck["optim_state"])
    assert float(opt_tr.state[ref_like[0]]["step"]) == 1               # one tracker step per update (ppo.py:235)
    assert torch.equal(opt_tr.state[ref_like[0]]["exp_avg"], trk.layout.unpack(trk.exp_avg)[
        "embedding_dict.feat_user.weight"])
    # (b) round trip into fresh objects: parameters, moments, counters, return statistics
    trk2, pol2 = build()
    cb.load_checkpoint(path, pol2, trk2)
    for a, b in ((pol.flat, pol2.flat), (pol.exp_avg, pol2.exp_avg), (pol.exp_avg_sq, pol2.exp_avg_sq),
                 (pol.opt_state, pol2.opt_state), (trk.flat, trk2.flat), (trk.exp_avg, trk2.exp_avg),
                 (trk.exp_avg_sq, trk2.exp_avg_sq), (trk.opt_state, trk2.opt_state), (pol.ret_rms.t, pol2.ret_rms.t)):
        assert torch.equal(a, b)
    # (c) an entry that is not an Adam state_dict is refused, not silently skipped
    ck["optim_RL"] = {"exp_avg": torch.zeros(3)}
    torch.save(ck, path)
    with pytest.raises(ValueError):
        cb.load_checkpoint(path, pol2, trk2)


def pol_split(n, size):
    from cirs_codes_b200.parallel import split_sizes
    return split_sizes(n, size)


@pytest.mark.parametrize("B", [16384, 3000])
def test_rollout_inverse_cdf_sampler_distribution(H, B):
    """The persistent rollout's tensor-core head samples with the two-level sampler (actor_tc_dev.cuh: inverse CDF inside
    a 40-column unit of the catalogue, exponential race across the units) instead of one race draw per item: the first
    actions of many environments that share one user (hence one state) must follow the oracle's softmax probabilities,
    be reproducible for a fixed seed / counter, and change when the counter advances.  16384 rows = 128 pipelined row
    tiles per CTA; 3000 rows = a ragged last tile."""
    import cirs_codes_b200 as cb
    from oracle import nets
    z = G.load("kuaishou_N1")
    c = G.cfg(z)
    outs = []
    for rep in range(2):
        env, trk = H.make_env(z, c, B=B), H.make_tracker(z, c, B=B)
        pol = H.make_policy(z, c, None, seed=321)
        buf = cb.VectorReplayBuffer(B * c["T"], B)
        col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, force_length=1)
        assert col.fused and col.persistent
        users = np.full(B, int(z["it0/users"][0]))
        col.collect(n_episode=B, users=users)
        first = buf.act.reshape(B, buf.sub_size)[:, 0].copy()
        col.collect(n_episode=B, users=users)          # the device-side counter advanced: different draws
        second = buf.act.reshape(B, buf.sub_size)[:, 0].copy()
        outs.append((first, second))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert not np.array_equal(outs[0][0], outs[0][1])
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    s0 = torch.tensor(z["it0/s0"][:1])
    p = nets.actor_probs(R, s0).detach().numpy()[0]
    acts = np.concatenate([outs[0][0], outs[0][1]])
    n = len(acts)
    cnt = np.bincount(acts, minlength=c["I"])
    assert cnt.sum() == n and acts.min() >= 0 and acts.max() < c["I"]
    err = np.abs(cnt / n - p)
    assert err.max() < 5 * np.sqrt(p.max() / n) + 1e-3, err.max()
    # chi-square over the well-populated items
    big = p * n >= 20
    chi2 = float((((cnt - p * n) ** 2) / (p * n))[big].sum())
    dof = int(big.sum())
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
