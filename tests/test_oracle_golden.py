"""Pin the CPU oracle (oracle/*.py) against outputs of the reference itself (tests/golden/*.npz, produced by
oracle/make_golden.py executing /root/reference) and against tianshou's own known-answer tests.

Tolerances: actions / done / lengths exact; rewards, states, probabilities, returns, losses <= 1e-5 relative
(north_star).  Quantities that are differences of O(1) numbers get an absolute floor of 1e-5 x that scale:
  * the PPO clip loss is a mean of ratio * (zero-mean, unit-std advantage): scale 1;
  * parameters after Adam: one step moves a weight by ~lr = 1e-3 regardless of gradient size, and for
    gradients within a few orders of Adam's eps (1e-8) the step is ill-conditioned in the gradient's rounding
    noise, so the floor is PARAM_ATOL = 1e-5 (1 % of one step) -- see DESIGN.md "tolerances".
"""
import numpy as np
import pytest
import torch

from oracle import env as oenv, nets, pipeline, ppo
from tests import goldutil as G


# ------------------------------------------------------------------ GAE known answers
def _ret(done, rew, v=None, gamma=0.1, lam=1.0):
    # tianshou/test/base/test_returns.py:21-72: v is passed as v_s_, v_s = roll(v_s_, 1); the buffer's
    # unfinished_index is the last transition when it is not done.
    done = np.array(done, dtype=bool)
    rew = np.array(rew, dtype=np.float64)
    v_ = np.zeros_like(rew) if v is None else np.array(v, dtype=np.float64)
    unf = np.zeros_like(done)
    unf[-1] = not done[-1]
    r, _ = ppo.episodic_return(rew, done, unf, v_, np.roll(v_ * (~done), 1) if v is None else np.roll(v_, 1),
                               gamma, lam)
    return r


def test_gae_known_answers_tianshou():
    # test_returns.py:36-46
    assert np.allclose(_ret([0, 1, 0, 1, 0, 1, 0.], [7, 6, 1, 2, 3, 4, 5.]), [7.6, 6, 1.2, 2, 3.4, 4, 5])
    # test_returns.py:47-57
    assert np.allclose(_ret([0, 1, 0, 1, 0, 0, 1.], [7, 6, 1, 2, 3, 4, 5.]), [7.6, 6, 1.2, 2, 3.45, 4.5, 5])
    # test_returns.py:58-72 (gamma .99, lambda .95, explicit values)
    done = [0, 0, 0, 1., 0, 0, 0, 1, 0, 0, 0, 1]
    rew = [101, 102, 103., 200, 104, 105, 106, 201, 107, 108, 109, 202]
    v = [2., 3., 4, -1, 5., 6., 7, -2, 8., 9., 10, -3]
    truth = [454.8344, 376.1143, 291.298, 200., 464.5610, 383.1085, 295.387, 201., 474.2876, 390.1027, 299.476, 202.]
    assert np.allclose(_ret(done, rew, v, gamma=0.99, lam=0.95), truth)


def test_running_mean_std_matches_numpy():
    rng = np.random.default_rng(0)
    xs = [rng.normal(3, 2, size=n) for n in (5, 17, 1, 40)]
    r = ppo.RunningMeanStd()
    for x in xs:
        r.update(x)
    allx = np.concatenate(xs)
    # statistics.py:74-78: starts from mean 0 / var 1 with count 0, which the first merge erases
    assert np.isclose(r.mean, allx.mean()) and np.isclose(r.var, allx.var()) and r.count == len(allx)


def test_exit_window_quirk():
    # N=5 -> window item indices for t=1..8
    want = {1: [0], 2: [0, 1], 3: [1, 2], 4: [3], 5: [0, 1, 2, 3, 4], 6: [1, 2, 3, 4, 5], 7: [2, 3, 4, 5, 6],
            8: [3, 4, 5, 6, 7]}
    for t, w in want.items():
        assert list(range(t))[t - 5:t] == w


# ------------------------------------------------------------------ replay of the reference's runs
def _make(z):
    c = G.cfg(z)
    env = oenv.KuaishouSimOracle(z["mat"], z["normed_mat"], z["cats"], z["alpha_u"] if c["use_ab"] else None,
                                 z["beta_i"] if c["use_ab"] else None, max_turn=c["T"], num_leave_compute=c["N"],
                                 leave_threshold=c["thr"], tau=c["tau"], gamma_exposure=c["gamma_exposure"],
                                 r_decay=c["r_decay"], version=c["version"])
    P = nets.to_params(z, "init/tracker/")
    P = {k: v.clone().requires_grad_(k != "pos_encoder.pe") for k, v in P.items()}
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    R = {k: v.clone() for k, v in R.items()}
    return c, env, P, R


@pytest.mark.parametrize("name", G.KUAISHOU_CASES)
def test_replay_reference_run(name):
    z = G.load(name)
    c, env, P, R = _make(z)
    tracker = nets.TrackerOracle(P, c["nhead"], c["T"])
    opt_rl, opt_tr, rms = ppo.AdamDup(), ppo.AdamDup(), ppo.RunningMeanStd()
    tparams = [v for k, v in P.items() if k != "pos_encoder.pe"]
    for it in range(c["iters"]):
        gt = G.turns(z, it)
        rec = []
        traj, res = pipeline.collect(env, tracker, R, z[f"it{it}/users"], actions=[t["obs_next_raw"] for t in gt],
                                     record=rec)
        assert len(rec) == len(gt)
        for t, (mine, ref) in enumerate(zip(rec, gt)):
            assert np.array_equal(mine["env_id"], ref["env_id"]), f"ready set, turn {t}"
            assert np.array_equal(mine["done"], ref["done"]), f"done, turn {t}"
            G.assert_close(mine["rew"], ref["rew"], 1e-5, what=f"rew turn {t}")
            G.assert_close(mine["state"], ref["state"], 1e-5, 1e-6, what=f"state turn {t}")
            G.assert_close(mine["state_next"], ref["state_next"], 1e-5, 1e-6, what=f"state_next turn {t}")
            G.assert_close(mine["probs"], ref["probs"], 1e-5, what=f"probs turn {t}")
            # the sampler: exponential race on the reference's own probabilities and noise
            assert np.array_equal(nets.sample_race(ref["probs"], ref["q"]).numpy(), ref["obs_next_raw"][:, 0])
        # buffer content in sample(0) order and collect statistics
        assert np.array_equal(traj.lengths, z[f"it{it}/buf/lengths"])
        assert np.array_equal(traj.act, z[f"it{it}/buf/act"])
        assert np.array_equal(traj.done, z[f"it{it}/buf/done"])
        G.assert_close(traj.rew, z[f"it{it}/buf/rew"], 1e-5, what="buf rew")
        G.assert_close(traj.obs.detach(), z[f"it{it}/buf/obs"], 1e-5, 1e-6, what="buf obs")
        G.assert_close(traj.obs_next.detach(), z[f"it{it}/buf/obs_next"], 1e-5, 1e-6, what="buf obs_next")
        assert res["n/st"] == int(z[f"it{it}/res/n_st"]) and res["n/ep"] == int(z[f"it{it}/res/n_ep"])
        assert np.array_equal(res["lens"], z[f"it{it}/res/lens"])
        G.assert_close(res["rews"], z[f"it{it}/res/rews"], 1e-5, what="episode rewards")
        # slot layout: env i owns [i*L, i*L+len_i)  (vecbuf.py:26-30)
        L = int(z[f"it{it}/buf/sub_size"])
        want = np.concatenate([np.arange(l) + i * L for i, l in enumerate(traj.lengths)])
        assert np.array_equal(want, z[f"it{it}/buf/index"])
        # update
        out = {}
        losses = pipeline.update(traj, R, opt_rl, tparams, opt_tr, rms, G.perms(z, it, len(traj.act)),
                                 c["batch_size"], out=out)
        for k in ("v_s", "returns", "adv", "logp_old"):
            G.assert_close(out[k], z[f"it{it}/upd/{k}"], 1e-5, 1e-6, what=k)
        G.assert_close(losses["loss/clip"], z[f"it{it}/upd/loss_clip"], 1e-5, 1e-5, what="clip loss")
        G.assert_close(losses["loss/vf"], z[f"it{it}/upd/loss_vf"], 1e-5, what="vf loss")
        G.assert_close(losses["loss/ent"], z[f"it{it}/upd/loss_ent"], 1e-5, what="entropy")
        G.assert_close(losses["loss"], z[f"it{it}/upd/loss"], 1e-5, 1e-5, what="loss")
        G.assert_close([rms.mean, rms.var, rms.count], z[f"it{it}/upd/ret_rms"], 1e-6, what="ret_rms")
        # parameters after the update (Adam with the duplicated trunk; tracker stepped once)
        after_a = nets.to_params(z, f"it{it}/after/actor/")
        after_c = nets.to_params(z, f"it{it}/after/critic/")
        Rref = nets.rl_params(after_a, after_c)
        for k in R:
            G.assert_close(R[k].detach(), Rref[k], 1e-5, G.PARAM_ATOL, what=f"RL param {k}")
        Pref = nets.to_params(z, f"it{it}/after/tracker/")
        for k in P:
            mine, ref = P[k].detach().numpy(), Pref[k].numpy()
            if k.endswith("in_proj_bias"):
                # the key bias has an identically-zero gradient (softmax is invariant to a per-query shift);
                # Adam turns the reference's rounding noise there into +-lr steps, so it cannot be compared
                mine, ref = G.drop_key_bias(mine), G.drop_key_bias(ref)
            G.assert_close(mine, ref, 1e-5, G.PARAM_ATOL, what=f"tracker param {k}")


# ------------------------------------------------------------------ VirtualTaobao (continuous actions, dense features)
def _make_taobao(z):
    c = G.taobao_cfg(z)
    UM = nets.to_params(z, "usermodel/")
    env = oenv.TaobaoSimOracle(lambda x: nets.mmoe_forward(UM, x).reshape(-1).numpy(), max_turn=c["T"],
                               num_leave_compute=c["N"], leave_threshold=c["thr"], tau=c["tau"],
                               gamma_exposure=c["gamma_exposure"], version=c["version"])
    P = nets.to_params(z, "init/tracker/")
    P = {k: v.clone().requires_grad_(k != "pos_encoder.pe") for k, v in P.items()}
    R = nets.rl_params_continuous(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    R = {k: v.clone() for k, v in R.items()}
    return c, UM, env, P, R


@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_taobao_user_model_known_answers(name):
    z = G.load(name)
    UM = nets.to_params(z, "usermodel/")
    G.assert_close(nets.mmoe_forward(UM, z["usermodel_x"]).numpy(), z["usermodel_y"], 1e-5, 1e-6, what="MMOE forward")


@pytest.mark.parametrize("name", G.TAOBAO_CASES)
def test_replay_reference_run_taobao(name):
    z = G.load(name)
    c, UM, env, P, R = _make_taobao(z)
    space = (z["action_low"], z["action_high"])
    tracker = nets.TrackerOracle(P, c["nhead"], c["T"], dense=True)
    opt_rl, opt_tr, rms = ppo.AdamDup(), ppo.AdamDup(), ppo.RunningMeanStd()
    tparams = [v for k, v in P.items() if k != "pos_encoder.pe"]
    for it in range(c["iters"]):
        gt = G.turns(z, it)
        rec = []
        noise = lambda turn, n, A: gt[turn]["eps"]  # noqa: E731  the reference's own N(0,1) draws
        traj, res = pipeline.collect(env, tracker, R, z[f"it{it}/users"], noise=noise, record=rec,
                                     action_space=space)
        assert len(rec) == len(gt)
        for t, (mine, ref) in enumerate(zip(rec, gt)):
            assert np.array_equal(mine["env_id"], ref["env_id"]), f"ready set, turn {t}"
            assert np.array_equal(mine["done"], ref["done"]), f"done, turn {t}"
            G.assert_close(mine["rew"], ref["rew"], 1e-5, 1e-6, what=f"rew turn {t}")
            G.assert_close(mine["state"], ref["state"], 1e-5, 1e-6, what=f"state turn {t}")
            G.assert_close(mine["state_next"], ref["state_next"], 1e-5, 1e-6, what=f"state_next turn {t}")
            G.assert_close(mine["probs"][:, :27], ref["mu"], 1e-5, 1e-6, what=f"mu turn {t}")
            G.assert_close(mine["probs"][:, 27:], ref["sigma"], 1e-5, what=f"sigma turn {t}")
            # sampler and action mapping on the reference's own mu / sigma / eps: bit exact
            a_ref = nets.sample_normal(torch.tensor(ref["mu"]), torch.tensor(ref["sigma"]), ref["eps"]).numpy()
            assert np.array_equal(nets.map_action(a_ref, *space).astype(np.float64), ref["obs_next_raw"][:, :27])
            # observation layout [a(27), r, 0, t+1]  (simulated_env.py:50)
            assert np.array_equal(ref["obs_next_raw"][:, 28], np.zeros(len(ref["rew"])))
            assert np.array_equal(ref["obs_next_raw"][:, 29], np.full(len(ref["rew"]), t + 1.0))
        assert np.array_equal(traj.lengths, z[f"it{it}/buf/lengths"])
        assert np.array_equal(traj.done, z[f"it{it}/buf/done"])
        G.assert_close(traj.act, z[f"it{it}/buf/act"], 1e-5, 1e-6, what="buf act")
        G.assert_close(traj.rew, z[f"it{it}/buf/rew"], 1e-5, 1e-6, what="buf rew")
        G.assert_close(traj.obs.detach(), z[f"it{it}/buf/obs"], 1e-5, 1e-6, what="buf obs")
        G.assert_close(traj.obs_next.detach(), z[f"it{it}/buf/obs_next"], 1e-5, 1e-6, what="buf obs_next")
        assert res["n/st"] == int(z[f"it{it}/res/n_st"]) and res["n/ep"] == int(z[f"it{it}/res/n_ep"])
        assert np.array_equal(res["lens"], z[f"it{it}/res/lens"])
        G.assert_close(res["rews"], z[f"it{it}/res/rews"], 1e-5, 1e-6, what="episode rewards")
        out = {}
        losses = pipeline.update(traj, R, opt_rl, tparams, opt_tr, rms, G.perms(z, it, len(traj.act)),
                                 c["batch_size"], out=out)
        for k in ("v_s", "returns", "adv", "logp_old"):
            G.assert_close(out[k], z[f"it{it}/upd/{k}"], 1e-5, 1e-5, what=k)
        G.assert_close(losses["loss/clip"], z[f"it{it}/upd/loss_clip"], 1e-5, 1e-5, what="clip loss")
        G.assert_close(losses["loss/vf"], z[f"it{it}/upd/loss_vf"], 1e-5, 1e-6, what="vf loss")
        G.assert_close(losses["loss/ent"], z[f"it{it}/upd/loss_ent"], 1e-5, what="entropy")
        G.assert_close(losses["loss"], z[f"it{it}/upd/loss"], 1e-5, 1e-5, what="loss")
        G.assert_close([rms.mean, rms.var, rms.count], z[f"it{it}/upd/ret_rms"], 1e-6, what="ret_rms")
        Rref = nets.rl_params_continuous(nets.to_params(z, f"it{it}/after/actor/"),
                                         nets.to_params(z, f"it{it}/after/critic/"))
        for k in R:
            G.assert_close(R[k].detach(), Rref[k], 1e-5, G.PARAM_ATOL, what=f"RL param {k}")
        Pref = nets.to_params(z, f"it{it}/after/tracker/")
        for k in P:
            mine, ref = P[k].detach().numpy(), Pref[k].numpy()
            if k.endswith("in_proj_bias"):
                mine, ref = G.drop_key_bias(mine), G.drop_key_bias(ref)
            G.assert_close(mine, ref, 1e-5, G.PARAM_ATOL, what=f"tracker param {k}")


def test_user_model_oracle_vs_golden():
    """oracle/user_model.py against the reference's own UserModel_Pairwise.forward / KuaishouEnv.compute_normed_reward
    outputs (tests/golden/user_model_deepfm.npz, recorded by oracle/make_golden_user_model.py)."""
    import os
    from oracle import user_model as um
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "user_model_deepfm.npz"))
    P = {k[3:]: g[k] for k in g.files if k.startswith("sd.")}
    pm = um.predict_mat(P, g["users"], g["items"], g["item_feat"], g["item_dense"])
    np.testing.assert_allclose(pm[g["raw_rows"]], g["raw_pred"], rtol=1e-5, atol=1e-5)
    nm = um.compute_normed_reward(P, g["users"], g["items"], g["item_feat"], g["item_dense"])
    assert nm.min() == 0.0 and nm.max() == 1.0
    np.testing.assert_allclose(nm, g["normed_mat"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ label encoders, test collectors (kuaishou_testcol)
def _testcol():
    z = G.load("kuaishou_testcol")
    U, I, B, T, N, thr, d, nhead, force_length, seed = (int(x) for x in z["cfg"])
    tau, gamma_e, r_decay = (float(x) for x in z["cfg_f"])
    return z, dict(U=U, I=I, B=B, T=T, N=N, thr=thr, d=d, nhead=nhead, force_length=force_length, seed=seed, tau=tau,
                   gamma_exposure=gamma_e, r_decay=r_decay)


def test_oracle_raw_id_alpha_beta_vs_golden():
    """SimulatedEnv whose alpha_u / beta_i are indexed by RAW id through NON-identity label encoders
    (simulated_env.py:157-161), teacher-forced with the reference's recorded actions."""
    z, c = _testcol()
    env = oenv.KuaishouSimOracle(z["mat"], z["normed_mat"], z["cats"], z["alpha_raw"], z["beta_raw"],
                                 max_turn=c["T"], num_leave_compute=c["N"], leave_threshold=c["thr"], tau=c["tau"],
                                 gamma_exposure=c["gamma_exposure"], r_decay=c["r_decay"], version="v1",
                                 raw_user=z["raw_user"], raw_item=z["raw_item"])
    assert np.isnan(z["alpha_raw"]).any(), "the raw-id tables must have holes an encoded index would hit"
    for ep in range(int(z["sim/n_ep"])):
        env.reset(z[f"sim/ep{ep}/user"])
        for t, a in enumerate(z[f"sim/ep{ep}/act"]):
            _, rew, done = env.step([a])
            assert bool(done[0]) == bool(z[f"sim/ep{ep}/done"][t])
            G.assert_close(rew[0], z[f"sim/ep{ep}/rew"][t], 1e-9, what=f"ep{ep} turn {t}")


@pytest.mark.parametrize("cname", ["FB", "NX_0", "NX_5"])
def test_oracle_test_collectors_vs_golden(cname):
    """core/collector_set.py:13-77 on raw KuaishouEnv: free browsing, remove_recommended_ids, + force_length, replayed
    with the reference's recorded race noise: actions / done / lengths exact, rewards and states 1e-5."""
    z, c = _testcol()
    env = oenv.KuaishouSimOracle(z["mat"], None, z["cats"], max_turn=c["T"], num_leave_compute=c["N"],
                                 leave_threshold=c["thr"], simulated=False)
    P = nets.to_params(z, "init/tracker/")
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    tracker = nets.TrackerOracle(P, c["nhead"], c["T"], keep_graph=False)
    n_turns = int(z[f"{cname}/n_turns"])
    gt = [{k: z[f"{cname}/turn{t}/{k}"] for k in ("state", "q", "act", "env_id", "rew", "done", "state_next",
                                                    "n_masked")} for t in range(n_turns)]
    rec = []
    traj, res = pipeline.collect(env, tracker, R, z[f"{cname}/users"], noise=lambda t, n, A: gt[t]["q"],
                                 force_length=c["force_length"] if cname == "NX_5" else 0, record=rec,
                                 remove_recommended=cname != "FB")
    assert len(rec) == n_turns
    for t, (mine, ref) in enumerate(zip(rec, gt)):
        assert np.array_equal(mine["env_id"], ref["env_id"]), f"ready set, turn {t}"
        assert np.array_equal(mine["act"], ref["act"]), f"actions, turn {t}"
        assert np.array_equal(mine["done"], ref["done"]), f"done, turn {t}"
        assert int(ref["n_masked"]) == (t if cname != "FB" else 0)
        G.assert_close(mine["rew"], ref["rew"], 1e-6, what=f"rew turn {t}")
        G.assert_close(mine["state"], ref["state"], 1e-5, 1e-6, what=f"state turn {t}")
    assert np.array_equal(traj.lengths, z[f"{cname}/buf/lengths"])
    assert np.array_equal(traj.act, z[f"{cname}/buf/act"])
    assert np.array_equal(res["lens"], z[f"{cname}/res/lens"])
    G.assert_close(res["rews"], z[f"{cname}/res/rews"], 1e-6, what="episode rewards")
    if cname != "FB":
        off = np.concatenate([[0], np.cumsum(traj.lengths)])
        assert all(len(set(traj.act[off[e]:off[e + 1]])) == off[e + 1] - off[e] for e in range(c["B"]))


# ------------------------------------------------------------------ raw VirtualTB: generator and click model
def _sub(z, prefix):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def test_oracle_virtualtb_generator_and_click_model_vs_golden():
    """UserModel.generate and ActionModel.predict with the shipped weights and the reference's recorded seeds / race
    noise (oracle/make_golden_extra.py usergen): the generated one-hot users and the (a, b) click results are exact."""
    z = G.load("taobao_usergen")
    o = oenv.VirtualTBOracle(_sub(z, "generator/"), _sub(z, "action/"))
    users = o.generate(z["gen/z"], z["gen/q"])
    assert np.array_equal(users, z["gen/user"]) and np.all(users.sum(1) == 11)
    got = o.click(z["click/user"], z["click/page"], z["click/act"], z["click/q"])
    assert np.array_equal(got, z["click/result"])
