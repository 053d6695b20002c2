"""GPU parity of the test-time rollout and of the label-encoder boundary against
tests/golden/kuaishou_testcol.npz, which was recorded from the reference's own CollectorSet / Callback_Coverage_Count /
SimulatedEnv (oracle/make_golden_extra.py):
  * SimulatedEnv whose alpha_u / beta_i are indexed by RAW ids through non-identity label encoders, built through the
    drop-in ``register / make / DummyVectorEnv`` calls of CIRS-RL-kuaishou.py:173-221;
  * the three collectors FB / NX_0 / NX_x with the reference's recorded race noise: actions, done, lengths exact;
  * the fused persistent rollout of the masked collectors (its own Philox noise) replayed through the CPU oracle;
  * coverage / dominated-category metrics as a device reduction, both branches of get_feat_dominate_dict."""
import numpy as np
import pytest
import torch

from tests import goldutil as G

pytestmark = pytest.mark.gpu


class _Lbe:
    """sklearn.preprocessing.LabelEncoder stand-in: classes_ + inverse_transform."""

    def __init__(self, classes):
        self.classes_ = np.asarray(classes)

    def inverse_transform(self, y):
        return self.classes_[np.asarray(y, dtype=np.int64)]


def _case():
    z = G.load("kuaishou_testcol")
    U, I, B, T, N, thr, d, nhead, force_length, seed = (int(x) for x in z["cfg"])
    tau, gamma_e, r_decay = (float(x) for x in z["cfg_f"])
    c = dict(U=U, I=I, B=B, T=T, N=N, thr=thr, d=d, nhead=nhead, force_length=force_length, seed=seed, tau=tau,
             gamma_exposure=gamma_e, r_decay=r_decay, version="v1", use_ab=True, batch_size=16, repeat=2)
    raw_item = z["raw_item"]
    list_feat = [[] for _ in range(int(raw_item.max()) + 1)]          # indexed by RAW item id, like the reference's
    for j, r in enumerate(raw_item):
        list_feat[r] = [int(x) for x in z["cats"][j] if x > 0]
    return z, c, list_feat, _Lbe(z["raw_user"]), _Lbe(raw_item)


def _register(z, c, list_feat, lbe_user, lbe_photo):
    from cirs_codes_b200 import env as E
    E.register(id="KuaishouEnv-v0", entry_point="environments.KuaishouRec.env.kuaishouEnv:KuaishouEnv",
               kwargs=dict(mat=z["mat"], lbe_user=lbe_user, lbe_photo=lbe_photo, num_leave_compute=c["N"],
                           leave_threshold=c["thr"], max_turn=c["T"], list_feat=list_feat, df_photo_env=None,
                           df_dist_small=None))
    E.register(id="SimulatedEnv-v0", entry_point="core.env.simulatedEnv.simulated_env:SimulatedEnv",
               kwargs=dict(user_model=None, task_name="KuaishouEnv-v0", version="v1", tau=c["tau"],
                           alpha_u=z["alpha_raw"], beta_i=z["beta_raw"], normed_mat=z["normed_mat"],
                           gamma_exposure=c["gamma_exposure"], r_decay=c["r_decay"]))
    return E


def test_raw_id_alpha_beta_through_drop_in_construction():
    z, c, list_feat, lbe_user, lbe_photo = _case()
    E = _register(z, c, list_feat, lbe_user, lbe_photo)
    spec = E.make("KuaishouEnv-v0")
    assert spec.mat.shape == (c["U"], c["I"]) and spec.lbe_photo is lbe_photo
    envs = E.DummyVectorEnv([lambda: E.make("SimulatedEnv-v0") for _ in range(1)])
    assert envs.mat[0].shape[1] == c["I"] and len(envs.mat) == 1 and envs.lbe_photo is lbe_photo
    for ep in range(int(z["sim/n_ep"])):
        obs = envs.reset(users=z[f"sim/ep{ep}/user"])
        assert obs.reshape(-1)[0] == z[f"sim/ep{ep}/user"][0]
        for t, a in enumerate(z[f"sim/ep{ep}/act"]):
            _, rew, done, _ = envs.step([a], [0])
            assert bool(done[0]) == bool(z[f"sim/ep{ep}/done"][t])
            G.assert_close(rew[0], z[f"sim/ep{ep}/rew"][t], 1e-5, what=f"ep{ep} turn {t}")


def _objects(H, z, c, list_feat, lbe_user, lbe_photo, B, **kw):
    import cirs_codes_b200 as cb
    env = cb.KuaishouVectorEnv(B, z["mat"], list_feat, simulated=False, max_turn=c["T"], num_leave_compute=c["N"],
                               leave_threshold=c["thr"], lbe_user=lbe_user, lbe_photo=lbe_photo)
    trk = H.make_tracker(z, c, B=B)
    pol = H.make_policy(z, c, None)
    pol.eval()
    buf = cb.VectorReplayBuffer(B * (c["T"] + 2), B)
    col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state, **kw)
    return env, trk, pol, buf, col


@pytest.mark.parametrize("cname", ["FB", "NX_0", "NX_5"])
def test_test_collectors_vs_golden(cname):
    """The reference's loop with the reference's recorded race noise (generic path: host-built seen bitset,
    cirs_actor_sample with the mask): every action, done flag, episode length and reward of the golden run."""
    from tests import gpu_harness as H
    z, c, list_feat, lbe_user, lbe_photo = _case()
    n_turns = int(z[f"{cname}/n_turns"])
    q = [z[f"{cname}/turn{t}/q"] for t in range(n_turns)]
    env, trk, pol, buf, col = _objects(H, z, c, list_feat, lbe_user, lbe_photo, c["B"], fused=False,
                                       remove_recommended_ids=cname != "FB",
                                       force_length=c["force_length"] if cname == "NX_5" else 0)
    res = col.collect(n_episode=c["B"], users=z[f"{cname}/users"], noise_fn=lambda t, n: q[t])
    idx = buf.sample_index(0)
    assert np.array_equal(buf._lengths, z[f"{cname}/buf/lengths"])
    assert np.array_equal(buf.act[idx], z[f"{cname}/buf/act"])
    assert np.array_equal(buf.done[idx], z[f"{cname}/buf/done"])
    G.assert_close(buf.rew[idx], z[f"{cname}/buf/rew"], 1e-5, what="rewards")
    assert np.array_equal(res["lens"], z[f"{cname}/res/lens"])
    G.assert_close(res["rews"], z[f"{cname}/res/rews"], 1e-5, what="episode rewards")


@pytest.mark.parametrize("force_length", [0, 5])
def test_fused_masked_collector_vs_oracle(force_length):
    """The persistent rollout kernel with remove_recommended_ids (seen bitset masked inside the tensor-core / FFMA head,
    own Philox noise), replayed through the CPU oracle's masked collect with the kernel's actions: no action may repeat
    inside an episode (the oracle asserts it against ITS recommended sets), done / lengths exact, rewards and states 1e-5."""
    from oracle import env as oenv, nets, pipeline
    from tests import gpu_harness as H
    z, c, list_feat, lbe_user, lbe_photo = _case()
    B = 64
    users = np.random.default_rng(4).integers(0, c["U"], size=B)
    env, trk, pol, buf, col = _objects(H, z, c, list_feat, lbe_user, lbe_photo, B, remove_recommended_ids=True,
                                       force_length=force_length)
    pol.train()          # sample (deterministic_eval is off anyway): exercises the race under the mask
    assert col.fused and col.persistent
    res = col.collect(n_episode=B, users=users)
    L, lens = buf.sub_size, buf._lengths.copy()
    acts = buf.act.reshape(B, L)
    o_env = oenv.KuaishouSimOracle(z["mat"], None, z["cats"], max_turn=c["T"], num_leave_compute=c["N"],
                                   leave_threshold=c["thr"], simulated=False)
    P = nets.to_params(z, "init/tracker/")
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    o_trk = nets.TrackerOracle(P, c["nhead"], c["T"], keep_graph=False)
    actions, ready = [], np.arange(B)
    for t in range(int(lens.max())):
        actions.append(acts[ready, t])
        ready = ready[lens[ready] > t + 1]
    traj, ores = pipeline.collect(o_env, o_trk, R, users, actions=actions, force_length=force_length,
                                  remove_recommended=True)
    assert np.array_equal(traj.lengths, lens)
    if force_length:
        assert np.all(lens == force_length)
    idx = buf.sample_index(0)
    it = torch.as_tensor(idx, device="cuda")
    assert np.array_equal(traj.done, buf.done[idx])
    G.assert_close(buf.rew[idx], traj.rew, 1e-5, what="rewards")
    G.assert_close(buf.obs[it].cpu().numpy(), traj.obs.numpy(), 1e-5, 1e-6, what="states")
    assert res["n/st"] == ores["n/st"]
    # the sampler really samples under the mask: across 64 environments the first two actions are not all the argmax
    assert len(np.unique(acts[:, 0])) > 4


def test_coverage_metrics_on_device_vs_golden():
    """Callback_Coverage_Count (evaluation.py:286-371) on buffers the fused collectors filled: CV / CV_turn / ifeat_*
    from one device reduction must equal what the reference's callback printed for the same actions -- here the golden
    run's actions are replayed (recorded noise, generic path), uploaded, and reduced on the device."""
    import pandas as pd
    from cirs_codes_b200.evaluation import Callback_Coverage_Count
    from tests import gpu_harness as H
    z, c, list_feat, lbe_user, lbe_photo = _case()

    class _Set:
        pass

    cset = _Set()
    cset.collector_dict = {}
    for cname in ("FB", "NX_0", "NX_5"):
        n_turns = int(z[f"{cname}/n_turns"])
        q = [z[f"{cname}/turn{t}/q"] for t in range(n_turns)]
        env, trk, pol, buf, col = _objects(H, z, c, list_feat, lbe_user, lbe_photo, c["B"], fused=False,
                                           remove_recommended_ids=cname != "FB",
                                           force_length=c["force_length"] if cname == "NX_5" else 0)
        col.collect(n_episode=c["B"], users=z[f"{cname}/users"], noise_fn=lambda t, n: q[t])
        buf.sync_device()
        buf.plan_device()                 # the state a fused collect leaves behind: device arrays + device plan
        cset.collector_dict[cname] = col
        cset.env = env
    df = pd.DataFrame({f"feat{k}": z["cats"][:, k] for k in range(4)}, index=z["raw_item"])
    dom_feat = {"feat": [(int(v), int(n)) for v, n in z["dom_feat"]]}
    got = Callback_Coverage_Count(cset, df, True, dom_feat, lbe_photo, 0.6).on_epoch_end(1, {})
    for k in [f for f in z.files if f.startswith("cov_feat/")]:
        assert abs(got[k[len("cov_feat/"):]] - float(z[k])) < 1e-12, k
    dom_each = {k[len("dom_each/"):]: [(int(v), int(n)) for v, n in z[k]] for k in z.files if k.startswith("dom_each/")}
    cb2 = Callback_Coverage_Count(cset, df, True, dom_each, lbe_photo, 0.5)
    got2 = cb2.on_epoch_end(1, {})
    for k in [f for f in z.files if f.startswith("cov_each/")]:
        assert abs(got2[k[len("cov_each/"):]] - float(z[k])) < 1e-12, k
    assert cb2._dev, "the device reduction (cirs_coverage_count) was not used"
