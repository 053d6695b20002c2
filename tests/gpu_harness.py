"""Builders shared by the GPU parity tests, __graft_entry__.smoke() and bench.py: construct the CUDA-path objects
(KuaishouVectorEnv, StateTrackerTransformer, PPOPolicy, Collector) from a golden file's tables and initial
weights, or from synthetic tables."""
import numpy as np
import torch

import cirs_codes_b200 as cb
from cirs_codes_b200 import synth
from tests import goldutil as G


def make_env(z, c, B=None, simulated=True, **over):
    kw = dict(normed_mat=z["normed_mat"], alpha_u=z["alpha_u"] if c["use_ab"] else None,
              beta_i=z["beta_i"] if c["use_ab"] else None, simulated=simulated, max_turn=c["T"],
              num_leave_compute=c["N"], leave_threshold=c["thr"], tau=c["tau"], gamma_exposure=c["gamma_exposure"],
              r_decay=c["r_decay"], version=c["version"])
    kw.update(over)
    return cb.KuaishouVectorEnv(B or c["B"], z["mat"], z["cats"], **kw)


def make_tracker(z, c, B=None, prefix="init/tracker/"):
    class _Env:
        mat = np.zeros((c["U"], c["I"]), dtype=np.float32)

    cols = cb.get_dataset_columns(c["d"], envname="KuaishouEnv-v0", env=_Env)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trk = cb.StateTrackerTransformer(cols[0], cols[1], cols[2], dim_model=c["d"], dim_state=20,
                                         dim_max_batch=B or c["B"], dataset="KuaishouEnv-v0",
                                         has_user_embedding=cols[3], has_action_embedding=cols[4],
                                         has_feedback_embedding=cols[5], nhead=c["nhead"], d_hid=128, nlayers=2,
                                         dropout=0.0, device="cuda", seed=c["seed"], MAX_TURN=c["T"])
    if z is not None and prefix is not None:
        trk.load_state_dict({k[len(prefix):]: torch.tensor(np.asarray(z[k])) for k in z.files if k.startswith(prefix)})
    return trk


def make_policy(z, c, tracker, prefix="init/", **over):
    net = cb.Net(20, hidden_sizes=[64, 64])
    actor, critic = cb.Actor(net, c["I"]), cb.Critic(net)
    if z is not None:
        actor.load_state_dict({k[len(prefix + "actor/"):]: torch.tensor(np.asarray(z[k])) for k in z.files
                               if k.startswith(prefix + "actor/")})
        critic.load_state_dict({k[len(prefix + "critic/"):]: torch.tensor(np.asarray(z[k])) for k in z.files
                                if k.startswith(prefix + "critic/")})
    else:
        torch.manual_seed(c["seed"])
        cb.orthogonal_init(actor, critic)
    optim_rl = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=1e-3)
    optim = [optim_rl]
    if tracker is not None:
        optim.append(torch.optim.Adam(tracker.parameters(), lr=1e-3))
    kw = dict(discount_factor=0.95, max_grad_norm=0.5, eps_clip=0.2, vf_coef=0.25, ent_coef=0.0,
              reward_normalization=1, advantage_normalization=1, recompute_advantage=0, value_clip=1,
              gae_lambda=0.95, action_bound_method="", action_scaling=False)
    kw.update(over)
    return cb.PPOPolicy(actor, critic, optim, torch.distributions.Categorical, **kw)


# ---------------------------------------------------------------- VirtualTaobao builders
def taobao_user_model(z):
    return {k[len("usermodel/"):]: torch.tensor(np.asarray(z[k])) for k in z.files if k.startswith("usermodel/")}


def make_taobao_env(z, c, B=None, **over):
    kw = dict(max_turn=c["T"], num_leave_compute=c["N"], leave_threshold=c["thr"], tau=c["tau"],
              gamma_exposure=c["gamma_exposure"], version=c["version"])
    kw.update(over)
    return cb.TaobaoVectorEnv(B or c["B"], taobao_user_model(z), **kw)


def make_taobao_tracker(z, c, B=None, prefix="init/tracker/"):
    cols = cb.get_dataset_columns(c["d"], envname="VirtualTB-v0")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trk = cb.StateTrackerTransformer(cols[0], cols[1], cols[2], dim_model=c["d"], dim_state=20,
                                         dim_max_batch=B or c["B"], dataset="VirtualTB-v0",
                                         has_user_embedding=cols[3], has_action_embedding=cols[4],
                                         has_feedback_embedding=cols[5], nhead=c["nhead"], d_hid=128, nlayers=2,
                                         dropout=0.0, device="cuda", seed=c["seed"], MAX_TURN=c["T"])
    if z is not None and prefix is not None:
        trk.load_state_dict({k[len(prefix):]: torch.tensor(np.asarray(z[k])) for k in z.files if k.startswith(prefix)})
    return trk


def make_taobao_policy(z, c, tracker, prefix="init/", load=True, **over):
    from cirs_codes_b200.env import Box
    net = cb.Net(20, hidden_sizes=[64, 64])
    actor, critic = cb.ActorProb(net, (27,), max_action=1.0), cb.Critic(net)
    if load:
        actor.load_state_dict({k[len(prefix + "actor/"):]: torch.tensor(np.asarray(z[k])) for k in z.files
                               if k.startswith(prefix + "actor/")})
        critic.load_state_dict({k[len(prefix + "critic/"):]: torch.tensor(np.asarray(z[k])) for k in z.files
                                if k.startswith(prefix + "critic/")})
    else:
        torch.manual_seed(c["seed"])
        cb.orthogonal_init(actor, critic)
    optim = [torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=1e-3)]
    if tracker is not None:
        optim.append(torch.optim.Adam(tracker.parameters(), lr=1e-3))

    def dist(*logits):
        return torch.distributions.Independent(torch.distributions.Normal(*logits), 1)

    kw = dict(discount_factor=0.95, max_grad_norm=0.5, eps_clip=0.2, vf_coef=0.25, ent_coef=0.0,
              reward_normalization=1, advantage_normalization=1, recompute_advantage=0, value_clip=1,
              gae_lambda=0.95, action_space=Box(-1, 1, (27,)))
    kw.update(over)
    return cb.PPOPolicy(actor, critic, optim, dist, **kw)


def synthetic_case(U=64, I=300, B=16, T=10, N=3, thr=1, d=32, nhead=4, seed=5, **kw):
    tb = synth.kuaishou_tables(U, I, seed=seed)
    c = dict(U=U, I=I, B=B, T=T, N=N, thr=thr, d=d, nhead=nhead, seed=seed, tau=100.0, gamma_exposure=10.0,
             r_decay=1.0, version="v1", use_ab=True, batch_size=64, repeat=2)
    c.update(kw)

    class Z(dict):
        files = []

    z = Z(tb)
    return z, c


def smoke_check():
    """Fused rollout + one PPO update on cuda:0 for a golden case; rollout replayed through the CPU oracle with the
    same actions (teacher forcing) and compared: done / lengths exact, rewards and states <= 1e-5."""
    from oracle import env as oenv, nets, pipeline
    z = G.load("kuaishou_N5")
    c = G.cfg(z)
    env, trk = make_env(z, c), make_tracker(z, c)
    pol = make_policy(z, c, trk)
    buf = cb.VectorReplayBuffer(c["B"] * c["T"], c["B"])
    col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state)
    users = np.asarray(z["it0/users"])
    res = col.collect(n_episode=c["B"], users=users)
    torch.cuda.synchronize()
    L, lens = buf.sub_size, buf._lengths
    acts = buf.act.reshape(c["B"], L)
    # oracle replay with the CUDA path's actions
    o_env = oenv.KuaishouSimOracle(z["mat"], z["normed_mat"], z["cats"], z["alpha_u"] if c["use_ab"] else None,
                                   z["beta_i"] if c["use_ab"] else None, max_turn=c["T"], num_leave_compute=c["N"],
                                   leave_threshold=c["thr"], tau=c["tau"], gamma_exposure=c["gamma_exposure"],
                                   r_decay=c["r_decay"], version=c["version"])
    P = nets.to_params(z, "init/tracker/")
    R = nets.rl_params(nets.to_params(z, "init/actor/"), nets.to_params(z, "init/critic/"))
    o_trk = nets.TrackerOracle(P, c["nhead"], c["T"], keep_graph=False)
    actions, ready = [], np.arange(c["B"])
    for t in range(int(lens.max())):
        actions.append(acts[ready, t])
        ready = ready[lens[ready] > t + 1]
    traj, ores = pipeline.collect(o_env, o_trk, R, users, actions=actions)
    assert np.array_equal(traj.lengths, lens), (traj.lengths, lens)
    idx = buf.sample_index(0)
    assert np.array_equal(traj.done, buf.done[idx])
    G.assert_close(buf.rew[idx], traj.rew, 1e-5, what="smoke rewards")
    G.assert_close(buf.obs[torch.as_tensor(idx, device="cuda")].cpu().numpy(), traj.obs.numpy(), 1e-5, 1e-6,
                   what="smoke states")
    assert res["n/st"] == ores["n/st"]
    out = pol.update(0, buf, batch_size=c["batch_size"], repeat=1)
    assert len(out["loss"]) > 0 and all(np.isfinite(out["loss"]))
    return res, out
