"""Parity at BASELINE.json's full size (configs[1]: 7176 users x 10728 items, 512 environments, d = 32): the fused
persistent rollout (tensor-core head, warp-group tracker) replayed through the CPU oracle with the CUDA path's own
actions (teacher forcing): episode lengths / done flags exact, rewards, states and the sampled actions' log-probs
<= 1e-5; then one PPO update on that collect against the oracle's update (losses <= 1e-5, value / return arrays)."""
import numpy as np
import pytest
import torch

from tests import goldutil as G

pytestmark = pytest.mark.gpu


def test_full_size_collect_and_update_match_oracle():
    import bench
    import cirs_codes_b200 as cb
    from oracle import env as oenv, nets, pipeline, ppo
    cfg = dict(bench.CONFIGS["configs1"])
    tb = bench.tables(cfg)
    dev = torch.device("cuda", 0)
    env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
    # "trained-like" embeddings: N(0, 0.1) instead of the N(0, 1e-4) initialisation, so states move
    sd = trk.state_dict()
    g = torch.Generator().manual_seed(7)
    sd["embedding_dict.feat_user.weight"] = torch.randn(cfg["U"], cfg["d"], generator=g) * 0.1
    sd["embedding_dict.feat_item.weight"] = torch.randn(cfg["I"], cfg["d"], generator=g) * 0.1
    trk.load_state_dict(sd)
    B, T = cfg["B"], cfg["T"]
    users = np.random.default_rng(11).integers(0, cfg["U"], size=B)
    res = col.collect(n_episode=B, users=users)
    torch.cuda.synchronize()
    L, lens = buf.sub_size, buf._lengths.copy()
    assert res["n/st"] == lens.sum() and lens.min() >= 1 and lens.max() <= T
    acts = buf.act.reshape(B, L)
    assert acts[np.arange(B), 0].min() >= 0 and acts.max() < cfg["I"]

    # ---- oracle replay with the CUDA path's actions
    R_ = bench.REF
    o_env = oenv.KuaishouSimOracle(tb["mat"], tb["normed_mat"], tb["cats"], tb["alpha_u"], tb["beta_i"], max_turn=T,
                                   num_leave_compute=cfg["N"], leave_threshold=cfg["thr"], tau=R_["tau"],
                                   gamma_exposure=R_["gamma_exposure"], r_decay=R_["r_decay"], version=R_["version"])
    P = {k: v.clone() for k, v in trk.state_dict().items()}
    psd = pol.state_dict()
    a_sd = {k[len("actor."):]: v for k, v in psd.items() if k.startswith("actor.")}
    c_sd = {k[len("critic."):]: v for k, v in psd.items() if k.startswith("critic.")}
    R = {k: v.clone() for k, v in nets.rl_params(a_sd, c_sd).items()}
    o_trk = nets.TrackerOracle(P, cfg["nhead"], T, keep_graph=False)
    actions, ready = [], np.arange(B)
    for t in range(int(lens.max())):
        actions.append(acts[ready, t])
        ready = ready[lens[ready] > t + 1]
    traj, ores = pipeline.collect(o_env, o_trk, R, users, actions=actions)
    assert np.array_equal(traj.lengths, lens)
    idx = buf.sample_index(0)
    it = torch.as_tensor(idx, device="cuda")
    assert np.array_equal(traj.done, buf.done[idx])
    G.assert_close(buf.rew[idx], traj.rew, 1e-5, what="rewards")
    G.assert_close(buf.obs[it].cpu().numpy(), traj.obs.numpy(), 1e-5, 1e-6, what="states")
    G.assert_close(buf.obs_next[it].cpu().numpy(), traj.obs_next.numpy(), 1e-5, 1e-6, what="next states")
    assert res["n/st"] == ores["n/st"]
    # log-probs the rollout kernel reported for its own samples == Categorical.log_prob under the oracle's softmax
    with torch.no_grad():
        s0 = traj.obs[np.concatenate([[0], np.cumsum(lens)[:-1]])]       # first state of every episode
        logits, _ = nets.categorical_logits(nets.actor_probs(R, s0))
    a0 = torch.as_tensor(acts[:, 0].astype(np.int64))
    want = logits.gather(1, a0[:, None]).flatten().numpy()

    # ---- one update on the device vs the oracle's update on the oracle's (matching) trajectory
    n = len(idx)
    rng = np.random.default_rng(3)
    perms = [rng.permutation(n) for _ in range(cfg["repeat"])]
    out = pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"], perms=perms)
    G.assert_close(pol.logp_old[it].cpu().numpy()[np.concatenate([[0], np.cumsum(lens)[:-1]])], want, 1e-5,
                   what="log-prob of the first actions")
    # (the tracker's own step comes after all minibatches, core/policy/ppo.py:235: the losses do not depend on it)
    o_out = pipeline.update(traj, R, ppo.AdamDup(), [], None, ppo.RunningMeanStd(), perms, cfg["batch_size"],
                            gamma=R_["gamma"],
                            gae_lambda=R_["gae_lambda"], eps_clip=R_["eps_clip"], vf_coef=R_["vf_coef"],
                            ent_coef=R_["ent_coef"], max_grad_norm=R_["max_grad_norm"])
    G.assert_close(out["loss/vf"], o_out["loss/vf"], 1e-5, what="vf loss")
    G.assert_close(out["loss/ent"], o_out["loss/ent"], 1e-5, what="entropy")
    G.assert_close(out["loss/clip"], o_out["loss/clip"], 1e-5, 1e-5, what="clip loss")
    G.assert_close(out["loss"], o_out["loss"], 1e-5, 1e-5, what="loss")
