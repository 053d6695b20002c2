"""CPU restatement of the reference's rollout driver and update pipeline (the control flow around the kernels).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Parity pinned: tests/test_oracle_golden.py replays tests/golden/*.npz (reference outputs) through these functions.

  collect   core/collector.py:147-367 with n_episode == env_num: reset everything, loop
            policy -> env.step -> tracker.build_state -> buffer.add, drop finished envs from the ready set
            (no reset on done, :294-311), stop when every env has finished one episode.
  update    tianshou/policy/base.py:219-244 (update) -> core/policy/ppo.py:96-109 (process_fn) ->
            a2c.py:80-109 (_compute_returns) -> ppo.py:166-246 (learn)
Transitions are kept per environment and flattened env-major, which is the order VectorReplayBuffer.sample(0)
returns them in (tianshou/data/buffer/manager.py:144-169; env i owns slots [i*L, (i+1)*L)).
"""
import numpy as np
import torch

from . import nets, ppo


class Trajectory:
    """Flat env-major view of one collect."""

    def __init__(self, obs, obs_next, act, rew, done, lengths):
        self.obs, self.obs_next, self.act, self.rew, self.done, self.lengths = obs, obs_next, act, rew, done, lengths

    @property
    def unfinished(self):
        last = np.cumsum(self.lengths) - 1
        u = np.zeros(len(self.act), dtype=bool)
        u[last] = ~self.done[last]
        return u


def collect(env, tracker, R, users, *, actions=None, noise=None, force_length=0, record=None, action_space=None,
            remove_recommended=False):
    """One Collector.collect(n_episode=B).  Actions come from, in order of preference: ``actions`` (teacher
    forcing: list per turn of arrays aligned with the ready set), ``noise`` (callable(turn, n, A) -> q, Exp(1)
    race noise) or argmax of the probabilities.  Returns (Trajectory, result-dict like collector.py:352-362).
    ``remove_recommended`` (the NX_* test collectors, core/collector_set.py:19): the items already recommended in the
    running episode (get_recommended_ids, core/policy/utils.py:7-27) are removed from the actor's probabilities
    before Categorical renormalises and samples (removed_recommended_id_from_embedding :30-58, ppo.py:134-160);
    ``noise`` stays indexed by ORIGINAL catalogue column (the masked columns' draws are unused)."""
    B = len(users)
    obs0 = env.reset(users)
    tracker.reset(B)
    ready = np.arange(B)
    s = tracker.first(obs0, ready)  # [B,S]
    per_env = [dict(obs=[], obs_next=[], act=[], rew=[], done=[]) for _ in range(B)]
    ep_rews, ep_lens, order = [], [], []
    turn = 0
    continuous = "actor.sigma_param" in R
    while True:
        if continuous:
            # ActorProb + Independent(Normal) (ppo.py:144-156); ``noise(turn, n, 27)`` -> N(0,1) draws; the buffer keeps
            # the raw sample, the environment receives map_action(act) (collector.py:246-250, base.py:143-173)
            with torch.no_grad():
                mu, sigma = nets.actor_mu_sigma(R, s.detach())
                p = torch.cat([mu, sigma], -1)
            if actions is not None:
                act = np.asarray(actions[turn], dtype=np.float32)
            elif noise is not None:
                act = nets.sample_normal(mu, sigma, noise(turn, len(ready), mu.shape[1])).numpy()
            else:
                act = mu.numpy()
            obs_next_raw, rew, done = env.step(nets.map_action(act, *action_space), ready)
        else:
            with torch.no_grad():
                p = nets.actor_probs(R, s.detach())
            if remove_recommended:
                keep = torch.ones_like(p, dtype=torch.bool)
                for k, e in enumerate(ready):
                    if per_env[e]["act"]:
                        keep[k, torch.as_tensor(np.asarray(per_env[e]["act"], dtype=np.int64))] = False
                p = torch.where(keep, p, torch.zeros_like(p))    # masked_select + Categorical's renormalisation
                p = p / p.sum(-1, keepdim=True)
            if actions is not None:
                act = np.asarray(actions[turn]).reshape(-1)
                if remove_recommended:
                    assert all(int(a) not in per_env[e]["act"] for a, e in zip(act, ready)), \
                        "a forced action repeats an item of its episode"
            elif noise is not None:
                q = torch.as_tensor(noise(turn, len(ready), p.shape[1]))
                if remove_recommended:   # p = 0 on masked columns: p / q = 0 there, never the argmax
                    act = torch.argmax(p / q, -1).numpy()
                else:
                    act = nets.sample_race(p, q).numpy()
            else:
                act = torch.argmax(p, -1).numpy()
            obs_next_raw, rew, done = env.step(act, ready)
        turn += 1
        if force_length > 0:  # collector.py:253-258
            done = np.full_like(done, turn >= force_length)
        s_next = tracker.step(obs_next_raw, rew, ready)
        if record is not None:
            record.append(dict(env_id=ready.copy(), state=s.detach().numpy().copy(), probs=p.numpy().copy(),
                               act=act.copy(), rew=rew.copy(), done=done.copy(),
                               state_next=s_next.detach().numpy().copy()))
        for k, e in enumerate(ready):
            pe = per_env[e]
            pe["obs"].append(s[k]); pe["obs_next"].append(s_next[k])
            pe["act"].append(act[k]); pe["rew"].append(rew[k]); pe["done"].append(done[k])
        if done.any():
            for k in np.where(done)[0]:
                e = ready[k]
                ep_rews.append(float(np.sum(per_env[e]["rew"]))); ep_lens.append(len(per_env[e]["rew"]))
                order.append(e)
            keep = ~done
            ready, s_next = ready[keep], s_next[keep]
        s = s_next
        if len(ready) == 0:
            break
    lengths = np.array([len(pe["act"]) for pe in per_env])
    flat = lambda key: [x for pe in per_env for x in pe[key]]  # noqa: E731
    traj = Trajectory(torch.stack(flat("obs")), torch.stack(flat("obs_next")), np.array(flat("act")),
                      np.array(flat("rew"), dtype=np.float64), np.array(flat("done"), dtype=bool), lengths)
    rews, lens = np.array(ep_rews), np.array(ep_lens)
    res = {"n/ep": len(rews), "n/st": int(lengths.sum()), "rews": rews, "lens": lens,
           "rew": rews.mean(), "len": lens.mean(), "rew_std": rews.std(), "len_std": lens.std()}
    return traj, res


def update(traj, R, opt_rl, tracker_params, opt_tracker, rms, perms, batch_size, *, gamma=0.95, gae_lambda=0.95,
           eps_clip=0.2, vf_coef=0.25, ent_coef=0.0, max_grad_norm=0.5, out=None):
    """policy.update(0, buffer, batch_size=, repeat=len(perms)).  ``out`` (dict) receives the intermediate
    batch arrays (v_s, returns, adv, logp_old) for parity checks."""
    with torch.no_grad():
        v_s = nets.critic_value(R, traj.obs.detach()).numpy()
        v_next = nets.critic_value(R, traj.obs_next.detach()).numpy()
    returns, adv = ppo.compute_returns(v_s, v_next, traj.rew, traj.done, traj.unfinished, rms, gamma, gae_lambda)
    with torch.no_grad():
        if "actor.sigma_param" in R:
            logp_old = nets.normal_log_prob(*nets.actor_mu_sigma(R, traj.obs.detach()), traj.act).numpy()
        else:
            logp_old = nets.log_prob(nets.actor_probs(R, traj.obs.detach()), traj.act).numpy()
    if out is not None:
        out.update(v_s=v_s, returns=returns, adv=adv, logp_old=logp_old)
    return ppo.ppo_learn(R, opt_rl, tracker_params, opt_tracker, traj.obs, traj.act, adv, returns, v_s, logp_old,
                         perms, batch_size, eps_clip=eps_clip, vf_coef=vf_coef, ent_coef=ent_coef,
                         max_grad_norm=max_grad_norm)
