"""CPU restatement of the reference's return computation and PPO update (numpy f64 + torch CPU fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Parity pinned: GAE against tianshou/test/base/test_returns.py:21-90 known answers (copied as numbers into
tests/test_oracle_golden.py) and the whole update against tests/golden/*.npz (reference outputs).

  gae_return            tianshou/policy/base.py:380-396 (_gae_return), :272-313 (compute_episodic_return)
  RunningMeanStd        tianshou/utils/statistics.py:66-95
  compute_returns       tianshou/policy/modelfree/a2c.py:80-109
  clip_grad_norm_dup / adam_step_dup   torch.nn.utils.clip_grad_norm_ and torch.optim.Adam (single-tensor CPU
                        loop) applied to a parameter LIST THAT CONTAINS THE SHARED TRUNK TWICE
                        (CIRS-RL-kuaishou.py:256-258, core/policy/ppo.py:221-226)
  ppo_learn             core/policy/ppo.py:166-246
"""
import numpy as np
import torch

from . import nets


def gae_return(v_s, v_s_, rew, end_flag, gamma, gae_lambda):
    """base.py:380-396.  All float64; reverse scan over the flat env-major array."""
    v_s, v_s_, rew = (np.asarray(x, dtype=np.float64) for x in (v_s, v_s_, rew))
    delta = rew + v_s_ * gamma - v_s
    m = (1.0 - np.asarray(end_flag, dtype=np.float64)) * (gamma * gae_lambda)
    out = np.zeros(rew.shape)
    gae = 0.0
    for i in range(len(rew) - 1, -1, -1):
        gae = delta[i] + m[i] * gae
        out[i] = gae
    return out


def episodic_return(rew, done, unfinished, v_s_, v_s, gamma, gae_lambda):
    """base.py:272-313.  ``unfinished`` marks transitions that are the last written slot of a sub-buffer whose
    episode is still running (buffer.unfinished_index()).  Returns (returns, advantage), float64."""
    v_s_ = np.asarray(v_s_, dtype=np.float64) * (~np.asarray(done, dtype=bool))  # value_mask, base.py:246-269
    end = np.asarray(done, dtype=bool) | np.asarray(unfinished, dtype=bool)
    adv = gae_return(v_s, v_s_, rew, end, gamma, gae_lambda)
    return adv + np.asarray(v_s, dtype=np.float64), adv


class RunningMeanStd:
    """statistics.py:66-95 (parallel-variance merge)."""

    def __init__(self):
        self.mean, self.var, self.count = 0.0, 1.0, 0

    def update(self, x):
        bm, bv, bc = np.mean(x), np.var(x), len(x)
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def compute_returns(v_s32, v_next32, rew, done, unfinished, rms, gamma, gae_lambda, eps=1e-8):
    """a2c.py:80-109 with reward_normalization=1.  v_* are the critic's float32 outputs.  Returns float32
    (returns_normalised, adv) as stored in the batch, and updates ``rms`` in place."""
    scale = np.sqrt(rms.var + eps)
    v_s = np.asarray(v_s32, dtype=np.float32) * scale  # f32 * np.float64 -> float64 under numpy 2
    v_s_ = np.asarray(v_next32, dtype=np.float32) * scale
    ret, adv = episodic_return(rew, done, unfinished, v_s_, v_s, gamma, gae_lambda)
    ret_n = ret / scale
    rms.update(ret)
    return ret_n.astype(np.float32), adv.astype(np.float32)


def split_indices(n, size, perm):
    """tianshou/data/batch.py:721-744 with merge_last=True."""
    merge = n % size > 0
    out = []
    for idx in range(0, n, size):
        if merge and idx + size + size >= n:
            out.append(perm[idx:])
            break
        out.append(perm[idx:idx + size])
    return out


def clip_grad_norm_dup(grads_with_dups, max_norm):
    """torch.nn.utils.clip_grad_norm_ over a list in which the trunk tensors appear twice: the norm counts
    them twice and the in-place scaling hits them twice (coef^2)."""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads_with_dups))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads_with_dups:
        g.mul_(coef)
    return float(total)


class AdamDup:
    """torch.optim.Adam (betas .9/.999, eps 1e-8, no weight decay, single-tensor CPU path) over a param list that
    may contain the same tensor more than once: each occurrence performs a full update and bumps that tensor's
    step counter."""

    def __init__(self, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.state = {}

    def step(self, params_with_dups, grads_with_dups):
        for p, g in zip(params_with_dups, grads_with_dups):
            st = self.state.setdefault(id(p), dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)))
            st["step"] += 1
            st["m"].lerp_(g, 1 - self.b1)
            st["v"].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** st["step"]
            bc2 = 1 - self.b2 ** st["step"]
            denom = (st["v"].sqrt() / (bc2 ** 0.5)).add_(self.eps)
            p.data.addcdiv_(st["m"], denom, value=-(self.lr / bc1))


RL_ORDER = ["preprocess.model.model.0.weight", "preprocess.model.model.0.bias",
            "preprocess.model.model.2.weight", "preprocess.model.model.2.bias"]


def rl_param_list(R):
    """optim_RL's list: actor.parameters() + critic.parameters() -- trunk appears twice.  ActorProb registers
    sigma_param on the module itself, so nn.Module.parameters() yields it first (continuous.py:160-173)."""
    trunk = [R[k] for k in RL_ORDER]
    if "actor.sigma_param" in R:
        return [R["actor.sigma_param"]] + trunk + [R["actor.mu.weight"], R["actor.mu.bias"]] + trunk + \
            [R["critic.last.weight"], R["critic.last.bias"]]
    return trunk + [R["actor.last.weight"], R["actor.last.bias"]] + trunk + \
        [R["critic.last.weight"], R["critic.last.bias"]]


def ppo_learn(R, opt_rl, tracker_params, opt_tracker, obs, act, adv, returns, v_old, logp_old, perms, batch_size,
              eps_clip=0.2, vf_coef=0.25, ent_coef=0.0, max_grad_norm=0.5):
    """core/policy/ppo.py:166-246.  ``obs`` [TB,S] carries the autograd graph back to ``tracker_params`` (or is a
    leaf when the tracker is not trained).  ``perms`` = one permutation of range(TB) per repeat.
    Returns dict of per-minibatch loss lists."""
    out = {"loss": [], "loss/clip": [], "loss/vf": [], "loss/ent": []}
    continuous = "actor.sigma_param" in R
    act_t = torch.as_tensor(act, dtype=torch.float32) if continuous else torch.as_tensor(act, dtype=torch.long)
    adv_t, ret_t = torch.as_tensor(adv), torch.as_tensor(returns)
    vold_t, lpo_t = torch.as_tensor(v_old), torch.as_tensor(logp_old)
    plist = rl_param_list(R)
    uniq = list({id(p): p for p in plist}.values())
    for p in uniq:
        p.requires_grad_(True)
    for perm in perms:
        for p in tracker_params:
            p.grad = None  # optim_state.zero_grad(), ppo.py:174
        for idx in split_indices(len(perm), batch_size, np.asarray(perm)):
            idx_t = torch.as_tensor(idx, dtype=torch.long)
            s = obs[idx_t]
            if continuous:
                mu, sigma = nets.actor_mu_sigma(R, s)
                logp = nets.normal_log_prob(mu, sigma, act_t[idx_t])
            else:
                p = nets.actor_probs(R, s)
                logp = nets.log_prob(p, act_t[idx_t])
            a = adv_t[idx_t]
            a = (a - a.mean()) / a.std()  # ppo.py:185-186 (unbiased std)
            ratio = (logp - lpo_t[idx_t]).exp()
            clip_loss = -torch.min(ratio * a, ratio.clamp(1 - eps_clip, 1 + eps_clip) * a).mean()
            value = nets.critic_value(R, s)
            v_clip = vold_t[idx_t] + (value - vold_t[idx_t]).clamp(-eps_clip, eps_clip)
            vf_loss = torch.max((ret_t[idx_t] - value) ** 2, (ret_t[idx_t] - v_clip) ** 2).mean()
            ent_loss = (nets.normal_entropy(sigma) if continuous else nets.entropy(p)).mean()
            loss = clip_loss + vf_coef * vf_loss - ent_coef * ent_loss
            for q in uniq:
                q.grad = None  # optim_RL.zero_grad()
            loss.backward(retain_graph=True)
            grads = [q.grad for q in plist]
            clip_grad_norm_dup(grads, max_grad_norm)
            opt_rl.step(plist, grads)
            out["loss"].append(loss.item())
            out["loss/clip"].append(clip_loss.item())
            out["loss/vf"].append(vf_loss.item())
            out["loss/ent"].append(ent_loss.item())
    tp = [p for p in tracker_params if p.grad is not None]
    if opt_tracker is not None and tp:
        opt_tracker.step(tp, [p.grad for p in tp])  # optim_state.step(), ppo.py:235
    return out
