"""CPU restatement (torch CPU fp32, autograd for gradients) of the reference's networks on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parameters are plain dicts {reference state_dict name ->
torch tensor}; nothing here instantiates torch.nn modules or imports the reference: every operation is written
out so that each CUDA kernel has a line-by-line checker.
Parity pinned: tests/test_oracle_golden.py checks these functions against tests/golden/*.npz (reference outputs).

  StateTracker  core/state_tracker.py:170-250 (+ PositionalEncoding :255-279); torch.nn.TransformerEncoderLayer
                semantics (post-norm, ReLU, batch_first=False, eps 1e-5) restated from SURVEY §10.3 / §9-A6
  heads         tianshou/utils/net/common.py:87-92,178-197; discrete.py:56-67,109-114
  Categorical   torch.distributions.Categorical: probs/sum, log(clamp(eps, 1-eps))
"""
import math

import numpy as np
import torch

EPS_F32 = float(torch.finfo(torch.float32).eps)  # 1.1920929e-07, clamp_probs


def to_params(npz, prefix):
    """{name: tensor} for every key under ``prefix`` of a golden file."""
    return {k[len(prefix):]: torch.tensor(np.asarray(npz[k])) for k in npz.files if k.startswith(prefix)}


def positional_encoding(max_len, d):
    """state_tracker.py:261-269: sin on even dims, cos on odd dims (cos block truncated when d is odd)."""
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
    pe = torch.zeros(max_len, d)
    pe[:, 0::2] = torch.sin(pos * div)
    n_odd = pe[:, 1::2].shape[-1]
    pe[:, 1::2] = torch.cos(pos * div)[:, :n_odd]
    return pe


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def encoder_layer(x, P, l, nhead):
    """One post-norm TransformerEncoderLayer on x[L, n, d] with a causal mask (state_tracker.py:154-156,243)."""
    L, n, d = x.shape
    dh = d // nhead
    pre = f"transformer_encoder.layers.{l}."
    qkv = x @ P[pre + "self_attn.in_proj_weight"].T + P[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]

    def heads(t):  # [L,n,d] -> [n,h,L,dh]
        return t.reshape(L, n, nhead, dh).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    sc = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)
    mask = torch.triu(torch.full((L, L), float("-inf")), diagonal=1)
    att = torch.softmax(sc + mask, dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(L, n, d)
    o = o @ P[pre + "self_attn.out_proj.weight"].T + P[pre + "self_attn.out_proj.bias"]
    x = layer_norm(x + o, P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    f = torch.relu(x @ P[pre + "linear1.weight"].T + P[pre + "linear1.bias"])
    f = f @ P[pre + "linear2.weight"].T + P[pre + "linear2.bias"]
    return layer_norm(x + f, P[pre + "norm2.weight"], P[pre + "norm2.bias"])


def encode(X, P, nhead, nlayers=2, all_positions=False):
    """state_tracker.py:170-186 with dropout=0.  X[L, n, d] token sequence -> s[n, S] from the LAST position
    (or s[L, n, S] for every position when ``all_positions``: by causality position t equals the reference's
    prefix-t recompute)."""
    L, n, d = X.shape
    x = X * math.sqrt(d) + positional_encoding(L, d).unsqueeze(1)
    for l in range(nlayers):
        x = encoder_layer(x, P, l, nhead)
    out = x if all_positions else x[-1]
    return out @ P["decoder.weight"].T + P["decoder.bias"]


def user_token(P, users=None, user_dense=None):
    """state_tracker.py:205-215.  Kuaishou: embedding gather; Taobao: dense 88-vector pass-through."""
    e_u = P["embedding_dict.feat_user.weight"][torch.as_tensor(users, dtype=torch.long)] if users is not None \
        else torch.as_tensor(user_dense, dtype=torch.float32)
    return e_u @ P["ffn_user.weight"].T + P["ffn_user.bias"]


def action_token(P, rew, acts=None, act_dense=None):
    """state_tracker.py:225-242: g = sigmoid(W_g [r ; a] + b_g); token = g * a."""
    a = P["embedding_dict.feat_item.weight"][torch.as_tensor(acts, dtype=torch.long)] if acts is not None \
        else torch.as_tensor(act_dense, dtype=torch.float32)
    r = torch.as_tensor(np.asarray(rew), dtype=torch.float32).reshape(-1, 1)
    g = torch.sigmoid(torch.cat([r, a], -1) @ P["fnn_gate.weight"].T + P["fnn_gate.bias"])
    return g * a


class TrackerOracle:
    """Stateful restatement of StateTrackerTransformer.build_state (state_tracker.py:188-250): keeps the
    (MAX_TURN+1, B, d) token buffer and RECOMPUTES THE WHOLE PREFIX every step exactly like the reference.
    With ``keep_graph`` the returned states carry autograd history (the reference trains the tracker by
    back-propagating through the observations stored in the replay buffer)."""

    def __init__(self, P, nhead, max_turn, dense=False, keep_graph=True):
        self.P, self.nhead, self.max_len, self.dense, self.keep_graph = P, nhead, max_turn + 1, dense, keep_graph
        self.d = P["ffn_user.weight"].shape[0]

    def reset(self, B):
        self.data = torch.zeros(self.max_len, B, self.d)
        self.len = np.zeros(B, dtype=np.int64)

    def _ctx(self):
        return torch.enable_grad() if self.keep_graph else torch.no_grad()

    def first(self, obs, env_id):
        with self._ctx():
            tok = user_token(self.P, user_dense=obs[:, :-3]) if self.dense else \
                user_token(self.P, users=np.asarray(obs).reshape(-1))
            self.len[env_id] = 1
            data = self.data.clone()
            data[0, env_id] = tok
            self.data = data
            return encode(self.data[:1, env_id], self.P, self.nhead)

    def step(self, obs_next, rew, env_id):
        with self._ctx():
            tok = action_token(self.P, rew, act_dense=obs_next[:, :-3]) if self.dense else \
                action_token(self.P, rew, acts=np.asarray(obs_next).reshape(-1))
            self.len[env_id] += 1
            length = int(self.len[env_id[0]])  # lock-step assumption, state_tracker.py:232
            data = self.data.clone()
            data[length - 1, env_id] = tok
            self.data = data
            return encode(self.data[:length, env_id], self.P, self.nhead)


# ---------------------------------------------------------------- policy / value heads
def trunk(R, s):
    """Net: 20 -> 64 -> 64 ReLU MLP shared by actor and critic (common.py:87-92; CIRS-RL-kuaishou.py:245-247)."""
    h = torch.relu(s @ R["preprocess.model.model.0.weight"].T + R["preprocess.model.model.0.bias"])
    return torch.relu(h @ R["preprocess.model.model.2.weight"].T + R["preprocess.model.model.2.bias"])


def actor_probs(R, s):
    """discrete.py:56-67: softmax(W3 h + b3)."""
    return torch.softmax(trunk(R, s) @ R["actor.last.weight"].T + R["actor.last.bias"], dim=-1)


def critic_value(R, s):
    """discrete.py:109-114."""
    return (trunk(R, s) @ R["critic.last.weight"].T + R["critic.last.bias"]).reshape(-1)


def rl_params(actor_sd, critic_sd):
    """Merge actor/critic state dicts into one dict; the trunk is ONE shared tensor set."""
    R = {k: v for k, v in actor_sd.items() if k.startswith("preprocess.")}
    R["actor.last.weight"], R["actor.last.bias"] = actor_sd["last.model.0.weight"], actor_sd["last.model.0.bias"]
    R["critic.last.weight"], R["critic.last.bias"] = critic_sd["last.model.0.weight"], critic_sd["last.model.0.bias"]
    return R


def categorical_logits(p):
    """torch.distributions.Categorical(probs=p).logits: renormalise, clamp to [eps, 1-eps], log (§9-A4)."""
    p = p / p.sum(-1, keepdim=True)
    return torch.log(torch.clamp(p, EPS_F32, 1.0 - EPS_F32)), p


def log_prob(p, act):
    lg, _ = categorical_logits(p)
    return lg.gather(-1, torch.as_tensor(act, dtype=torch.long).reshape(-1, 1)).reshape(-1)


def entropy(p):
    lg, pn = categorical_logits(p)
    return -(lg * pn).sum(-1)


def sample_race(p, q):
    """Categorical.sample == torch.multinomial(p, 1, True) == argmax(p / q), q ~ Exp(1)."""
    return torch.argmax(torch.as_tensor(p) / torch.as_tensor(q), dim=-1)


# ---------------------------------------------------------------- continuous actor (VirtualTaobao)
def rl_params_continuous(actor_sd, critic_sd):
    """ActorProb + Critic state dicts -> one dict (tianshou/utils/net/continuous.py:120-199, :60-110)."""
    R = {k: v for k, v in actor_sd.items() if k.startswith("preprocess.")}
    R["actor.mu.weight"], R["actor.mu.bias"] = actor_sd["mu.model.0.weight"], actor_sd["mu.model.0.bias"]
    R["actor.sigma_param"] = actor_sd["sigma_param"]
    R["critic.last.weight"], R["critic.last.bias"] = critic_sd["last.model.0.weight"], critic_sd["last.model.0.bias"]
    return R


def actor_mu_sigma(R, s, max_action=1.0):
    """continuous.py:179-199: mu = max_action * tanh(W h + b); sigma = exp(sigma_param) broadcast over rows."""
    mu = max_action * torch.tanh(trunk(R, s) @ R["actor.mu.weight"].T + R["actor.mu.bias"])
    sigma = (R["actor.sigma_param"].view(1, -1) + torch.zeros_like(mu)).exp()
    return mu, sigma


def normal_log_prob(mu, sigma, act):
    """Independent(Normal(mu, sigma), 1).log_prob (torch/distributions/normal.py log_prob, summed over the last dim)."""
    act = torch.as_tensor(act, dtype=torch.float32)
    var = sigma ** 2
    return (-((act - mu) ** 2) / (2 * var) - sigma.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)


def normal_entropy(sigma):
    """Independent(Normal).entropy: sum of 0.5 + 0.5 log(2 pi) + log sigma."""
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(-1)


def sample_normal(mu, sigma, eps):
    """Normal.sample == torch.normal(mu, sigma) == N(0,1) * sigma + mu (ATen normal_out_impl: mul_ then add_)."""
    return torch.as_tensor(eps, dtype=torch.float32) * sigma + mu


def map_action(act, low, high):
    """tianshou/policy/base.py:143-173, action_bound_method="clip", action_scaling=True: float32 throughout, so the
    scaling to [low, high] = [-1, 1] is NOT an exact identity (act + 1 rounds)."""
    act = np.clip(np.asarray(act, dtype=np.float32), -1.0, 1.0)
    low, high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
    return low + (high - low) * (act + 1.0) / 2.0


# ---------------------------------------------------------------- Taobao reward model
def mmoe_forward(UM, x):
    """UserModel_MMOE.forward (core/user_model_mmoe.py:144-220) for the Taobao columns: dense features only (no
    sparse embeddings -> no FM term, :186), one regression task.  x[n,118] = [user 88, prev_r, 0, turn, item 27]
    (simulated_env.py:79-80).  Linear part core/layers.py:67-70; DNN deepctr layers/core.py:120-134 (ReLU, no BN, no
    dropout); MMOE layer core/layers.py:107-116 (experts reshaped [output_dim, num_experts], softmax gate);
    tower Linear(8 -> 1, no bias); PredictionLayer 'regression' adds a bias (layers/core.py:155-161)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    lin = x @ UM["linear_model_task.0.weight"]
    h = torch.relu(x @ UM["dnn.linears.0.weight"].T + UM["dnn.linears.0.bias"])
    h = torch.relu(h @ UM["dnn.linears.1.weight"].T + UM["dnn.linears.1.bias"])
    n_exp = UM["mmoe_layer.gating_networks.0.weight"].shape[0]
    e = (h @ UM["mmoe_layer.expert_network.weight"].T + UM["mmoe_layer.expert_network.bias"])
    e = e.reshape(x.shape[0], -1, n_exp)
    g = torch.softmax(h @ UM["mmoe_layer.gating_networks.0.weight"].T, dim=1).unsqueeze(-1)
    m = torch.bmm(e, g).squeeze(-1)
    dnn = m @ UM["tower_network.0.weight"].T
    return (lin + dnn) + UM["out.0.bias"]
