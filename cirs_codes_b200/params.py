"""Flat device parameter buffers and their conversion from / to the reference's torch ``state_dict`` layout.

Every network on the hot path keeps its parameters, their gradients and the two Adam moments as three / four
parallel flat float32 CUDA buffers with an identical layout, so that clip + Adam (K7) and the gradient
all-reduce are single passes over one contiguous range.  Inside the buffer a ``Linear(in -> out)`` is stored
k-major, ``Wt[in][ldo]`` with ``ldo = round_up(out, pad)``; the reference stores ``[out][in]``
(torch.nn.Linear).  Padding columns are zero, receive zero gradient and therefore stay zero under Adam.

Reference names (checkpoint compatibility, CIRS-RL-kuaishou.py:340-345):
  policy   actor.*  = tianshou Actor(Net)  (utils/net/discrete.py:11-67), critic.* = Critic(Net) (:70-114)
  tracker  StateTrackerTransformer.state_dict()  (core/state_tracker.py:128-168)
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

ALIGN = 32  # floats: every segment starts on a 128-byte boundary


def _up(x, m):
    return (x + m - 1) // m * m


class Segment:
    def __init__(self, name, kind, rows, cols, ld, offset):
        self.name, self.kind, self.rows, self.cols, self.ld, self.offset = name, kind, rows, cols, ld, offset
        self.size = _up(rows * ld, ALIGN)


class FlatLayout:
    """Ordered list of segments.  kinds: 'wt' (Linear weight, ref [cols, rows] -> stored [rows][ld]),
    'vec' (bias / LayerNorm vector, [cols] padded to ld), 'table' (embedding table [rows, cols], ld = cols)."""

    def __init__(self):
        self.segs = OrderedDict()
        self.total = 0

    def add(self, name, kind, rows, cols, pad=32):
        ld = cols if kind == "table" else _up(cols, pad)
        seg = Segment(name, kind, rows, cols, ld, self.total)
        self.segs[name] = seg
        self.total += seg.size
        return seg

    def pack(self, sd, device):
        """reference state_dict (torch tensors or numpy) -> flat float32 tensor on ``device``."""
        flat = torch.zeros(self.total, dtype=torch.float32)
        for seg in self.segs.values():
            t = torch.as_tensor(np.asarray(sd[seg.name]) if not torch.is_tensor(sd[seg.name]) else sd[seg.name])
            t = t.detach().to(torch.float32).cpu()
            view = flat[seg.offset:seg.offset + seg.rows * seg.ld].view(seg.rows, seg.ld)
            if seg.kind == "wt":
                assert tuple(t.shape) == (seg.cols, seg.rows), (seg.name, tuple(t.shape), (seg.cols, seg.rows))
                view[:, :seg.cols] = t.t()
            elif seg.kind == "vec":
                view[0, :seg.cols] = t.reshape(-1)
            else:
                assert tuple(t.shape) == (seg.rows, seg.cols), (seg.name, tuple(t.shape))
                view[:, :] = t
        return flat.to(device)

    def unpack(self, flat):
        """flat tensor -> {reference name: torch CPU tensor in the reference's layout}."""
        flat = flat.detach().cpu()
        out = OrderedDict()
        for seg in self.segs.values():
            view = flat[seg.offset:seg.offset + seg.rows * seg.ld].view(seg.rows, seg.ld)
            if seg.kind == "wt":
                out[seg.name] = view[:, :seg.cols].t().contiguous()
            elif seg.kind == "vec":
                out[seg.name] = view[0, :seg.cols].clone()
            else:
                out[seg.name] = view.clone()
        return out

    def ptr(self, flat, name):
        return flat.data_ptr() + 4 * self.segs[name].offset


# ---------------------------------------------------------------------------------------------- policy heads
TRUNK = "preprocess.model.model."


def policy_layout(dim_state, n_action, hidden=_lib.HIDDEN, continuous=False, max_action=1.0):
    """Trunk first (the duplicated tensors of optim_RL), then actor.last, then critic.last.
    ``continuous``: tianshou ActorProb (utils/net/continuous.py:120-199) -- "actor.last" is its ``mu`` layer
    (n_action <= 32) and ``sigma_param`` [n_action] follows the critic head."""
    L = FlatLayout()
    L.add("trunk.0.weight", "wt", dim_state, hidden)
    L.add("trunk.0.bias", "vec", 1, hidden)
    L.add("trunk.2.weight", "wt", hidden, hidden)
    L.add("trunk.2.bias", "vec", 1, hidden)
    L.n_trunk = L.total
    pad = 32 if continuous else 128
    L.add("actor.last.weight", "wt", hidden, n_action, pad=pad)
    L.add("actor.last.bias", "vec", 1, n_action, pad=pad)
    L.add("critic.last.weight", "vec", 1, hidden)  # reference shape [1, 64] -> contiguous wv[64]
    L.add("critic.last.bias", "vec", 1, 1)
    if continuous:
        assert n_action <= 32, "the continuous actor kernels hold one action component per lane"
        L.add("actor.sigma_param", "vec", 1, n_action)   # reference shape [n_action, 1]
    L.dim_state, L.n_action, L.continuous, L.max_action = dim_state, n_action, bool(continuous), float(max_action)
    return L


def policy_sd_from_reference(actor_sd, critic_sd):
    """Merge the reference's actor / critic state_dicts (the trunk is one shared tensor set)."""
    g = lambda sd, k: sd[k]  # noqa: E731
    if "sigma_param" in actor_sd:   # ActorProb: mu layer + state-independent log-std
        return {"trunk.0.weight": g(actor_sd, TRUNK + "0.weight"), "trunk.0.bias": g(actor_sd, TRUNK + "0.bias"),
                "trunk.2.weight": g(actor_sd, TRUNK + "2.weight"), "trunk.2.bias": g(actor_sd, TRUNK + "2.bias"),
                "actor.last.weight": g(actor_sd, "mu.model.0.weight"), "actor.last.bias": g(actor_sd, "mu.model.0.bias"),
                "actor.sigma_param": g(actor_sd, "sigma_param"),
                "critic.last.weight": g(critic_sd, "last.model.0.weight"),
                "critic.last.bias": g(critic_sd, "last.model.0.bias")}
    return {"trunk.0.weight": g(actor_sd, TRUNK + "0.weight"), "trunk.0.bias": g(actor_sd, TRUNK + "0.bias"),
            "trunk.2.weight": g(actor_sd, TRUNK + "2.weight"), "trunk.2.bias": g(actor_sd, TRUNK + "2.bias"),
            "actor.last.weight": g(actor_sd, "last.model.0.weight"), "actor.last.bias": g(actor_sd, "last.model.0.bias"),
            "critic.last.weight": g(critic_sd, "last.model.0.weight"),
            "critic.last.bias": g(critic_sd, "last.model.0.bias")}


def policy_sd_to_reference(sd):
    trunk = {TRUNK + "0.weight": sd["trunk.0.weight"], TRUNK + "0.bias": sd["trunk.0.bias"],
             TRUNK + "2.weight": sd["trunk.2.weight"], TRUNK + "2.bias": sd["trunk.2.bias"]}
    if "actor.sigma_param" in sd:
        actor = dict(trunk, **{"mu.model.0.weight": sd["actor.last.weight"], "mu.model.0.bias": sd["actor.last.bias"],
                               "sigma_param": sd["actor.sigma_param"].reshape(-1, 1)})
    else:
        actor = dict(trunk, **{"last.model.0.weight": sd["actor.last.weight"],
                               "last.model.0.bias": sd["actor.last.bias"]})
    critic = dict(trunk, **{"last.model.0.weight": sd["critic.last.weight"].reshape(1, -1),
                            "last.model.0.bias": sd["critic.last.bias"]})
    return actor, critic


def policy_struct(L, flat):
    s = _lib.PolicyWeightsStruct()
    s.dim_state, s.n_action, s.ld_action = L.dim_state, L.n_action, L.segs["actor.last.weight"].ld
    s.w1t, s.b1 = L.ptr(flat, "trunk.0.weight"), L.ptr(flat, "trunk.0.bias")
    s.w2t, s.b2 = L.ptr(flat, "trunk.2.weight"), L.ptr(flat, "trunk.2.bias")
    s.w3t, s.b3 = L.ptr(flat, "actor.last.weight"), L.ptr(flat, "actor.last.bias")
    s.wv, s.bv = L.ptr(flat, "critic.last.weight"), L.ptr(flat, "critic.last.bias")
    s.flat, s.n_flat, s.n_trunk = flat.data_ptr(), L.total, L.n_trunk
    s.sigma = L.ptr(flat, "actor.sigma_param") if getattr(L, "continuous", False) else None
    s.max_action = getattr(L, "max_action", 1.0)
    return s


# ---------------------------------------------------------------------------------------------- Taobao reward model
def mmoe_layout(sd):
    """UserModel_MMOE parameters used by forward() for dense feature columns and one regression task
    (core/user_model_mmoe.py:80-98, 144-220).  ``sd``: the reference model's state_dict."""
    w1, w2 = sd["dnn.linears.0.weight"], sd["dnn.linears.1.weight"]
    we, wg = sd["mmoe_layer.expert_network.weight"], sd["mmoe_layer.gating_networks.0.weight"]
    n_in, h1, h2, n_exp = int(w1.shape[1]), int(w1.shape[0]), int(w2.shape[0]), int(wg.shape[0])
    L = FlatLayout()
    L.add("linear_model_task.0.weight", "vec", 1, n_in)
    L.add("dnn.linears.0.weight", "wt", n_in, h1)
    L.add("dnn.linears.0.bias", "vec", 1, h1)
    L.add("dnn.linears.1.weight", "wt", h1, h2)
    L.add("dnn.linears.1.bias", "vec", 1, h2)
    L.add("mmoe_layer.expert_network.weight", "wt", h2, int(we.shape[0]))
    L.add("mmoe_layer.expert_network.bias", "vec", 1, int(we.shape[0]))
    L.add("mmoe_layer.gating_networks.0.weight", "wt", h2, n_exp)
    L.add("gate_zero_bias", "vec", 1, n_exp)
    L.add("tower_network.0.weight", "vec", 1, int(we.shape[0]) // n_exp)
    L.cfg = dict(n_in=n_in, h1=h1, h2=h2, n_expert=n_exp, expert_dim=int(we.shape[0]) // n_exp)
    return L


def mmoe_pack(L, sd, device):
    sd = dict(sd)
    sd["gate_zero_bias"] = torch.zeros(L.cfg["n_expert"])
    return L.pack(sd, device), float(torch.as_tensor(sd["out.0.bias"]).reshape(-1)[0])


def mmoe_fill(s, L, flat, out_bias):
    """Fill a _lib.MMOEStruct in place."""
    for k, v in L.cfg.items():
        setattr(s, k, v)
    s.lin_w = L.ptr(flat, "linear_model_task.0.weight")
    s.w1t, s.b1 = L.ptr(flat, "dnn.linears.0.weight"), L.ptr(flat, "dnn.linears.0.bias")
    s.w2t, s.b2 = L.ptr(flat, "dnn.linears.1.weight"), L.ptr(flat, "dnn.linears.1.bias")
    s.wet, s.be = L.ptr(flat, "mmoe_layer.expert_network.weight"), L.ptr(flat, "mmoe_layer.expert_network.bias")
    s.wgt, s.bg = L.ptr(flat, "mmoe_layer.gating_networks.0.weight"), L.ptr(flat, "gate_zero_bias")
    s.tower = L.ptr(flat, "tower_network.0.weight")
    s.out_bias = out_bias


# ---------------------------------------------------------------------------------------------- state tracker
def tracker_layout(d, nhead, d_hid, nlayers, dim_state, max_len, n_user=0, n_item=0, d_user_in=None,
                   d_item_in=None):
    """StateTrackerTransformer parameters (core/state_tracker.py:128-168).  n_user / n_item > 0 -> embedding
    tables (KuaishouEnv, core/inputs.py:36-42); 0 -> dense pass-through inputs (VirtualTB, :27-35)."""
    L = FlatLayout()
    d_user_in = d if d_user_in is None else d_user_in
    d_item_in = d if d_item_in is None else d_item_in
    if n_user:
        L.add("embedding_dict.feat_user.weight", "table", n_user, d)
    if n_item:
        L.add("embedding_dict.feat_item.weight", "table", n_item, d)
    L.add("ffn_user.weight", "wt", d_user_in, d)
    L.add("ffn_user.bias", "vec", 1, d)
    L.add("fnn_gate.weight", "wt", 1 + d_item_in, d)
    L.add("fnn_gate.bias", "vec", 1, d)
    for l in range(nlayers):
        p = f"transformer_encoder.layers.{l}."
        L.add(p + "self_attn.in_proj_weight", "wt", d, 3 * d)
        L.add(p + "self_attn.in_proj_bias", "vec", 1, 3 * d)
        L.add(p + "self_attn.out_proj.weight", "wt", d, d)
        L.add(p + "self_attn.out_proj.bias", "vec", 1, d)
        L.add(p + "linear1.weight", "wt", d, d_hid)
        L.add(p + "linear1.bias", "vec", 1, d_hid)
        L.add(p + "linear2.weight", "wt", d_hid, d)
        L.add(p + "linear2.bias", "vec", 1, d)
        for n in ("norm1", "norm2"):
            L.add(p + n + ".weight", "vec", 1, d)
            L.add(p + n + ".bias", "vec", 1, d)
    L.add("decoder.weight", "wt", d, dim_state)
    L.add("decoder.bias", "vec", 1, dim_state)
    L.cfg = dict(d=d, nhead=nhead, d_hid=d_hid, nlayers=nlayers, dim_state=dim_state, max_len=max_len,
                 n_user=n_user, n_item=n_item, d_user_in=d_user_in, d_item_in=d_item_in)
    return L


def positional_encoding(max_len, d):
    """PositionalEncoding table (core/state_tracker.py:261-269): sin on even columns, cos on odd columns (the cos
    block loses its last column when d is odd)."""
    pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-math.log(10000.0) / d))
    pe = torch.zeros(max_len, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)[:, :pe[:, 1::2].shape[-1]]
    return pe


def tracker_struct(L, flat, pe):
    c = L.cfg
    s = _lib.TrackerWeightsStruct()
    for k in ("d", "nhead", "d_hid", "nlayers", "dim_state", "max_len", "d_user_in", "d_item_in", "n_user", "n_item"):
        setattr(s, k, c[k])
    s.emb_user = L.ptr(flat, "embedding_dict.feat_user.weight") if c["n_user"] else None
    s.emb_item = L.ptr(flat, "embedding_dict.feat_item.weight") if c["n_item"] else None
    s.user_wt, s.user_b = L.ptr(flat, "ffn_user.weight"), L.ptr(flat, "ffn_user.bias")
    s.gate_wt, s.gate_b = L.ptr(flat, "fnn_gate.weight"), L.ptr(flat, "fnn_gate.bias")
    s.pe = pe.data_ptr() if pe is not None else None
    for l in range(c["nlayers"]):
        p = f"transformer_encoder.layers.{l}."
        ly = s.layer[l]
        ly.in_wt, ly.in_b = L.ptr(flat, p + "self_attn.in_proj_weight"), L.ptr(flat, p + "self_attn.in_proj_bias")
        ly.out_wt, ly.out_b = L.ptr(flat, p + "self_attn.out_proj.weight"), L.ptr(flat, p + "self_attn.out_proj.bias")
        ly.l1_wt, ly.l1_b = L.ptr(flat, p + "linear1.weight"), L.ptr(flat, p + "linear1.bias")
        ly.l2_wt, ly.l2_b = L.ptr(flat, p + "linear2.weight"), L.ptr(flat, p + "linear2.bias")
        ly.n1_w, ly.n1_b = L.ptr(flat, p + "norm1.weight"), L.ptr(flat, p + "norm1.bias")
        ly.n2_w, ly.n2_b = L.ptr(flat, p + "norm2.weight"), L.ptr(flat, p + "norm2.bias")
    s.dec_wt, s.dec_b = L.ptr(flat, "decoder.weight"), L.ptr(flat, "decoder.bias")
    s.flat, s.n_flat = flat.data_ptr(), L.total
    return s


# ---------------------------------------------------------------------------------------------- raw VirtualTaobao
def virtualtb_pack(generator_sd=None, action_sd=None, device="cuda"):
    """The shipped VirtualTB networks (virtualTB/data/generator_model.pt = UserModel.generator_model.state_dict(),
    action_model.pt = ActionModel.model.state_dict(); keys "0.weight", "0.bias", "2.weight", ...) -> one flat device
    buffer + a filled _lib.VirtualTBStruct.  Either may be None."""
    L = FlatLayout()
    sd = {}
    if generator_sd is not None:
        L.add("g.0.weight", "wt", 128, 128); L.add("g.0.bias", "vec", 1, 128)
        L.add("g.2.weight", "wt", 128, 88); L.add("g.2.bias", "vec", 1, 88)
        sd.update({"g." + k: v for k, v in generator_sd.items()})
    if action_sd is not None:
        L.add("a.0.weight", "wt", 116, 128); L.add("a.0.bias", "vec", 1, 128)
        L.add("a.2.weight", "wt", 128, 256); L.add("a.2.bias", "vec", 1, 256)
        L.add("a.4.weight", "wt", 256, 21); L.add("a.4.bias", "vec", 1, 21)
        sd.update({"a." + k: v for k, v in action_sd.items()})
    flat = L.pack(sd, device)
    s = _lib.VirtualTBStruct()
    if generator_sd is not None:
        s.g1t, s.g1b = L.ptr(flat, "g.0.weight"), L.ptr(flat, "g.0.bias")
        s.g2t, s.g2b = L.ptr(flat, "g.2.weight"), L.ptr(flat, "g.2.bias")
    if action_sd is not None:
        s.a1t, s.a1b = L.ptr(flat, "a.0.weight"), L.ptr(flat, "a.0.bias")
        s.a2t, s.a2b = L.ptr(flat, "a.2.weight"), L.ptr(flat, "a.2.bias")
        s.a3t, s.a3b = L.ptr(flat, "a.4.weight"), L.ptr(flat, "a.4.bias")
    return flat, s
