"""StateTrackerTransformer: the embedding-sequence encoder, rollout step (K2) and training pass (K6).

Host-side mirror of core/state_tracker.py:128-250 (same constructor arguments, same ``build_state`` protocol used
as the Collector's ``preprocess_fn``, same ``state_dict`` names so reference checkpoints load).

Differences in HOW (not WHAT):
  * rollout keeps a per-environment K/V cache and encodes ONE new token per step (csrc/tracker_step.cu) instead
    of re-encoding the whole prefix (state_tracker.py:246) -- exact because the mask is causal;
  * the returned states do not carry an autograd graph.  The reference trains the tracker by back-propagating
    through the observations stored in the replay buffer; here PPOPolicy.learn accumulates
    d loss / d obs per buffer slot and ``backward_from_buffer`` runs one full-sequence forward + backward
    (csrc/tracker_train.cu) -- the same gradient, computed once;
  * dropout: the reference's dropout is always active (nobody calls .eval(), state_tracker.py:154-155) and draws
    from the global torch RNG; this implementation computes the dropout = 0 function (the parity configuration,
    SURVEY §7.3-5).  A non-zero ``dropout`` argument is accepted for signature compatibility and ignored.
"""
import ctypes as C
import warnings

import numpy as np
import torch

from . import _lib, params
from .inputs import DenseFeat, SparseFeat, compute_input_dim


class StateTrackerTransformer:
    def __init__(self, user_columns, action_columns, feedback_columns, dim_model, dim_state, dim_max_batch,
                 dropout=0.1, dataset="VirtualTB-v0", has_user_embedding=True, has_action_embedding=True,
                 has_feedback_embedding=False, nhead=8, d_hid=128, nlayers=2, device="cuda", seed=2021,
                 init_std=0.0001, padding_idx=None, MAX_TURN=100, lr=1e-3):
        _lib.require_cuda()
        _lib.load()
        if dropout:
            warnings.warn("cirs_codes_b200 StateTrackerTransformer computes the dropout=0 function "
                          "(see module docstring); the dropout argument is ignored", stacklevel=2)
        self.device = torch.device(device if str(device) != "cpu" else "cuda")
        self.dataset, self.dim_model, self.dim_state = dataset, int(dim_model), int(dim_state)
        self.MAX_TURN = int(MAX_TURN) + 1                                  # state_tracker.py:144
        self.nhead, self.d_hid, self.nlayers = int(nhead), int(d_hid), int(nlayers)
        self.user_columns, self.action_columns, self.feedback_columns = user_columns, action_columns, feedback_columns
        # has_*_embedding == True means "the observation already IS the embedding" (dense pass-through)
        n_user = n_item = 0
        if not has_user_embedding:
            assert len(user_columns) == 1 and isinstance(user_columns[0], SparseFeat)
            n_user = user_columns[0].vocabulary_size
        if not has_action_embedding:
            assert len(action_columns) == 1 and isinstance(action_columns[0], SparseFeat)
            n_item = action_columns[0].vocabulary_size
        d_user_in, d_item_in = compute_input_dim(user_columns), compute_input_dim(action_columns)
        assert d_item_in == self.dim_model, "the gate multiplies the item vector elementwise: width must be dim_model"
        self.layout = params.tracker_layout(self.dim_model, self.nhead, self.d_hid, self.nlayers, self.dim_state,
                                            self.MAX_TURN, n_user, n_item, d_user_in, d_item_in)
        self.pe = params.positional_encoding(self.MAX_TURN, self.dim_model).to(self.device).contiguous()
        self.flat = self.layout.pack(self._init_state_dict(seed, init_std, d_user_in, d_item_in, n_user, n_item),
                                     self.device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.opt_state = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.opt_scratch = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.lr = float(lr)
        self._w = params.tracker_struct(self.layout, self.flat, self.pe)
        self._g = params.tracker_struct(self.layout, self.grad, None)
        self._ws = None
        self.build_state(dim_batch=int(dim_max_batch), reset=True)

    # ------------------------------------------------------------------ parameters
    def _init_state_dict(self, seed, init_std, d_user_in, d_item_in, n_user, n_item):
        """Same initialisers as the reference (torch.nn defaults on CPU under torch.manual_seed(seed),
        state_tracker.py:57,146-168; embeddings N(0, init_std), core/user_model.py:574-579)."""
        import torch.nn as nn
        gen_state = torch.get_rng_state()
        torch.manual_seed(seed)
        sd = {}
        if n_user:
            sd["embedding_dict.feat_user.weight"] = torch.empty(n_user, self.dim_model).normal_(0, init_std)
        if n_item:
            sd["embedding_dict.feat_item.weight"] = torch.empty(n_item, self.dim_model).normal_(0, init_std)
        ffn_user, fnn_gate = nn.Linear(d_user_in, self.dim_model), nn.Linear(1 + d_item_in, self.dim_model)
        layer = nn.TransformerEncoderLayer(self.dim_model, self.nhead, self.d_hid, 0.0)
        enc = nn.TransformerEncoder(layer, self.nlayers, enable_nested_tensor=False)
        dec = nn.Linear(self.dim_model, self.dim_state)
        dec.bias.data.zero_()
        dec.weight.data.uniform_(-0.1, 0.1)                                 # state_tracker.py:162-168
        for name, mod in (("ffn_user", ffn_user), ("fnn_gate", fnn_gate), ("transformer_encoder", enc),
                          ("decoder", dec)):
            for k, v in mod.state_dict().items():
                sd[f"{name}.{k}"] = v.detach().clone()
        torch.set_rng_state(gen_state)
        return sd

    def state_dict(self):
        sd = self.layout.unpack(self.flat)
        sd["pos_encoder.pe"] = self.pe.detach().cpu().unsqueeze(1)          # [max_len, 1, d] like the reference
        return sd

    def load_state_dict(self, sd, strict=True):
        sd = {k: v for k, v in sd.items() if k != "pos_encoder.pe"}
        self.flat.copy_(self.layout.pack(sd, self.device))

    def parameters(self):
        """One host Parameter per reference tensor, in the reference module's registration order, so that
        ``torch.optim.Adam(tracker.parameters(), lr=...)`` (CIRS-RL-kuaishou.py:259) is constructed unchanged and its
        ``state_dict()`` has the reference's layout (checkpoints, CIRS-RL-kuaishou.py:340-358).  PPOPolicy reads
        lr / betas / eps from that optimiser and finds the tracker through ``_cirs_owner``; the parameters are keys
        and checkpoint views only -- the update itself runs in csrc/optim.cu on the flat device buffer."""
        if getattr(self, "_host_params", None) is None:
            sd = self.layout.unpack(self.flat)
            self._host_params = []
            for k, v in sd.items():
                p = torch.nn.Parameter(v.clone(), requires_grad=False)
                p._cirs_owner, p._cirs_key = self, k
                self._host_params.append(p)
        return list(self._host_params)

    def bridge_optim(self, opt):
        """``opt.state_dict()`` / ``opt.load_state_dict()`` carry the Adam moments of the flat device buffer in
        torch.optim.Adam's per-tensor layout (see PPOPolicy._bridge_optim)."""
        if getattr(opt, "_cirs_bridged", False):
            return
        orig_sd, orig_load = opt.state_dict, opt.load_state_dict
        keyed = [p for g in opt.param_groups for p in g["params"] if hasattr(p, "_cirs_key")]

        def state_dict():
            step = int(self.opt_state[0])
            opt.state.clear()
            if step:
                m, v = self.layout.unpack(self.exp_avg), self.layout.unpack(self.exp_avg_sq)
                for p in keyed:
                    opt.state[p] = {"step": torch.tensor(float(step)), "exp_avg": m[p._cirs_key].clone(),
                                    "exp_avg_sq": v[p._cirs_key].clone()}
            return orig_sd()

        def load_state_dict(sd):
            orig_load(sd)
            m, v = self.layout.unpack(self.exp_avg), self.layout.unpack(self.exp_avg_sq)
            step = 0
            for p in keyed:
                st = opt.state.get(p)
                if st:
                    m[p._cirs_key], v[p._cirs_key] = st["exp_avg"], st["exp_avg_sq"]
                    step = int(float(st["step"]))
            self.exp_avg.copy_(self.layout.pack(m, self.device))
            self.exp_avg_sq.copy_(self.layout.pack(v, self.device))
            self.opt_state.copy_(torch.tensor([step, 2 * step], dtype=torch.int32))

        opt.state_dict, opt.load_state_dict, opt._cirs_bridged = state_dict, load_state_dict, True

    def to(self, *a, **k):
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ rollout (K2)
    def build_state(self, obs=None, env_id=None, obs_next=None, rew=None, done=None, info=None, policy=None,
                    dim_batch=None, reset=False, zero_len=True):
        """state_tracker.py:188-250.  Returns {} / {"obs": s0} / {"obs_next": s_t}; states are float32 CUDA
        tensors [len(env_id), dim_state]."""
        if reset and dim_batch:
            B, c = int(dim_batch), self.layout.cfg
            if getattr(self, "n_env", None) != B:
                # one K/V cache per environment count, kept: train and test collectors of different sizes alternate
                # (onpolicy_trainer) without reallocating, and pointers captured in CUDA graphs stay valid
                kv = self.__dict__.setdefault("_kv", {})
                if B not in kv:
                    k = torch.zeros(c["nlayers"], B, c["max_len"], c["d"], dtype=torch.float32, device=self.device)
                    kv[B] = (k, torch.zeros_like(k), torch.zeros(B, dtype=torch.int32, device=self.device))
                self.n_env = B
                self.kcache, self.vcache, self.len_data = kv[B]
            if zero_len:   # (the persistent rollout kernel keeps its own positions: nothing to clear)
                self.len_data.zero_()
            return None
        res = {}
        if obs is not None:
            ids = np.asarray(env_id, dtype=np.int64)
            d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
            self.len_data.index_fill_(0, d_ids.long(), 0)
            res = {"obs": self._step_rows(d_ids, self._inputs(obs, "user"), None)}
        elif obs_next is not None:
            ids = np.asarray(env_id, dtype=np.int64)
            d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
            d_rew = torch.as_tensor(np.asarray(rew, dtype=np.float32).reshape(-1), device=self.device)
            res = {"obs_next": self._step_rows(d_ids, self._inputs(obs_next, "action"), d_rew)}
        return res

    def _inputs(self, x, kind):
        """Sparse ids -> int32 [n]; dense features -> float32 [n, width] (VirtualTB: obs[:, :-3], :208,227)."""
        has_table = self._w.emb_user if kind == "user" else self._w.emb_item
        if has_table:
            return torch.as_tensor(np.asarray(x).reshape(len(x), -1)[:, 0].astype(np.int32), device=self.device)
        x = np.asarray(x, dtype=np.float32)
        if self.dataset == "VirtualTB-v0":
            x = x[:, :-3]
        return torch.as_tensor(np.ascontiguousarray(x), device=self.device)

    def _step_rows(self, d_ids, x, d_rew):
        n = d_ids.numel()
        out = torch.empty(n, self.dim_state, dtype=torch.float32, device=self.device)
        idx, dense = (x, None) if x.dtype == torch.int32 else (None, x)
        self.step_device(n, d_ids, None, self.len_data, -1, idx, dense, d_rew, state_out=out)
        self.len_data.index_add_(0, d_ids.long(), torch.ones(n, dtype=torch.int32, device=self.device))
        return out

    def step_device(self, n_rows, env_id, active, pos, expect_pos, idx, dense, rew, state_out=None, cur_state=None,
                    traj=None):
        """One launch of csrc/tracker_step.cu.  traj = (L, obs, obs_next) replay-buffer arrays or None."""
        L, tobs, tnext = traj if traj is not None else (0, None, None)
        _lib.call("cirs_tracker_step", C.byref(self._w), self.n_env, int(n_rows), _lib.ptr(env_id),
                  _lib.ptr(active), _lib.ptr(pos), int(expect_pos), _lib.ptr(idx), _lib.ptr(dense), _lib.ptr(rew),
                  _lib.ptr(self.kcache), _lib.ptr(self.vcache), _lib.ptr(state_out),
                  self.dim_state if state_out is not None else 0, _lib.ptr(cur_state), int(L), _lib.ptr(tobs),
                  _lib.ptr(tnext), _lib.stream())

    # ------------------------------------------------------------------ training (K6 + K7)
    def zero_grad(self):
        _lib.call("cirs_zero", _lib.ptr(self.grad), self.grad.numel() * 4, _lib.stream())

    def forward_async(self, buffer, users, tok_slot=None):
        """Issue the forward half of the training pass (it needs no upstream gradient) on a side stream, so that it
        runs beside the PPO minibatches; ``backward_from_buffer(..., after_forward=True)`` later joins and runs the
        backward half.  Tracker weights and buffer must not change in between (they do not: the tracker's Adam step is
        the last thing an update does)."""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._fwd_done = torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.backward_from_buffer(buffer, None, users, tok_slot=tok_slot, phase=1)
            self._fwd_done.record(self._side)

    def backward_from_buffer(self, buffer, d_obs, users, obs_check=None, compact=True, tok_slot=None, phase=0,
                             after_forward=False):
        """Accumulate d loss / d tracker-params into self.grad given d_obs[B*L, S] (zero rows = no gradient).
        ``compact``: process only the stored transitions' tokens (rows = buffer.sample_index(0)) instead of all B*L
        padded slots."""
        lib = _lib.load()
        if after_forward:
            torch.cuda.current_stream().wait_event(self._fwd_done)
            phase = 2
        B, L = buffer.buffer_num, buffer.sub_size
        env_off = None
        n_tok = 0
        if compact:
            n_tok = int(buffer._lengths.sum())
            # first compact row of every environment, computed on the device from the lengths the rollout wrote (no
            # host -> device copy, hence no synchronisation in front of the training pass)
            buffer.sync_device()
            if not getattr(buffer, "_plan_ok", False):
                buffer.plan_device()        # env_off / sample_index(0) on the device (csrc/util.cu)
            env_off = buffer.d_env_off
            if tok_slot is None:
                tok_slot = buffer.d_index[:n_tok]
        else:
            tok_slot = None
        n_rows = n_tok if compact else B * L
        need = lib.cirs_tracker_train_workspace_bytes(C.byref(self._w), B, n_rows)
        if self._ws is None or self._ws.numel() < need:   # sized for a full buffer once: no allocation in later updates
            need = max(need, lib.cirs_tracker_train_workspace_bytes(C.byref(self._w), B, B * L))
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
        dense = not self._w.emb_user
        _lib.call("cirs_tracker_train", C.byref(self._w), C.byref(self._g), B, L,
                  None if dense else _lib.ptr(users), None if dense else _lib.ptr(buffer.d_act),
                  _lib.ptr(buffer.d_rew), _lib.ptr(buffer.d_len),
                  _lib.ptr(buffer.d_users_dense) if dense else None, _lib.ptr(buffer.d_act_env) if dense else None,
                  n_tok,
                  _lib.ptr(tok_slot), _lib.ptr(env_off), int(buffer._lengths.max()) if compact else 0,
                  _lib.ptr(d_obs), _lib.ptr(obs_check), _lib.ptr(self._ws),
                  int(self._ws.numel()), int(phase), _lib.stream())

    def optim_step(self, cfg_struct):
        """optim_state.step() (core/policy/ppo.py:235): plain Adam, no clipping."""
        _lib.call("cirs_clip_adam", _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), self.layout.total, 0, C.byref(cfg_struct), _lib.ptr(self.opt_state),
                  _lib.ptr(self.opt_scratch), _lib.stream())
