"""PPOPolicy: actor / critic forward with fused sampling (K3), returns (K4), the PPO update (K5, K7) and the
tracker's training pass (K6), behind the reference's policy interface.

Host-side mirror of (SURVEY §8b "Policy"):
  core/policy/ppo.py:14-246                         PPOPolicy.__init__ / process_fn / forward / learn
  tianshou/policy/modelfree/a2c.py:11-148           A2CPolicy (_compute_returns)
  tianshou/policy/modelfree/pg.py:10-139            PGPolicy (ret_rms, action scaling)
  tianshou/policy/base.py:13-423                    BasePolicy.update / map_action / compute_episodic_return
Same constructor keywords, same ``update(0, buffer, batch_size=, repeat=) -> {"loss": [...], ...}`` result.
Both actors of the reference run on the device: the discrete actor over the item catalogue (KuaishouEnv,
tianshou Actor + Categorical) and the continuous one (VirtualTaobao, tianshou ActorProb + Independent(Normal),
CIRS-RL-taobao.py:205-246) -- selected by the type of ``actor``.

Multi-GPU: environments are sharded over ranks; each global minibatch is the union of the ranks' local minibatches.
Gradients are summed with ONE all-reduce per minibatch (torch.distributed, NCCL over NVLink); the advantage
moments of all minibatches of a repeat, the return moments and the losses are reduced once per repeat / update.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, params
from .data import Batch


def split_indices(n, size, perm):
    """tianshou/data/batch.py:721-744 with merge_last=True."""
    merge = n % size > 0
    out = []
    for i in range(0, n, size):
        if merge and i + size + size >= n:
            out.append(perm[i:])
            break
        out.append(perm[i:i + size])
    return out


def split_sizes(n, size):
    """Chunk sizes of split_indices (they depend on n and size only)."""
    return [len(c) for c in split_indices(n, size, np.arange(n))]


class RunningMeanStd:
    """Host view of the device-resident return statistics (tianshou/utils/statistics.py:66-95)."""

    def __init__(self, dev):
        self.t = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device=dev)

    mean = property(lambda self: float(self.t[0]))
    var = property(lambda self: float(self.t[1]))
    count = property(lambda self: float(self.t[2]))


class PPOPolicy:
    def __init__(self, actor, critic, optim, dist_fn=None, eps_clip=0.2, dual_clip=None, value_clip=False,
                 advantage_normalization=True, recompute_advantage=False, vf_coef=0.5, ent_coef=0.01,
                 max_grad_norm=None, gae_lambda=0.95, max_batchsize=256, discount_factor=0.99,
                 reward_normalization=False, action_scaling=True, action_bound_method="clip", action_space=None,
                 lr_scheduler=None, deterministic_eval=False, state_tracker=None, device="cuda", seed=0,
                 process_group=None, **kwargs):
        _lib.require_cuda()
        _lib.load()
        assert dual_clip is None, "dual_clip is not used by CIRS (CIRS-RL-kuaishou.py:277-279)"
        assert not recompute_advantage, "recompute_advantage is not used by CIRS (default 0)"
        assert lr_scheduler is None
        if not reward_normalization:
            assert not value_clip, "value clip is available only when `reward_normalization` is True"  # ppo.py:87-89
        self.device = torch.device(device)
        self.actor, self.critic = actor, critic
        self.optim = optim if isinstance(optim, (list, tuple)) else [optim]
        self.dist_fn = dist_fn
        self.training, self.updating = True, False
        self.callbacks = []
        self._gamma, self._lambda = float(discount_factor), float(gae_lambda)
        self._rew_norm, self._batch = bool(reward_normalization), int(max_batchsize)
        self._deterministic_eval = bool(deterministic_eval)
        self.continuous = hasattr(actor, "sigma_param")                     # tianshou ActorProb
        self.action_type = "continuous" if self.continuous else "discrete"  # pg.py:56-61
        self.action_space = action_space
        self.action_scaling, self.action_bound_method = bool(action_scaling), action_bound_method
        self.seed, self._calls = int(seed), 0
        self.group = process_group
        self.use_c_comm = True   # NCCL process groups: gradient all-reduces issued from C inside cirs_ppo_learn
        self.c_loop = True       # False: per-minibatch entry points driven from Python (tests; gloo groups)
        # K6's forward half on a side stream beside the PPO minibatches (cirs_tracker_train phase 1 / 2).  Measured on
        # B200 at configs[1]: 1.672 vs 1.654 ms per iteration -- the stream fork / join costs what the overlap hides, so off
        self.overlap_tracker_forward = os.environ.get("CIRS_TRK_OVERLAP") == "1"
        self.h2d_bytes = self.d2h_bytes = 0          # host<->device traffic of the last update()
        # The reference builds ONE Net shared by actor and critic (CIRS-RL-kuaishou.py:245-247) and lists its tensors
        # twice in optim_RL / clip_grad_norm_; this implementation reproduces exactly that structure.
        # A separate critic trunk, or an optimizer over de-duplicated parameters, would train differently: refuse it.
        if actor.preprocess is not critic.preprocess:
            a_sd, c_sd = actor.preprocess.state_dict(), critic.preprocess.state_dict()
            same = a_sd.keys() == c_sd.keys() and all(torch.equal(a_sd[k], c_sd[k]) for k in a_sd)
            raise ValueError("PPOPolicy: actor.preprocess must BE critic.preprocess (the reference's shared trunk); "
                             + ("the two trunks hold equal tensors but are distinct modules" if same
                                else "the critic has its own trunk weights, which this path would discard"))
        trunk_ids = {id(p) for p in actor.preprocess.parameters()}
        first = (optim[0] if isinstance(optim, (list, tuple)) else optim)
        listed = [id(p) for g in first.param_groups for p in g["params"]]
        dup = {i for i in trunk_ids if listed.count(i) == 2}
        if trunk_ids and dup != trunk_ids:
            raise ValueError("PPOPolicy: the policy optimizer must list the shared trunk's parameters twice, i.e. "
                             "Adam(list(actor.parameters()) + list(critic.parameters())) as in "
                             "CIRS-RL-kuaishou.py:256-258 (duplicate-trunk clip / Adam semantics, SURVEY 7.3-2)")

        dim_state = actor.preprocess.input_dim
        n_action = actor.output_dim
        self.layout = params.policy_layout(dim_state, n_action, continuous=self.continuous,
                                           max_action=getattr(actor, "_max", 1.0))
        self.dim_state, self.n_action = dim_state, n_action
        sd = params.policy_sd_from_reference(actor.state_dict(), critic.state_dict())
        self.flat = self.layout.pack(sd, self.device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.opt_state = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.opt_scratch = torch.zeros(16, dtype=torch.float64, device=self.device)
        self._w = params.policy_struct(self.layout, self.flat)
        self._g = params.policy_struct(self.layout, self.grad)

        def hyper(opt):
            g = opt.param_groups[0]
            if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
                raise ValueError("PPOPolicy: weight_decay / amsgrad / maximize are not implemented by csrc/optim.cu "
                                 "(the reference uses plain Adam, CIRS-RL-kuaishou.py:256-259)")
            return float(g["lr"]), tuple(g.get("betas", (0.9, 0.999))), float(g.get("eps", 1e-8))

        lr, betas, eps = hyper(self.optim[0])
        self.cfg = _lib.PPOConfigStruct(float(eps_clip), float(vf_coef), float(ent_coef),
                                        float(max_grad_norm) if max_grad_norm else 0.0, int(bool(value_clip)),
                                        int(bool(advantage_normalization)), lr, betas[0], betas[1], eps)
        self.state_tracker = state_tracker
        self.cfg_tracker = None
        if len(self.optim) > 1:
            lr2, b2, e2 = hyper(self.optim[1])
            self.cfg_tracker = _lib.PPOConfigStruct(0, 0, 0, 0.0, 0, 0, lr2, b2[0], b2[1], e2)
            if self.state_tracker is None:
                for p in self.optim[1].param_groups[0]["params"]:
                    self.state_tracker = getattr(p, "_cirs_owner", self.state_tracker)
        self.ret_rms = RunningMeanStd(self.device)
        self._bridge_optim()
        if self.state_tracker is not None and len(self.optim) > 1:
            self.state_tracker.bridge_optim(self.optim[1])
        self._ws_actor = self._ws_ppo = None
        self._ws_actor_rows = self._ws_ppo_rows = 0

    # ------------------------------------------------------------------ module-like surface
    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, *a, **k):
        return self

    def cpu(self):
        return self

    def state_dict(self):
        """{"actor.<name>", "critic.<name>"} in the reference's naming (loads into tianshou's PPOPolicy)."""
        a, c = params.policy_sd_to_reference(self.layout.unpack(self.flat))
        out = {"actor." + k: v for k, v in a.items()}
        out.update({"critic." + k: v for k, v in c.items()})
        return out      # tianshou keeps ret_rms outside the state_dict too (a plain attribute); see trainer.save_checkpoint

    def load_state_dict(self, sd, strict=True):
        a = {k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")}
        c = {k[len("critic."):]: v for k, v in sd.items() if k.startswith("critic.")}
        self.flat.copy_(self.layout.pack(params.policy_sd_from_reference(a, c), self.device))
        self._pre = None   # evaluations queued by post_collect used the previous weights
        if "ret_rms" in sd:   # checkpoints written by round 1 of this package
            self.ret_rms.t.copy_(torch.as_tensor(sd["ret_rms"], dtype=torch.float64))

    # ------------------------------------------------------------------ optimizer state <-> torch.optim.Adam
    def _optim_names(self):
        """{id(parameter): key of the flat layout} for the tensors ``optim[0]`` lists (actor / critic modules)."""
        names = {}
        for n, p in self.actor.named_parameters():
            if n.startswith("preprocess."):
                key = "trunk." + n[len("preprocess.model.model."):]
            elif n == "sigma_param":
                key = "actor.sigma_param"
            else:
                key = "actor.last." + n.rsplit(".", 1)[1]
            names[id(p)] = key
        for n, p in self.critic.named_parameters():
            if not n.startswith("preprocess."):
                names[id(p)] = "critic.last." + n.rsplit(".", 1)[1]
        return names

    def _bridge_optim(self):
        """Make ``optim[0].state_dict()`` / ``.load_state_dict()`` -- what the reference's checkpoint code calls
        (CIRS-RL-kuaishou.py:340-358) -- carry the Adam moments that live in this policy's flat device buffers, in
        torch.optim.Adam's own layout (per-tensor exp_avg / exp_avg_sq / step, the shared trunk stepping twice per
        update).  The torch optimizer never steps; it is the hyper-parameter carrier and the checkpoint format."""
        opt = self.optim[0]
        if getattr(opt, "_cirs_bridged", False):
            return
        orig_sd, orig_load = opt.state_dict, opt.load_state_dict

        def state_dict():
            self.push_optim_state()
            return orig_sd()

        def load_state_dict(sd):
            orig_load(sd)
            self.pull_optim_state()

        opt.state_dict, opt.load_state_dict, opt._cirs_bridged = state_dict, load_state_dict, True

    def push_optim_state(self):
        """device moments -> optim[0].state (reference tensor shapes)."""
        names, opt = self._optim_names(), self.optim[0]
        m, v = self.layout.unpack(self.exp_avg), self.layout.unpack(self.exp_avg_sq)
        steps = self.opt_state.cpu().numpy()
        if int(steps[0]) == 0:
            opt.state.clear()
            return
        for g in opt.param_groups:
            for p in g["params"]:
                key = names[id(p)]
                step = steps[1] if key.startswith("trunk.") else steps[0]
                opt.state[p] = {"step": torch.tensor(float(step)), "exp_avg": m[key].reshape(p.shape).clone(),
                                "exp_avg_sq": v[key].reshape(p.shape).clone()}

    def pull_optim_state(self):
        """optim[0].state (e.g. just loaded from a reference checkpoint) -> device moments and step counters."""
        names, opt = self._optim_names(), self.optim[0]
        m, v = self.layout.unpack(self.exp_avg), self.layout.unpack(self.exp_avg_sq)
        steps = [0, 0]
        for g in opt.param_groups:
            for p in g["params"]:
                st = opt.state.get(p)
                if not st:
                    continue
                key = names[id(p)]
                m[key], v[key] = st["exp_avg"].reshape(m[key].shape), st["exp_avg_sq"].reshape(v[key].shape)
                steps[1 if key.startswith("trunk.") else 0] = int(float(st["step"]))
        self.exp_avg.copy_(self.layout.pack(m, self.device))
        self.exp_avg_sq.copy_(self.layout.pack(v, self.device))
        self.opt_state.copy_(torch.tensor(steps, dtype=torch.int32))

    def map_action(self, act):
        """policy/base.py:143-173: identity for a discrete action space; for a Box space clip to [-1, 1]
        (action_bound_method="clip") and scale to [low, high] (action_scaling) in float32 like numpy does."""
        if torch.is_tensor(act):
            act = act.detach().cpu().numpy()
        if self.continuous and isinstance(act, np.ndarray):
            if self.action_bound_method == "clip":
                act = np.clip(act, -1.0, 1.0)
            elif self.action_bound_method == "tanh":
                act = np.tanh(act)
            if self.action_scaling:
                low = np.asarray(getattr(self.action_space, "low", -1.0), dtype=np.float32)
                high = np.asarray(getattr(self.action_space, "high", 1.0), dtype=np.float32)
                act = low + (high - low) * (act + 1.0) / 2.0
        return act

    def exploration_noise(self, act, batch):
        return act

    def set_collector(self, train_collector):
        self.train_collector = train_collector

    # ------------------------------------------------------------------ forward (K3)
    def _actor_ws(self, n):
        if n > self._ws_actor_rows:
            need = _lib.load().cirs_actor_workspace_bytes(n, self.n_action)
            self._ws_actor = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_actor_rows = n
        return self._ws_actor

    def sample_device(self, n_rows, state, state_stride, act, logp, value, env_id=None, active=None, noise_q=None,
                      mode=None, seen=None, workspace=None, rng_counter=None):
        """One fused trunk + head + softmax + sample launch (csrc/actor.cu).  ``rng_counter`` (device int64[1]) makes
        the Philox stream advance on the device, so the launch can be replayed from a CUDA graph."""
        if mode is None:
            mode = 1 if (self._deterministic_eval and not self.training) else 0
        if rng_counter is None:
            self._calls += 1
        ws = workspace if workspace is not None else self._actor_ws(n_rows)
        _lib.call("cirs_actor_sample", C.byref(self._w), int(n_rows), _lib.ptr(env_id), _lib.ptr(active),
                  _lib.ptr(state), int(state_stride), _lib.ptr(noise_q), self.seed,
                  self._calls if rng_counter is None else (1 << 40), _lib.ptr(rng_counter), int(mode),
                  _lib.ptr(seen), _lib.ptr(act), _lib.ptr(logp), _lib.ptr(value), _lib.ptr(ws), _lib.stream())

    def actor_workspace(self, n_rows):
        """A dedicated sampler workspace (the fused Collector keeps its own so that captured graphs stay valid)."""
        need = _lib.load().cirs_actor_workspace_bytes(n_rows, self.n_action)
        return torch.empty(need, dtype=torch.uint8, device=self.device)

    def forward(self, batch, buffer=None, remove_recommended_ids=False, state=None, noise_q=None, **kwargs):
        """core/policy/ppo.py:111-163.  batch.obs: float32 CUDA tensor [n, dim_state].  Returns Batch(act, logp,
        value, logits=None, state=None, dist=None): the [n, n_action] probabilities are never materialised."""
        obs = batch.obs if not isinstance(batch, torch.Tensor) else batch
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device).contiguous()
        n = obs.shape[0]
        if self.continuous:
            return self._forward_continuous(obs, n, noise_q)
        act = torch.empty(n, dtype=torch.int32, device=self.device)
        logp = torch.empty(n, dtype=torch.float32, device=self.device)
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        seen = None
        if remove_recommended_ids and buffer is not None and len(buffer) > 0:
            seen = self._seen_bitset(buffer, n)
        if noise_q is not None:
            noise_q = torch.as_tensor(noise_q, dtype=torch.float32, device=self.device).contiguous()
        self.sample_device(n, obs, self.dim_state, act, logp, value, noise_q=noise_q, seen=seen)
        return Batch(logits=None, act=act.long(), state=None, dist=None, logp=logp, value=value)

    def _forward_continuous(self, obs, n, noise_eps):
        """ppo.py:144-156 with dist_fn = Independent(Normal): act = eps * sigma + mu (raw, unclipped -- the buffer stores
        this; the environment receives map_action(act)).  ``noise_eps``: N(0,1) draws [n, n_action] (parity runs)."""
        A = self.n_action
        act = torch.empty(n, A, dtype=torch.float32, device=self.device)
        mu = torch.empty(n, A, dtype=torch.float32, device=self.device)
        logp = torch.empty(n, dtype=torch.float32, device=self.device)
        value = torch.empty(n, dtype=torch.float32, device=self.device)
        if noise_eps is not None:
            noise_eps = torch.as_tensor(noise_eps, dtype=torch.float32, device=self.device).contiguous()
        mode = 1 if (self._deterministic_eval and not self.training) else 0
        self._calls += 1
        _lib.call("cirs_actorprob_sample", C.byref(self._w), n, None, None, _lib.ptr(obs), self.dim_state,
                  _lib.ptr(noise_eps), self.seed, self._calls, None, mode, _lib.ptr(act), _lib.ptr(logp),
                  _lib.ptr(value), _lib.ptr(mu), _lib.stream())
        sigma = self.flat[self.layout.segs["actor.sigma_param"].offset:][:A].exp().expand(n, A)
        return Batch(logits=(mu, sigma), act=act, state=None, dist=None, logp=logp, value=value)

    __call__ = forward

    def _seen_bitset(self, buffer, n):
        """get_recommended_ids (core/policy/utils.py:7-27) as a per-row bitset over the catalogue."""
        last = buffer.last_index
        indices = last[~buffer.done[last]]
        words = (self.n_action + 31) // 32
        bits = np.zeros((len(indices), words), dtype=np.uint32)
        rows = np.arange(len(indices))
        while len(indices):
            acts = buffer.act[indices]
            np.bitwise_or.at(bits, (rows, acts >> 5), (np.uint32(1) << (acts & 31).astype(np.uint32)))
            prev = buffer.prev(indices)
            if np.all(prev == indices):
                break
            indices = prev
        assert bits.shape[0] == n, "remove_recommended_ids: ready set and unfinished episodes differ"
        return torch.from_numpy(bits.view(np.int32)).to(self.device)

    # ------------------------------------------------------------------ update
    def _h2d_i32(self, arr):
        """Host int array -> device int32 tensor through a small ring of pinned staging buffers (asynchronous copy on
        the current stream; a buffer is reused only after the copy that read it has completed)."""
        n = int(len(arr))
        ring = getattr(self, "_pin_ring", None)
        if ring is None or ring[0][0].numel() < n:
            cap = max(n, 2 * (ring[0][0].numel() if ring else 0), 4096)
            ring = [[torch.empty(cap, dtype=torch.int32).pin_memory(), None] for _ in range(4)]
            self._pin_ring, self._pin_next = ring, 0
        slot = self._pin_ring[self._pin_next]
        self._pin_next = (self._pin_next + 1) % len(self._pin_ring)
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0][:n].numpy()[:] = arr
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        out.copy_(slot[0][:n], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return out

    def _ppo_ws(self, n):
        if n > self._ws_ppo_rows:
            need = _lib.load().cirs_ppo_workspace_bytes(n, self.n_action)
            self._ws_ppo = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_ppo_rows = n
        return self._ws_ppo

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size(self.group)
        return None, 1

    # ------------------------------------------------------------------ multi-GPU plumbing
    def _comm(self):
        """The C-side communicator of this policy's process group (csrc/comm.cu; NCCL), created on first use: rank 0
        draws the NCCL unique id and torch.distributed broadcasts it.  None for a single process or a CPU (gloo) group
        -- the latter keeps the per-minibatch Python loop with torch.distributed collectives (CPU tests)."""
        dist, world = self._world()
        if world == 1 or not self.use_c_comm:
            return None
        if getattr(self, "_c_comm", None) is None:
            if dist.get_backend(self.group) != "nccl":
                self.use_c_comm = False
                return None
            from .parallel import create_comm
            self._c_comm = create_comm(dist, self.group, self.device)
        return self._c_comm

    # ---- process_fn's kernels queued BEFORE the collect's read-back
    pre_eval = True     # set False (or CIRS_NO_PRE_EVAL=1) to queue them from process_fn as the reference's order has it

    def pre_update(self, buffer):
        """Called by the fused Collector after it has queued its read-back copies and before it waits for them.  The
        collect ends with a host synchronisation (lengths, rewards, transition count) and the update's first kernel used
        to be launched ~50 us after it: sync wake-up, bookkeeping, argument marshalling -- with the GPU idle.
        process_fn's kernels (V(obs) + old log-probs, V(obs'), GAE) only need the transition count to size their grids,
        so they are queued here, behind the copies, with the count still on the device (cirs_policy_eval_dev; capacity = the previous update's count + 12.5 %, at least 1024 rows) and they
        run while the host wakes up.  process_fn skips its own launches when the count fits the capacity and repeats
        them otherwise.  Same kernels on the same inputs: only the catalogue-split count of pass F (planned for the
        capacity) may differ from the in-order run, i.e. the summation order of the soft-max partials."""
        self._pre = None
        n_prev = getattr(self, "_n_prev", 0)
        if (not self.pre_eval or not self.training or os.environ.get("CIRS_NO_PRE_EVAL") == "1" or self.continuous
                or n_prev <= 0 or not self._tc_eval_ok()):
            return   # (test collectors run under policy.eval(): their buffers are never updated on)
        n_slots = buffer.maxsize
        cap = min(n_slots, max(1024, -(-(n_prev + n_prev // 8) // 128) * 128))
        self._process_kernels(buffer, cap, buffer.d_env_off[buffer.buffer_num:buffer.buffer_num + 1])
        self._pre = (buffer, cap)

    def _tc_eval_ok(self):
        if getattr(self, "_tc_ok", None) is None:
            na = int(self._w.n_action)
            self._tc_ok = (na >= 64 and int(self._w.dim_state) <= 32 and os.environ.get("CIRS_NO_TC", "0") in ("", "0")
                           and os.environ.get("CIRS_NO_TMA", "0") in ("", "0"))
        return self._tc_ok

    def _process_kernels(self, buffer, n, n_dev=None, indices=None):
        """V(obs) and the old log-probs, V(obs'), then GAE / returns (a2c.py:80-109, ppo.py:96-109): three C calls.
        ``n_dev`` (device int32[1]): the row count is still on the device and ``n`` is a capacity."""
        n_slots, dev = buffer.maxsize, self.device
        B, L = buffer.buffer_num, buffer.sub_size
        if getattr(self, "_slot_n", 0) != n_slots:
            self._slot_n = n_slots
            z = lambda dt=torch.float32: torch.zeros(n_slots, dtype=dt, device=dev)  # noqa: E731
            self.v_s, self.v_next, self.logp_old, self.returns, self.adv = z(), z(), z(), z(), z()
            self.d_obs = torch.zeros(n_slots, self.dim_state, dtype=torch.float32, device=dev)
            self._gae_scratch = torch.zeros(2 * B, dtype=torch.float64, device=dev)
            self._moments = torch.zeros(3, dtype=torch.float64, device=dev)
        indices = buffer.d_index if indices is None else indices
        aws = self._actor_ws(n_slots)   # sized for a full buffer once: no allocation inside later updates
        st = _lib.stream()
        if self.continuous:
            _lib.call("cirs_actorprob_eval", C.byref(self._w), n, _lib.ptr(indices), _lib.ptr(buffer.obs),
                      _lib.ptr(buffer.d_act), _lib.ptr(self.v_s), _lib.ptr(self.logp_old), st)
            _lib.call("cirs_actorprob_eval", C.byref(self._w), n, _lib.ptr(indices), _lib.ptr(buffer.obs_next), None,
                      _lib.ptr(self.v_next), None, st)
        elif n_dev is not None:
            _lib.call("cirs_policy_eval_dev", C.byref(self._w), n, _lib.ptr(n_dev), _lib.ptr(indices),
                      _lib.ptr(buffer.obs), _lib.ptr(buffer.d_act), _lib.ptr(self.v_s), _lib.ptr(self.logp_old),
                      _lib.ptr(aws), st)
            _lib.call("cirs_policy_eval_dev", C.byref(self._w), n, _lib.ptr(n_dev), _lib.ptr(indices),
                      _lib.ptr(buffer.obs_next), None, _lib.ptr(self.v_next), None, _lib.ptr(aws), st)
        else:
            _lib.call("cirs_policy_eval", C.byref(self._w), n, _lib.ptr(indices), _lib.ptr(buffer.obs),
                      _lib.ptr(buffer.d_act), _lib.ptr(self.v_s), _lib.ptr(self.logp_old), _lib.ptr(aws), st)
            _lib.call("cirs_policy_eval", C.byref(self._w), n, _lib.ptr(indices), _lib.ptr(buffer.obs_next), None,
                      _lib.ptr(self.v_next), None, _lib.ptr(aws), st)
        _lib.call("cirs_compute_returns", B, L, _lib.ptr(buffer.d_len), _lib.ptr(self.v_s), _lib.ptr(self.v_next),
                  _lib.ptr(buffer.d_rew), _lib.ptr(buffer.d_done), self._gamma, self._lambda,
                  _lib.ptr(self.ret_rms.t) if self._rew_norm else None, _lib.ptr(self._gae_scratch),
                  _lib.ptr(self._moments) if self._rew_norm else None, _lib.ptr(self.returns), _lib.ptr(self.adv), st)

    def post_collect(self, buffer):
        """Called by the fused Collector between the rollout kernel and its read-back: every rank's transition count
        (buffer.d_env_off[B], written by cirs_update_plan) is summed into slot ``rank`` of a [world] vector and copied
        to pinned host memory, so the collect's ONE synchronisation also delivers the minibatch plan of the following
        update (parallel.plan_from_counts) -- no host sync inside update()."""
        dist, world = self._world()
        self._n_all_pin = None
        if world == 1:
            return
        rank = dist.get_rank(self.group)
        if getattr(self, "_cnt_dev", None) is None:
            self._cnt_dev = torch.zeros(world, dtype=torch.int32, device=self.device)
            self._cnt_pin = torch.zeros(world, dtype=torch.int32).pin_memory()
        _lib.call("cirs_zero", _lib.ptr(self._cnt_dev), world * 4, _lib.stream())
        self._cnt_dev[rank:rank + 1].copy_(buffer.d_env_off[buffer.buffer_num:buffer.buffer_num + 1])
        comm = self._comm()
        if comm is not None:
            _lib.call("cirs_comm_allreduce", comm, _lib.ptr(self._cnt_dev), world, 2, _lib.stream())
        else:
            dist.all_reduce(self._cnt_dev, group=self.group)
        self._cnt_pin.copy_(self._cnt_dev, non_blocking=True)
        self._n_all_pin = (self._cnt_pin, buffer)

    def _allreduce(self, t):
        dist, world = self._world()
        if world == 1:
            return
        comm = self._comm()
        if comm is not None and t.dtype in (torch.float32, torch.float64, torch.int32):
            code = {torch.float32: 0, torch.float64: 1, torch.int32: 2}[t.dtype]
            _lib.call("cirs_comm_allreduce", comm, _lib.ptr(t), t.numel(), code, _lib.stream())
        else:
            dist.all_reduce(t, group=self.group)

    def process_fn(self, buffer, indices):
        """core/policy/ppo.py:96-109 + a2c.py:80-109: critic values, GAE / returns, old log-probs -- all on the
        device, results stay per buffer slot."""
        n = indices.numel()
        dev = self.device
        pre, self._pre = getattr(self, "_pre", None), None
        self._n_prev = n
        d_index = getattr(buffer, "d_index", None)
        if not (pre is not None and pre[0] is buffer and n <= pre[1] and d_index is not None
                and indices.data_ptr() == d_index.data_ptr()):
            # not queued by post_collect (first update, foreign collector, count above the capacity): in order, here
            self._process_kernels(buffer, n, indices=indices)
        st = _lib.stream()
        dist, world = self._world()
        self._n_all = None
        pin = getattr(self, "_n_all_pin", None)
        if world > 1 and pin is not None and pin[1] is buffer and int(pin[0][dist.get_rank(self.group)]) == n:
            # the counts travelled with the collect's read-back (post_collect): no collective, no host sync here
            self._n_all = pin[0].numpy().astype(np.int64)
            if self._rew_norm:
                if self.c_loop and self._comm() is not None:
                    self._defer_moments = True      # they ride on learn()'s advantage-statistics all-reduce
                else:
                    self._allreduce(self._moments)
        elif world > 1:
            # ONE collective: the raw return moments and, in slot 3 + rank, this rank's transition count; the counts
            # give every rank the whole minibatch plan of this update (parallel.plan_from_counts)
            rank = dist.get_rank(self.group)
            m = torch.zeros(3 + world, dtype=torch.float64, device=dev)
            if self._rew_norm:
                m[:3] = self._moments
            m[3 + rank] = float(n)
            self._allreduce(m)
            if self._rew_norm:
                self._moments.copy_(m[:3])
            self._n_all = m[3:].round().to(torch.int64).cpu().numpy()
        self._n_all_pin = None
        if self._rew_norm and not getattr(self, "_defer_moments", False):
            _lib.call("cirs_rms_update", _lib.ptr(self.ret_rms.t), _lib.ptr(self._moments), st)

    def update(self, sample_size, buffer, batch_size=None, repeat=1, perms=None, mb_sizes=None, **kwargs):
        """policy/base.py:219-244: sample(0) -> process_fn -> learn.  ``perms`` (one permutation of range(n) per
        repeat: host arrays, int32 CUDA tensors, or one [repeat, n] int32 CUDA tensor already resident in HBM) overrides np.random.permutation for
        replayable parity runs and for the resident-input benchmark arm.  ``mb_sizes`` (tests only) replaces the
        minibatch sizes Batch.split would produce, so that one process can replay a data-parallel run's plan."""
        if buffer is None or len(buffer) == 0:
            return {}
        assert sample_size == 0, "on-policy: the whole buffer is used (core/trainer/onpolicy.py:199)"
        self.updating = True
        buffer.sync_device()
        n = len(buffer)
        self.h2d_bytes, self.d2h_bytes = 0, 0
        if getattr(buffer, "_plan_ok", False):
            indices = buffer.d_index[:n]       # sample_index(0) built on the device by the collect (cirs_update_plan)
        else:
            indices = self._h2d_i32(buffer.sample_index(0))
            self.h2d_bytes += 4 * n
        # the tracker's training forward needs neither returns nor gradients: start it now on a side stream, beside
        # the head passes of process_fn / learn (joined before the tracker's backward in learn)
        self._trk_fwd = False
        if self.overlap_tracker_forward and self.state_tracker is not None and self.cfg_tracker is not None \
                and getattr(buffer, "_plan_ok", False):
            self.state_tracker.forward_async(buffer, getattr(buffer, "d_users", None), tok_slot=indices)
            self._trk_fwd = True
        self.process_fn(buffer, indices)
        result = self.learn(buffer, n, indices, batch_size or n, repeat, perms=perms, mb_sizes=mb_sizes)
        self.updating = False
        return result

    def learn(self, buffer, n, indices, batch_size, repeat, perms=None, mb_sizes=None):
        """core/policy/ppo.py:166-246."""
        dev, st = self.device, _lib.stream()
        dist, world = self._world()
        tracker = self.state_tracker if self.cfg_tracker is not None else None
        from .parallel import plan_from_counts, sharded_sizes
        n_glob_plan = None
        if mb_sizes is not None:
            assert world == 1 and sum(mb_sizes) == n
            sizes = [int(x) for x in mb_sizes]
        elif world > 1 and getattr(self, "_n_all", None) is not None:
            sizes, n_glob_plan = plan_from_counts(self._n_all, dist.get_rank(self.group), batch_size)
        else:
            sizes = sharded_sizes(n, batch_size, dist if world > 1 else None, self.group, dev)
        n_mb = len(sizes)
        # ---- minibatch offsets and the repeats' permutations travel in ONE pinned staging buffer and ONE asynchronous
        # copy: [offs (n_mb + 1, padded to 64) | perm_0 (n) | perm_1 (n) | ...]; the previous update's copy has completed
        # (every update ends with a stream synchronisation for the losses), so the buffer can be rewritten
        stacked = torch.is_tensor(perms) and perms.is_cuda     # [repeat, n] int32 already resident in HBM
        host_perms = perms is None or not (stacked or (torch.is_tensor(perms[0]) and perms[0].is_cuda))
        o_pad = (n_mb + 1 + 63) & ~63
        need = o_pad + (repeat * n if host_perms else 0)
        if getattr(self, "_stage_cap", 0) < need:
            self._stage_cap = max(need, 2 * getattr(self, "_stage_cap", 0), 4096)
            self._stage_pin = torch.empty(self._stage_cap, dtype=torch.int32).pin_memory()
            self._stage_np = self._stage_pin.numpy()
            self._stage_dev = torch.empty(self._stage_cap, dtype=torch.int32, device=dev)
        sp = self._stage_np
        sp[0] = 0
        sp[1:n_mb + 1] = np.cumsum(sizes)
        offs = sp[:n_mb + 1].copy()                  # host copy for the C loop's bounds
        if host_perms:
            for step in range(repeat):               # minibatch order of repeat ``step``: indices[perm]  (batch.py:736)
                sp[o_pad + step * n:o_pad + (step + 1) * n] = \
                    np.random.permutation(n) if perms is None else np.asarray(perms[step])
            self.h2d_bytes += 4 * n * repeat
        self._stage_dev[:need].copy_(self._stage_pin[:need], non_blocking=True)
        self.h2d_bytes += 4 * (n_mb + 1)
        d_offs_ptr = self._stage_dev.data_ptr()
        if getattr(self, "_learn_cap", (0, 0)) < (repeat * buffer.maxsize, repeat * (n_mb + 1)):
            self._learn_cap = (repeat * buffer.maxsize, repeat * (n_mb + 1))
            self._slots = torch.zeros(repeat * buffer.maxsize, dtype=torch.int32, device=dev)
            self._stats = torch.zeros(repeat * (n_mb + 1) * 3 + 3, dtype=torch.float64, device=dev)
            self._losses = torch.zeros(repeat * (n_mb + 1) * 4, dtype=torch.float32, device=dev)
        slots_ptr, stats_ptr, losses_ptr = self._slots.data_ptr(), self._stats.data_ptr(), self._losses.data_ptr()
        if host_perms:                               # every repeat's slots in one launch
            _lib.call("cirs_gather_i32", slots_ptr, _lib.ptr(indices), d_offs_ptr + 4 * o_pad, repeat * n, st)
        elif stacked:
            assert perms.shape == (repeat, n) and perms.dtype == torch.int32 and perms.is_contiguous()
            _lib.call("cirs_gather_i32", slots_ptr, _lib.ptr(indices), _lib.ptr(perms), repeat * n, st)
        else:
            for step in range(repeat):
                assert perms[step].numel() == n and perms[step].dtype == torch.int32
                _lib.call("cirs_gather_i32", slots_ptr + 4 * n * step, _lib.ptr(indices), _lib.ptr(perms[step]), n, st)

        # a chunk of Batch.split(merge_last) never exceeds 2 * batch_size - 1 rows: size the workspace once
        ws = self._ppo_ws(max(int(max(sizes)), min(buffer.maxsize, 2 * int(batch_size) - 1)))
        d_obs = self.d_obs if tracker is not None else None
        comm = self._comm()
        if self.c_loop and (world == 1 or comm is not None):
            # the whole repeat x minibatch loop is ONE C call (csrc/ppo.cu cirs_ppo_learn); with a communicator the
            # gradient all-reduce of every minibatch is issued from C between its kernels and clip + Adam
            n_glob, tail = None, 0
            if world > 1:
                assert n_glob_plan is not None
                n_glob = np.ascontiguousarray(n_glob_plan, dtype=np.int32)
                if getattr(self, "_defer_moments", False):   # return moments behind the statistics: one collective
                    tail = 3
                    self._stats[repeat * n_mb * 3:repeat * n_mb * 3 + 3].copy_(self._moments)
            _lib.call("cirs_ppo_learn", C.byref(self._w), C.byref(self._g), _lib.ptr(self.exp_avg),
                      _lib.ptr(self.exp_avg_sq), C.byref(self.cfg), repeat, n_mb, offs.ctypes.data, d_offs_ptr,
                      slots_ptr, _lib.ptr(buffer.obs), _lib.ptr(buffer.d_act), _lib.ptr(self.adv),
                      _lib.ptr(self.returns), _lib.ptr(self.v_s), _lib.ptr(self.logp_old), stats_ptr,
                      _lib.ptr(d_obs), d_obs.numel() if d_obs is not None else 0, losses_ptr,
                      _lib.ptr(self.opt_state), _lib.ptr(self.opt_scratch), _lib.ptr(ws), comm,
                      n_glob.ctypes.data if n_glob is not None else None, tail, st)
            if tail:
                self._moments.copy_(self._stats[repeat * n_mb * 3:repeat * n_mb * 3 + 3])
                _lib.call("cirs_rms_update", _lib.ptr(self.ret_rms.t), _lib.ptr(self._moments), st)
                self._defer_moments = False
        else:
            # CPU-side process group (gloo; tests): per-minibatch entry points with torch.distributed collectives.
            # advantage moments of every minibatch of every repeat (they depend on the permutations only): ONE collective
            for step in range(repeat):
                _lib.call("cirs_adv_stats", n_mb, d_offs_ptr, slots_ptr + 4 * n * step,
                          _lib.ptr(self.adv), stats_ptr + 24 * n_mb * step, st)
            stats = self._stats[:repeat * n_mb * 3]
            self._allreduce(stats)
            for step in range(repeat):
                n_glob = n_glob_plan if n_glob_plan is not None else \
                    stats.view(repeat, n_mb, 3)[step, :, 0].round().to(torch.int64).cpu().numpy()
                if tracker is not None:                                              # optim_state.zero_grad(), :174
                    _lib.call("cirs_zero", _lib.ptr(self.d_obs), self.d_obs.numel() * 4, st)
                for j in range(n_mb):
                    b, e = int(offs[j]), int(offs[j + 1])
                    _lib.call("cirs_ppo_minibatch", C.byref(self._w), C.byref(self._g), C.byref(self.cfg), e - b,
                              int(n_glob[j]), slots_ptr + 4 * (n * step + b), _lib.ptr(buffer.obs),
                              _lib.ptr(buffer.d_act), _lib.ptr(self.adv), _lib.ptr(self.returns), _lib.ptr(self.v_s),
                              _lib.ptr(self.logp_old), stats_ptr + 24 * (n_mb * step + j), _lib.ptr(d_obs),
                              losses_ptr + 16 * (n_mb * step + j), _lib.ptr(ws), st)
                    self._allreduce(self.grad)                                       # ONE collective per minibatch
                    _lib.call("cirs_clip_adam", _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg),
                              _lib.ptr(self.exp_avg_sq), self.layout.total, self.layout.n_trunk, C.byref(self.cfg),
                              _lib.ptr(self.opt_state), _lib.ptr(self.opt_scratch), st)
        losses = self._losses[:repeat * n_mb * 4]
        if tracker is not None:
            tracker.zero_grad()
            tracker.backward_from_buffer(buffer, self.d_obs, getattr(buffer, "d_users", None), tok_slot=indices,
                                         after_forward=getattr(self, "_trk_fwd", False))
            if comm is not None:       # tracker gradient + losses: one fused NCCL operation
                _lib.call("cirs_comm_group_begin", comm)
            self._allreduce(tracker.grad)
            self._allreduce(losses)
            if comm is not None:
                _lib.call("cirs_comm_group_end", comm)
            tracker.optim_step(self.cfg_tracker)                                     # optim_state.step(), :235
        else:
            self._allreduce(losses)
        if getattr(self, "_tc_flag", None) is None:
            self._tc_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        _lib.call("cirs_head_tc_timeout_peek", self._tc_flag.data_ptr(), st)
        lh = self._d2h_f32(losses).reshape(-1, 4).astype(np.float64)                 # the update's only D2H read
        self.d2h_bytes += losses.numel() * 4 + 4
        self.check_timeouts()
        return {"loss": lh[:, 0].tolist(), "loss/clip": lh[:, 1].tolist(), "loss/vf": lh[:, 2].tolist(),
                "loss/ent": lh[:, 3].tolist()}

    def _d2h_f32(self, t):
        """Device float32 vector -> host numpy through a pinned buffer (one asynchronous copy + one stream sync)."""
        pin = getattr(self, "_pin_f32", None)
        if pin is None or pin.numel() < t.numel():
            pin = self._pin_f32 = torch.empty(max(t.numel(), 256), dtype=torch.float32).pin_memory()
        pin[:t.numel()].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pin[:t.numel()].numpy().copy()

    def check_timeouts(self):
        """A tensor-core kernel that gave up waiting on an mbarrier (never expected) computed with incomplete data: the
        host raises instead of carrying on with silently wrong gradients.  The flag lives in managed / device memory
        was copied to pinned memory in front of the update's read-back (cirs_head_tc_timeout_peek)."""
        if int(self._tc_flag[0]):
            _lib.load().cirs_head_tc_timeout()   # clears the device-side flag
            raise _lib.CirsError("a tcgen05 head kernel timed out waiting on an mbarrier: results are invalid")
