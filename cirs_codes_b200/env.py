"""Vectorised KuaishouEnv / SimulatedEnv: B environments stepped by ONE kernel launch (csrc/env_kuaishou.cu).

Host-side mirror of the reference's vector-env interface (SURVEY §8b "Vector env"):
  tianshou/env/venvs.py:153-252        BaseVectorEnv.reset(id) / step(action, id) / seed / __len__ / is_async
  environments/KuaishouRec/env/kuaishouEnv.py:30-235   KuaishouEnv (reward = mat[u, a], category-overlap exit)
  core/env/simulatedEnv/simulated_env.py:17-193        SimulatedEnv (reward = normed_mat / (1 + exposure))
The reference holds one Python object per environment and loops over them (venvs.py:212-220); here the state of
all environments lives in device arrays (user, turn, hist[B, T], cum_rew) and ``step`` is a single launch.
``reset`` / ``step`` keep the numpy-in / numpy-out contract; ``reset_device`` / ``step_device`` are the
zero-copy entry points used by the fused Collector.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Discrete:
    """gym.spaces.Discrete stand-in (gym is not a dependency)."""

    def __init__(self, n, seed=None):
        self.n, self.shape, self.dtype = int(n), (), np.int64
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(0, self.n))

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def contains(self, x):
        return 0 <= int(x) < self.n


def cats_to_mask(cats):
    """list_feat (list of lists of category ids 1..31, kuaishouEnv.py:92-96) or int array [I, <=4] zero padded ->
    uint32 bitmask per item."""
    if isinstance(cats, (list, tuple)):
        m = np.zeros(len(cats), dtype=np.uint32)
        for i, row in enumerate(cats):
            for c in row:
                if c > 0:
                    m[i] |= np.uint32(1) << np.uint32(c)
        return m
    cats = np.asarray(cats)
    assert cats.max() <= 31 and cats.min() >= 0, "category ids must be in 1..31 (0 = padding)"
    m = np.zeros(cats.shape[0], dtype=np.uint32)
    for k in range(cats.shape[1]):
        c = cats[:, k].astype(np.uint32)
        m |= np.where(c > 0, np.uint32(1) << c, np.uint32(0)).astype(np.uint32)
    return m


class _PerEnv(list):
    """What ``DummyVectorEnv.__getattr__`` returns in the reference -- a list with one entry per environment
    (tianshou/env/venvs.py:113-131) -- for an attribute every environment shares: ``envs.mat[0].shape[1]``
    (evaluation.py:289) works unchanged, without B copies."""

    def __init__(self, value, n):
        super().__init__([value])
        self._n = n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return list.__getitem__(self, 0)

    def __iter__(self):
        return (list.__getitem__(self, 0) for _ in range(self._n))


class KuaishouVectorEnv:
    """``simulated=True``  -> SimulatedEnv over KuaishouEnv (training env, CIRS-RL-kuaishou.py:187-210)
    ``simulated=False`` -> raw KuaishouEnv (test envs, CIRS-RL-kuaishou.py:214-221).

    ``lbe_user`` / ``lbe_photo`` (sklearn LabelEncoders, ``classes_`` = raw ids in encoded order): as in the reference,
    ``list_feat`` is then indexed by RAW item id (kuaishouEnv.py:52: list_feat_small = list_feat[lbe_photo.classes_])
    and ``alpha_u`` / ``beta_i`` by RAW user / item id (simulated_env.py:157-161 goes through
    lbe_*.inverse_transform at every step); both are gathered into encoded order once, here.  Without encoders every
    table is taken to be in encoded order already."""

    is_async = False

    def __init__(self, env_num, mat, list_feat, *, normed_mat=None, alpha_u=None, beta_i=None, df_dist_small=None,
                 simulated=True, max_turn=30, num_leave_compute=1, leave_threshold=0, tau=100.0,
                 gamma_exposure=10.0, r_decay=1.0, version="v1", track_seen=False, device="cuda", seed=None,
                 lbe_user=None, lbe_photo=None, df_photo_env=None):
        _lib.require_cuda()
        _lib.load()
        self.device = torch.device(device)
        self.env_num, self.max_turn = int(env_num), int(max_turn)
        dev = self.device
        self.lbe_user, self.lbe_photo, self.df_photo_env = lbe_user, lbe_photo, df_photo_env
        if lbe_photo is not None:
            raw_items = np.asarray(lbe_photo.classes_)
            list_feat = [list_feat[int(x)] for x in raw_items]
            if beta_i is not None:
                beta_i = np.asarray(beta_i).reshape(-1)[raw_items]
        if lbe_user is not None and alpha_u is not None:
            alpha_u = np.asarray(alpha_u).reshape(-1)[np.asarray(lbe_user.classes_)]
        self.list_feat_small = list_feat

        def f32(x):
            return None if x is None else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)

        self._mat_host = mat if not torch.is_tensor(mat) else None
        self.d_mat = None if mat is None else (mat if torch.is_tensor(mat) else f32(mat))
        self.normed_mat = normed_mat if torch.is_tensor(normed_mat) else f32(normed_mat)
        ref = self.d_mat if self.d_mat is not None else self.normed_mat
        self.n_user, self.n_item = int(ref.shape[0]), int(ref.shape[1])
        self.simulated = bool(simulated)
        if self.simulated:
            assert self.normed_mat is not None, "SimulatedEnv needs normed_mat (simulated_env.py:100)"
        else:
            assert self.d_mat is not None
        self.cat_mask = torch.from_numpy(cats_to_mask(list_feat).view(np.int32)).to(dev)
        assert self.cat_mask.numel() == self.n_item
        self.alpha_u = None if alpha_u is None else f32(np.asarray(alpha_u).reshape(-1))
        self.beta_i = None if beta_i is None else f32(np.asarray(beta_i).reshape(-1))
        self.dist = None if df_dist_small is None else f32(np.asarray(df_dist_small))
        B, T = self.env_num, self.max_turn
        self.user = torch.zeros(B, dtype=torch.int32, device=dev)
        self.turn = torch.zeros(B, dtype=torch.int32, device=dev)
        self.hist = torch.zeros(B, T, dtype=torch.int32, device=dev)
        self.cum_rew = torch.zeros(B, dtype=torch.float64, device=dev)
        self.active = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.seen = torch.zeros(B, (self.n_item + 31) // 32, dtype=torch.int32, device=dev) if track_seen else None
        self.rew = torch.zeros(B, dtype=torch.float32, device=dev)     # last step's outputs, per env slot
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.cfg = dict(num_leave_compute=int(num_leave_compute), leave_threshold=float(leave_threshold),
                        tau=float(tau), gamma_exposure=float(gamma_exposure), r_decay=float(r_decay),
                        version=1 if version == "v1" else 2)
        s = _lib.KuaishouEnvStruct()
        s.n_env, s.max_turn, s.num_leave_compute = B, T, int(num_leave_compute)
        s.n_user, s.n_item, s.simulated, s.version = self.n_user, self.n_item, int(self.simulated), self.cfg["version"]
        s.leave_threshold, s.tau = float(leave_threshold), float(tau)
        s.gamma_exposure, s.r_decay = float(gamma_exposure), float(r_decay)
        s.normed_mat, s.mat, s.cat_mask = _lib.ptr(self.normed_mat), _lib.ptr(self.d_mat), _lib.ptr(self.cat_mask)
        s.alpha_u, s.beta_i, s.dist = _lib.ptr(self.alpha_u), _lib.ptr(self.beta_i), _lib.ptr(self.dist)
        s.user, s.turn, s.hist = _lib.ptr(self.user), _lib.ptr(self.turn), _lib.ptr(self.hist)
        s.cum_rew, s.seen = _lib.ptr(self.cum_rew), _lib.ptr(self.seen)
        self._struct = s
        self.action_space = [Discrete(self.n_item, None if seed is None else seed + i) for i in range(min(B, 8))]
        self._rng = np.random.default_rng(seed)

    @property
    def mat(self):
        """Per-environment view like DummyVectorEnv's attribute passthrough: ``envs.mat[0]`` is the [U, I] matrix."""
        m = self._mat_host if self._mat_host is not None else self.d_mat
        return _PerEnv(m, self.env_num)

    def enable_seen(self):
        """Allocate the per-environment bitset of already-recommended items (remove_recommended_ids on the device);
        the step kernel sets the bit of every action and reset clears it."""
        if self.seen is None:
            self.seen = torch.zeros(self.env_num, (self.n_item + 31) // 32, dtype=torch.int32,
                                    device=self.user.device)
            self._struct.seen = _lib.ptr(self.seen)
        return self.seen

    # ------------------------------------------------------------------ gym / tianshou surface
    def __len__(self):
        return self.env_num

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed if not isinstance(seed, (list, tuple)) else seed[0])
        return [seed] * self.env_num

    def render(self, **kwargs):
        return None

    def close(self):
        return None

    def draw_users(self, n):
        """kuaishouEnv.py:155-159 draws random.randint(0, U-1) per environment; here one vectorised draw."""
        return self._rng.integers(0, self.n_user, size=n, dtype=np.int64)

    def _ids(self, id):
        if id is None:
            return np.arange(self.env_num, dtype=np.int64)
        return np.atleast_1d(np.asarray(id, dtype=np.int64))

    def reset(self, id=None, users=None):
        """venvs.py:153-173 -> obs ndarray [len(id), 1] (the user id, kuaishouEnv.py:147-153)."""
        ids = self._ids(id)
        users = self.draw_users(len(ids)) if users is None else np.asarray(users, dtype=np.int64).reshape(-1)
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        d_users = torch.as_tensor(users.astype(np.int32), device=self.device)
        self.reset_device(d_users, d_ids)
        return users.reshape(-1, 1).copy()

    def step(self, action, id=None):
        """venvs.py:175-252 -> (obs_next [n,1] int64, rew float64 [n], done bool [n], info).  One launch."""
        ids = self._ids(id)
        act = np.asarray(action).reshape(len(ids), -1)[:, 0].astype(np.int32)
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        d_act = torch.as_tensor(act, device=self.device)
        rew = torch.empty(len(ids), dtype=torch.float32, device=self.device)
        done = torch.empty(len(ids), dtype=torch.uint8, device=self.device)
        self.step_device(d_act, rew, done, env_id=d_ids, use_active=False)
        rew_h, done_h = rew.cpu().numpy().astype(np.float64), done.cpu().numpy().astype(bool)
        info = {"env_id": ids}
        cum = self.cum_rew[torch.as_tensor(ids, device=self.device)].cpu().numpy()
        if self.simulated:
            turn = self.turn[torch.as_tensor(ids, device=self.device)].cpu().numpy()
            info["CTR"] = cum / np.maximum(turn, 1) / 10.0                     # simulated_env.py:141
        else:
            info["cum_reward"] = cum                                           # kuaishouEnv.py:178
        return act.astype(np.int64).reshape(-1, 1), rew_h, done_h, info

    # ------------------------------------------------------------------ device entry points
    def reset_device(self, d_users, d_ids=None):
        n = d_users.numel()
        _lib.call("cirs_kuaishou_reset", C.byref(self._struct), n, _lib.ptr(d_ids), _lib.ptr(d_users),
                  _lib.ptr(self.active), _lib.stream())

    def step_device(self, d_act, rew, done, env_id=None, use_active=True, traj=None, ep_len=None, force_length=0):
        """traj = (L, traj_act, traj_rew, traj_done) device arrays of the replay buffer, or None."""
        n = d_act.numel()
        L, ta, tr, td = traj if traj is not None else (0, None, None, None)
        _lib.call("cirs_kuaishou_step", C.byref(self._struct), n, _lib.ptr(env_id),
                  _lib.ptr(self.active) if use_active else None, _lib.ptr(d_act), _lib.ptr(rew), _lib.ptr(done),
                  int(L), _lib.ptr(ta), _lib.ptr(tr), _lib.ptr(td), _lib.ptr(ep_len), int(force_length),
                  _lib.stream())


class Box:
    """gym.spaces.Box stand-in (VirtualTB.action_space, virtualTB.py:24)."""

    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.shape, self.dtype = tuple(shape), dtype
        self.low = np.full(self.shape, low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype)
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class TaobaoVectorEnv:
    """B SimulatedEnv(VirtualTB) training environments (CIRS-RL-taobao.py:152-179) stepped by ONE kernel launch
    (csrc/env_taobao.cu): Euclidean exit test on the last min(t, N-1) actions (virtualTB.py:126-133), exposure effect
    over the whole action history (simulated_env.py:147-168) and the reward model ``user_model.forward`` evaluated
    inside the step (simulated_env.py:77-109; UserModel_MMOE, core/user_model_mmoe.py).

    ``user_model``: the reference's UserModel_MMOE (anything with ``state_dict()``) or that state_dict itself.
    Observations follow the reference: reset -> float64 [n, 91] = [user 88, 0, 0, 0] (virtualTB.py:54-55); step ->
    float64 [n, 30] = [action 27, reward, 0, turn] (simulated_env.py:50).  ``step`` expects the action already mapped
    by ``policy.map_action`` (the Collector does that, collector.py:246-250).
    The real VirtualTB's click model and new-user generator only consume torch RNG during training and their values
    are discarded by SimulatedEnv (simulated_env.py:114,138), so they are not evaluated; users are drawn uniformly
    per one-hot group unless injected (``reset(users=...)``)."""

    is_async = False
    GROUPS = (8, 8, 11, 11, 11, 11, 2, 2, 3, 18, 3)   # model/UserModel.py:22-32

    def __init__(self, env_num, user_model, *, max_turn=50, num_leave_compute=5, leave_threshold=3.0, tau=10.0,
                 gamma_exposure=10.0, version="v1", device="cuda", seed=None, generator=None):
        """``generator``: state_dict of VirtualTB's user generator (virtualTB/data/generator_model.pt).  With it,
        ``reset()`` draws users exactly like the reference (UserModel.generate, model/UserModel.py:40-60) on the device;
        without it users are uniform one-hot draws (the reference's data file is not part of this repository)."""
        from . import params
        _lib.require_cuda()
        _lib.load()
        self.device = torch.device(device)
        self.env_num, self.max_turn = int(env_num), int(max_turn)
        dev, B, T = self.device, self.env_num, self.max_turn
        sd = user_model.state_dict() if hasattr(user_model, "state_dict") else user_model
        sd = {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().float().cpu()
              for k, v in sd.items()}
        self._um_layout = params.mmoe_layout(sd)
        self._um_flat, out_bias = params.mmoe_pack(self._um_layout, sd, dev)
        self.user = torch.zeros(B, 88, dtype=torch.float32, device=dev)
        self.turn = torch.zeros(B, dtype=torch.int32, device=dev)
        self.hist = torch.zeros(B, T, 27, dtype=torch.float32, device=dev)
        self.prev_rew = torch.zeros(B, dtype=torch.float64, device=dev)
        self.cum_rew = torch.zeros(B, dtype=torch.float64, device=dev)
        self.active = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.rew = torch.zeros(B, dtype=torch.float32, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.action_space = [Box(-1, 1, (27,), np.float32, None if seed is None else seed + i)
                             for i in range(min(B, 8))]

        def make(map_action):
            s = _lib.TaobaoEnvStruct()
            s.n_env, s.max_turn, s.num_leave_compute = B, T, int(num_leave_compute)
            s.version, s.map_action = (1 if version == "v1" else 2), int(map_action)
            s.act_low, s.act_high = -1.0, 1.0
            s.leave_threshold, s.tau, s.gamma_exposure = float(leave_threshold), float(tau), float(gamma_exposure)
            params.mmoe_fill(s.um, self._um_layout, self._um_flat, out_bias)
            s.user, s.turn, s.hist = _lib.ptr(self.user), _lib.ptr(self.turn), _lib.ptr(self.hist)
            s.prev_rew, s.cum_rew = _lib.ptr(self.prev_rew), _lib.ptr(self.cum_rew)
            return s

        self._struct = make(0)          # step(): the caller has applied policy.map_action
        self._struct_raw = make(1)      # fused rollout: raw policy samples, mapped inside the kernel
        self._rng = np.random.default_rng(seed)
        self._gen = None
        if generator is not None:
            self._gen = params.virtualtb_pack(generator_sd=_cpu_sd(generator), device=dev)
            self._gen_seed, self._gen_calls = int(seed or 0) + 7919, 0

    def __len__(self):
        return self.env_num

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed if not isinstance(seed, (list, tuple)) else seed[0])
        return [seed] * self.env_num

    def render(self, **kwargs):
        return None

    def close(self):
        return None

    def draw_users(self, n, z=None, q=None):
        """New users [n, 88].  With the generator network: UserModel.generate on the device (``z`` [n, 128] uniform
        seeds and ``q`` [n, 88] Exp(1) race draws may be injected for parity runs)."""
        if self._gen is not None:
            return generate_users(self._gen, n, self.device, self._gen_seed, self._next_gen_call(), z, q).cpu().numpy()
        out = np.zeros((n, 88), dtype=np.float32)
        off = 0
        for g in self.GROUPS:
            out[np.arange(n), off + self._rng.integers(0, g, size=n)] = 1.0
            off += g
        return out

    def _next_gen_call(self):
        self._gen_calls += 1
        return self._gen_calls

    def _ids(self, id):
        if id is None:
            return np.arange(self.env_num, dtype=np.int64)
        return np.atleast_1d(np.asarray(id, dtype=np.int64))

    def reset(self, id=None, users=None):
        ids = self._ids(id)
        users = self.draw_users(len(ids)) if users is None else np.asarray(users, dtype=np.float32).reshape(len(ids), 88)
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        self.reset_device(torch.as_tensor(np.ascontiguousarray(users), device=self.device), d_ids)
        return np.concatenate([users.astype(np.float64), np.zeros((len(ids), 3))], axis=1)

    def step(self, action, id=None):
        ids = self._ids(id)
        n = len(ids)
        act = np.ascontiguousarray(np.asarray(action, dtype=np.float32).reshape(n, 27))
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        d_act = torch.as_tensor(act, device=self.device)
        rew = torch.empty(n, dtype=torch.float32, device=self.device)
        done = torch.empty(n, dtype=torch.uint8, device=self.device)
        _lib.call("cirs_taobao_step", C.byref(self._struct), n, _lib.ptr(d_ids), None, _lib.ptr(d_act), None,
                  _lib.ptr(rew), _lib.ptr(done), 0, None, None, None, None, None, 0, _lib.stream())
        sel = torch.as_tensor(ids, device=self.device)
        rew_h = self.prev_rew[sel].cpu().numpy()            # float64 like the reference
        done_h = done.cpu().numpy().astype(bool)
        turn = self.turn[sel].cpu().numpy()
        obs = np.concatenate([act.astype(np.float64), rew_h[:, None], np.zeros((n, 1)), turn[:, None].astype(np.float64)],
                             axis=1)
        info = {"env_id": ids, "CTR": self.cum_rew[sel].cpu().numpy() / np.maximum(turn, 1) / 10.0}
        return obs, rew_h, done_h, info

    def reset_device(self, d_users, d_ids=None):
        _lib.call("cirs_taobao_reset", C.byref(self._struct), int(d_users.shape[0]), _lib.ptr(d_ids),
                  _lib.ptr(d_users), _lib.ptr(self.active), _lib.stream())


def _cpu_sd(sd):
    sd = sd.state_dict() if hasattr(sd, "state_dict") else sd
    return {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().float().cpu()
            for k, v in sd.items()}


def generate_users(packed, n, device, seed=0, offset=0, z=None, q=None):
    """UserModel.generate (virtualTB/model/UserModel.py:40-60) for n users on the device -> float32 CUDA [n, 88]."""
    flat, w = packed
    f = lambda x: None if x is None else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=device)  # noqa: E731
    d_z, d_q = f(z), f(q)
    out = torch.empty(n, 88, dtype=torch.float32, device=device)
    _lib.call("cirs_virtualtb_generate_users", C.byref(w), int(n), _lib.ptr(d_z), _lib.ptr(d_q), int(seed), int(offset),
              _lib.ptr(out), _lib.stream())
    return out


class VirtualTBVectorEnv:
    """B raw VirtualTB environments (environments/VirtualTaobao/virtualTB/envs/virtualTB.py:13-133; the test
    environments of CIRS-RL-taobao.py:181-183) stepped by one kernel launch (csrc/virtualtb.cu): users from the
    generator network, reward = clicks predicted by the click model, Euclidean exit test.

    ``generator`` / ``action_model``: state_dicts of the shipped networks (virtualTB/data/generator_model.pt,
    action_model.pt).  Observations like the reference: reset -> [user 88, 0, 0, 0]; step -> [action 27, a, b, turn].
    ``noise`` hooks for parity runs: ``reset(z=, q=)``, ``step(action, id, q=)``."""

    is_async = False

    def __init__(self, env_num, generator, action_model, *, max_turn=100, num_leave_compute=5, leave_threshold=4.5,
                 device="cuda", seed=None):
        from . import params
        _lib.require_cuda()
        _lib.load()
        self.device = torch.device(device)
        self.env_num, self.max_turn = int(env_num), int(max_turn)
        dev, B, T = self.device, self.env_num, self.max_turn
        self._packed = params.virtualtb_pack(_cpu_sd(generator) if generator is not None else None,
                                             _cpu_sd(action_model), dev)
        self.has_generator = generator is not None
        self.user = torch.zeros(B, 88, dtype=torch.float32, device=dev)
        self.turn = torch.zeros(B, dtype=torch.int32, device=dev)
        self.hist = torch.zeros(B, T, 27, dtype=torch.float32, device=dev)
        self.prev_rew = torch.zeros(B, dtype=torch.float64, device=dev)
        self.cum_rew = torch.zeros(B, dtype=torch.float64, device=dev)
        self.active = torch.zeros(B, dtype=torch.uint8, device=dev)
        s = _lib.TaobaoEnvStruct()
        s.n_env, s.max_turn, s.num_leave_compute = B, T, int(num_leave_compute)
        s.version, s.map_action, s.act_low, s.act_high = 1, 0, -1.0, 1.0
        s.leave_threshold, s.tau, s.gamma_exposure = float(leave_threshold), 0.0, 0.0
        s.user, s.turn, s.hist = _lib.ptr(self.user), _lib.ptr(self.turn), _lib.ptr(self.hist)
        s.prev_rew, s.cum_rew = _lib.ptr(self.prev_rew), _lib.ptr(self.cum_rew)
        self._struct = s
        self.action_space = [Box(-1, 1, (27,), np.float32, None if seed is None else seed + i) for i in range(min(B, 8))]
        self._seed, self._calls = int(seed or 0) + 104729, 0
        self._rng = np.random.default_rng(seed)

    def __len__(self):
        return self.env_num

    def seed(self, seed=None):
        self._seed = int(seed if not isinstance(seed, (list, tuple)) else seed[0] or 0) + 104729
        return [seed] * self.env_num

    def render(self, **kwargs):
        return None

    def close(self):
        return None

    def _ids(self, id):
        if id is None:
            return np.arange(self.env_num, dtype=np.int64)
        return np.atleast_1d(np.asarray(id, dtype=np.int64))

    def reset(self, id=None, users=None, z=None, q=None):
        ids = self._ids(id)
        n = len(ids)
        self._calls += 1
        if users is None:
            assert self.has_generator, "VirtualTBVectorEnv.reset() needs the generator network or injected users"
            d_users = generate_users(self._packed, n, self.device, self._seed, self._calls, z, q)
        else:
            d_users = torch.as_tensor(np.ascontiguousarray(users, dtype=np.float32).reshape(n, 88), device=self.device)
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        _lib.call("cirs_taobao_reset", C.byref(self._struct), n, _lib.ptr(d_ids), _lib.ptr(d_users),
                  _lib.ptr(self.active), _lib.stream())
        return np.concatenate([d_users.cpu().numpy().astype(np.float64), np.zeros((n, 3))], axis=1)

    def step(self, action, id=None, q=None):
        ids = self._ids(id)
        n = len(ids)
        self._calls += 1
        act = np.ascontiguousarray(np.asarray(action, dtype=np.float32).reshape(n, 27))
        d_ids = torch.as_tensor(ids.astype(np.int32), device=self.device)
        d_act = torch.as_tensor(act, device=self.device)
        d_q = None if q is None else torch.as_tensor(np.ascontiguousarray(q, dtype=np.float32), device=self.device)
        rew = torch.empty(n, dtype=torch.float32, device=self.device)
        done = torch.empty(n, dtype=torch.uint8, device=self.device)
        click = torch.empty(n, 2, dtype=torch.int32, device=self.device)
        _lib.call("cirs_virtualtb_step", C.byref(self._struct), C.byref(self._packed[1]), n, _lib.ptr(d_ids),
                  _lib.ptr(d_act), _lib.ptr(d_q), self._seed, self._calls, _lib.ptr(rew), _lib.ptr(done),
                  _lib.ptr(click), 0, _lib.stream())
        sel = torch.as_tensor(ids, device=self.device)
        rew_h, done_h, click_h = rew.cpu().numpy().astype(np.float64), done.cpu().numpy().astype(bool), click.cpu().numpy()
        turn = self.turn[sel].cpu().numpy()
        click_obs = np.where(done_h[:, None], 0.0, click_h.astype(np.float64))   # lst_action is cleared on done (:99)
        obs = np.concatenate([act.astype(np.float64), click_obs, turn[:, None].astype(np.float64)], axis=1)
        info = {"env_id": ids, "CTR": self.cum_rew[sel].cpu().numpy() / np.maximum(turn, 1) / 10.0}
        return obs, rew_h, done_h, info


# ---------------------------------------------------------------------------------------------------------------
# Drop-in construction (CIRS-RL-kuaishou.py:173-221, CIRS-RL-taobao.py:152-185): the reference registers its
# environments with gym (``register(id, entry_point, kwargs)``), builds one with ``gym.make`` to read shapes, and
# wraps B lambdas in tianshou's ``DummyVectorEnv``.  The three functions below accept exactly those calls and hand
# back this package's device-resident vector environments, so the launch script changes only its imports:
#     from cirs_codes_b200.env import register, make, DummyVectorEnv
_REGISTRY = {}


def register(id, entry_point=None, kwargs=None, **_):
    """gym.envs.registration.register: remember the constructor keywords of an environment id."""
    _REGISTRY[id] = (entry_point or "", dict(kwargs or {}))


class EnvSpec:
    """What ``gym.make(id)`` returns here: a handle holding the registered keywords with the attributes the launch
    scripts read (``mat``, ``lbe_user``, ``lbe_photo``, ``action_space``, ``observation_space``, ``max_turn``)."""

    def __init__(self, id, entry_point, kwargs):
        self.id, self.entry_point, self.kwargs = id, entry_point, kwargs
        self.simulated = "simulated_env" in entry_point.lower() or "user_model" in kwargs
        base = kwargs
        if self.simulated:
            base = _REGISTRY[kwargs.get("task_name", "VirtualTB-v0")][1]
        self.base_kwargs = base
        self.taobao = "mat" not in base
        for k in ("mat", "lbe_user", "lbe_photo", "list_feat", "df_photo_env", "df_dist_small"):
            setattr(self, k, base.get(k))
        self.max_turn = base.get("max_turn", 100)
        if self.taobao:
            self.action_space = Box(-1, 1, (27,), np.float32)
            self.observation_space = Box(0, 100, (91,), np.float32)
        else:
            n_user, n_item = np.asarray(self.mat).shape
            self.action_space = Box(0, n_item - 1, (1,), np.int32)
            self.observation_space = Box(0, n_user - 1, (1,), np.int32)


def make(id, **kw):
    entry, kwargs = _REGISTRY[id]
    k = dict(kwargs)
    k.update(kw)
    return EnvSpec(id, entry, k)


def DummyVectorEnv(env_fns, device="cuda", seed=None):
    """tianshou.env.DummyVectorEnv([lambda: gym.make(id) for _ in range(B)]) -> ONE vector environment of B slots."""
    spec = env_fns[0]()
    assert isinstance(spec, EnvSpec), "DummyVectorEnv expects lambdas returning cirs_codes_b200.env.make(id)"
    B, kw, base = len(env_fns), spec.kwargs, spec.base_kwargs
    if spec.taobao:
        assert spec.simulated, "the raw VirtualTB environment is built with TaobaoVectorEnv(simulated=False, ...)"
        return TaobaoVectorEnv(B, kw["user_model"], max_turn=base.get("max_turn", 100),
                               num_leave_compute=base.get("num_leave_compute", 5),
                               leave_threshold=base.get("leave_threshold", 4.5), tau=kw.get("tau", 1.0),
                               gamma_exposure=kw.get("gamma_exposure", 1), version=kw.get("version", "v1"),
                               device=device, seed=seed)
    common = dict(max_turn=base.get("max_turn", 100), num_leave_compute=base.get("num_leave_compute", 5),
                  leave_threshold=base.get("leave_threshold", 1), df_dist_small=_frame(base.get("df_dist_small")),
                  lbe_user=base.get("lbe_user"), lbe_photo=base.get("lbe_photo"),
                  df_photo_env=base.get("df_photo_env"), device=device, seed=seed)
    if not spec.simulated:
        return KuaishouVectorEnv(B, base["mat"], base["list_feat"], simulated=False, **common)
    return KuaishouVectorEnv(B, base["mat"], base["list_feat"], normed_mat=kw["normed_mat"], alpha_u=kw.get("alpha_u"),
                             beta_i=kw.get("beta_i"), simulated=True, tau=kw.get("tau", 1.0),
                             gamma_exposure=kw.get("gamma_exposure", 1), r_decay=kw.get("r_decay", 1),
                             version=kw.get("version", "v1"), **common)


def _frame(x):
    """df_dist_small is a pandas DataFrame in the reference (util.py:33-36 ``.iloc[action, hist]``)."""
    if x is None:
        return None
    return x.to_numpy() if hasattr(x, "to_numpy") else np.asarray(x)
