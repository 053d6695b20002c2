"""Collector: the rollout driver (core/collector.py:20-367, CIRS's fork of tianshou 0.4.2's Collector).

Same constructor and ``collect`` result as the reference: ``Collector(policy, env, buffer, preprocess_fn,
exploration_noise, remove_recommended_ids, force_length)``, ``collect(n_episode=...) -> {"n/ep", "n/st", "rews",
"lens", "idxs", "rew", "len", "rew_std", "len_std"}``; the fork's behaviour is kept: everything is reset at the
start of every collect (:201), the buffer is emptied (:113-121), finished environments are dropped from the ready
set and never reset (:294-311), ``force_length`` overrides ``done`` (:253-258).

Two execution paths with identical results:
  * generic  -- the reference's loop, one call per component per turn through the numpy / Batch interfaces
               (policy() -> env.step() -> preprocess_fn() -> buffer.add()); any duck-typed env / policy works;
  * fused    -- when env, tracker, policy and buffer are this package's device-resident objects and
               n_episode == env_num: per turn three kernel launches (actor sample -> env step -> tracker step)
               on per-slot device arrays with an ``active`` mask; trajectories are written by the kernels straight
               into the replay buffer's env-major slots.  Three launch strategies, same device code and results:
               ``persistent`` (default) -- the whole rollout is ONE cooperative kernel with grid-wide barriers
               between the phases of a turn, ending when a device counter says every episode is over;
               ``use_graph`` -- reset, user token and max_turn turns captured once into a CUDA graph and replayed
               (the sampler's Philox counter lives on the device); otherwise the host issues the turns and stops
               through a non-blocking poll of "how many environments are still running".
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .data import Batch, VectorReplayBuffer
from .env import KuaishouVectorEnv, TaobaoVectorEnv
from .state_tracker import StateTrackerTransformer


class _LazyResult(dict):
    """The fused collect's result dict (core/collector.py:345-367 keys).  ``n/ep``, ``n/st`` and ``turns`` are there at
    once; the completion-ordered ``rews / lens / idxs`` (by turn, then environment id, like the reference's loop) and
    their means / deviations are filled in at the first access of any other key or of the dict as a whole."""
    _LAZY = ("rews", "lens", "idxs", "rew", "len", "rew_std", "len_std")

    def __init__(self, rews, lens, B, L):
        super().__init__({"n/ep": len(lens), "n/st": int(lens.sum()), "turns": int(lens.max()) if len(lens) else 0})
        self._raw = (rews, lens, B, L)

    def _fill(self):
        raw, self._raw = self._raw, None
        if raw is None:
            return
        rews, lens, B, L = raw
        order = np.lexsort((np.arange(B), lens))                    # completion order: by turn, then env id
        full = Collector._result(rews[order], lens[order], (np.arange(B) * L)[order])
        for k in self._LAZY:
            dict.__setitem__(self, k, full[k])

    def __getitem__(self, k):
        if self._raw is not None and k in self._LAZY:
            self._fill()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        if self._raw is not None and k in self._LAZY:
            self._fill()
        return dict.get(self, k, default)

    def __contains__(self, k):
        return k in self._LAZY or dict.__contains__(self, k)

    def _all(self):
        self._fill()
        return self

    def keys(self): return dict.keys(self._all())          # noqa: E704
    def items(self): return dict.items(self._all())        # noqa: E704
    def values(self): return dict.values(self._all())      # noqa: E704
    def __iter__(self): return dict.__iter__(self._all())  # noqa: E704
    def __len__(self): return dict.__len__(self._all())    # noqa: E704
    def __repr__(self): return dict.__repr__(self._all())  # noqa: E704
    def copy(self): return dict(self._all())               # noqa: E704


class Collector:
    def __init__(self, policy, env, buffer=None, preprocess_fn=None, exploration_noise=False,
                 remove_recommended_ids=False, force_length=0, fused=True, use_graph=True, persistent=True):
        self.policy, self.env = policy, env
        self.env_num = len(env)
        self.exploration_noise = exploration_noise
        self.preprocess_fn = preprocess_fn
        self.remove_recommended_ids = remove_recommended_ids
        self.force_length = int(force_length)
        self._action_space = getattr(env, "action_space", None)
        if buffer is None:
            buffer = VectorReplayBuffer(self.env_num * (getattr(env, "max_turn", 100) + 1), self.env_num)
        assert buffer.buffer_num >= self.env_num
        self.buffer = buffer
        self.tracker = getattr(preprocess_fn, "__self__", None)
        self.taobao = isinstance(env, TaobaoVectorEnv)
        self.fused = bool(fused and isinstance(env, (KuaishouVectorEnv, TaobaoVectorEnv))
                          and isinstance(self.tracker, StateTrackerTransformer)
                          and hasattr(policy, "sample_device") and isinstance(buffer, VectorReplayBuffer)
                          and buffer.buffer_num == self.env_num
                          and not (remove_recommended_ids and self.taobao)
                          and getattr(policy, "continuous", False) == self.taobao)
        if self.fused and remove_recommended_ids:
            env.enable_seen()   # the device-side set of already-recommended items (core/policy/utils.py:7-27)
        if remove_recommended_ids:
            persistent = True                 # the per-turn kernel chain has no device-side seen mask
        self.persistent = bool(persistent)    # fused rollout as ONE persistent cooperative kernel (csrc/rollout.cu)
        self.use_graph = bool(use_graph)      # else: replay the per-turn kernels from one captured CUDA graph
        self.data = Batch()
        self.h2d_bytes = self.d2h_bytes = 0   # host<->device traffic of the last fused collect()
        self.reset_stat()

    # ------------------------------------------------------------------ resets (collector.py:99-134)
    def reset(self, users=None):
        self.data = Batch(obs={}, act={}, rew={}, done={}, obs_next={}, info={}, policy={})
        self.reset_env(users)
        self.reset_buffer()
        self.reset_stat()

    def reset_stat(self):
        self.collect_step, self.collect_episode, self.collect_time = 0, 0, 0.0

    def reset_buffer(self, keep_statistics=False):
        self.buffer.reset()

    def reset_env(self, users=None):
        if self.preprocess_fn:
            self.preprocess_fn(dim_batch=self.env_num, reset=True)
        obs = self.env.reset(users=users) if users is not None else self.env.reset()
        self._reset_obs = obs
        if self.preprocess_fn:
            obs = self.preprocess_fn(obs=obs, env_id=np.arange(self.env_num)).get("obs", obs)
        self.data.obs = obs

    # ------------------------------------------------------------------ collect
    def collect(self, n_step=None, n_episode=None, random=False, render=None, no_grad=True, users=None,
                noise_fn=None):
        """``users`` (ndarray [env_num]) injects the episode's users instead of drawing them; ``noise_fn(turn, n)``
        supplies the Exp(1) race noise q[n, n_action] of the sampler (parity runs; generic path only)."""
        assert not getattr(self.env, "is_async", False)
        assert n_step is None and n_episode is not None and n_episode > 0, \
            "CIRS collects whole episodes (n_episode == env_num, SURVEY §9 invariants)"
        assert n_episode == self.env_num, "n_episode must equal the number of environments (collector.py:198-220)"
        start = time.time()
        if self.fused and not random and noise_fn is None:
            res = self._collect_fused_taobao(users) if self.taobao else self._collect_fused(users)
        else:
            res = self._collect_generic(n_episode, random, users, noise_fn)
        self.collect_step += res["n/st"]
        self.collect_episode += res["n/ep"]
        self.collect_time += max(time.time() - start, 1e-9)
        return res

    @staticmethod
    def _result(rews, lens, idxs):
        if len(lens):
            rm, rs, lm, ls = rews.mean(), rews.std(), lens.mean(), lens.std()
        else:
            rm = rs = lm = ls = 0
        return {"n/ep": len(lens), "n/st": int(np.sum(lens)), "rews": rews, "lens": lens, "idxs": idxs, "rew": rm,
                "len": lm, "rew_std": rs, "len_std": ls}

    # ---- generic path: the reference's loop (collector.py:219-320)
    def _collect_generic(self, n_episode, random, users, noise_fn):
        ready = np.arange(min(self.env_num, n_episode))
        self.reset(users)
        if isinstance(self.buffer, VectorReplayBuffer) and \
                np.issubdtype(np.asarray(self._reset_obs).dtype, np.integer):
            self.buffer._alloc(self.data.obs.shape[-1])
            self.buffer.d_users.copy_(torch.as_tensor(np.asarray(self._reset_obs).reshape(-1).astype(np.int32)))
        elif isinstance(self.buffer, VectorReplayBuffer) and self.taobao:
            self.buffer._alloc(self.data.obs.shape[-1], act_dim=27, user_dim=88)
            self.buffer.d_users_dense.copy_(torch.as_tensor(np.asarray(self._reset_obs)[:, :88].astype(np.float32)))
        step_count = episode_count = cnt_loop = 0
        ep_rews, ep_lens, ep_idxs = [], [], []
        while True:
            assert len(self.data.obs) == len(ready)
            if random:
                act = np.array([self._action_space[i % len(self._action_space)].sample() for i in ready])
                self.data.update(act=act)
            else:
                kw = {} if noise_fn is None else {"noise_q": noise_fn(cnt_loop, len(ready))}
                result = self.policy(self.data, self.buffer, state=None,
                                     remove_recommended_ids=self.remove_recommended_ids, **kw)
                act = result.act
                act = act.detach().cpu().numpy() if torch.is_tensor(act) else np.asarray(act)
                if self.exploration_noise:
                    act = self.policy.exploration_noise(act, self.data)
                self.data.update(policy=result.get("policy", Batch()), act=act)
            action_remap = self.policy.map_action(self.data.act)
            if self.taobao:
                self.data.update(act_env=action_remap)      # the tracker's training pass needs the mapped action
            obs_next, rew, done, info = self.env.step(action_remap, ready)
            cnt_loop += 1
            if self.force_length > 0:
                done = np.full_like(done, cnt_loop >= self.force_length, dtype=bool)
            self.data.update(obs_next=obs_next, rew=rew, done=done, info=info)
            if self.preprocess_fn:
                self.data.update(self.preprocess_fn(obs_next=self.data.obs_next, rew=self.data.rew,
                                                    done=self.data.done, info=self.data.info,
                                                    policy=self.data.policy, env_id=ready))
            ptr, ep_rew, ep_len, ep_idx = self.buffer.add(self.data, buffer_ids=ready)
            step_count += len(ready)
            if np.any(done):
                local = np.where(done)[0]
                episode_count += len(local)
                ep_lens.append(ep_len[local]); ep_rews.append(ep_rew[local]); ep_idxs.append(ep_idx[local])
                surplus = len(ready) - (n_episode - episode_count)
                if surplus > 0:
                    mask = np.ones_like(ready, dtype=bool)
                    mask[local[:surplus]] = False
                    ready = ready[mask]
                    self.data = Batch(obs=self.data.obs, obs_next=self.data.obs_next)[mask]
            self.data.obs = self.data.obs_next
            if episode_count >= n_episode:
                break
        if episode_count:
            rews, lens, idxs = map(np.concatenate, (ep_rews, ep_lens, ep_idxs))
        else:
            rews, lens, idxs = np.array([]), np.array([], int), np.array([], int)
        res = self._result(rews, lens, idxs)
        res["n/st"] = step_count
        return res

    # ---- fused path
    def _fused_state(self):
        env, trk, pol, dev = self.env, self.tracker, self.policy, self.env.device
        B, T = self.env_num, env.max_turn
        if not hasattr(self, "_f"):
            z = lambda dt: torch.zeros(B, dtype=dt, device=dev)  # noqa: E731
            self._f = dict(act=z(torch.int32), logp=z(torch.float32), value=z(torch.float32),
                           cur=torch.zeros(B, trk.dim_state, dtype=torch.float32, device=dev),
                           d_users=z(torch.int32), rng=torch.zeros(1, dtype=torch.int64, device=dev),
                           ws=pol.actor_workspace(B),
                           ws_roll=torch.empty(_lib.load().cirs_rollout_workspace_bytes(B, pol.n_action),
                                               dtype=torch.uint8, device=dev),
                           pin=torch.zeros(2 * T + 8, dtype=torch.int32).pin_memory(),
                           pin_users=torch.zeros(B, dtype=torch.int32).pin_memory(),
                           ev=[torch.cuda.Event() for _ in range(2 * T + 8)])
            self._f["pin_users_np"] = self._f["pin_users"].numpy()
            self._graph = None
        return self._f

    def _rollout_body(self, max_steps, poll):
        """reset -> user token -> max_steps x (sample -> env step -> tracker step), all on per-slot device arrays.
        With ``poll`` the host stops issuing turns once a non-blocking read says every episode has ended; without it
        (CUDA-graph capture) all turns are issued and finished environments are skipped on the device."""
        env, trk, pol, buf, f = self.env, self.tracker, self.policy, self.buffer, self._f
        B, L = self.env_num, buf.sub_size
        trk.build_state(dim_batch=B, reset=True)
        env.reset_device(f["d_users"])                               # sets active[:] = 1, turn = 0
        buf.d_users.copy_(f["d_users"])
        buf.d_len.zero_()
        trk.step_device(B, None, None, env.turn, 0, f["d_users"], None, None, cur_state=f["cur"],
                        traj=(L, buf.obs, buf.obs_next))             # user token -> s0 = obs[e, 0]
        traj = (L, buf.d_act, buf.d_rew, buf.d_done)
        turns = 0
        for t in range(max_steps):
            pol.sample_device(B, f["cur"], trk.dim_state, f["act"], f["logp"], f["value"], active=env.active,
                              workspace=f["ws"], rng_counter=f["rng"])
            env.step_device(f["act"], env.rew, env.done, traj=traj, ep_len=buf.d_len, force_length=self.force_length)
            trk.step_device(B, None, None, env.turn, t + 1, f["act"], None, env.rew, cur_state=f["cur"],
                            traj=(L, buf.obs, buf.obs_next))
            turns = t + 1
            if poll:   # non-blocking: number of environments still running after this turn, read two turns later
                f["pin"][t:t + 1].copy_(env.active.sum(dtype=torch.int32).reshape(1), non_blocking=True)
                f["ev"][t].record()
                if t >= 2 and f["ev"][t - 2].query() and int(f["pin"][t - 2]) == 0:
                    break
        return turns

    def _collect_fused(self, users):
        env, trk, buf, dev = self.env, self.tracker, self.buffer, self.env.device
        B, T = self.env_num, env.max_turn
        buf._alloc(trk.dim_state)
        L = buf.sub_size
        max_steps = self.force_length if self.force_length > 0 else T
        assert L >= max_steps, "buffer_size must be >= env_num * max(max_turn, force_length) (SURVEY §9 invariants)"
        assert self.force_length <= T, "force_length must not exceed the environments' max_turn"
        f = self._fused_state()
        self.h2d_bytes = self.d2h_bytes = 0
        if torch.is_tensor(users) and users.is_cuda:                # inputs already resident in HBM
            f["d_users"].copy_(users)
        else:
            users = env.draw_users(B) if users is None else np.asarray(users).reshape(-1)
            f["pin_users_np"][:] = users                            # pinned staging buffer (numpy view, no torch op)
            f["d_users"].copy_(f["pin_users"], non_blocking=True)   # pinned host -> device
            self.h2d_bytes += 4 * B
        self.data = Batch()
        buf.reset(device_only=True)
        trk.build_state(dim_batch=B, reset=True, zero_len=not self.persistent)   # K/V caches sized for THIS collector
        if self.persistent:
            # the whole rollout in ONE persistent cooperative kernel (csrc/rollout.cu)
            pol = self.policy
            mode = 1 if (pol._deterministic_eval and not pol.training) else 0
            if self.remove_recommended_ids:
                mode |= 4
            _lib.call("cirs_rollout_kuaishou", C.byref(env._struct), C.byref(trk._w), C.byref(pol._w),
                      _lib.ptr(f["d_users"]), _lib.ptr(env.active), _lib.ptr(f["act"]), _lib.ptr(f["logp"]),
                      _lib.ptr(f["value"]), _lib.ptr(f["cur"]), _lib.ptr(env.rew), _lib.ptr(env.done), L,
                      _lib.ptr(buf.obs), _lib.ptr(buf.obs_next), _lib.ptr(buf.d_act), _lib.ptr(buf.d_rew),
                      _lib.ptr(buf.d_done), _lib.ptr(buf.d_len), _lib.ptr(trk.kcache), _lib.ptr(trk.vcache),
                      int(trk.kcache.shape[1]), pol.seed, _lib.ptr(f["rng"]), mode, max_steps, self.force_length,
                      _lib.ptr(f["ws_roll"]), _lib.stream())
            buf.d_users.copy_(f["d_users"])
        elif self.use_graph:
            if self._graph is None:
                self._rollout_body(max_steps, poll=False)           # eager warm-up (function attributes, allocations)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._rollout_body(max_steps, poll=False)
                self._graph = g
            self._graph.replay()                                     # the whole rollout: ONE graph launch
        else:
            self._rollout_body(max_steps, poll=True)
        buf.plan_device()                                           # sample_index(0) / offsets for the update, on device
        if hasattr(self.policy, "post_collect"):
            self.policy.post_collect(buf)                           # multi-GPU: transition counts ride on the read-back
        flag = f["ws_roll"][128:132].view(torch.int32) if self.persistent else None
        # the collect's D2H read (one synchronisation); the policy may queue the update's front behind the copies
        lens, rews = self._read_back(buf.d_len, env.cum_rew, flag, getattr(self.policy, "pre_update", None), buf)
        self.d2h_bytes += 4 * B + 8 * B + 4
        buf.set_from_device(lens)
        # everything else the reference's result dict carries (completion-ordered copies, means / deviations) is
        # computed when it is first read: it is logging data, and the update that follows must not wait for it
        return _LazyResult(rews, lens, B, L)

    def _read_back(self, d_len, d_rew, d_flag=None, queue_behind=None, buf=None):
        """Episode lengths (i32) and cumulative rewards (f64) -> pinned host buffers, asynchronous copies and ONE
        synchronisation (on an event recorded right behind the copies).  ``d_flag``: the rollout kernel's "an mbarrier
        wait gave up" word (never expected): a set flag means the head phase computed with incomplete data, so the
        collect raises instead of returning.  ``queue_behind(buf)``: called after the copies are queued and before the
        host waits -- PPOPolicy.pre_update queues process_fn's kernels there, so the GPU has work while the host wakes
        up and marshals the update's first launches."""
        B = self.env_num
        if not hasattr(self, "_pin_out"):
            self._pin_out = (torch.zeros(B, dtype=torch.int32).pin_memory(),
                             torch.zeros(B, dtype=torch.float64).pin_memory(),
                             torch.zeros(1, dtype=torch.int32).pin_memory())
            self._pin_np = tuple(t.numpy() for t in self._pin_out)
        p_len, p_rew, p_flag = self._pin_out
        p_len.copy_(d_len[:B], non_blocking=True)
        p_rew.copy_(d_rew[:B], non_blocking=True)
        if d_flag is not None:
            p_flag.copy_(d_flag, non_blocking=True)
        if queue_behind is None:
            torch.cuda.current_stream().synchronize()
        else:
            if not hasattr(self, "_rb_event"):
                self._rb_event = torch.cuda.Event()
            self._rb_event.record()
            queue_behind(buf)
            self._rb_event.synchronize()
        n_len, n_rew, n_flag = self._pin_np
        if d_flag is not None and n_flag[0]:
            raise _lib.CirsError("cirs_rollout_kuaishou: a tcgen05 mbarrier wait timed out; the rollout is invalid")
        return n_len.astype(np.int64), n_rew.copy()

    # ---- fused path, VirtualTaobao: the whole collect is ONE kernel, one warp per environment (csrc/rollout_taobao.cu)
    def _collect_fused_taobao(self, users):
        env, trk, buf, pol, dev = self.env, self.tracker, self.buffer, self.policy, self.env.device
        B, T = self.env_num, env.max_turn
        buf._alloc(trk.dim_state, act_dim=27, user_dim=88)
        L = buf.sub_size
        max_steps = self.force_length if self.force_length > 0 else T
        assert L >= max_steps, "buffer_size must be >= env_num * max(max_turn, force_length) (SURVEY §9 invariants)"
        assert self.force_length <= T, "force_length must not exceed the environments' max_turn"
        if not hasattr(self, "_f"):
            self._f = dict(cur=torch.zeros(B, trk.dim_state, dtype=torch.float32, device=dev),
                           rng=torch.zeros(1, dtype=torch.int64, device=dev),
                           pin_users=torch.zeros(B, 88, dtype=torch.float32).pin_memory())
        f = self._f
        self.h2d_bytes = self.d2h_bytes = 0
        if torch.is_tensor(users) and users.is_cuda:
            buf.d_users_dense.copy_(users)
        else:
            users = env.draw_users(B) if users is None else np.asarray(users, dtype=np.float32).reshape(B, 88)
            f["pin_users"].copy_(torch.from_numpy(np.ascontiguousarray(users)))
            buf.d_users_dense.copy_(f["pin_users"], non_blocking=True)
            self.h2d_bytes += 4 * 88 * B
        self.data = Batch()
        buf.reset()
        trk.build_state(dim_batch=B, reset=True)
        mode = 1 if (pol._deterministic_eval and not pol.training) else 0
        _lib.call("cirs_rollout_taobao", C.byref(env._struct_raw), C.byref(trk._w), C.byref(pol._w),
                  _lib.ptr(buf.d_users_dense), _lib.ptr(env.active), _lib.ptr(f["cur"]), L, _lib.ptr(buf.obs),
                  _lib.ptr(buf.obs_next), _lib.ptr(buf.d_act), _lib.ptr(buf.d_act_env), _lib.ptr(buf.d_rew),
                  _lib.ptr(buf.d_done), _lib.ptr(buf.d_len), _lib.ptr(trk.kcache), _lib.ptr(trk.vcache),
                  int(trk.kcache.shape[1]), pol.seed, _lib.ptr(f["rng"]), mode, max_steps, self.force_length,
                  _lib.stream())
        buf.plan_device()
        if hasattr(pol, "post_collect"):
            pol.post_collect(buf)
        lens, rews = self._read_back(buf.d_len, env.cum_rew)
        self.d2h_bytes += 4 * B + 8 * B
        buf.set_from_device(lens)
        order = np.lexsort((np.arange(B), lens))
        res = self._result(rews[order], lens[order], (np.arange(B) * L)[order])
        res["turns"] = int(lens.max()) if len(lens) else 0
        return res
