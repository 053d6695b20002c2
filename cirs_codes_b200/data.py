"""Batch and VectorReplayBuffer with tianshou's attribute surface, backed by device-resident trajectory arrays.

Mirrors (host-side interface, SURVEY §8b "Buffer"):
  tianshou/data/batch.py:164-745            Batch        -- only what the hot path and its callers use
  tianshou/data/buffer/vecbuf.py:8-30       VectorReplayBuffer(total_size, buffer_num)
  tianshou/data/buffer/manager.py:9-169     add / sample_index / prev / next / unfinished_index / last_index
  tianshou/data/buffer/base.py:163-347      ring sub-buffers, ``add`` bookkeeping (ptr, ep_rew, ep_len, ep_idx)

Layout (identical to the reference): environment i owns slots [i*L, (i+1)*L), L = ceil(total_size / buffer_num).
obs / obs_next are float32 [B*L, S] CUDA tensors (the reference keeps torch tensors there too);
act (int32 on the device, int64 view on the host), rew (float32 device / float64 host), done (uint8 / bool).
The fused rollout writes the device arrays directly from the kernels; host views are synchronised lazily the
first time a callback touches ``buffer.act`` / ``.rew`` / ``.done`` (evaluation.py:309-354,
core/policy/utils.py:11-23).
"""
import numpy as np
import torch


class Batch:
    """Minimal dict-of-arrays with attribute access, batched indexing and update()."""

    def __init__(self, batch_dict=None, **kwargs):
        if batch_dict is not None:
            kwargs = dict(batch_dict, **kwargs)
        for k, v in kwargs.items():
            self.__dict__[k] = Batch(v) if isinstance(v, dict) else v

    def __setattr__(self, k, v):
        self.__dict__[k] = Batch(v) if isinstance(v, dict) else v

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.__dict__[idx]
        out = Batch()
        for k, v in self.__dict__.items():
            if isinstance(v, Batch):
                out.__dict__[k] = v[idx] if not v.is_empty() else Batch()
            elif torch.is_tensor(v):
                ii = idx
                if isinstance(idx, np.ndarray):
                    ii = torch.as_tensor(idx, device=v.device)
                out.__dict__[k] = v[ii]
            elif v is None:
                out.__dict__[k] = None
            else:
                out.__dict__[k] = np.asarray(v)[idx]
        return out

    def __setitem__(self, index, value):
        """tianshou/data/batch.py:244-267: ``b["key"] = v`` assigns a key; ``b[index] = other`` assigns the rows
        ``index`` of every key from a Batch / dict with a subset of this Batch's keys (keys missing there are filled
        with 0 / None / an empty Batch, like the reference); creating keys by item assignment is refused."""
        if isinstance(index, str):
            self.__setattr__(index, value)
            return
        if isinstance(value, dict):
            value = Batch(value)
        if not isinstance(value, Batch):
            raise ValueError("Batch does not supported tensor assignment. Use a compatible Batch or dict instead.")
        if not set(value.keys()).issubset(self.__dict__.keys()):
            raise ValueError("Creating keys is not supported by item assignment.")
        for key, val in self.__dict__.items():
            if isinstance(val, Batch) and val.is_empty():
                continue
            ii = index
            if torch.is_tensor(val) and isinstance(index, np.ndarray):
                ii = torch.as_tensor(index, device=val.device)
            if key in value.__dict__:
                src = value.__dict__[key]
                if isinstance(val, Batch):
                    val[index] = src
                elif torch.is_tensor(val):
                    val[ii] = torch.as_tensor(src, dtype=val.dtype, device=val.device)
                else:
                    val[index] = src
            elif isinstance(val, Batch):
                pass
            elif torch.is_tensor(val) or (isinstance(val, np.ndarray) and
                                          issubclass(val.dtype.type, (np.bool_, np.number))):
                val[ii] = 0
            elif val is not None:
                val[index] = None

    def split(self, size, shuffle=True, merge_last=False):
        """tianshou/data/batch.py:721-744: yield minibatches of ``size`` rows (the whole Batch when it is shorter);
        ``merge_last`` folds a short last chunk into the previous one.  np.random.permutation like the reference."""
        length = len(self)
        assert size >= 1
        indices = np.random.permutation(length) if shuffle else np.arange(length)
        merge_last = merge_last and length % size > 0
        for idx in range(0, length, size):
            if merge_last and idx + size + size >= length:
                yield self[indices[idx:]]
                break
            yield self[indices[idx:idx + size]]

    def __contains__(self, k):
        return k in self.__dict__

    def keys(self):
        return self.__dict__.keys()

    def items(self):
        return self.__dict__.items()

    def get(self, k, default=None):
        return self.__dict__.get(k, default)

    def update(self, batch=None, **kwargs):
        if batch is not None:
            src = batch.__dict__ if isinstance(batch, Batch) else batch
            for k, v in src.items():
                self.__setattr__(k, v)
        for k, v in kwargs.items():
            self.__setattr__(k, v)

    def is_empty(self):
        return len(self.__dict__) == 0

    def __len__(self):
        for v in self.__dict__.values():
            if isinstance(v, Batch):
                if not v.is_empty():
                    return len(v)
            elif v is not None and hasattr(v, "__len__"):
                return len(v)
        return 0

    def __repr__(self):
        return "Batch(" + ", ".join(f"{k}={type(v).__name__}" for k, v in self.__dict__.items()) + ")"


class VectorReplayBuffer:
    def __init__(self, total_size, buffer_num, device="cuda", dim_state=None):
        assert buffer_num > 0
        self.buffer_num = int(buffer_num)
        self.sub_size = int(np.ceil(total_size / buffer_num))           # vecbuf.py:27
        self.maxsize = self.sub_size * self.buffer_num                  # manager.py:29-37
        self.device = torch.device(device)
        self.dim_state = dim_state
        self._offset = np.arange(self.buffer_num, dtype=np.int64) * self.sub_size
        self.obs = self.obs_next = None
        self._alloc_done = False
        self.reset()

    # ------------------------------------------------------------------ storage
    def _alloc(self, dim_state, act_dim=0, user_dim=0):
        """act_dim > 0: continuous actions float32 [n, act_dim] (VirtualTaobao) plus the mapped action the environment
        used (``d_act_env``, the tracker's token input) and dense user features [B, user_dim]."""
        if self._alloc_done:
            return
        self.dim_state, self.act_dim = int(dim_state), int(act_dim)
        n, dev = self.maxsize, self.device
        if self.act_dim:
            self._h_act = np.zeros((n, self.act_dim), dtype=np.float32)
            self._h_act_env = np.zeros((n, self.act_dim), dtype=np.float32)
            self.d_act_env = torch.zeros(n, self.act_dim, dtype=torch.float32, device=dev)
            self.d_users_dense = torch.zeros(self.buffer_num, max(int(user_dim), 1), dtype=torch.float32, device=dev)
        self.obs = torch.zeros(n, self.dim_state, dtype=torch.float32, device=dev)
        self.obs_next = torch.zeros(n, self.dim_state, dtype=torch.float32, device=dev)
        self.d_act = torch.zeros((n, self.act_dim) if self.act_dim else n,
                                 dtype=torch.float32 if self.act_dim else torch.int32, device=dev)
        self.d_rew = torch.zeros(n, dtype=torch.float32, device=dev)
        self.d_done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.d_len = torch.zeros(self.buffer_num, dtype=torch.int32, device=dev)   # transitions per environment
        self.d_users = torch.zeros(self.buffer_num, dtype=torch.int32, device=dev)  # episode's user (tracker bwd)
        # sample_index(0) and the per-environment offsets of the last fused collect, built on the device
        self.d_index = torch.zeros(n, dtype=torch.int32, device=dev)
        self.d_env_off = torch.zeros(self.buffer_num + 1, dtype=torch.int32, device=dev)
        self._alloc_done = True

    def reset(self, keep_statistics=False, device_only=False):
        B = self.buffer_num
        self._lengths = np.zeros(B, dtype=np.int64)        # stored transitions per sub-buffer
        self._index = np.zeros(B, dtype=np.int64)          # next write position inside the sub-buffer
        self._last = np.zeros(B, dtype=np.int64)           # last written global slot (base.py:_index bookkeeping)
        self._ep_rew = np.zeros(B, dtype=np.float64)
        self._ep_len = np.zeros(B, dtype=np.int64)
        self._ep_idx = self._offset.copy()
        self._host_valid, self._dev_valid = True, True     # which side holds the truth for act / rew / done
        if device_only:
            # the fused rollout writes act / rew / done on the device: the host mirrors (0.4 MB of memset per reset at
            # 512 x 30 slots, on the critical path in front of the rollout launch) are only materialised when someone
            # reads them (_sync_host)
            self._host_valid = False
            return
        ad = getattr(self, "act_dim", 0)
        self._h_act = np.zeros((self.maxsize, ad), dtype=np.float32) if ad else np.zeros(self.maxsize, dtype=np.int64)
        if ad:
            self._h_act_env = np.zeros((self.maxsize, ad), dtype=np.float32)
        self._h_rew = np.zeros(self.maxsize, dtype=np.float64)
        self._h_done = np.zeros(self.maxsize, dtype=bool)
        self._plan_ok = False                              # d_index / d_env_off describe the stored transitions
        if self._alloc_done:
            self.d_len.zero_()

    # ------------------------------------------------------------------ host <-> device coherence
    def _sync_host(self):
        if not self._host_valid:
            self._h_act = self.d_act.cpu().numpy().astype(np.float32 if self.act_dim else np.int64)
            if self.act_dim:
                self._h_act_env = self.d_act_env.cpu().numpy()
            self._h_rew = self.d_rew.cpu().numpy().astype(np.float64)
            self._h_done = self.d_done.cpu().numpy().astype(bool)
            self._host_valid = True

    def sync_device(self):
        """Upload host-side act / rew / done / lengths (written by ``add``) before a device-side update."""
        if not self._dev_valid:
            self.d_act.copy_(torch.from_numpy(self._h_act.astype(np.float32 if self.act_dim else np.int32)))
            if self.act_dim:
                self.d_act_env.copy_(torch.from_numpy(self._h_act_env))
            self.d_rew.copy_(torch.from_numpy(self._h_rew.astype(np.float32)))
            self.d_done.copy_(torch.from_numpy(self._h_done.astype(np.uint8)))
            self.d_len.copy_(torch.from_numpy(self._lengths.astype(np.int32)))
            self._dev_valid = True

    def plan_device(self):
        """sample_index(0) on the device from the lengths the rollout kernel wrote (csrc/util.cu cirs_update_plan):
        ``d_index[:n]`` = stored slots env-major, ``d_env_off[e]`` = first compact row of environment e,
        ``d_env_off[B]`` = n.  Stream-ordered; the update reads them without any host round trip."""
        from . import _lib
        _lib.call("cirs_update_plan", self.buffer_num, self.sub_size, _lib.ptr(self.d_len), _lib.ptr(self.d_index),
                  _lib.ptr(self.d_env_off), _lib.stream())
        self._plan_ok = True

    def set_from_device(self, lengths):
        """Called by the fused rollout: the kernels wrote obs / obs_next / act / rew / done for env-major slots
        [e*L, e*L + lengths[e]).  ``lengths`` is the host copy of the per-environment episode lengths."""
        self._lengths = np.asarray(lengths, dtype=np.int64).copy()
        self._index = self._lengths % self.sub_size
        self._last = self._offset + np.maximum(self._lengths - 1, 0)
        self._host_valid, self._dev_valid = False, True

    @property
    def act(self):
        self._sync_host()
        return self._h_act

    @property
    def rew(self):
        self._sync_host()
        return self._h_rew

    @property
    def done(self):
        self._sync_host()
        return self._h_done

    # ------------------------------------------------------------------ tianshou interface
    def __len__(self):
        return int(self._lengths.sum())

    @property
    def last_index(self):
        return self._last.copy()                                            # manager.py:56-58

    def add(self, batch, buffer_ids=None):
        """manager.py:91-142 (vectorised over the ready environments instead of a Python loop).  ``batch`` holds
        obs / obs_next (torch [n, S] on the device), act, rew, done (numpy).  Returns (ptr, ep_rew, ep_len, ep_idx)."""
        ids = np.arange(self.buffer_num) if buffer_ids is None else np.asarray(buffer_ids, dtype=np.int64)
        obs, obs_next = batch.obs, batch.obs_next
        raw = np.asarray(batch.act)
        cont = raw.ndim == 2 and np.issubdtype(raw.dtype, np.floating)
        self._alloc(obs.shape[-1], act_dim=raw.shape[1] if cont else 0)
        self._sync_host()
        act = raw.astype(np.float32) if cont else raw.reshape(len(ids), -1)[:, 0].astype(np.int64)
        rew = np.asarray(batch.rew, dtype=np.float64).reshape(-1)
        done = np.asarray(batch.done, dtype=bool).reshape(-1)
        ptr = self._offset[ids] + self._index[ids]                          # base.py:163-183 _add_index
        self._h_act[ptr], self._h_rew[ptr], self._h_done[ptr] = act, rew, done
        if cont:   # the action the environment saw (policy.map_action), stashed by the Collector
            env_act = batch.get("act_env") if hasattr(batch, "get") else None
            self._h_act_env[ptr] = act if env_act is None else np.asarray(env_act, dtype=np.float32)
        pt = torch.as_tensor(ptr, device=self.device)
        self.obs.index_copy_(0, pt, obs.detach().to(torch.float32))
        self.obs_next.index_copy_(0, pt, obs_next.detach().to(torch.float32))
        self._last[ids] = ptr
        self._index[ids] = (self._index[ids] + 1) % self.sub_size
        self._lengths[ids] = np.minimum(self._lengths[ids] + 1, self.sub_size)
        self._ep_rew[ids] += rew
        self._ep_len[ids] += 1
        ep_rew = np.where(done, self._ep_rew[ids], 0.0)
        ep_len = np.where(done, self._ep_len[ids], 0)
        ep_idx = self._ep_idx[ids].copy()
        fin = ids[done]
        self._ep_rew[fin], self._ep_len[fin] = 0.0, 0
        self._ep_idx[fin] = self._offset[fin] + self._index[fin]
        self._dev_valid = False
        self._plan_ok = False
        return ptr, ep_rew, ep_len, ep_idx

    def sample_index(self, batch_size):
        """manager.py:144-169 with batch_size == 0: every stored transition, env-major, oldest first."""
        assert batch_size == 0, "only sample(0) (all data, on-policy) is on the hot path"
        lens = self._lengths
        wrapped = (lens == self.sub_size) & (self._index != 0)              # wrapped ring: base.py:267-275
        n = int(lens.sum())
        if n == 0:
            return np.zeros(0, dtype=np.int64)
        start = np.cumsum(lens) - lens                                      # first flat position of every sub-buffer
        loc = np.arange(n) - np.repeat(start, lens)                         # 0 .. len-1 inside each sub-buffer
        if wrapped.any():                                                   # oldest first: start reading at the write index
            loc = (loc + np.repeat(np.where(wrapped, self._index, 0), lens)) % self.sub_size
        return loc + np.repeat(self._offset, lens)

    def sample(self, batch_size):
        idx = self.sample_index(batch_size)
        return self[idx], idx

    def __getitem__(self, index):
        index = np.asarray(index)
        it = torch.as_tensor(index, device=self.device)
        return Batch(obs=self.obs[it], obs_next=self.obs_next[it], act=self.act[index], rew=self.rew[index],
                     done=self.done[index])

    def _sub(self, index):
        index = np.asarray(index, dtype=np.int64)
        return index, index // self.sub_size, index % self.sub_size

    def prev(self, index):
        """manager.py:194-213 (_prev_index): previous slot of the same episode, or itself at an episode start."""
        index, e, loc = self._sub(index)
        p = (loc - 1) % np.where(self._lengths[e] > 0, self._lengths[e], 1) + self._offset[e]
        first = self.done[p] | (p == self._last[e])
        return np.where(first, index, p)

    def next(self, index):
        """manager.py:215-232 (_next_index): next slot of the same episode, or itself at an episode end."""
        index, e, loc = self._sub(index)
        end = self.done[index] | (index == self._last[e])
        n = (loc + 1) % np.where(self._lengths[e] > 0, self._lengths[e], 1) + self._offset[e]
        return np.where(end, index, n)

    def unfinished_index(self):
        """manager.py:60-66: last written slot of every sub-buffer whose episode is still running."""
        has = self._lengths > 0
        last = self._last[has]
        return last[~self.done[last]]
