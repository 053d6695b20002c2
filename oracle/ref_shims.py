"""Import shims that let the UNMODIFIED reference (/root/reference) import in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (run in the build
container, where /root/reference exists) to execute the reference's own hot
path and record golden vectors under ``tests/golden/``.  Nothing in the product
package, ``bench.py`` or the ``-m gpu`` tests imports this file: the reference
tree does not exist on the GPU box.

The reference needs gym, h5py, tensorflow and logzero, none of which are
installed here and none of which do arithmetic on the hot path; they are
replaced by inert stand-ins.  No reference source is copied.
"""
import sys
import types
import importlib.machinery

REF = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    if getattr(install, "_done", False):
        return
    # these two probe sys.modules['tensorflow'] on import: import them before faking it
    import torch.utils.tensorboard  # noqa: F401
    import torch._dynamo  # noqa: F401
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int  # simulated_env.py:176
    if not hasattr(np, "bool"):
        np.bool = bool

    # ---- gym ----------------------------------------------------------
    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = shape
            self.dtype = dtype
            self._rng = np.random.RandomState(0)

        def seed(self, s=None):
            self._rng = np.random.RandomState(s)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(tuple(shape), dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = n

        def sample(self):
            return int(self._rng.randint(self.n))

    class _Other(Space):
        def __init__(self, *a, **k):
            super().__init__()

    class Env:
        metadata = {}
        reward_range = (-float("inf"), float("inf"))
        spec = None
        action_space = None
        observation_space = None

        def seed(self, seed=None):
            return [seed]

        def close(self):
            pass

        def render(self, mode="human"):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

    registry = {}

    def register(id, entry_point=None, kwargs=None, **_):
        registry[id] = (entry_point, kwargs or {})

    def make(id, **kw):
        entry, kwargs = registry[id]
        if isinstance(entry, str):
            modname, clsname = entry.split(":")
            entry = getattr(importlib.import_module(modname), clsname)
        k = dict(kwargs)
        k.update(kw)
        return entry(**k)

    spaces = _mod("gym.spaces", Box=Box, Discrete=Discrete, Space=Space, MultiDiscrete=_Other,
                  MultiBinary=_Other, Dict=_Other, Tuple=_Other)
    reg = _mod("gym.envs.registration", register=register)
    envs = _mod("gym.envs", registration=reg)
    _mod("gym", Env=Env, Space=Space, Wrapper=Wrapper, spaces=spaces, envs=envs, make=make,
         register=register, registry=registry)

    # ---- h5py / logzero / tensorflow ---------------------------------
    _mod("h5py", Group=type("Group", (), {}), File=type("File", (), {}), Dataset=type("Dataset", (), {}))

    class _Logger:
        def __getattr__(self, k):
            return lambda *a, **kw: None

    _mod("logzero", logger=_Logger(), logfile=lambda *a, **k: None)
    cb = _mod("tensorflow.python.keras.callbacks", History=type("History", (), {}),
              CallbackList=type("CallbackList", (), {}), EarlyStopping=type("EarlyStopping", (), {}),
              ModelCheckpoint=type("ModelCheckpoint", (), {}))
    keras = _mod("tensorflow.python.keras", callbacks=cb)
    py = _mod("tensorflow.python", keras=keras)
    _mod("tensorflow", python=py)

    # deepctr_torch spawns a version-check thread that calls requests.get
    import requests

    def _no_net(*a, **k):
        raise RuntimeError("no network")

    requests.get = _no_net

    for p in (REF + "/environments/VirtualTaobao", REF + "/DeepCTR-Torch", REF + "/tianshou", REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # this repository's root stays in front: DeepCTR-Torch ships a top-level ``tests`` package that would otherwise
    # shadow ours (multiprocessing children re-import ``tests.*`` through the parent's sys.path)
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    install._done = True
