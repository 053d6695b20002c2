"""torch.ops.cirs_b200.* -- the C ABI registered as PyTorch operators (csrc/torch_ops.cpp, TORCH_LIBRARY): the operator
form SURVEY 8b proposes for the six per-step entry points of the path (env_step_kuaishou, tracker_step, actor_sample,
gae, ppo_minibatch, adam_clip).  A thin layer over libcirs_b200.so -- no kernel lives here, and the host classes of this
package keep calling the C ABI directly through ctypes; the operators exist for callers that want at::Tensor in /
at::Tensor out on the current CUDA stream with errors as RuntimeError.

    from cirs_codes_b200 import torch_ops
    ops = torch_ops.load()                                    # torch.ops.cirs_b200
    rew, done = ops.env_step_kuaishou(torch_ops.handle(env._struct), act_i32, None, env.active, 0)

Descriptor structs (environment tables, network weights, PPO configuration) stay owned by the host objects and are
passed as opaque integer handles (``handle(struct)`` = the address of the ctypes struct)."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "torch_ops.cpp")
LIB = os.path.join(HERE, "libcirs_b200_torch.so")
_loaded = False


def build(force=False):
    """g++ over csrc/torch_ops.cpp against this interpreter's torch headers -> libcirs_b200_torch.so (in-tree)."""
    from . import _lib
    if not os.path.exists(_lib.LIB_PATH):
        raise _lib.CirsError(f"{_lib.LIB_PATH} is missing: build the CUDA library first (__graft_entry__.build())")
    hdr = os.path.join(os.path.dirname(HERE), "include", "cirs_b200.h")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return LIB
    import torch
    from torch.utils import cpp_extension as ext
    inc = [f"-I{p}" for p in ext.include_paths()]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    tlib = ext.library_paths()[0]
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + inc +
           [f"-I{os.path.join(cuda_home, 'include')}", SRC, "-o", LIB, f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10",
            "-ltorch_cuda", "-lc10_cuda", f"-L{HERE}", "-lcirs_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"])
    subprocess.run(cmd, check=True)
    return LIB


def load():
    """Register the operators (once) and return ``torch.ops.cirs_b200``.  Raises if the library has not been built."""
    global _loaded
    import torch
    if not _loaded:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} is missing: build it with cirs_codes_b200.torch_ops.build() "
                               "(__graft_entry__.build() does)")
        torch.ops.load_library(LIB)
        _loaded = True
    return torch.ops.cirs_b200


def handle(struct):
    """Opaque descriptor handle of a ctypes struct owned by a host object (env._struct, tracker._w, policy._w /
    policy._g / policy.cfg): its address.  The struct must outlive the call."""
    return ctypes.addressof(struct)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
