"""Environment-sharded data parallelism: one process per GPU, environments partitioned over ranks, parameters
replicated.  The reference has no multi-GPU path; this is the new part of the design.

Rollout needs no communication.  During the update every GLOBAL minibatch j is the union over ranks of the ranks'
local chunk j, so all ranks must agree on the number of minibatches even though they hold different numbers of
transitions (episode lengths differ).  Collectives per update:
  * 1 x all-reduce(SUM) of the return moments + every rank's transition count (policy.process_fn); the minibatch
    plan follows from the counts without further communication (plan_from_counts)
  * per repeat 1 x all-reduce(SUM) of the advantage moments of all minibatches   (policy.learn)
  * per minibatch ONE all-reduce(SUM) of the flat actor/critic gradient    (policy.learn; NCCL over NVLink)
  * 1 x all-reduce(SUM) of the tracker's flat gradient, 1 x of the losses  (policy.learn)
All helpers take the torch.distributed module explicitly so that the same code runs on NCCL (GPU) and gloo (CPU
tests, world_size 2).
"""
import numpy as np
import torch


def split_sizes(n, size):
    """Chunk sizes of tianshou's Batch.split(size, merge_last=True) (tianshou/data/batch.py:721-744)."""
    if n <= 0:
        return []
    k = n // size
    if n % size == 0:
        return [size] * k
    if k <= 1:
        return [n]
    return [size] * (k - 1) + [size + n % size]


def even_sizes(n, parts):
    """np.array_split sizes: ``parts`` chunks whose sizes differ by at most one (empty chunks allowed)."""
    base, extra = divmod(n, parts)
    return [base + (1 if i < extra else 0) for i in range(parts)]


def sharded_sizes(n_local, batch_size, dist=None, group=None, device="cpu"):
    """Local chunk sizes for this rank.  Single process: the reference's split.  Multi-process: every rank cuts its
    local transitions into the same number of chunks, max over ranks of its own reference-split count."""
    local = split_sizes(n_local, batch_size)
    if dist is None:
        return local
    t = torch.tensor([len(local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return even_sizes(n_local, int(t.item()))


def plan_from_counts(n_all, rank, batch_size):
    """Minibatch plan of one update when every rank knows every rank's transition count (``n_all``, gathered once
    per update together with the return moments): the number of global minibatches is the max over ranks of the
    reference's own split count, every rank cuts its transitions into that many near-equal chunks, and the size of
    global minibatch j is the sum of the ranks' j-th chunks -- all without further communication or host syncs.
    Returns (local chunk sizes of ``rank``, global minibatch sizes)."""
    n_all = [int(x) for x in n_all]
    n_mb = max(len(split_sizes(n, batch_size)) for n in n_all)
    n_mb = max(n_mb, 1)
    per_rank = [even_sizes(n, n_mb) for n in n_all]
    n_glob = [sum(pr[j] for pr in per_rank) for j in range(n_mb)]
    return per_rank[rank], n_glob


def init_from_env(backend="nccl"):
    """torch.distributed init from torchrun's environment (RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world)."""
    import os
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def create_comm(dist, group, device):
    """The C-side NCCL communicator (csrc/comm.cu) for ``group``: rank 0 draws the unique id, torch.distributed
    broadcasts its 128 bytes, every rank initialises with its CUDA device current.  Returns the opaque handle."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.call("cirs_comm_unique_id", C.cast(ident, C.c_void_p))
    device = torch.device("cuda", torch.cuda.current_device())   # one process per GPU: the current device is the rank's
    t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=device)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    raw = bytes(t.cpu().tolist())
    ident = (C.c_ubyte * 128).from_buffer_copy(raw)
    handle = C.c_void_p()
    _lib.call("cirs_comm_create", C.cast(ident, C.c_void_p), rank, world, C.byref(handle))
    assert lib is not None and handle.value
    return handle
