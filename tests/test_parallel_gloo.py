"""world_size-2 gloo tests (CPU) of the host-side logic of the env-sharded data-parallel update: all ranks agree on
the number of global minibatches, local chunks partition the local transitions, and the all-reduced raw moments
reproduce the single-process statistics of the union (advantage mean / unbiased std per minibatch, return mean /
var for RunningMeanStd) -- the quantities the CUDA path reduces with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, batch_size, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from cirs_codes_b200 import parallel
    r, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n = n_per_rank[rank]
    sizes = parallel.sharded_sizes(n, batch_size, dist, None, "cpu")
    # the same plan from the gathered counts (what policy.process_fn / learn use): no MAX all-reduce, no host sync
    cnt = torch.zeros(world, dtype=torch.float64)
    cnt[rank] = n
    dist.all_reduce(cnt)
    sizes2, n_glob = parallel.plan_from_counts(cnt.numpy(), rank, batch_size)
    assert sizes2 == sizes and sum(n_glob) == sum(n_per_rank) and len(n_glob) == len(sizes)
    rng = np.random.default_rng(100 + rank)
    adv = rng.normal(1.0, 2.0, size=n)
    perm = rng.permutation(n)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    stats = torch.zeros(len(sizes), 3, dtype=torch.float64)
    for j in range(len(sizes)):
        a = adv[perm[offs[j]:offs[j + 1]]]
        stats[j] = torch.tensor([len(a), a.sum(), (a * a).sum()])
    dist.all_reduce(stats)
    ret = rng.normal(3.0, 1.5, size=n)
    mom = torch.tensor([ret.sum(), (ret * ret).sum(), float(n)], dtype=torch.float64)
    dist.all_reduce(mom)
    q.put((rank, sizes, stats.numpy(), mom.numpy(), adv, perm, ret))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_minibatches_and_moments_world2():
    world, batch_size = 2, 16
    n_per_rank = [37, 90]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, batch_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, st0, m0, adv0, perm0, ret0), (_, s1, st1, m1, adv1, perm1, ret1) = got
    from cirs_codes_b200 import parallel
    # same number of minibatches = max of the ranks' own reference-split counts; chunks partition local data
    assert len(s0) == len(s1) == max(len(parallel.split_sizes(n, batch_size)) for n in n_per_rank)
    assert sum(s0) == n_per_rank[0] and sum(s1) == n_per_rank[1] and max(s0) - min(s0) <= 1
    assert np.allclose(st0, st1) and np.allclose(m0, m1)
    o0, o1 = np.concatenate([[0], np.cumsum(s0)]), np.concatenate([[0], np.cumsum(s1)])
    for j in range(len(s0)):
        union = np.concatenate([adv0[perm0[o0[j]:o0[j + 1]]], adv1[perm1[o1[j]:o1[j + 1]]]])
        cnt, s, ss = st0[j]
        mean = s / cnt
        std = np.sqrt((ss - cnt * mean * mean) / (cnt - 1))
        assert cnt == len(union) and np.isclose(mean, union.mean()) and np.isclose(std, union.std(ddof=1))
    allret = np.concatenate([ret0, ret1])
    bm, bv = m0[0] / m0[2], m0[1] / m0[2] - (m0[0] / m0[2]) ** 2
    assert m0[2] == len(allret) and np.isclose(bm, allret.mean()) and np.isclose(bv, allret.var())


def test_single_process_split_is_the_reference_split():
    from cirs_codes_b200 import parallel
    assert parallel.sharded_sizes(10, 3) == [3, 3, 4]
    assert parallel.sharded_sizes(12, 4) == [4, 4, 4]
    assert parallel.sharded_sizes(7, 16) == [7]
    assert parallel.even_sizes(10, 4) == [3, 3, 2, 2] and parallel.even_sizes(2, 4) == [1, 1, 0, 0]
