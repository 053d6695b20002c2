"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU, NCCL): the data-parallel
PPO update (environments sharded over ranks, gradient all-reduce per minibatch issued from C inside
cirs_ppo_learn, tracker-gradient / moment / loss reductions) must reproduce the single-process update on the union
of the shards -- losses, actor / critic parameters, tracker parameters and return statistics.

Every rank builds the SAME 2B-environment collect in its own process (deterministic seeds), runs the single-process
update on it as the reference, then keeps only its own B environments' slice of the replay buffer and runs the
distributed update with local permutations; the single-process run is fed the matching global permutation and
minibatch sizes (global minibatch j = union over ranks of local chunk j)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cirs_codes_b200 as cb  # noqa: E402
from cirs_codes_b200 import parallel  # noqa: E402
from tests import goldutil as G, gpu_harness as H  # noqa: E402


def build(z, c, B):
    env, trk = H.make_env(z, c, B=B), H.make_tracker(z, c, B=B, prefix=None)
    pol = H.make_policy(None, c, trk)
    buf = cb.VectorReplayBuffer(B * c["T"], B)
    col = cb.Collector(pol, env, buf, preprocess_fn=trk.build_state)
    return env, trk, pol, buf, col


def main():
    rank, world = parallel.init_from_env("nccl")
    import torch.distributed as dist
    assert world >= 2
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    B = 24
    z, c = H.synthetic_case(U=64, I=300, B=B * world, T=10, N=3, thr=1, d=32, nhead=4, seed=5, batch_size=64,
                            repeat=2)
    BA = B * world
    # ---- the union collect, identical in every process
    env, trk, pol, buf, col = build(z, c, BA)
    sd = trk.state_dict()
    g = torch.Generator().manual_seed(3)
    sd["embedding_dict.feat_user.weight"] = torch.randn(c["U"], c["d"], generator=g) * 0.1
    sd["embedding_dict.feat_item.weight"] = torch.randn(c["I"], c["d"], generator=g) * 0.1
    trk.load_state_dict(sd)
    users = np.random.default_rng(1).integers(0, c["U"], size=BA)
    col.collect(n_episode=BA, users=users)
    lens = buf._lengths.copy()
    L = buf.sub_size
    n_all = [int(lens[r * B:(r + 1) * B].sum()) for r in range(world)]
    n = int(lens.sum())
    # ---- plan: local chunks per rank, global minibatch = union of the ranks' chunks
    plans = [parallel.plan_from_counts(n_all, r, c["batch_size"]) for r in range(world)]
    n_glob = plans[0][1]
    starts = np.concatenate([[0], np.cumsum(n_all)])
    prng = np.random.default_rng(17)
    local_perms = [[prng.permutation(n_all[r]) for r in range(world)] for _ in range(c["repeat"])]
    global_perms = []
    for rep in range(c["repeat"]):
        chunks = []
        offs = [np.concatenate([[0], np.cumsum(plans[r][0])]) for r in range(world)]
        for j in range(len(n_glob)):
            for r in range(world):
                chunks.append(starts[r] + local_perms[rep][r][offs[r][j]:offs[r][j + 1]])
        global_perms.append(np.concatenate(chunks))
    assert all(sorted(p.tolist()) == list(range(n)) for p in global_perms)

    # ---- reference: single-process update on the union (process group hidden from the policy)
    pol._world = lambda: (None, 1)
    ref = pol.update(0, buf, batch_size=c["batch_size"], repeat=c["repeat"], perms=global_perms, mb_sizes=n_glob)
    ref_flat, ref_trk, ref_rms = pol.flat.clone(), trk.flat.clone(), pol.ret_rms.t.clone()

    # ---- distributed: this rank's shard of the same buffer
    env2, trk2, pol2, buf2, col2 = build(z, c, B)
    trk2.load_state_dict(sd)
    buf2._alloc(trk2.dim_state)
    sl = slice(rank * B * L, (rank + 1) * B * L)
    for name in ("obs", "obs_next", "d_act", "d_rew", "d_done"):
        getattr(buf2, name).copy_(getattr(buf, name)[sl])
    buf2.d_len.copy_(buf.d_len[rank * B:(rank + 1) * B])
    buf2.d_users.copy_(buf.d_users[rank * B:(rank + 1) * B])
    buf2.set_from_device(lens[rank * B:(rank + 1) * B])
    out = pol2.update(0, buf2, batch_size=c["batch_size"], repeat=c["repeat"],
                      perms=[local_perms[rep][rank] for rep in range(c["repeat"])])
    assert pol2._comm() is not None, "the NCCL communicator of csrc/comm.cu was not used"
    torch.cuda.synchronize()
    for k in ("loss", "loss/clip", "loss/vf", "loss/ent"):
        G.assert_close(out[k], ref[k], 1e-5, 1e-5, what=f"rank {rank} {k}")
    G.assert_close(pol2.ret_rms.t.cpu().numpy(), ref_rms.cpu().numpy(), 1e-9, what="ret_rms")
    G.assert_close(pol2.flat.cpu().numpy(), ref_flat.cpu().numpy(), 1e-5, 2e-5, what="actor / critic parameters")
    # the key bias of self-attention has an identically zero gradient (softmax shift invariance): Adam turns rounding
    # noise there into +-lr steps, in the reference too -- excluded
    a, b = trk2.flat.cpu().numpy().copy(), ref_trk.cpu().numpy().copy()
    d = c["d"]
    for l in range(2):
        seg = trk2.layout.segs[f"transformer_encoder.layers.{l}.self_attn.in_proj_bias"]
        a[seg.offset + d:seg.offset + 2 * d] = 0
        b[seg.offset + d:seg.offset + 2 * d] = 0
    G.assert_close(a, b, 1e-5, 2e-5, what="tracker parameters")
    # every rank ends with bit-identical parameters (replicas must not drift)
    mine = torch.cat([pol2.flat, trk2.flat])
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), "replicas diverged"
    # ---- the collect's count exchange (PPOPolicy.post_collect: the counts ride on the collect's read-back)
    col2.collect(n_episode=B, users=users[rank * B:(rank + 1) * B])
    assert pol2._n_all_pin is not None
    counts = [None] * world
    dist.all_gather_object(counts, int(buf2._lengths.sum()))
    assert pol2._n_all_pin[0].numpy().tolist() == counts, (pol2._n_all_pin[0].numpy().tolist(), counts)
    out2 = pol2.update(0, buf2, batch_size=c["batch_size"], repeat=c["repeat"])
    torch.cuda.synchronize()
    assert np.isfinite(np.asarray(out2["loss"])).all() and pol2._n_all.tolist() == counts
    mine = torch.cat([pol2.flat, trk2.flat])
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), "replicas diverged after the collect-driven update"
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world} n={n} n_all={n_all} minibatches={len(n_glob)} "
              f"loss={out['loss'][0]:.6f} ref={ref['loss'][0]:.6f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
