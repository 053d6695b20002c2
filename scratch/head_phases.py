"""Per-phase cycle counters of the tensor-core head passes F (128-column tiles) and B2 (bulk-copy fed) over a few configs[1] updates
(cirs_head_tc_debug_phases).  Usage: python scratch/head_phases.py [configs1|configs2]"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from cirs_codes_b200 import _lib
cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "configs1"])
dev = torch.device("cuda", 0)
tb = bench.tables(cfg)
env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
B = cfg["B"]
rng = np.random.default_rng(0)
def step():
    col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B))
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
for _ in range(3): step()
lib = _lib.load()
out = (ctypes.c_int64 * 64)()
lib.cirs_head_tc_debug_phases(1, None, 1)
N = 5
for _ in range(N): step()
lib.cirs_head_tc_debug_phases(0, out, 1)
c = np.array(list(out), dtype=np.float64)
def show(title, names, base, tiles):
    print(title, "(cycles per tile of one CTA; tiles counted: %d)" % tiles)
    tot = 0.0
    for i, n in enumerate(names):
        if n: print(f"   {n:34s} {c[base + i] / max(tiles, 1):9.0f}"); tot += c[base + i]
    print(f"   {'sum':34s} {tot / max(tiles, 1):9.0f}")
tF = c[4]
show("pass F issuer", ["wait tma_b", "wait dfree", "issue MMA + commit", "wait mma(t-1) + copy"], 0, tF)
show("pass F worker warp 0", ["wait mma", "tmem_ld", "bias+max, arrive", "exp-sum"], 8, tF)
tB = c[25]
show("pass B2 issuer", ["(start)", "wait dlr[nb]", "wait tma_n", "issue MMA1", "wait mma1 + copy_n", "wait mma2 + copy_k",
                        "wait dlr[b]", "wait tma_k", "issue MMA2"], 16, tB)
show("pass B2 worker warp 0", ["wait mma1", "tmem_ld", "compute + store_dl", "arrive + wait mma2"], 32, tB)
show("pass B2 worker warp 15", ["wait mma1", "tmem_ld", "compute + store_dl", "arrive + wait mma2"], 40, tB)
