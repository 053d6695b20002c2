"""Which resource paces pass F's MMAs: phase counters of the stamped ring kernel with (1) no epilogue arithmetic, (2) no
ring refills, (3) both.  Runs cirs_policy_eval-like passes through a PPO update.  python scratch/f_experiments.py"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from cirs_codes_b200 import _lib
cfg = dict(bench.CONFIGS["configs1"])
dev = torch.device("cuda", 0)
tb = bench.tables(cfg)
env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
B = cfg["B"]
rng = np.random.default_rng(0)
users = rng.integers(0, cfg["U"], size=B)
col.collect(n_episode=B, users=users)
fz = bench.Frozen(pol, trk, col)
def step():
    fz.restore()
    col.collect(n_episode=B, users=users)
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
for _ in range(3): step()
lib = _lib.load()
out = (ctypes.c_int64 * 64)()
names_i = ["wait tma_b", "wait dfree", "issue MMA + commit", "wait mma(t-1) + copy"]
names_w = ["wait mma", "tmem_ld", "bias+max, arrive", "exp-sum"]
for mode, label in ((0, "normal"), (1, "no epilogue arithmetic"), (2, "no ring refills"), (3, "neither")):
    lib.cirs_head_tc_debug_phases(1 + 16 * mode, None, 1)
    for _ in range(3): step()
    lib.cirs_head_tc_debug_phases(0, out, 1)
    c = np.array(list(out), dtype=np.float64)
    T = max(c[4], 1)
    print(f"mode {mode} ({label}): issuer " + ", ".join(f"{n} {c[i] / T:.0f}" for i, n in enumerate(names_i)) + f" | sum {c[0:4].sum() / T:.0f}")
    print(f"        worker " + ", ".join(f"{n} {c[8 + i] / T:.0f}" for i, n in enumerate(names_w)) + f" | sum {c[8:12].sum() / T:.0f}")
