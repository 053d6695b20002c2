"""Host-side timeline of one benchmark iteration: every C-ABI call and stream synchronisation with its start / end
(perf_counter, us), so that the Python time BETWEEN them is visible.  python scratch/host_trace.py [config]"""
import sys, time
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
users = rng.integers(0, cfg["U"], size=B)
col.collect(n_episode=B, users=users)
fr = bench.Frozen(pol, trk, col)
def step():
    fr.restore()
    res = col.collect(n_episode=B, users=users)
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
for _ in range(10): step()
torch.cuda.synchronize()
log = []
orig_call = _lib.call
def call(name, *a):
    t0 = time.perf_counter_ns(); r = orig_call(name, *a); log.append((name, t0, time.perf_counter_ns())); return r
_lib.call = call
import cirs_codes_b200.collector as cc, cirs_codes_b200.policy as pp, cirs_codes_b200.state_tracker as ss, cirs_codes_b200.data as dd
for m in (cc, pp, ss, dd):
    if hasattr(m, "_lib"): m._lib.call = call
orig_sync = torch.cuda.Stream.synchronize
def sync(self):
    t0 = time.perf_counter_ns(); r = orig_sync(self); log.append(("SYNC", t0, time.perf_counter_ns())); return r
torch.cuda.Stream.synchronize = sync
orig_esync = torch.cuda.Event.synchronize
def esync(self):
    t0 = time.perf_counter_ns(); r = orig_esync(self); log.append(("EVENT SYNC", t0, time.perf_counter_ns())); return r
torch.cuda.Event.synchronize = esync
for it in range(3):
    fr.restore(); torch.cuda.synchronize()
    log.clear()
    t00 = time.perf_counter_ns()
    res = col.collect(n_episode=B, users=users)
    log.append(("-- collect returned", time.perf_counter_ns(), time.perf_counter_ns()))
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
    t11 = time.perf_counter_ns()
    if it == 2:
        prev = t00
        for name, a, b in log:
            print(f"{(a - t00) / 1e3:9.1f} us  +{(a - prev) / 1e3:7.1f} python | {name:34s} {(b - a) / 1e3:8.1f} us")
            prev = b
        print(f"total {(t11 - t00) / 1e3:.1f} us")
