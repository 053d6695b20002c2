import sys, time, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, '.')
import bench
cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "configs1"])
dev = torch.device("cuda", 0)
tb = bench.tables(cfg)
env, trk, pol, buf, col = bench.setup_workload(cfg, tb, dev)
B = cfg["B"]
rng = np.random.default_rng(0)
def step():
    res = col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B))
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
    return res
for _ in range(5): step()
torch.cuda.synchronize()
# phase timing
tc = tu = 0.0
for _ in range(20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pol.update(0, buf, batch_size=cfg["batch_size"], repeat=cfg["repeat"])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tc += t1 - t0; tu += t2 - t1
print(f"collect {tc/20*1e3:.3f} ms  update {tu/20*1e3:.3f} ms  n/st {res['n/st']} turns {res['turns']}")
# graph replay alone
g = col._graph
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): g.replay()
torch.cuda.synchronize(); print(f"graph replay alone {(time.perf_counter()-t0)/20*1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
