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
col.persistent=False; col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B)); g = col._graph
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): g.replay()
torch.cuda.synchronize(); print(f"graph replay alone {(time.perf_counter()-t0)/20*1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
# rollout variants
import time
for name, kw in (("persistent", dict(persistent=True)), ("graph", dict(persistent=False, use_graph=True)), ("eager", dict(persistent=False, use_graph=False))):
    col.persistent, col.use_graph = kw.get("persistent", False), kw.get("use_graph", False)
    for _ in range(3): col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B))
    torch.cuda.synchronize(); t0 = time.perf_counter(); st = 0; tr = 0
    for _ in range(20):
        r = col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B)); st += r["n/st"]; tr += r["turns"]
    torch.cuda.synchronize(); print(f"collect[{name}] {(time.perf_counter()-t0)/20*1e3:.3f} ms  mean steps {st/20:.0f} mean turns {tr/20:.1f}")
col.persistent = True
r = col.collect(n_episode=B, users=rng.integers(0, cfg["U"], size=B))
dbg = col._f["ws_roll"][256:256 + 8 * (1 + 6 * 512)].view(torch.int64).cpu().numpy()
nt = int(dbg[0]); print("persistent turns", nt, "lens max", r["turns"])
for t in range(nt): print(f"  turn {t:2d} n_act {dbg[1+3*t]:5d}  phaseA {dbg[2+3*t]/1e3:7.1f} us  phaseB {dbg[3+3*t]/1e3:7.1f} us")
names = ["combine+env", "(token entry)", "tok-in"] + [f"L{l}:{n}" for l in range(2) for n in ("inproj", "attn", "outproj", "ln1", "l1", "l2", "ln2")] + ["dec+store", "(dup)"]
for label, base in (("last turn", 1 + 3 * 512), ("turn 0", 1 + 3 * 512 + 64)):
    tq = dbg[base: base + 23]
    print(f"phase-B stage times of CTA 0, {label} (us):")
    print("   " + "  ".join(f"{names[i-1]} {(tq[i]-tq[i-1])/1e3:.2f}" for i in range(1, 20)))
    print(f"   trunk: {(tq[22]-tq[19])/1e3:.2f}   total {(tq[22]-tq[0])/1e3:.2f}")

ta = dbg[1 + 3 * 512 + 32: 1 + 3 * 512 + 32 + 7]
print("phase-A stamps of CTA 0, last turn (us): turn start -> entry %.2f, stage h2 tile %.2f, MMA %.2f, epilogue %.2f, merge+partials %.2f, fence+grid.sync %.2f" % (
    (ta[0]-ta[6])/1e3, (ta[1]-ta[0])/1e3, (ta[2]-ta[1])/1e3, (ta[3]-ta[2])/1e3, (ta[4]-ta[3])/1e3, (ta[5]-ta[4])/1e3))
