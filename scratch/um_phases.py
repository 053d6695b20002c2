"""Per-phase cycle breakdown of the user-model pair kernel (CIRS_UM_FLAGS=4: thread 0 of CTA 0, clock() per phase)."""
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cirs_codes_b200 import _lib, user_model as um
from oracle import user_model as oum
lib = _lib.load()
U, I = 7176, 10728
rng = np.random.Generator(np.random.PCG64(2023))
P = oum.synth_params(U, I + 1, 32, seed=2023)
w = um.UserModelWeights(P)
dev = w.device
users = torch.arange(U, dtype=torch.int32, device=dev); items = torch.arange(1, I + 1, dtype=torch.int32, device=dev)
feat = torch.from_numpy(rng.integers(0, 32, (I, 4)).astype(np.int32)).to(dev)
dense = torch.from_numpy(rng.uniform(3, 60, (I, 1)).astype(np.float32)).to(dev)
out = torch.empty((U, I), device=dev)
ws = torch.empty(lib.cirs_user_model_workspace_bytes(U, I, 16), dtype=torch.uint8, device=dev)
def run(flags, reps=5):
    os.environ["CIRS_UM_FLAGS"] = str(flags)
    for _ in range(2):
        um.predict_all(w, users, items, feat, dense, normalise=False, out=out, workspace=ws)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); um.predict_all(w, users, items, feat, dense, normalise=False, out=out, workspace=ws); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))
res = {"ms": run(0), "ms_two_groups": run(8)}
for fl in (4, 12):
    run(fl, 1)
    buf = (C.c_int64 * 8)()
    lib.cirs_user_model_debug_phases(buf)
    v = list(buf); n = max(v[7], 1)
    res["phases_flags%d" % fl] = dict(zip(["stage", "sync1", "issue", "fm_prefetch", "mma_wait", "epilogue", "sync2"], [round(x / n, 1) for x in v[:7]]), tiles=v[7], per_tile=round(sum(v[:7]) / n, 1))
print(json.dumps(res))
