B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-user-model"
timeout 300 $B > gpurun_out/r6d_bench_base.json 2> gpurun_out/r6d_bench_base.err
CIRS_TRK_OVERLAP=1 timeout 300 $B > gpurun_out/r6d_bench_ovl.json 2> gpurun_out/r6d_bench_ovl.err
timeout 300 $B > gpurun_out/r6d_bench_base2.json 2> gpurun_out/r6d_bench_base2.err
CIRS_TRK_OVERLAP=1 timeout 300 $B > gpurun_out/r6d_bench_ovl2.json 2> gpurun_out/r6d_bench_ovl2.err
python - <<'PY'
import json
for c in ("base","ovl","base2","ovl2"):
    try:
        d=json.loads(open(f"gpurun_out/r6d_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["gpu_launches"])
    except Exception as e: print(c, "ERR", e)
PY
tail -3 gpurun_out/r6d_bench_ovl.err
