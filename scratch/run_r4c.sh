set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4d_tests.log 2>&1; tail -4 gpurun_out/r4d_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-user-model"
timeout 300 $B > gpurun_out/r4d_bench_new.json 2> gpurun_out/r4d_bench_new.err
CIRS_F_ATM=0 timeout 300 $B > gpurun_out/r4d_bench_noatm.json 2> gpurun_out/r4d_bench_noatm.err
timeout 300 $B --config configs2 --steps 30 > gpurun_out/r4d_bench_c2.json 2> gpurun_out/r4d_bench_c2.err
python - <<'PY'
import json
for c in ("new","noatm","c2"):
    try:
        d=json.loads(open(f"gpurun_out/r4d_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), [(k[:16], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:8]])
    except Exception as e: print(c, "ERR", e)
PY
timeout 300 python scratch/head_phases.py configs1 > gpurun_out/r4d_head_phases_c1.txt 2>&1; head -14 gpurun_out/r4d_head_phases_c1.txt
