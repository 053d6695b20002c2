set -x
timeout 120 scratch/bulk_bench > gpurun_out/r4_bulk_bench.txt 2>&1; tail -5 gpurun_out/r4_bulk_bench.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_tests.log 2>&1; tail -4 gpurun_out/r4a_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-user-model"
timeout 300 $B > gpurun_out/r4a_bench_new.json 2> gpurun_out/r4a_bench_new.err
CIRS_F_RING=0 timeout 300 $B > gpurun_out/r4a_bench_noring.json 2> gpurun_out/r4a_bench_noring.err
CIRS_TC_PLAN=0 timeout 300 $B > gpurun_out/r4a_bench_oldplan.json 2> gpurun_out/r4a_bench_oldplan.err
timeout 300 $B --config configs2 --steps 30 > gpurun_out/r4a_bench_c2.json 2> gpurun_out/r4a_bench_c2.err
CIRS_TC_PLAN=0 CIRS_F_RING=0 timeout 300 $B --config configs2 --steps 30 > gpurun_out/r4a_bench_c2_old.json 2> gpurun_out/r4a_bench_c2_old.err
python - <<'PY'
import json
for c in ("new","noring","oldplan","c2","c2_old"):
    try:
        d=json.loads(open(f"gpurun_out/r4a_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), [(k[:16], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:8]])
    except Exception as e: print(c, "ERR", e)
PY
