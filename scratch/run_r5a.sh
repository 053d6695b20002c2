timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r5a_tests.log 2>&1; tail -4 gpurun_out/r5a_tests.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-user-model"
timeout 300 $B > gpurun_out/r5a_bench_pdl.json 2> gpurun_out/r5a_bench_pdl.err
CIRS_NO_PDL=1 timeout 300 $B > gpurun_out/r5a_bench_nopdl.json 2> gpurun_out/r5a_bench_nopdl.err
timeout 300 $B --config configs2 --steps 30 > gpurun_out/r5a_bench_c2.json 2> gpurun_out/r5a_bench_c2.err
python - <<'PY'
import json
for c in ("pdl","nopdl","c2"):
    try:
        d=json.loads(open(f"gpurun_out/r5a_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), [(k[:16], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:8]])
    except Exception as e: print(c, "ERR", e)
PY
tail -3 gpurun_out/r5a_bench_pdl.err
