timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r6a_tests.log 2>&1; tail -6 gpurun_out/r6a_tests.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-user-model"
timeout 300 $B > gpurun_out/r6a_bench_pre.json 2> gpurun_out/r6a_bench_pre.err
CIRS_NO_PRE_EVAL=1 timeout 300 $B > gpurun_out/r6a_bench_nopre.json 2> gpurun_out/r6a_bench_nopre.err
timeout 300 $B --config configs2 --steps 30 > gpurun_out/r6a_bench_c2.json 2> gpurun_out/r6a_bench_c2.err
CIRS_NO_PRE_EVAL=1 timeout 300 $B --config configs2 --steps 30 > gpurun_out/r6a_bench_c2_nopre.json 2> gpurun_out/r6a_bench_c2_nopre.err
python - <<'PY'
import json
for c in ("pre","nopre","c2","c2_nopre"):
    try:
        d=json.loads(open(f"gpurun_out/r6a_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"], [(k[:16], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:6]])
    except Exception as e: print(c, "ERR", e)
PY
tail -3 gpurun_out/r6a_bench_pre.err
