set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; tail -4 gpurun_out/r3a_tests.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-user-model > gpurun_out/r3a_bench_configs1.json 2> gpurun_out/r3a_bench_configs1.err; tail -c 600 gpurun_out/r3a_bench_configs1.err
timeout 300 python bench.py --config configs2 --steps 50 --warmup 5 --no-cpu-baseline --no-user-model > gpurun_out/r3a_bench_configs2.json 2> gpurun_out/r3a_bench_configs2.err
timeout 300 python scratch/host_profile.py configs1 > gpurun_out/r3a_hostprof.txt 2>&1
python - <<'PY'
import json
for c in ("configs1","configs2"):
    try:
        d=json.loads(open(f"gpurun_out/r3a_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), [(k, round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:4]])
    except Exception as e: print(c, "ERR", e)
PY
timeout 300 python scratch/host_trace.py configs1 > gpurun_out/r3_host_trace.txt 2>&1; tail -22 gpurun_out/r3_host_trace.txt
