timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r8_tests.log 2>&1; tail -3 gpurun_out/r8_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r8_smoke.log 2>&1; tail -1 gpurun_out/r8_smoke.log
timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-user-model > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; tail -c 300 gpurun_out/r8_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r8_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"], [(k[:18], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:7]])
PY
