timeout 900 python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; tail -c 300 gpurun_out/r2d_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_ref.err; tail -c 200 gpurun_out/r2d_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"], d["steps"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["clocks"])
print(sorted(d.keys()))
r=json.loads(open("gpurun_out/r2d_bench_ref.json").read().strip().splitlines()[-1]); print(r["value"], r["impl"], r["e2e"])
PY
