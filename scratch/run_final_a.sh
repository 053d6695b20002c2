set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; tail -3 gpurun_out/r2c_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; tail -2 gpurun_out/r2c_smoke.log
timeout 600 python bench.py > gpurun_out/r2c_bench_configs1.json 2> gpurun_out/r2c_bench_configs1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_ref.err
for c in configs2 configs4 configs3 configs0; do
  timeout 600 python bench.py --config $c --steps 60 --warmup 5 --no-user-model > gpurun_out/r2c_bench_$c.json 2> gpurun_out/r2c_bench_$c.err
done
python - <<'PY'
import json
for c in ("configs1","configs2","configs4","configs3","configs0","reference_arm"):
    try:
        d=json.loads(open(f"gpurun_out/r2c_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d.get("ms_per_step"), round(d["e2e"]["value"]), d.get("gpu_launches"), d["config"].get("env_steps_per_step"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(c, "ERR", e)
PY
