CIRS_B2_WIDE=1 timeout 150 python -m pytest tests/test_gpu_head_tc.py -x -q > gpurun_out/r11_b2w_tests.log 2>&1; tail -4 gpurun_out/r11_b2w_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-user-model"
if grep -q " passed" gpurun_out/r11_b2w_tests.log && ! grep -q "failed" gpurun_out/r11_b2w_tests.log; then
  CIRS_B2_WIDE=1 timeout 120 $B > gpurun_out/r11_bench_wide.json 2> gpurun_out/r11_bench_wide.err
  timeout 120 $B > gpurun_out/r11_bench_base.json 2> gpurun_out/r11_bench_base.err
  python - <<'PY'
import json
for c in ("wide","base"):
    try:
        d=json.loads(open(f"gpurun_out/r11_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], [(k[:22], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:6]])
    except Exception as e: print(c, "ERR", e)
PY
fi
