B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-user-model"
cp cirs_codes_b200/libcirs_b200.so /tmp/default.so
for v in default mbar1 mbar2; do
  if [ $v = default ]; then cp /tmp/default.so cirs_codes_b200/libcirs_b200.so; else cp scratch/variants/libcirs_b200_$v.so cirs_codes_b200/libcirs_b200.so; fi
  timeout 300 $B > gpurun_out/r4h_bench_$v.json 2> gpurun_out/r4h_bench_$v.err
  timeout 300 python scratch/head_phases.py configs1 > gpurun_out/r4h_phases_$v.txt 2>&1
done
cp /tmp/default.so cirs_codes_b200/libcirs_b200.so
python - <<'PY'
import json
for c in ("default","mbar1","mbar2"):
    try:
        d=json.loads(open(f"gpurun_out/r4h_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), [(k[:16], round(v["us_per_step"])) for k,v in list(d["kernels"].items())[:8]])
    except Exception as e: print(c, "ERR", e)
PY
for v in default mbar1 mbar2; do echo "== $v"; sed -n 2,40p gpurun_out/r4h_phases_$v.txt; done
