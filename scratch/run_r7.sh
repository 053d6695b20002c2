timeout 600 python bench.py --steps 100 --no-cpu-baseline --no-user-model > gpurun_out/r7_bench_n1.json 2> gpurun_out/r7_bench_n1.err; tail -c 300 gpurun_out/r7_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-user-model > gpurun_out/r7_bench_n2.json 2> gpurun_out/r7_bench_n2.err; tail -c 300 gpurun_out/r7_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r7_multi_tests.log 2>&1; tail -2 gpurun_out/r7_multi_tests.log
python - <<'PY'
import json
for f in ("r7_bench_n1","r7_bench_n2"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"]["env_steps_per_step"])
    print([(k, round(v["us_per_step"])) for k,v in d["kernels"].items() if "nccl" in k or "rollout" in k])
PY
