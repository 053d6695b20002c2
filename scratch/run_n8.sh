timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --no-user-model > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err; tail -c 300 gpurun_out/r2c_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["config"]["env_steps_per_step"])
print([(k, round(v["us_per_step"])) for k,v in d["kernels"].items() if "nccl" in k or "rollout" in k])
PY
