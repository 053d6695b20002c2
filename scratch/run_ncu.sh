set -x
B="python bench.py --no-cpu-baseline --no-user-model"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv $B --steps 3 --warmup 3 > gpurun_out/r2c_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'rollout_kuaishou|tracker_chunk|head_tc_|trunk_bwd|clip_adam|tracker_dw|gae_moments|row_loss|loss_reduce|eval_merge|chunk_plan|update_plan' -s 120 -c 34 -f -o /tmp/r2c_full $B --steps 2 --warmup 3 > gpurun_out/r2c_ncu_full.log 2>&1
ncu -i /tmp/r2c_full.ncu-rep --page raw --csv > gpurun_out/r2c_full_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/ | tail -8
