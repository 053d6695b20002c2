rm -f gpurun_out/r3_timeline.txt
CIRS_PROFILE_TIMELINE=gpurun_out/r3_timeline.txt timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-user-model > gpurun_out/r3_tl_bench.json 2> gpurun_out/r3_tl_bench.err
tail -c 300 gpurun_out/r3_tl_bench.err
wc -l gpurun_out/r3_timeline.txt
