timeout 300 python scratch/host_profile.py configs1 > gpurun_out/r5b_hostprof.txt 2>&1; tail -32 gpurun_out/r5b_hostprof.txt | cut -c1-400
rm -f gpurun_out/r5b_timeline.txt
CIRS_PROFILE_TIMELINE=gpurun_out/r5b_timeline.txt timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-user-model > gpurun_out/r5b_tl_bench.json 2> gpurun_out/r5b_tl_bench.err
wc -l gpurun_out/r5b_timeline.txt
