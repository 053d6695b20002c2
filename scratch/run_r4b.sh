set -x
timeout 300 python scratch/head_phases.py configs1 > gpurun_out/r4b_head_phases_c1.txt 2>&1; cat gpurun_out/r4b_head_phases_c1.txt | tail -50
timeout 300 python scratch/head_phases.py configs2 > gpurun_out/r4b_head_phases_c2.txt 2>&1; cat gpurun_out/r4b_head_phases_c2.txt | tail -50
