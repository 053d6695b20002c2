timeout 300 python -m pytest tests/test_torch_ops.py -x -q > gpurun_out/r10_ops.log 2>&1; tail -25 gpurun_out/r10_ops.log
