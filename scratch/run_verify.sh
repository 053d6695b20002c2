timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r12_tests.log 2>&1; tail -3 gpurun_out/r12_tests.log
