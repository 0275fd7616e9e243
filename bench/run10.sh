timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench/sweep_tma.py --stages 0,2,3 > gpurun_out/sweep_tma.log 2>&1; cat gpurun_out/sweep_tma.log
