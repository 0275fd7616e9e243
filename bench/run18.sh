timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_tasks.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu3.log; tail -25 gpurun_out/pytest_gpu3.log
