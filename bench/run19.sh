timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench/bench_k2k3.py > gpurun_out/bench_k2k3.log 2>&1; cat gpurun_out/bench_k2k3.log
