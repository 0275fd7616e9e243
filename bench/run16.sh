timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench/sweep_tma.py --stages 0 --flags 0x0,0x1f,0x3f > gpurun_out/sweep_tma.log 2>&1; cat gpurun_out/sweep_tma.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_r01_e.err; tail -c 900 gpurun_out/bench_r01_e.json
