timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r01_2gpu.json 2> gpurun_out/bench_r01_2gpu.err; tail -c 1500 gpurun_out/bench_r01_2gpu.json; grep -v "^$" gpurun_out/bench_r01_2gpu.err | tail -5
timeout 600 python bench/sweep_tma.py --stages 0 --flags 0x0,0x1f,0x3f > gpurun_out/sweep_tma.log 2>&1; cat gpurun_out/sweep_tma.log
