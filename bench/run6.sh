python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench/diag_plogp.py > gpurun_out/diag_plogp.log 2>&1
python bench/sweep_k1.py --out gpurun_out/sweep_nostats.json > gpurun_out/sweep_nostats.log 2>&1
python bench/sweep_k1.py --stats --out gpurun_out/sweep_stats.json > gpurun_out/sweep_stats.log 2>&1
for f in 0x01 0x07 0x0f 0x17 0x27; do python bench/sweep_k1.py --flags $f --configs 2,5 --variants 2,10 >> gpurun_out/flags_breakdown.log 2>&1; done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_b.json 2> gpurun_out/bench_r01_b.err; tail -c 1500 gpurun_out/bench_r01_b.json; tail -5 gpurun_out/bench_r01_b.err
