bench/diag_div > gpurun_out/diag_div.log 2>&1; cat gpurun_out/diag_div.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_r01_d.err; tail -c 1200 gpurun_out/bench_r01_d.json; tail -3 gpurun_out/bench_r01_d.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain13.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu13a.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain13.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_tma -s 3 -c 1 -o gpurun_out/prof_r01b_k1_tma_bench python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu13b.log 2>&1
tail -2 gpurun_out/ncu13b.log
