set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_r01_c.err; tail -c 600 gpurun_out/bench_r01_c.json; tail -3 gpurun_out/bench_r01_c.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain9.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu9a.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain9.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 3 -c 1 -o gpurun_out/prof_r01_k1_bench python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu9b.log 2>&1
tail -2 gpurun_out/ncu9b.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2>&1; cat gpurun_out/bench_r01_ref.json | cut -c1-400
