#!/bin/bash
# Evidence for profiles/ (run through gpurun on one B200):  bash bench/profile_round.sh <tag>
# 1. the bench line, 2. every launch of a short bench run with its device time, 3. ncu --set full of one k1_tma launch
# of the bench workload.  Each ncu command runs only after the same command exited 0 without ncu.
tag=${1:-rXX}
out=gpurun_out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err || exit 1
short="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
$short > $out/${tag}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file $out/${tag}_launches_bench.csv $short > $out/${tag}_ncu_list.log 2>&1
$short > $out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_tma -s 3 -c 1 \
    -o $out/${tag}_k1_tma_bench $short > $out/${tag}_ncu_full.log 2>&1
python bench/bench_configs.py --logits > $out/${tag}_configs.jsonl 2>&1
python bench/bench_k2k3.py > $out/${tag}_k2_k3_k4_timing.txt 2>&1
tail -c 700 $out/${tag}_bench_line.json
