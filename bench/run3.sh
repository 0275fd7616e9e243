set -x
for f in 0x01 0x03 0x07 0x0f 0x17 0x27 0x3f; do python bench/sweep_k1.py --flags $f --configs 2,5 --variants 2,16 >> gpurun_out/flags_breakdown.log 2>&1; done
python bench/prof_one.py 2 0x3f 2 3 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 1 -c 1 -o gpurun_out/prof_cfg2_stats python bench/prof_one.py 2 0x3f 2 3 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
