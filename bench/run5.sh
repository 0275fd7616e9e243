python bench/prof_multi.py 5 10 0x0,0x1,0x1f > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 3 -c 3 -o gpurun_out/prof_cfg5_v10 python bench/prof_multi.py 5 10 0x0,0x1,0x1f > gpurun_out/ncu5.log 2>&1
tail -3 gpurun_out/ncu5.log
