python bench/prof_multi.py 5 12 0x0,0x17 > gpurun_out/plain7.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_fast -s 2 -c 2 -o gpurun_out/prof_cfg5_v12 python bench/prof_multi.py 5 12 0x0,0x17 > gpurun_out/ncu7.log 2>&1
tail -3 gpurun_out/ncu7.log
