python bench/sweep_k1.py --configs 1,2,5 --variants 0,4,10,12,17 --out gpurun_out/sweep_nostats_pf.json > gpurun_out/sweep_nostats_pf.log 2>&1
python bench/sweep_k1.py --stats --configs 1,2,5 --variants 2,4,10,12,17 --out gpurun_out/sweep_stats_pf.json > gpurun_out/sweep_stats_pf.log 2>&1
cat gpurun_out/sweep_nostats_pf.log gpurun_out/sweep_stats_pf.log
