#!/bin/bash
# ncu captures of the tensor-core kernels of the headline chain (B200_PROFILING.md): the same command runs first
# without ncu and must exit 0.
#   prof_r2        --set full, default cache control (caches flushed before every replay: cold-cache DRAM bytes)
#   prof_r2_warm   --cache-control none: the L2 state the chain really leaves (serpentine tile order), DRAM bytes only
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --no-sweep --no-configs --train-steps 0"
USF_GRAPHS=0 timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1
USF_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:usf_tc_ -s 34 -c 4 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1 || { tail -5 gpurun_out/ncu_full.log; exit 1; }
USF_GRAPHS=0 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:usf_ -s 36 -c 18 --csv --log-file gpurun_out/ncu_warm_dram.csv $CMD > gpurun_out/ncu_warm.log 2>&1 || { tail -5 gpurun_out/ncu_warm.log; exit 1; }
USF_GRAPHS=0 USF_TC_SERPENTINE=0 timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:usf_ -s 36 -c 18 --csv --log-file gpurun_out/ncu_warm_dram_noserp.csv $CMD > gpurun_out/ncu_warm2.log 2>&1 || true
ls -la gpurun_out/prof_r2.ncu-rep gpurun_out/ncu_warm_dram.csv
