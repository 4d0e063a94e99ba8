#!/bin/bash
# ncu captures of the headline launch chain (B200_PROFILING.md): the same command runs first without ncu and must
# exit 0.  One gpurun call, one GPU.
#   launches_steady.csv  every launch of two steady steps with its device time (cold-cache, serialised: compare SHARES)
#   prof_r2.ncu-rep      --set full + source counters of the two tensor-core kernels (default cache control: caches
#                        flushed before every replay -> cold-cache DRAM bytes)
#   ncu_warm_dram*.csv   --cache-control none: DRAM bytes with the L2 state the chain really leaves behind, with and
#                        without the serpentine tile order
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --no-sweep --no-configs --train-steps 0"
CHAIN='regex:usf_tc_gemm|usf_tc_mlp|usf_convert_rows'
USF_GRAPHS=0 timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1
# chain launches before the two timed steps: calibration 18 (bf16 on 256 rows) + 3 warm-up steps x 18 = 72
USF_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$CHAIN" -s 72 -c 36 --csv --log-file gpurun_out/launches_steady.csv $CMD > gpurun_out/ncu_launches.log 2>&1 || { tail -5 gpurun_out/ncu_launches.log; exit 1; }
USF_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:usf_tc_gemm|usf_tc_mlp" -s 70 -c 4 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1 || { tail -5 gpurun_out/ncu_full.log; exit 1; }
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
USF_GRAPHS=0 timeout 600 ncu --cache-control none --clock-control none --metrics $M -k "$CHAIN" -s 72 -c 18 --csv --log-file gpurun_out/ncu_warm_dram.csv $CMD > gpurun_out/ncu_warm.log 2>&1 || { tail -5 gpurun_out/ncu_warm.log; exit 1; }
USF_GRAPHS=0 USF_TC_SERPENTINE=0 timeout 600 ncu --cache-control none --clock-control none --metrics $M -k "$CHAIN" -s 72 -c 18 --csv --log-file gpurun_out/ncu_warm_dram_noserp.csv $CMD > gpurun_out/ncu_warm2.log 2>&1 || true
ls -la gpurun_out/prof_r2.ncu-rep gpurun_out/ncu_warm_dram.csv gpurun_out/launches_steady.csv
