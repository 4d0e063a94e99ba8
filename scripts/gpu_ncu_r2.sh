#!/bin/bash
# One ncu capture of the two tensor-core kernels of the headline chain (ncu --set full, source-level counters).
# The same command runs first without ncu and must exit 0 (B200_PROFILING.md).
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --no-sweep --no-configs --train-steps 0"
USF_GRAPHS=0 $CMD > gpurun_out/ncu_plain.log 2>&1
USF_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:usf_tc_ -s 34 -c 4 -o gpurun_out/prof_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1 || { tail -5 gpurun_out/ncu_full.log; exit 1; }
ls -la gpurun_out/prof_r2.ncu-rep
