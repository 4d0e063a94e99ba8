#!/bin/bash
mkdir -p gpurun_out
PB="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --train-steps 0"
$PB > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:usf_tc_mlp -s 20 -c 1 -f -o gpurun_out/prof_mlp $PB > gpurun_out/ncu_mlp.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_mlp.log
