#!/bin/bash
mkdir -p gpurun_out
T="timeout 1200 python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider -s"
$T tests/test_gpu_flow.py > gpurun_out/t2_flow.log 2>&1; echo "t2 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_bf16_short.log 2>&1; echo "bench bf16 short exit $?" >> gpurun_out/summary.txt
if grep -q "bench bf16 exit 0" gpurun_out/summary.txt; then
  PB="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline"
  $PB > gpurun_out/plain_prof.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"usf_tc_gemm|usf_convert_rows" -s 68 -c 68 --csv --log-file gpurun_out/launches_steady.csv $PB > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/t2_flow.log
tail -n 1 gpurun_out/bench_bf16.log | cut -c1-1800
tail -n 1 gpurun_out/bench_bf16_short.log | cut -c1-400
