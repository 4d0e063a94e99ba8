#!/bin/bash
mkdir -p gpurun_out
T="timeout 1200 python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider -s"
$T tests/test_gpu_kernels.py > gpurun_out/t1_kernels.log 2>&1; echo "t1 exit $?" >> gpurun_out/summary.txt
$T tests/test_gpu_flow.py > gpurun_out/t2_flow.log 2>&1; echo "t2 exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --precision fp32 --rows 16384 --steps 20 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?" >> gpurun_out/summary.txt
# profiler passes (only after the plain runs above exited 0)
if grep -q "bench bf16 exit 0" gpurun_out/summary.txt; then
  PB="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline"
  $PB > gpurun_out/plain_prof.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PB > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" >> gpurun_out/summary.txt
  $PB > gpurun_out/plain_prof2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:usf_tc_gemm -s 99 -c 4 -o gpurun_out/prof_tc $PB > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
for f in t1_kernels t2_flow; do tail -n 4 gpurun_out/$f.log; done
tail -n 1 gpurun_out/bench_bf16.log | cut -c1-1500
