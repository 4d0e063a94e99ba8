#!/bin/bash
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider"
$T tests/test_gpu_kernels.py -k tcgen05 > gpurun_out/t_tc_cg2.log 2>&1; echo "tc cg2 exit $?" >> gpurun_out/summary.txt
USF_TC_CTA_GROUP=1 $T tests/test_gpu_kernels.py -k tcgen05 > gpurun_out/t_tc_cg1.log 2>&1; echo "tc cg1 exit $?" >> gpurun_out/summary.txt
$T tests/test_gpu_flow.py > gpurun_out/t_flow_cg2.log 2>&1; echo "flow cg2 exit $?" >> gpurun_out/summary.txt
USF_TC_CTA_GROUP=1 $T tests/test_gpu_flow.py -k "against_oracle or golden_bf16" > gpurun_out/t_flow_cg1.log 2>&1; echo "flow cg1 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --train-steps 0 --no-cpu-baseline > gpurun_out/bench_cg2.log 2>&1; echo "bench cg2 exit $?" >> gpurun_out/summary.txt
USF_TC_CTA_GROUP=1 timeout 600 python bench.py --train-steps 0 --no-cpu-baseline > gpurun_out/bench_cg1.log 2>&1; echo "bench cg1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/t_tc_cg2.log gpurun_out/t_flow_cg2.log | grep -E "passed|failed"
for f in bench_cg2 bench_cg1; do tail -n 1 gpurun_out/$f.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print({k:j[k] for k in ('value','ms_per_step')}, j['roofline']['launch_ms_by_kind'], round(j['roofline']['frac'],3))
except Exception as e: print('no json', e)
"; done
