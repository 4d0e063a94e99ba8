#!/bin/bash
mkdir -p gpurun_out
T="timeout 1200 python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider"
$T tests/test_gpu_flow.py -k "against_oracle or edge" > gpurun_out/t2_flow.log 2>&1; echo "t2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/t2_flow.log
tail -n 1 gpurun_out/bench_bf16.log | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print({k:j[k] for k in ('value','ms_per_step','clocks','e2e','train') if k in j})
print(j.get('cpu_baseline'))
print(j['roofline']['launch_ms_by_kind'], j['roofline']['frac'])
"
