#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --train-steps 0 --no-cpu-baseline --steps 60 --warmup 10 > gpurun_out/bench_$tag.log 2>&1; tail -n 1 gpurun_out/bench_$tag.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); r=j['roofline']['launch_ms_by_kind']; print('$tag', round(j['ms_per_step'],3), {k:round(v,4) for k,v in r.items()}, j['clocks']['sm_mhz'])
except Exception as e: print('$tag', 'no json', e)
"; }
run cg2_s7 USF_TC_MAX_STAGES=7
run cg2_s6 USF_TC_MAX_STAGES=6
run cg2_s5 USF_TC_MAX_STAGES=5
run cg2_s4 USF_TC_MAX_STAGES=4
run cg2_s3 USF_TC_MAX_STAGES=3
run cg1_s4 USF_TC_CTA_GROUP=1 USF_TC_MAX_STAGES=4
run cg2_s7b USF_TC_MAX_STAGES=7
