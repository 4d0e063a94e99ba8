#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_ablate.py 2>&1 | tail -3 | cut -c1-200
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --train-steps 0 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['launch_ms_by_kind'], j['roofline']['frac'])"
