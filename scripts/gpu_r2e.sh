#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --train-steps 0 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['launch_ms_by_kind'], j['roofline']['frac'])"
timeout 300 python scripts/mlp_trace.py 2>&1 | tail -7 | cut -c1-1200
