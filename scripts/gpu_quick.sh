#!/bin/bash
# quick check: GPU tests, smoke, a short bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | grep smoke
timeout 600 python bench.py --steps 50 --warmup 10 --train-steps 0 --no-sweep 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['clocks'], j['roofline']['launch_ms_by_kind'], j['roofline']['frac'], j.get('cpu_baseline'))"
