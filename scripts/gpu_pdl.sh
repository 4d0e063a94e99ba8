#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for pdl in 0 1 0 1; do
USF_PDL=$pdl timeout 300 python bench.py --steps 50 --warmup 10 --train-steps 0 --no-sweep --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('pdl $pdl', round(j['value']/1e6,2), 'M/s', round(j['ms_per_step'],4), j['clocks']['sm_mhz'], j['clocks'].get('power_w'))"
done
