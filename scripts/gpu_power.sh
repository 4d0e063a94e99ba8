#!/bin/bash
# same box, alternating: does sleeping in the off-critical-path waits change sustained (power-capped) throughput?
for rep in 1 2; do
for ns in 0 100 400; do
USF_TC_BACKOFF_NS=$ns timeout 300 python bench.py --steps 400 --warmup 20 --prewarm-s 2 --train-steps 0 --no-sweep --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys,os
j=json.loads(sys.stdin.read()); print('backoff', os.environ.get('NS'), '$ns', round(j['value']/1e6,2), 'M/s', round(j['ms_per_step'],4), j['clocks'])"
done; done
