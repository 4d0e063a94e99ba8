#!/bin/bash
# A/B of several environment settings on one box: scripts/ab_env2.sh "A=1 B=0" "A=0" ...  -> per-kind launch times
mkdir -p gpurun_out
for rep in 1 2; do
  for cfg in "$@"; do
    env $cfg timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-sweep --no-configs --train-steps 0 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=j['roofline']
print('[$cfg] rep$rep value %.2f M/s step %.4f ms kinds %s sm %s' % (j['value']/1e6, j['ms_per_step'], {k: round(v,4) for k,v in r['launch_ms_by_kind'].items()}, j['clocks'].get('sm_mhz')))"
  done
done
