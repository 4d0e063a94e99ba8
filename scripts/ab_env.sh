#!/bin/bash
# A/B of an environment switch on one box: scripts/ab_env.sh VAR "0 1" [bench args...]  -> gpurun_out/ab_VAR.txt
# (alternating runs of the headline bench, resident-input value + per-kind launch times)
VAR=$1; VALUES=$2; shift 2
mkdir -p gpurun_out
OUT=gpurun_out/ab_${VAR}.txt
: > $OUT
for rep in 1 2; do
  for v in $VALUES; do
    env $VAR=$v python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-sweep --no-configs --train-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=j['roofline']
print('$VAR=$v rep$rep value %.2f M/s step %.4f ms e2e %.2f M/s  kinds %s clocks %s' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, {k: round(v,4) for k,v in r['launch_ms_by_kind'].items()}, j['clocks']))" >> $OUT
  done
done
cat $OUT
