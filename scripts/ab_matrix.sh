#!/bin/bash
# Per-kind launch times of the headline step under combinations of the tensor-core tuning switches (one box, alternating):
#   scripts/ab_matrix.sh "USF_TC_RAGGED=0 USF_TC_L2HINT=0 USF_TC_SERPENTINE=0" "USF_TC_RAGGED=1 ..." ...
mkdir -p gpurun_out
OUT=gpurun_out/ab_matrix.txt
: > $OUT
for rep in 1 2; do
  for combo in "$@"; do
    env $combo python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-sweep --no-configs --train-steps 0 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=j['roofline']
print('%-62s rep$rep value %.2f M/s step %.4f ms kinds %s sm %s MHz err %s' % ('$combo', j['value']/1e6, j['ms_per_step'], {k: round(v,4) for k,v in r['launch_ms_by_kind'].items()}, j['clocks']['sm_mhz'], j.get('bf16_calibration_err')))" >> $OUT
  done
done
cat $OUT
