#!/bin/bash
# Last GPU call of round 2 (one B200, ~4 min of box time): the bench with its two new lines -- the reference classes in
# torch eager ON the B200 (`cpu_baseline.torch_eager_b200`) and the training step at the reference's batch size
# (`train.small_batch`) -- then one `ncu --set full` capture of the default tier's kernel (`usf_tcb2_gemm_kernel`, bf16x2:
# affine GEMM, two hidden layers, last layer + coupling, next affine GEMM), after the same command has exited 0 without ncu.
mkdir -p gpurun_out
S=gpurun_out/last_summary.txt
: > $S
T0=$SECONDS
timeout 210 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench exit $? after $((SECONDS - T0)) s" >> $S
CMD="python bench.py --precision bf16x2 --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --no-sweep --no-configs --train-steps 0"
if [ $((SECONDS - T0)) -lt 215 ]; then
  USF_GRAPHS=0 timeout 60 $CMD > gpurun_out/last_ncu_plain.log 2>&1
  RC=$?
  echo "bf16x2 plain exit $RC after $((SECONDS - T0)) s" >> $S
  if [ $RC -eq 0 ]; then
    # 33 usf_tcb2 launches per step (8 blocks x [affine, hidden, hidden, last+coupling] + the final GEMM); skip the 3 warm-ups
    USF_GRAPHS=0 timeout 100 ncu --set full --clock-control none --import-source on -k "regex:usf_tcb2" -s 99 -c 5 \
      -o gpurun_out/prof_b2 -f $CMD > gpurun_out/last_ncu_full.log 2>&1
    echo "ncu bf16x2 exit $? after $((SECONDS - T0)) s" >> $S
  fi
fi
cat $S
python - <<'PY'
import json
try:
    j = json.loads([l for l in open('gpurun_out/last_bench.json') if l.startswith('{')][-1])
    print({k: j[k] for k in ('value', 'ms_per_step', 'clocks', 'gpu_launches')})
    print(j['e2e']['value'], j['train'].get('small_batch'), j['train']['ms_per_step'])
    print(j['cpu_baseline'].get('torch_eager_b200'))
    print(j['cpu_baseline']['value'], j['cpu_baseline']['kind'])
except Exception as e:
    print('no bench line:', e)
PY
tail -3 gpurun_out/last_bench.err
ls -la gpurun_out/prof_b2.ncu-rep 2>/dev/null
