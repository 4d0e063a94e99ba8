#!/bin/bash
# Round-2 end-of-round evidence on one B200 (one gpurun call): GPU tests, smoke, bench (ours + reference arm), the
# training step's timeline / captured-graph critical path, the LU inverse and small-stack checks, and the ncu launch list
# of two steady scoring steps.  Everything lands in gpurun_out/ (copied into profiles/r2/ afterwards).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
USF_GRAPH_DOT=gpurun_out/train_step.dot TIMELINE_DUMP=gpurun_out/train_step_events.txt timeout 200 python scripts/train_timeline.py 4096 > gpurun_out/train_step_timeline.txt 2>&1; echo "train timeline exit $?" >> gpurun_out/summary.txt
timeout 100 python scripts/graph_critical_path.py gpurun_out/train_step.dot gpurun_out/train_step_events.txt > gpurun_out/train_step_critical_path.txt 2>&1
timeout 100 python scripts/lu_inverse_check.py > gpurun_out/lu_inverse.txt 2>&1; echo "lu inverse exit $?" >> gpurun_out/summary.txt
timeout 200 python scripts/train_grad_diag.py > gpurun_out/train_grad_diag.txt 2>&1; echo "grad diag exit $?" >> gpurun_out/summary.txt
if grep -q "bench exit 0" gpurun_out/summary.txt; then
  CMD="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --no-sweep --no-configs --train-steps 0"
  CHAIN='regex:usf_tc_gemm|usf_tc_mlp|usf_convert_rows'
  USF_GRAPHS=0 timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  USF_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$CHAIN" -s 72 -c 36 --csv --log-file gpurun_out/launches_steady.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
tail -n 2 gpurun_out/gpu_tests.log
grep -i smoke gpurun_out/smoke.log | tail -2
grep "ms/step" gpurun_out/train_step_timeline.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','clocks','gpu_launches')})
print(j['e2e']); print(j['train']); print(j['roofline'].get('frac'), j['roofline'].get('by_kind', {}).keys())
r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print({k:r.get(k) for k in ('impl','value','unit','ms_per_step')})
PY
