#!/bin/bash
# The round's remaining ~3 GPU-minutes: the TF32 tensor peak BASELINE.md section 2 leaves to the builder, then the bench
# once more with the training baselines of BASELINE.md section 3 in its cpu_baseline leg (cpu_baseline.train).
mkdir -p gpurun_out
T0=$SECONDS
timeout 40 python scripts/tf32_peak.py 2 > gpurun_out/tf32_peak.json 2> gpurun_out/tf32_peak.err
echo "tf32 peak exit $? after $((SECONDS - T0)) s" > gpurun_out/last2_summary.txt
LEFT=$((160 - (SECONDS - T0)))
timeout $LEFT python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last2_bench.json 2> gpurun_out/last2_bench.err
echo "bench exit $? after $((SECONDS - T0)) s" >> gpurun_out/last2_summary.txt
cat gpurun_out/last2_summary.txt gpurun_out/tf32_peak.json
python - <<'PY'
import json
try:
    j = json.loads([l for l in open('gpurun_out/last2_bench.json') if l.startswith('{')][-1])
    print({k: j[k] for k in ('value', 'ms_per_step', 'clocks')})
    print(json.dumps(j['cpu_baseline'].get('train')))
    print(json.dumps(j['cpu_baseline'].get('torch_eager_b200')))
    print(json.dumps(j['train']))
except Exception as e:
    print('no bench line:', e)
PY
tail -3 gpurun_out/last2_bench.err
