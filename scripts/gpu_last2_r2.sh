#!/bin/bash
# The round's remaining ~3 GPU-minutes: the TF32 tensor peak BASELINE.md section 2 leaves to the builder; the training
# baselines of BASELINE.md section 3 (reference classes, CPU and torch eager on the B200) called directly; then, if the
# time is there, the whole bench once more with them in its cpu_baseline leg.
mkdir -p gpurun_out
T0=$SECONDS
S=gpurun_out/last2_summary.txt
timeout 70 python scripts/tf32_peak.py 2 > gpurun_out/tf32_peak.json 2> gpurun_out/tf32_peak.err
echo "tf32 peak exit $? after $((SECONDS - T0)) s" > $S
timeout 60 python - > gpurun_out/train_baselines.json 2> gpurun_out/train_baselines.err <<'PY'
import json, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
dev = torch.device('cuda', 0)
out = {"torch_eager_b200": {f"B{b}": bench.reference_train_run(dev, b, 10, 3) for b in (64, 4096)},
       "cpu": {"B64": bench.reference_train_run("cpu", 64, 10, 1), "B4096": bench.reference_train_run("cpu", 4096, 3, 1)},
       "cores": torch.get_num_threads()}
print(json.dumps(out))
PY
echo "train baselines exit $? after $((SECONDS - T0)) s" >> $S
LEFT=$((166 - (SECONDS - T0)))
if [ $LEFT -ge 100 ]; then
  timeout $LEFT python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last2_bench.json 2> gpurun_out/last2_bench.err
  echo "bench exit $? after $((SECONDS - T0)) s (limit $LEFT s)" >> $S
else
  echo "bench skipped: $LEFT s left" >> $S
fi
cat $S gpurun_out/tf32_peak.json gpurun_out/train_baselines.json
tail -2 gpurun_out/train_baselines.err gpurun_out/tf32_peak.err
