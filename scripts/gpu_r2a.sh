#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_ablate.py > gpurun_out/tc_ablate.log 2>&1; echo "ablate exit $?"
cat gpurun_out/tc_ablate.log | tail -5
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs exit $?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/configs.json')):
    print(r['config'], {p:(round(r[p]['samples_per_s']/1e6,2), r[p]['launches'], '%.1e'%r[p]['max_rel_err_vs_fp64_oracle']) for p in ('bf16','fp32')})
PY
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
