#!/bin/bash
# Round-end style run: GPU tests, smoke, bench (ours + reference arm), ncu launch list + full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
if grep -q "bench exit 0" gpurun_out/summary.txt; then
  PB="python bench.py --steps 2 --warmup 3 --prewarm-s 0 --no-cpu-baseline --train-steps 0"
  $PB > gpurun_out/plain_prof.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"usf_tc|usf_convert_rows" -s 36 -c 36 --csv --log-file gpurun_out/launches_steady.csv $PB > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" >> gpurun_out/summary.txt
  $PB > gpurun_out/plain_prof2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:usf_tc -s 51 -c 3 -o gpurun_out/prof_tc $PB > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" >> gpurun_out/summary.txt
fi
timeout 300 python scripts/host_narrow.py > gpurun_out/host_narrow.txt 2>&1; echo "host_narrow exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/pageable_e2e.py > gpurun_out/pageable_e2e.txt 2>&1; echo "pageable_e2e exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 2 gpurun_out/gpu_tests.log
cat gpurun_out/smoke.log | grep smoke
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','clocks','e2e','train','cpu_baseline','gpu_launches')})
print(j['roofline'])
PY
