#!/bin/bash
# First GPU contact: fp32 kernels, then fp32 flow parity, then the tcgen05 GEMM, then everything else.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout 900 python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider"
$T tests/test_gpu_kernels.py -k "not tcgen05" > gpurun_out/t1_kernels_fp32.log 2>&1; echo "t1 exit $?" >> gpurun_out/summary.txt
$T tests/test_gpu_flow.py -k "golden_fp32 or gradients or uniformly or adbench or reference_behaviour" > gpurun_out/t2_flow_fp32.log 2>&1; echo "t2 exit $?" >> gpurun_out/summary.txt
$T tests/test_gpu_kernels.py -k "tcgen05" > gpurun_out/t3_tc.log 2>&1; echo "t3 exit $?" >> gpurun_out/summary.txt
$T tests/test_gpu_flow.py -k "not (golden_fp32 or gradients or uniformly or adbench or reference_behaviour)" > gpurun_out/t4_flow_rest.log 2>&1; echo "t4 exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 --rows 16384 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/t1_kernels_fp32.log gpurun_out/t2_flow_fp32.log gpurun_out/t3_tc.log gpurun_out/t4_flow_rest.log
tail -2 gpurun_out/bench_bf16.log
