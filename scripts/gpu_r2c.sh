#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_trace2.py > gpurun_out/tc_trace2.log 2>&1; echo "trace exit $?"
