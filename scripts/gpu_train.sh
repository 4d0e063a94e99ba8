#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_flow.py -x -q -m gpu -p no:cacheprovider -k "training_step_graph or adbench" 2>&1 | tail -5
timeout 400 python scripts/train_small.py 2>&1 | grep "B="
