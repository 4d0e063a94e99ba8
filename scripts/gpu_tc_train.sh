#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_flow.py -x -q -m gpu -p no:cacheprovider -k "mixed_precision or training_step_graph or golden_gradients" 2>&1 | tail -15
