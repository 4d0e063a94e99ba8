#!/bin/bash
TRAIN_PREC=bf16 PROF_B=4096 timeout 400 python scripts/train_small.py 2>&1 | grep -v "^-\|Warn\|warn" | cut -c1-70,140-240 | tail -28
