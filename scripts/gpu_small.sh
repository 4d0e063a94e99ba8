#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/small_batch.py 2>&1 | grep "rows="
