#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_mn_train.py -x -q -k graph 2>&1 | grep -E "^E|passed|failed" | cut -c1-400 | head -20
