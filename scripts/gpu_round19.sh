#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_mn_train.py -x -q 2>&1 | grep -E "^E  .*(norm|tensors|assert)|passed|failed" | cut -c1-330 | head -12; done
