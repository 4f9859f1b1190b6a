#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dymn.py -x -q 2>&1 | tail -70 | cut -c1-260
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
