#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dymn.py -x -q -s 2>&1 | tail -40
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
