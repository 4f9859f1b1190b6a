#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -40 > gpurun_out/pytest_gemm.txt; cat gpurun_out/pytest_gemm.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
