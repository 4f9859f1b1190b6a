#!/bin/bash
# first GPU contact: device check, CPU+GPU tests, verbose on failure
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
