#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
EAT_BENCH_KERNELS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_v22_kernels.txt | tail -1 | tee gpurun_out/bench_v22_fp32_b256.json | cut -c1-300
grep "ms/step" gpurun_out/bench_v22_kernels.txt | head -8
