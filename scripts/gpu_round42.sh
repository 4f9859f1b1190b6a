#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
timeout 300 python scripts/bench_wgrad.py --batch 256 2>&1 | cut -c1-100 | tee gpurun_out/wgrad_b256_v2.txt | tail -22
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_v23_fp32_b256.json | cut -c1-300
