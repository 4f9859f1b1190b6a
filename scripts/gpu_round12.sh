#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
timeout 300 python scripts/bench_dw.py --batch 32 2>&1 | tee gpurun_out/dw_fp32_tiled.txt
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v4_fp32_b128.json | cut -c1-1800
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
