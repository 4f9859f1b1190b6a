#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
timeout 300 python scripts/bench_gemm.py --batch 32 2>&1 | tail -21 | tee gpurun_out/gemm_tc_fp32_v3.txt
timeout 300 python scripts/bench_gemm.py --batch 32 --train 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v11_fp32_b256.json | cut -c1-1300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode eval 2>&1 | tail -1 | cut -c1-300
