#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k wgrad 2>&1 | tail -15
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v2_fp32_b128.json | cut -c1-1800
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --precision bf16 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v2_bf16_b128.json | cut -c1-1800
