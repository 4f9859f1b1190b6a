#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -15
timeout 300 python scripts/bench_gemm.py --batch 32 2>&1 | tee gpurun_out/gemm_tc_fp32_v2.txt
timeout 300 python scripts/bench_gemm.py --batch 32 --dtype bf16 2>&1 | tail -22 | tee gpurun_out/gemm_tc_bf16_v2.txt
timeout 300 python scripts/bench_gemm.py --batch 32 --train 2>&1 | tail -22
timeout 300 python scripts/bench_gemm.py --batch 128 --dtype bf16 2>&1 | tail -1
timeout 300 python scripts/bench_gemm.py --batch 128 2>&1 | tail -1
