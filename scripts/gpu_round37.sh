#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn_train.py tests/test_gpu_mn.py tests/test_gpu_dymn.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
timeout 300 python scripts/bench_gemm.py --batch 256 --raw 2>&1 | cut -c1-100 | awk 'NR<=10 || /total/'
timeout 300 python scripts/bench_gemm.py --batch 256 --train 2>&1 | cut -c1-120 | tee gpurun_out/gemm_b256_train_v10.txt | tail -22
timeout 300 python scripts/bench_gemm.py --batch 256 2>&1 | cut -c1-120 | tee gpurun_out/gemm_b256_eval_v10.txt | tail -1
timeout 300 python scripts/bench_gemm.py --batch 32 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_v19_fp32_b256.json | cut -c1-300
