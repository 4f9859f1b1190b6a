#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
timeout 300 python scripts/bench_gemm.py --batch 256 --train 2>&1 | cut -c1-120 | tee gpurun_out/gemm_b256_train_v4.txt | tail -22
timeout 300 python scripts/bench_gemm.py --batch 256 2>&1 | cut -c1-120 | tee gpurun_out/gemm_b256_eval_v4.txt | tail -22
timeout 300 python scripts/bench_gemm.py --batch 32 2>&1 | tail -1
EAT_BENCH_KERNELS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_v13_kernels.txt | tail -1 | tee gpurun_out/bench_v13_fp32_b256.json | cut -c1-400
cat gpurun_out/bench_v13_kernels.txt | grep "ms/step" | head -50
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode eval 2>&1 | tail -1 | tee gpurun_out/bench_v13_eval_b256.json | cut -c1-300
