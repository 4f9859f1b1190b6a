#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn_train.py tests/test_gpu_mn.py tests/test_gpu_dymn.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -10
timeout 200 python scripts/timing/run_tc_timing.py 2>&1 | tee gpurun_out/tc_timing_v2.txt | grep -E "^M=|producer group 0|epilogue" | cut -c1-200
timeout 200 python scripts/bench_gemm.py --batch 256 --train 2>&1 | tee gpurun_out/gemm_b256_train_v11.txt | tail -1
timeout 200 python scripts/bench_gemm.py --batch 256 2>&1 | tee gpurun_out/gemm_b256_eval_v11.txt | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_v25_fp32_b256.json | cut -c1-200
