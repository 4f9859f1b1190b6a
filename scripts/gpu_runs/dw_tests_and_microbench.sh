#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dw.py tests/test_gpu_mn_train.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dw_slide5_b256.txt | tail -14
