#!/bin/bash
# round 2: pre-split weights for the TMA GEMM (workspace route used by the engine): parity, then mn10 / mn40 bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn.py tests/test_gpu_mn_train.py tests/test_gpu_train_step.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head
EAT_BENCH_KERNELS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/wpre1_bench.json 2> gpurun_out/wpre1_bench.err
head -4 gpurun_out/wpre1_bench.err
python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline --no-gpu-baseline > gpurun_out/wpre1_bench_mn40_b64.json 2>> gpurun_out/wpre1.err
for f in "" _mn40_b64; do python -c "
import json
d=json.load(open('gpurun_out/wpre1_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['frac'], d['kernel_time_shares'])
"; done
