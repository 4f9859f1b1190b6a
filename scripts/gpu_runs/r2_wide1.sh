#!/bin/bash
# pw_tma N tiles wider than 128 columns: parity, then per-layer tables with the wide tiles on (default) and off
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn.py tests/test_gpu_mn_train.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
for bn in 208 128; do
  export EAT_TMA_BNMAX=$bn
  echo "== EAT_TMA_BNMAX=$bn"
  EAT_BENCH_KERNELS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/wide1_bench_$bn.json 2> gpurun_out/wide1_bench_$bn.err
  cut -c1-250 gpurun_out/wide1_bench_$bn.json
  grep -E "eat_pw_tma_fwd |eat_pw_tc_fwd |eat_pw_tc_wgrad" gpurun_out/wide1_bench_$bn.err | head -4
  timeout 600 python bench.py --model mn40 --batch 64 --no-cpu-baseline --steps 10 --warmup 3 --no-gpu-baseline > gpurun_out/wide1_bench_mn40_$bn.json 2> /dev/null
  cut -c1-250 gpurun_out/wide1_bench_mn40_$bn.json
done
EAT_TMA_BNMAX=208 EAT_TMA_WIDE_MINKB=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>/dev/null | cut -c1-250
