#!/bin/bash
# dy_act_bwd with shuffle-combined shared atomics; wide-tile knobs on mn40
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dymn.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
EAT_BENCH_KERNELS=1 timeout 600 python bench.py --model dymn20 --batch 128 --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/dyn4_bench_dymn20_b128.json 2> gpurun_out/dyn4_bench_dymn20.err
cut -c1-250 gpurun_out/dyn4_bench_dymn20_b128.json
grep -E "^  eat_" gpurun_out/dyn4_bench_dymn20.err | head -14
for v in "EAT_TMA_BNMAX=256" "EAT_TMA_WIDE_MINKB=3" "EAT_TMA_WIDE_MINKB=8"; do
  echo "== mn40 $v"
  env $v timeout 600 python bench.py --model mn40 --batch 64 --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>/dev/null | cut -c1-200
done
