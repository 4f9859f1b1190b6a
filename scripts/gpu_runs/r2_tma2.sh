#!/bin/bash
# round 2: per-role cycle accounting of pw_tma_kernel + accumulator-count sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
for nacc in 2 4 8; do
  echo "===== EAT_TMA_NACC=$nacc" >> gpurun_out/tma2_timing.log
  EAT_TMA_NACC=$nacc timeout 300 python scripts/timing/run_tma_timing.py >> gpurun_out/tma2_timing.log 2>&1
done
for mode in "--train" "--raw"; do
  echo "== mode=$mode" >> gpurun_out/tma2_bench.log
  timeout 300 python scripts/bench_gemm.py --batch 256 $mode >> gpurun_out/tma2_bench.log 2>&1
done
grep -E "==|total" gpurun_out/tma2_bench.log
