#!/bin/bash
# round 2: pw_tma v3 (bf16x3 in place): parity, role timing with 1 and 2 CTAs/SM, microbench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
for c in 1 2; do
  echo "===== EAT_TMA_CTAS=$c" >> gpurun_out/tma4_timing.log
  EAT_TMA_CTAS=$c timeout 300 python scripts/timing/run_tma_timing.py >> gpurun_out/tma4_timing.log 2>&1
done
for c in 1 2; do
for mode in "--train" "--raw"; do
  echo "== ctas=$c mode=$mode" >> gpurun_out/tma4_bench.log
  EAT_TMA_CTAS=$c timeout 300 python scripts/bench_gemm.py --batch 256 $mode >> gpurun_out/tma4_bench.log 2>&1
done
done
grep -E "==|total" gpurun_out/tma4_bench.log
