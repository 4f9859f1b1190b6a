#!/bin/bash
# round 2: first run of the TMA-fed weight-gradient kernel (wgrad_tma.cu): parity, microbench vs wgrad_tcgen05.cu, dymn
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn_train.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head -12
for impl in tma tc; do
  echo "== EAT_WG_IMPL=$impl" >> gpurun_out/wg1_bench.log
  EAT_WG_IMPL=$impl timeout 300 python scripts/bench_wgrad.py --batch 256 >> gpurun_out/wg1_bench.log 2>&1
done
grep -E "==|total" gpurun_out/wg1_bench.log
timeout 600 python -m pytest tests/test_gpu_dymn.py -m gpu -x -q -k train 2>&1 | grep -E "assert|Error|passed|failed" | head
