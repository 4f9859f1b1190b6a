#!/bin/bash
# round 2: first run of the TMA-fed TF32x3 pointwise GEMM (pw_tma.cu): parity, then microbench vs the register-staged kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tma1_pytest.log
tail -4 gpurun_out/tma1_pytest.log
for mode in "--train" "--raw" ""; do
  for impl in tma tc; do
    echo "== impl=$impl mode=$mode" >> gpurun_out/tma1_bench.log
    EAT_PW_IMPL=$impl timeout 300 python scripts/bench_gemm.py --batch 256 $mode >> gpurun_out/tma1_bench.log 2>&1
  done
done
grep -E "==|total" gpurun_out/tma1_bench.log
