#!/bin/bash
# round 2: BatchNorm-backward reduce v2 (compile-time activation / gradient composition, 8 loads in flight) vs v1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mn_train.py tests/test_gpu_train_step.py tests/test_gpu_dymn.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head
for v in v2 v1; do
  echo "== EAT_BN_REDUCE=$v" >> gpurun_out/bn1_bench.log
  EAT_BN_REDUCE=$v timeout 300 python scripts/bench_bn.py --batch 256 >> gpurun_out/bn1_bench.log 2>&1
done
cat gpurun_out/bn1_bench.log
EAT_BENCH_KERNELS=1 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>&1 >/dev/null | head -8
