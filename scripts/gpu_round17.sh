#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1 2 3; do echo "== variant $v"; EAT_DW_VARIANT=$v timeout 300 python scripts/bench_dw.py --batch 32 2>&1 | grep -E "k=3|fwd_ms" | cut -c1-140; done
timeout 600 python -m pytest tests/test_gpu_dymn.py -x -q -k train 2>&1 | grep -E "norm|passed|failed|tensors" | cut -c1-300 | head -20
