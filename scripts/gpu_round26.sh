#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dw.py -x -q 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
for v in 0 1; do
  echo "== slide variant $v train"; EAT_DW_VARIANT=$v timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dw_slide_v${v}_b256.txt | tail -14
  echo "== slide variant $v eval"; EAT_DW_VARIANT=$v timeout 300 python scripts/bench_dw.py --batch 256 --eval 2>&1 | cut -c1-62 | tail -14
done
echo "== old eval"; EAT_DW_IMPL=old timeout 300 python scripts/bench_dw.py --batch 256 --eval 2>&1 | cut -c1-62 | tail -14
