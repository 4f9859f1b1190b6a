#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for d in 0 1 2 3 4; do
  echo "== DBG $d"; EAT_TC_DBG=$d timeout 300 python scripts/bench_gemm.py --batch 256 --raw 2>&1 | cut -c1-100 | tee gpurun_out/gemm_dbg$d.txt | awk 'NR<=10 || /total/'
done
