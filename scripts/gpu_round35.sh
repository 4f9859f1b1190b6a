#!/bin/bash
cd "$(dirname "$0")/.."
for d in 5 6; do
  echo "== DBG $d"; EAT_TC_DBG=$d timeout 300 python scripts/bench_gemm.py --batch 256 --raw 2>&1 | cut -c1-100 | awk 'NR<=10 || /total/'
done
