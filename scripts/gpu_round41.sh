#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/bench_wgrad.py --batch 256 2>&1 | cut -c1-100 | tee gpurun_out/wgrad_b256.txt | tail -22
