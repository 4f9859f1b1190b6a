#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tma, no narrow" > gpurun_out/wg2_bench.log
timeout 300 python scripts/bench_wgrad.py --batch 256 >> gpurun_out/wg2_bench.log 2>&1
head -3 gpurun_out/wg2_bench.log; tail -1 gpurun_out/wg2_bench.log
bash scripts/gpu_runs/r2_full1.sh
