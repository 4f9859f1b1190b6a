#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python scripts/timing/run_tc_timing.py 2>&1 | tee gpurun_out/tc_timing.txt | tail -48
