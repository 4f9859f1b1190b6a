#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; tail -15 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tc_fp32_b128.json | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --precision bf16 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tc_bf16_b128.json | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
