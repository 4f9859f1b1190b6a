#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_fp32_v3.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_kernel -s 30 -c 2 -o gpurun_out/prof_pw_tc_v2 python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 4 -c 2 -o gpurun_out/prof_dw_fwd python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 2 -o gpurun_out/prof_wgrad_tc python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -8
