#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_v6_fp32_b256.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 1 -c 1 -o gpurun_out/prof_dw_l0 python scripts/bench_dw.py --only 0 --batch 32 > gpurun_out/ncu6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 1 -c 1 -o gpurun_out/prof_dw_l2 python scripts/bench_dw.py --only 2 --batch 32 > gpurun_out/ncu7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_kernel -s 2 -c 1 -o gpurun_out/prof_pw_tc_v2_l1 python scripts/bench_gemm.py --only 1 --batch 32 --iters 1 > gpurun_out/ncu8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc_kernel -s 2 -c 1 -o gpurun_out/prof_pw_tc_v2_l1_train python scripts/bench_gemm.py --only 1 --batch 32 --iters 1 --train > gpurun_out/ncu9.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
