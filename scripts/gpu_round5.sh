#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/bench_gemm.py --batch 32 2>&1 | tee gpurun_out/gemm_tc_fp32.txt
timeout 300 python scripts/bench_gemm.py --batch 32 --dtype bf16 2>&1 | tail -22 | tee gpurun_out/gemm_tc_bf16.txt
timeout 300 python scripts/bench_gemm.py --batch 32 --train 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -s 2 -c 1 -o gpurun_out/prof_pw_tc_l1 python scripts/bench_gemm.py --batch 32 --only 1 --iters 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -s 2 -c 1 -o gpurun_out/prof_pw_tc_l19 python scripts/bench_gemm.py --batch 32 --only 19 --iters 1 > gpurun_out/ncu2.log 2>&1
timeout 300 python oracle/bench_gpu_port.py --batch 128 --steps 10 2>&1 | tail -1
timeout 300 python oracle/bench_gpu_port.py --batch 128 --steps 10 --bf16 2>&1 | tail -1
timeout 300 python oracle/bench_gpu_port.py --batch 256 --steps 10 --mode eval 2>&1 | tail -1
ls -la gpurun_out | tail -5
