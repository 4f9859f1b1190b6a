#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"dw_|pw_tc|wgrad_tc|bn_bwd" -o gpurun_out/prof_r25 -f python scripts/prof_kernels.py > gpurun_out/ncu25.log 2>&1
tail -3 gpurun_out/ncu25.log
ls -la gpurun_out/prof_r25.ncu-rep
python scripts/bench_dw.py --batch 256 > gpurun_out/dw_b256.txt 2>&1; tail -14 gpurun_out/dw_b256.txt | cut -c1-200
python scripts/bench_gemm.py --batch 256 --train > gpurun_out/gemm_b256_train.txt 2>&1; tail -22 gpurun_out/gemm_b256_train.txt | cut -c1-120
