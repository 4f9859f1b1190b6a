#!/bin/bash
# round-1 final data collection (1 GPU): tests, headline bench (+cpu baseline), reference arm, variants, ncu evidence
cd "$(dirname "$0")/../.."
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>$O/bench_default.err | tail -1 > $O/bench_fp32_b256.json; cut -c1-400 $O/bench_fp32_b256.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode eval 2>/dev/null | tail -1 > $O/bench_eval_b256.json; cut -c1-200 $O/bench_eval_b256.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 128 2>/dev/null | tail -1 > $O/bench_fp32_b128.json; cut -c1-200 $O/bench_fp32_b128.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16 2>/dev/null | tail -1 > $O/bench_bf16_b256.json; cut -c1-200 $O/bench_bf16_b256.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model dymn10 --batch 64 2>/dev/null | tail -1 > $O/bench_dymn10_b64.json; cut -c1-200 $O/bench_dymn10_b64.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model dymn10 --batch 128 --mode eval 2>/dev/null | tail -1 > $O/bench_dymn10_eval_b128.json; cut -c1-200 $O/bench_dymn10_eval_b128.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model mn40 --batch 64 2>/dev/null | tail -1 > $O/bench_mn40_b64.json; cut -c1-200 $O/bench_mn40_b64.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model mn04 --batch 256 2>/dev/null | tail -1 > $O/bench_mn04_b256.json; cut -c1-200 $O/bench_mn04_b256.json
timeout 300 python oracle/bench_gpu_port.py 2>/dev/null | tail -4 > $O/stock_torch_gpu.txt; cat $O/stock_torch_gpu.txt | cut -c1-200
timeout 300 python scripts/bench_gemm.py --batch 256 --train > $O/gemm_train_b256.txt 2>&1; tail -1 $O/gemm_train_b256.txt
timeout 300 python scripts/bench_gemm.py --batch 256 > $O/gemm_eval_b256.txt 2>&1; tail -1 $O/gemm_eval_b256.txt
timeout 300 python scripts/bench_dw.py --batch 256 > $O/dw_b256.txt 2>&1; tail -1 $O/dw_b256.txt
# ncu: launch list of one eager training step at B=32 (cold cache, serialised) and full captures of the hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ncu_launches_fp32_b32.csv python bench.py --batch 32 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $O/ncu_launch_run.log 2>&1; tail -1 $O/ncu_launch_run.log | cut -c1-120
PB=64 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"pw_tc|wgrad_tc|dw_slide|dw_wgrad_slide|dw_dgrad2|bn_bwd" -o $O/prof_final -f python scripts/prof_kernels.py > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ls -la $O | head -40
