#!/bin/bash
# round-1 final numbers after the last kernel changes (1 GPU)
cd "$(dirname "$0")/../.."
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2 | tee $O/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>$O/bench_default.err | tail -1 > $O/bench_fp32_b256.json; cut -c1-300 $O/bench_fp32_b256.json
EAT_BENCH_KERNELS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$O/bench_kernels.txt | tail -1 | cut -c1-120
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode eval 2>/dev/null | tail -1 > $O/bench_eval_b256.json; cut -c1-200 $O/bench_eval_b256.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 128 2>/dev/null | tail -1 > $O/bench_fp32_b128.json; cut -c1-200 $O/bench_fp32_b128.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model dymn10 --batch 64 2>/dev/null | tail -1 > $O/bench_dymn10_b64.json; cut -c1-200 $O/bench_dymn10_b64.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --model mn40 --batch 64 2>/dev/null | tail -1 > $O/bench_mn40_b64.json; cut -c1-200 $O/bench_mn40_b64.json
timeout 300 python scripts/bench_dw.py --batch 256 > $O/dw_b256.txt 2>&1; tail -1 $O/dw_b256.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ncu_launches_fp32_b32.csv python bench.py --batch 32 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $O/ncu_launch_run.log 2>&1; tail -1 $O/ncu_launch_run.log | cut -c1-120
