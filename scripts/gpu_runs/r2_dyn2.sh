#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for thr in 1000 100000; do
python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn2_bench_dymn20_$thr.json 2>> gpurun_out/dyn2.err
EAT_DYN_TMA_MIN_RPS=$thr python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn2_bench_dymn20_$thr.json 2>> gpurun_out/dyn2.err
python -c "
import json
d=json.load(open('gpurun_out/dyn2_bench_dymn20_$thr.json'))
print('$thr', round(d['value']), round(d['ms_per_step'],2), d['kernel_time_shares'])
"; done
EAT_DYN_TMA_MIN_RPS=400 python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn2_bench_dymn20_400.json 2>> gpurun_out/dyn2.err
python -c "
import json
d=json.load(open('gpurun_out/dyn2_bench_dymn20_400.json'))
print('400', round(d['value']), round(d['ms_per_step'],2), d['kernel_time_shares'])
"
