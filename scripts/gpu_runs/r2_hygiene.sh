#!/bin/bash
# round 2: sanitizer, ncu captures (launch list + --set full of the GEMM / mel kernels), extra bench configurations
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
# 1. compute-sanitizer over the GEMM parity tests (TMA kernels + the register-staged bf16 kernel)
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/hyg_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/hyg_sanitizer_$tool.log
done
# 2. launch list of two eager bench steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/hyg_launches.csv \
  python bench.py --steps 1 --warmup 3 --batch 64 --no-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/hyg_ncu_bench.log 2>&1
wc -l gpurun_out/hyg_launches.csv
# 3. full captures: one launch each of the pointwise GEMM (train variant), the weight gradient, the mel kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pw_tma_kernel|wgrad_tma_kernel|mel_kernel" -c 6 \
  -o gpurun_out/hyg_full python bench.py --steps 1 --warmup 3 --batch 64 --no-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/hyg_ncu_full.log 2>&1
ls -la gpurun_out/hyg_full.ncu-rep
# 4. other configurations of the bench
python bench.py --steps 10 --warmup 3 --mode eval --no-cpu-baseline > gpurun_out/hyg_bench_eval.json 2>> gpurun_out/hyg_bench.err
python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline --no-gpu-baseline > gpurun_out/hyg_bench_bf16.json 2>> gpurun_out/hyg_bench.err
python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline > gpurun_out/hyg_bench_dymn20_b128.json 2>> gpurun_out/hyg_bench.err
python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline > gpurun_out/hyg_bench_mn40_b64.json 2>> gpurun_out/hyg_bench.err
python bench.py --steps 10 --warmup 3 --batch 120 --no-cpu-baseline > gpurun_out/hyg_bench_b120.json 2>> gpurun_out/hyg_bench.err
tail -5 gpurun_out/hyg_bench.err
for f in eval bf16 dymn20_b128 mn40_b64 b120; do python -c "
import json,sys
d=json.load(open('gpurun_out/hyg_bench_$f.json'))
print('$f', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], {k:(round(v['value']) if 'value' in v else v) for k,v in d.get('gpu_baseline',{}).items()})
"; done
