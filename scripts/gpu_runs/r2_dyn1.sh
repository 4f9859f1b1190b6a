#!/bin/bash
# round 2: DynamicConv 1x1 on the TMA kernel (3-D tensor maps, per-sample pre-mixed weights): parity, dymn bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dymn.py tests/test_gpu_mn.py tests/test_gpu_mn_train.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head
python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn1_bench_dymn20_b128.json 2>> gpurun_out/dyn1.err
python bench.py --steps 5 --warmup 3 --model dymn10 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn1_bench_dymn10_b128.json 2>> gpurun_out/dyn1.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn1_bench.json 2>> gpurun_out/dyn1.err
tail -3 gpurun_out/dyn1.err
for f in _dymn20_b128 _dymn10_b128 ""; do python -c "
import json
d=json.load(open('gpurun_out/dyn1_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], d['kernel_time_shares'])
"; done
