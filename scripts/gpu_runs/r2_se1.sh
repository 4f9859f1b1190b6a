#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mn_train.py tests/test_gpu_train_step.py tests/test_gpu_refscripts.py tests/test_gpu_dymn.py -m gpu -q 2>&1 | grep -E "assert|Error|passed|failed" | head
python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline --no-gpu-baseline > gpurun_out/se1_bench_mn40_b64.json 2>> gpurun_out/se1.err
EAT_BENCH_KERNELS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/se1_bench.json 2> gpurun_out/se1_bench.err
grep -E "se_fc_bwd|simt" gpurun_out/se1_bench.err
for f in _mn40_b64 ""; do python -c "
import json
d=json.load(open('gpurun_out/se1_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['kernel_time_shares'])
"; done
