#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mel.py tests/test_gpu_refscripts.py -m gpu -q 2>&1 | grep -E "assert|Error|passed|failed" | head
EAT_BENCH_KERNELS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/mel1_bench.json 2> gpurun_out/mel1_bench.err
grep -E "mel|pw_tma" gpurun_out/mel1_bench.err
python bench.py --steps 10 --warmup 3 --mode eval --no-cpu-baseline --no-gpu-baseline > gpurun_out/mel1_bench_eval.json 2>> gpurun_out/mel1.err
for f in "" _eval; do python -c "
import json
d=json.load(open('gpurun_out/mel1_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'])
"; done
