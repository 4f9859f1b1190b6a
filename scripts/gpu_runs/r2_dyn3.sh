#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dymn.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head
python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline > gpurun_out/dyn3_bench_dymn20_b128.json 2>> gpurun_out/dyn3.err
python bench.py --steps 5 --warmup 3 --model dymn10 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/dyn3_bench_dymn10_b128.json 2>> gpurun_out/dyn3.err
tail -3 gpurun_out/dyn3.err
for f in _dymn20_b128 _dymn10_b128; do python -c "
import json
d=json.load(open('gpurun_out/dyn3_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], d['kernel_time_shares'], {k:(round(v['value']) if 'value' in v else v) for k,v in d.get('gpu_baseline',{}).items()})
"; done
