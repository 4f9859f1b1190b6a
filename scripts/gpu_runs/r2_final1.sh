#!/bin/bash
# round 2: full GPU test-suite + the bench configurations quoted in DESIGN.md / profiles/README.md
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
EAT_TEST_REPORT=gpurun_out/parity_report.txt timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | grep -v "^\s*$" | tail -8 > gpurun_out/final1_pytest.log
tail -3 gpurun_out/final1_pytest.log
EAT_BENCH_KERNELS=2 python bench.py --steps 20 --warmup 5 > gpurun_out/final1_bench.json 2> gpurun_out/final1_bench.err
python bench.py --steps 10 --warmup 3 --mode eval --no-cpu-baseline > gpurun_out/final1_bench_eval.json 2>> gpurun_out/final1_misc.err
python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline > gpurun_out/final1_bench_dymn20_b128.json 2>> gpurun_out/final1_misc.err
python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline > gpurun_out/final1_bench_mn40_b64.json 2>> gpurun_out/final1_misc.err
python bench.py --steps 10 --warmup 3 --batch 120 --no-cpu-baseline --no-gpu-baseline > gpurun_out/final1_bench_b120.json 2>> gpurun_out/final1_misc.err
python bench.py --steps 3 --warmup 3 --impl reference > gpurun_out/final1_bench_impl_reference.json 2>> gpurun_out/final1_misc.err
tail -3 gpurun_out/final1_misc.err
for f in "" _eval _dymn20_b128 _mn40_b64 _b120; do python -c "
import json
d=json.load(open('gpurun_out/final1_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], {k:(round(v['value']) if 'value' in v else v) for k,v in d.get('gpu_baseline',{}).items()}, d.get('cpu_baseline',{}).get('value'))
"; done
