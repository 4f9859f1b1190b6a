#!/bin/bash
# round 2, last pass on the final tree (3.9 GPU-minutes left): full GPU test-suite, smoke(), the default bench WITHOUT the
# baseline legs (reference-GPU / CPU arms are unchanged code, measured in r2_final3.sh: 1179 / 956 / 25 clips/s), then eval,
# mn40, dymn20 as far as the budget reaches
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
T0=$SECONDS
EAT_TEST_REPORT=gpurun_out/parity_report.txt timeout 200 python -m pytest tests -m gpu -q -rs 2>&1 | grep -v "^\s*$" | tail -12 > gpurun_out/final4_pytest.log
tail -3 gpurun_out/final4_pytest.log
echo "pytest took $((SECONDS-T0)) s"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
EAT_BENCH_KERNELS=2 timeout 100 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err
timeout 60 python bench.py --steps 10 --warmup 3 --mode eval --no-cpu-baseline --no-gpu-baseline > gpurun_out/final4_bench_eval.json 2>> gpurun_out/final4_misc.err
timeout 60 python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline --no-gpu-baseline > gpurun_out/final4_bench_mn40_b64.json 2>> gpurun_out/final4_misc.err
timeout 60 python bench.py --steps 5 --warmup 3 --model dymn20 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/final4_bench_dymn20_b128.json 2>> gpurun_out/final4_misc.err
for f in "" _eval _mn40_b64 _dymn20_b128; do python -c "
import json
d=json.load(open('gpurun_out/final4_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], d.get('gpu_launches'))
"; done
echo "total $((SECONDS-T0)) s"
