#!/bin/bash
# round 2, last pass: full GPU test-suite, smoke(), the bench configurations quoted in DESIGN.md / profiles/README.md,
# and the ncu launch list of a short bench run
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
EAT_TEST_REPORT=gpurun_out/parity_report.txt timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | grep -v "^\s*$" | tail -8 > gpurun_out/final2_pytest.log
tail -3 gpurun_out/final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== mn10 EAT_TMA_BNMAX=256"
EAT_TMA_BNMAX=256 timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline 2>/dev/null | cut -c1-200
EAT_BENCH_KERNELS=2 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode eval --no-cpu-baseline > gpurun_out/final2_bench_eval.json 2>> gpurun_out/final2_misc.err
timeout 600 python bench.py --steps 5 --warmup 3 --model mn40 --batch 64 --no-cpu-baseline > gpurun_out/final2_bench_mn40_b64.json 2>> gpurun_out/final2_misc.err
timeout 600 python bench.py --steps 5 --warmup 3 --model dymn10 --batch 128 --no-cpu-baseline --no-gpu-baseline > gpurun_out/final2_bench_dymn10_b128.json 2>> gpurun_out/final2_misc.err
timeout 600 python bench.py --steps 10 --warmup 3 --batch 120 --no-cpu-baseline --no-gpu-baseline > gpurun_out/final2_bench_b120.json 2>> gpurun_out/final2_misc.err
tail -3 gpurun_out/final2_misc.err
for f in "" _eval _mn40_b64 _dymn10_b128 _b120; do python -c "
import json
d=json.load(open('gpurun_out/final2_bench$f.json'))
print('$f', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else v) for k,v in d.get('gpu_baseline',{}).items()}, d.get('cpu_baseline',{}).get('value'), d.get('gpu_launches'))
"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final2_ncu_launches_fp32_b64.csv python bench.py --steps 2 --warmup 1 --batch 64 --no-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/final2_ncu_bench.log 2>&1
wc -l gpurun_out/final2_ncu_launches_fp32_b64.csv
