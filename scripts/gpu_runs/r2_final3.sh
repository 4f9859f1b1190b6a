#!/bin/bash
# round 2, re-entry pass at HEAD: full GPU test-suite, smoke(), the default bench with the kernel table (13.9 GPU-minutes left:
# everything else of r2_final2.sh is dropped)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
T0=$SECONDS
EAT_TEST_REPORT=gpurun_out/parity_report.txt timeout 400 python -m pytest tests -m gpu -q -rs --durations=8 2>&1 | grep -v "^\s*$" | tail -24 > gpurun_out/final3_pytest.log
tail -4 gpurun_out/final3_pytest.log
echo "pytest took $((SECONDS-T0)) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
T0=$SECONDS
EAT_BENCH_KERNELS=2 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
echo "bench took $((SECONDS-T0)) s"
python -c "
import json
d=json.load(open('gpurun_out/final3_bench.json'))
print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['kernel'], d['roofline']['frac'], {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else v) for k,v in d.get('gpu_baseline',{}).items()}, d.get('cpu_baseline',{}).get('value'), d.get('gpu_launches'), d.get('kernel_time_shares'))
"
