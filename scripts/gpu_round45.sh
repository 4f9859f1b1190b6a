#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_v24_fp32_b256.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['roofline']['frac'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode eval 2>/dev/null | tail -1 | tee gpurun_out/bench_v24_eval_b256.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
