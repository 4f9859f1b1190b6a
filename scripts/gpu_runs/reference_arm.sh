#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -c "import bench, os; print('usable', bench.usable_cores(), 'cpu_count', os.cpu_count())"
( time timeout 230 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference_v2.json | cut -c1-700 ) 2>&1 | tail -6
