#!/bin/bash
# round 2: weight gradients forked onto a side stream (graph branches): tests + bench with the fork on / off
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mn_train.py tests/test_gpu_train_step.py tests/test_gpu_dymn.py tests/test_gpu_refscripts.py -m gpu -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head
for f in 1 0; do
  EAT_FORK_WGRAD=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/fork1_bench_$f.json 2> gpurun_out/fork1_bench_$f.err
  python -c "
import json; d=json.load(open('gpurun_out/fork1_bench_$f.json')); print('fork=$f', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
