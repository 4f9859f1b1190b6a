#!/bin/bash
# depthwise kernels with the per-thread cp.async prefetch ring: parity tests, then the microbench with the ring off / on
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dw.py tests/test_gpu_dymn.py tests/test_gpu_mn_train.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
for d in 0 2 3 auto; do
  if [ $d = auto ]; then unset EAT_DW_RING; else export EAT_DW_RING=$d; fi
  echo "== EAT_DW_RING=$d"
  timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dwring1_micro_$d.txt | tail -14
done
unset EAT_DW_RING
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-baseline > gpurun_out/dwring1_bench.json 2> gpurun_out/dwring1_bench.err
cut -c1-400 gpurun_out/dwring1_bench.json
EAT_DW_RING=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-baseline > gpurun_out/dwring1_bench_off.json 2> gpurun_out/dwring1_bench_off.err
cut -c1-200 gpurun_out/dwring1_bench_off.json
