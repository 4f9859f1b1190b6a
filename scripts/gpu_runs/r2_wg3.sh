#!/bin/bash
# wgrad_tma: only the boxes inside N / K are loaded and fixed up; 128-row blocks for the first layers.  Also: the 5x5
# stride-1 depthwise data gradient on the sliding-window kernel (EAT_DW5_DGRAD=slide) against the tile kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mn_train.py tests/test_gpu_dymn.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
for mb in 64 auto; do
  if [ $mb = auto ]; then unset EAT_WG_MB; else export EAT_WG_MB=$mb; fi
  echo "== EAT_WG_MB=$mb"
  timeout 300 python scripts/bench_wgrad.py --batch 256 2>&1 | tee gpurun_out/wg3_micro_$mb.txt | head -12
  tail -1 gpurun_out/wg3_micro_$mb.txt
done
unset EAT_WG_MB
for d in tile slide; do
  echo "== EAT_DW5_DGRAD=$d"
  for l in 4 11; do EAT_DW5_DGRAD=$d timeout 120 python scripts/bench_dw.py --batch 256 --only $l 2>&1 | head -1 | cut -c1-150; done
done
EAT_DW5_DGRAD=slide timeout 300 python -m pytest tests/test_gpu_dw.py -x -q -m gpu 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/wg3_bench.json 2> gpurun_out/wg3_bench.err
cut -c1-250 gpurun_out/wg3_bench.json
