#!/bin/bash
# depthwise kernels: ring policy (depth 2, off where it costs a CTA) + FFMA2 tap loops, A/B against a scalar-fmaf build of
# dw_slide.cu (efficientat_b200/libeat_b200_scalarfma.so, built with -DEAT_DW_SCALAR_FMA)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dw.py tests/test_gpu_dymn.py -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -20
echo "== FFMA2"
timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dwring2_micro_ffma2.txt | tail -14
if [ -f efficientat_b200/libeat_b200_scalarfma.so ]; then
  cp efficientat_b200/libeat_b200.so /tmp/libeat_keep.so
  cp efficientat_b200/libeat_b200_scalarfma.so efficientat_b200/libeat_b200.so
  echo "== scalar fmaf"
  timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dwring2_micro_scalar.txt | tail -14
  cp /tmp/libeat_keep.so efficientat_b200/libeat_b200.so
fi
echo "== FFMA2 again (run-to-run noise)"
timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | cut -c1-150 | tee gpurun_out/dwring2_micro_ffma2_b.txt | tail -14
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-baseline > gpurun_out/dwring2_bench.json 2> gpurun_out/dwring2_bench.err
cut -c1-400 gpurun_out/dwring2_bench.json
