#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "--mode eval --batch 256" "--model dymn10 --batch 64" "--model dymn20 --batch 32" "--model mn40 --batch 64" "--model mn10 --batch 256" "--model dymn10 --mode eval --batch 128"; do
  name=$(echo $cfg | tr -d '-' | tr ' ' '_')
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg 2>&1 | tail -1 | tee gpurun_out/bench_$name.json | cut -c1-260
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("   ->", d["config"]["workload"][:60], "| value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],2), "|", {k:round(v,3) for k,v in list(d["kernel_time_shares"].items())[:6]})
except Exception as e: print("   -> parse failed", e)
PY
done
timeout 300 python oracle/bench_gpu_port.py --batch 256 --steps 10 --mode eval --bf16 2>&1 | tail -1
