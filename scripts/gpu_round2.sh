#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt | tail -40
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --batch 64 2>&1 | tail -3 | tee gpurun_out/bench_b64.json
python bench.py --steps 5 --warmup 3 --batch 64 --precision bf16 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_b64_bf16.json
