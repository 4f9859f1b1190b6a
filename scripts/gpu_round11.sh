#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dymn.py -x -q 2>&1 | tail -3
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_2gpu.json | cut -c1-1200
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_1gpu_default.json | cut -c1-2200
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-900
