#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/final; mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_fp32_b256_1gpu_samebox.json; cut -c1-200 $O/bench_fp32_b256_1gpu_samebox.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>$O/bench_2gpu.err | tail -1 > $O/bench_fp32_b256_2gpu.json; cut -c1-300 $O/bench_fp32_b256_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 -m pytest tests/test_parallel_gloo.py -q 2>&1 | tail -2
