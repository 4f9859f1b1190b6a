#!/bin/bash
# round 2, two GPUs: trainer DDP check (broadcast, bucketed overlapped all-reduce eager + in-graph, identical replicas), 2-GPU bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_ddp_trainer.py > gpurun_out/ddp2_check.json 2> gpurun_out/ddp2_check.err
tail -3 gpurun_out/ddp2_check.err; cat gpurun_out/ddp2_check.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/ddp2_bench.json 2> gpurun_out/ddp2_bench.err
tail -3 gpurun_out/ddp2_bench.err; python -c "
import json; d=json.load(open('gpurun_out/ddp2_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_baseline'))"
