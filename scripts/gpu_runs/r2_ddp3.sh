#!/bin/bash
# round 2, two GPUs: the multi-rank exit path (graphs closed, barrier, _exit) must return promptly
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
date +%s > gpurun_out/ddp3_t0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/ddp3_bench.json 2> gpurun_out/ddp3_bench.err
echo "bench rc=$? after $(( $(date +%s) - $(cat gpurun_out/ddp3_t0) )) s"
grep -c "^{" gpurun_out/ddp3_bench.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > gpurun_out/ddp3_ref.json 2> gpurun_out/ddp3_ref.err
echo "reference arm rc=$? after $(( $(date +%s) - $(cat gpurun_out/ddp3_t0) )) s"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_ddp_trainer.py > gpurun_out/ddp3_check.json 2> gpurun_out/ddp3_check.err
echo "check rc=$? after $(( $(date +%s) - $(cat gpurun_out/ddp3_t0) )) s"; grep "^{" gpurun_out/ddp3_check.json
