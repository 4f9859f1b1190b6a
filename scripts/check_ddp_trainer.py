"""Two-GPU check of AudioSetTrainer's data-parallel path on the CUDA side (NCCL), the part tests/test_parallel_gloo.py
cannot reach on CPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_ddp_trainer.py

 1. replicas that start from DIFFERENT parameters / BatchNorm buffers adopt rank 0's (constructor broadcast);
 2. the bucketed all-reduce that overlaps backward (side stream; also when captured in the CUDA graph) produces the same
    summed gradient arena as one plain all-reduce after backward, and that sum is the sum of the per-rank gradients;
 3. after optimiser steps the replicas hold bit-identical parameters;
 4. timing of a B=64/GPU mn10 step with and without the overlap.
Prints one JSON line on rank 0."""
import contextlib
import io
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficientat_b200.models.mn.model import get_model  # noqa: E402
from efficientat_b200.models.preprocess import AugmentMelSTFT  # noqa: E402
from efficientat_b200.synth import synth_labels, synth_state_, synth_waveform  # noqa: E402
from efficientat_b200.train import AudioSetTrainer  # noqa: E402


def build(seed, dev, graph, buckets, width=0.4):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = synth_state_(get_model(width_mult=width, verbose=False), seed=seed).to(dev)
        mel = AugmentMelSTFT(freqm=0, timem=0, fmin_aug_range=1, fmax_aug_range=1).to(dev)
    model.classifier[4].p = 0.0
    return model, AudioSetTrainer(model, mel, lr=1e-4, mixup_alpha=0.3, cuda_graph=graph, grad_buckets=buckets)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    def gathered(t):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t.contiguous())
        return buf

    # ---- 1. broadcast: every rank builds a different state (seed = 7 + rank)
    model, tr = build(7 + rank, dev, graph=False, buckets=3)
    ps = gathered(tr.flat_p)
    bufs = gathered(torch.cat([b.flatten().float() for b in model.buffers()]))
    out["params_equal_after_broadcast"] = all(torch.equal(ps[0], p) for p in ps)
    out["buffers_equal_after_broadcast"] = all(torch.equal(bufs[0], b) for b in bufs)

    # ---- 2. reduced gradients: bucketed + overlapped (eager and graph) vs one all-reduce vs sum of local gradients
    B = 4
    wave = synth_waveform(B, 32000, seed=100 + rank).to(dev)
    y = synth_labels(B, 527, seed=200 + rank).to(dev)
    teacher = torch.sigmoid(torch.randn(B, 527, generator=torch.Generator().manual_seed(300 + rank))).to(dev)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    lam = torch.rand(B, generator=torch.Generator().manual_seed(2)) * 0.5 + 0.5
    grads = {}
    for name, graph, buckets in (("plain", False, 0), ("bucketed_eager", False, 3), ("bucketed_graph", True, 3)):
        m, t = build(7, dev, graph, buckets)
        t.model.train(); t.mel.train()
        _, g = t.forward_backward(wave, y, teacher, perm, lam)
        if t.bucketer is None:
            local_g = g.clone()
            dist.all_reduce(g)
            both = gathered(local_g)
            out["plain_equals_sum_of_local"] = bool(torch.allclose(g, sum(both), rtol=1e-5, atol=1e-8))
        grads[name] = g.clone()
    ref = grads["plain"]
    for k in ("bucketed_eager", "bucketed_graph"):
        out[f"{k}_vs_plain_maxrel"] = float((grads[k] - ref).abs().max() / ref.abs().max())
        out[f"{k}_cos"] = float(torch.nn.functional.cosine_similarity(grads[k].double(), ref.double(), dim=0))

    # ---- 3. replicas stay identical through optimiser steps (graph + buckets)
    m, t = build(7 + rank, dev, True, 3)
    for i in range(3):
        t.step(wave, y, teacher)
    ps = gathered(t.flat_p)
    out["params_bit_identical_after_3_steps"] = all(torch.equal(ps[0], p) for p in ps)

    # ---- 4. step time with / without overlap (mn10, 64 ten-second clips per GPU)
    Bt = 64
    wave = synth_waveform(Bt, 320000, seed=100 + rank).to(dev)
    y = synth_labels(Bt, 527, seed=200 + rank).to(dev)
    teacher = torch.sigmoid(torch.randn(Bt, 527, generator=torch.Generator().manual_seed(300 + rank))).to(dev)
    for name, buckets in (("one_allreduce_after_backward", 0), ("bucketed_overlapped", 3)):
        m, t = build(7, dev, True, buckets, width=1.0)
        for _ in range(4):
            t.step(wave, y, teacher)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            t.step(wave, y, teacher)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[f"ms_per_step_{name}"] = float(ms)
    if rank == 0:
        print(json.dumps(out), flush=True)
    t.close()                         # graphs with NCCL kernels must be gone before the communicator is torn down
    del t, m
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)                       # (destroy_process_group() after graph-captured collectives was seen to hang at exit)


if __name__ == "__main__":
    main()
