"""world_size-2 gloo tests (CPU) of the data-parallel host logic that efficientat_b200.train.AudioSetTrainer calls
(efficientat_b200/parallel.py): sharding, the rank-0 broadcast of parameters and buffers, the bucketed gradient
all-reduce driven by backward's "parameters >= i are final" notifications (GradBucketer), and the property the trainer
relies on -- the mean of per-shard gradients of a mean loss equals the gradient of the global-batch mean loss when
shards are equal-sized.  The CUDA path of the same calls (NCCL, side stream, inside a CUDA graph) is
exercised on two GPUs by scripts/check_ddp_trainer.py (result under profiles/)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from efficientat_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = parallel.init_from_env(backend="gloo")
    torch.manual_seed(1234 + rank)                       # ranks start from DIFFERENT parameters ...
    params = [torch.randn(5, 3), torch.randn(7)]
    arena, views = parallel.flatten_like_arena(params)
    parallel.broadcast_from_rank0_(arena)                # ... and adopt rank 0's
    g = torch.Generator().manual_seed(7)                 # identical global batch on every rank
    X, Y = torch.randn(8, 3, generator=g), torch.randn(8, 5, generator=g)
    lo, hi = parallel.shard_range(8, r, w)
    Wm, b = views[0].clone().requires_grad_(True), views[1][:5].clone().requires_grad_(True)
    loss = ((X[lo:hi] @ Wm.t() + b - Y[lo:hi]) ** 2).mean()
    loss.backward()
    flat_g, _ = parallel.flatten_like_arena([Wm.grad, b.grad])
    # the trainer's path: buckets reduced as backward reports them (last parameter first), 1/world applied afterwards
    flat_b = flat_g.clone()
    bk = parallel.GradBucketer([Wm.grad.numel(), b.grad.numel()], n_buckets=4)      # target 5 elements: the bias is its own bucket
    bk.ready(flat_b, 1)                                   # "parameter 1 (the bias) is final": its bucket goes out
    part = flat_b.clone()
    bk.finish(flat_b)                                     # the rest + join
    flat_b.div_(w)
    parallel.allreduce_mean_(flat_g)
    t = parallel.max_over_ranks(float(rank + 1))
    if rank == 0:
        torch.save({"arena": arena, "grad": flat_g, "grad_bucketed": flat_b, "part": part, "tmax": t}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    for n in (1, 7, 8, 120, 256):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gradient_mean_matches_global_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    torch.manual_seed(1234)
    params = [torch.randn(5, 3), torch.randn(7)]
    assert torch.equal(res["arena"], torch.cat([p.reshape(-1) for p in params]))     # rank 0's parameters won
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(8, 3, generator=g), torch.randn(8, 5, generator=g)
    Wm, b = params[0].clone().requires_grad_(True), params[1][:5].clone().requires_grad_(True)
    ((X @ Wm.t() + b - Y) ** 2).mean().backward()
    want = torch.cat([Wm.grad.reshape(-1), b.grad.reshape(-1)])
    assert torch.allclose(res["grad"], want, atol=1e-6)
    assert torch.allclose(res["grad_bucketed"], want, atol=1e-6)
    # after the first notification only the bias bucket had been reduced (sum over 2 ranks = 2 x mean)
    assert torch.allclose(res["part"][15:], 2 * want[15:], atol=1e-6) and not torch.allclose(res["part"][:15], 2 * want[:15], atol=1e-6)
    assert res["tmax"] == 2.0


def test_bucket_bounds_partition_the_parameters():
    sizes = [3, 10, 10, 500, 20, 700, 1000]
    for nb in (1, 2, 3, 4, 9):
        bk = parallel.GradBucketer(sizes, nb)
        assert bk.bounds[0] == 0 and bk.bounds[-1] == len(sizes) and bk.bounds == sorted(set(bk.bounds))
        assert len(bk.bounds) - 1 <= max(1, nb)
    bk = parallel.GradBucketer(sizes, 3)
    calls = []
    bk._reduce = lambda flat, lo, hi: calls.append((lo, hi))
    flat = torch.zeros(sum(sizes))
    for i in (6, 5, 4, 2, 0):                              # backward reports from the last parameter to the first
        bk.ready(flat, i)
    bk.finish(flat)
    assert [c for c in calls] == sorted(calls, reverse=True) and calls[0][1] == len(sizes) and calls[-1][0] == 0
    assert all(calls[i][0] == calls[i + 1][1] for i in range(len(calls) - 1))
