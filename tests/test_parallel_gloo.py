"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding, rank-0 parameter broadcast, and
the property the trainer relies on -- the mean of per-shard gradients of a mean loss equals the gradient of
the global-batch mean loss when shards are equal-sized."""
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
    parallel.allreduce_mean_(flat_g)
    t = parallel.max_over_ranks(float(rank + 1))
    if rank == 0:
        torch.save({"arena": arena, "grad": flat_g, "tmax": t}, out)
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
    assert res["tmax"] == 2.0
