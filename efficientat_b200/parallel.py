"""Data parallelism over clip batches: one process per GPU, replicas of the model, per-replica BatchNorm,
one gradient all-reduce per step (the semantics the reference gets from Lightning DDP,
ex_pl_audioset.py:287-293; `ex_audioset.py` itself is single-device).

The helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU test-suite): they only move flat tensors.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None, device=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, local_rank, world_size); a no-op single-process group when WORLD_SIZE is 1 or unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of a global batch: the first n % world ranks get one extra clip."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def flatten_like_arena(tensors):
    """Copy a list of tensors into one contiguous fp32 arena; returns (arena, views) with views aliasing it."""
    n = sum(t.numel() for t in tensors)
    arena = torch.empty(n, dtype=torch.float32, device=tensors[0].device if tensors else "cpu")
    views, off = [], 0
    for t in tensors:
        k = t.numel()
        arena[off:off + k].copy_(t.reshape(-1))
        views.append(arena[off:off + k].view(t.shape))
        off += k
    return arena, views


def allreduce_mean_(flat, group=None):
    """In-place gradient averaging of a flat arena: ONE sum all-reduce, then 1/world.  (The CUDA trainer folds
    the 1/world into its Adam kernel instead and calls dist.all_reduce directly.)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


class GradBucketer:
    """Overlap of the gradient all-reduce with the backward pass (what DistributedDataParallel's buckets do for the
    reference, ex_pl_audioset.py:287-293): the flat gradient arena (parameter order) is cut into `n_buckets` contiguous
    ranges of roughly equal size; backward produces gradients from the LAST parameter to the first and reports
    `ready(first_param_index)` after each stage, and every bucket whose parameters are all complete is sum-all-reduced
    at once -- on a side CUDA stream for device tensors, so that it runs under the rest of backward (inside a CUDA graph
    the fork / join becomes graph edges); synchronously for CPU tensors (gloo, test-suite).  `finish()` joins."""

    def __init__(self, sizes, n_buckets=3, group=None):
        self.group = group
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + int(n))
        total = self.offsets[-1]
        # bucket boundaries in PARAMETER indices, cut from the end (the classifier's gradients come first)
        cuts, target, acc = [len(sizes)], total / max(1, n_buckets), 0
        for i in range(len(sizes) - 1, 0, -1):
            acc += int(sizes[i])
            if acc >= target and len(cuts) < n_buckets:
                cuts.append(i)
                acc = 0
        cuts.append(0)
        self.bounds = sorted(set(cuts))                   # [0, ..., n_params]
        self._side = None
        self.reset()

    def reset(self):
        self._next = len(self.bounds) - 1                  # index of the upper bound of the next bucket to send

    def _reduce(self, flat, lo, hi):
        view = flat[self.offsets[lo]:self.offsets[hi]]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if flat.is_cuda:
            main = torch.cuda.current_stream(flat.device)
            if self._side is None:
                self._side = torch.cuda.Stream(flat.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def ready(self, flat, first_param_index):
        """gradients of parameters >= first_param_index are final"""
        while self._next > 0 and self.bounds[self._next - 1] >= first_param_index:
            self._reduce(flat, self.bounds[self._next - 1], self.bounds[self._next])
            self._next -= 1

    def finish(self, flat):
        self.ready(flat, 0)
        if flat.is_cuda and self._side is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self._side)
        self.reset()


def broadcast_from_rank0_(flat, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, 0, group=group)
    return flat


def max_over_ranks(value, device=None, group=None):
    """Timing helper: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
