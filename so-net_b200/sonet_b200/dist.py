"""Batch-sharded multi-GPU forward: one process per GPU, weights replicated,
the global batch split contiguously over ranks, and ONE all-gather per forward of the per-shard
logits (or per-cloud losses) over NCCL/NVLink. Every op of the eval forward is per-cloud, so there
is no exchange step inside the path and the gathered result is bit-identical to a 1-GPU run.

The same code runs on CPU with the gloo backend (tests/test_dist_gloo.py, world_size 2) — only the
collective plumbing is exercised there, never the kernels.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, local_rank, world).
    With WORLD_SIZE=1 no process group is created."""
    rank, local_rank, world = env_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % max(torch.cuda.device_count(), 1))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_bounds(total, rank, world):
    """Contiguous split of `total` clouds: the first (total % world) ranks get one extra."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, total_rows=None, out=None):
    """Gather [rows_r, ...] shards from every rank into [sum rows, ...] in rank order.
    Equal shards use one all_gather_into_tensor (into `out` when given: a persistent buffer avoids
    a caching-allocator round trip per step); ragged shards pad to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    local = local.contiguous()
    if total_rows is None or total_rows % world == 0:
        if out is None:
            out = local.new_empty((local.shape[0] * world,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local)
        return out
    sizes = [shard_bounds(total_rows, r, world) for r in range(world)]
    rows = max(hi - lo for lo, hi in sizes)
    padded = local.new_zeros((rows,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((rows * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * rows:r * rows + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


class ShardedForward:
    """Runs `forward_fn(shard_inputs) -> [rows, ...]` on this rank's contiguous slice of a global
    batch and all-gathers the rows. `forward_fn` is e.g. a classifier.Model's set_input+test_model."""

    def __init__(self, forward_fn):
        self.forward_fn = forward_fn
        self.rank, _, self.world = env_world()

    def __call__(self, global_inputs, total_rows):
        lo, hi = shard_bounds(total_rows, self.rank, self.world)
        shard = {k: v[lo:hi] for k, v in global_inputs.items()}
        return all_gather_rows(self.forward_fn(shard), total_rows)
