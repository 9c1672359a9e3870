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


class SonetComm:
    """The C-ABI collective (sonet_comm_init / sonet_allgather, csrc/comm.cu): an NCCL communicator
    created through libsonet_b200 on the current device. torch.distributed must be initialised
    (any backend): its process group is only the transport of the 128-byte unique id."""

    def __init__(self):
        import ctypes
        from . import _C
        self._C = _C
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", torch.cuda.current_device())
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            _C.check(_C.lib().sonet_comm_unique_id(idt.data_ptr()), "sonet_comm_unique_id")
        idt = idt.to(dev) if dist.get_backend() == "nccl" else idt
        dist.broadcast(idt, src=0)
        idt = idt.cpu().contiguous()
        handle = ctypes.c_void_p()
        _C.check(_C.lib().sonet_comm_init(idt.data_ptr(), rank, world, ctypes.byref(handle)),
                 "sonet_comm_init")
        self.handle, self.rank, self.world = handle, rank, world

    def all_gather(self, local, out, stream=None):
        """out [world*rows, ...] <- every rank's local [rows, ...]; asynchronous on `stream`
        (default: the current stream); CUDA-graph capturable."""
        local = local.contiguous()
        st = (stream or torch.cuda.current_stream(local.device)).cuda_stream
        self._C.check(self._C.lib().sonet_allgather(self.handle, local.data_ptr(), out.data_ptr(),
                                                    local.numel() * local.element_size(), st),
                      "sonet_allgather")
        return out

    def destroy(self):
        if self.handle:
            self._C.check(self._C.lib().sonet_comm_destroy(self.handle), "sonet_comm_destroy")
            self.handle = None


class AsyncGather:
    """Per-step all-gather of result rows that completes under the NEXT step's forward.

    launch(i, rows): stage the rows (a graph's static output buffer is overwritten by the next
    replay) and start the gather of step i; wait_prev(i): make the current stream wait for the
    gather of step i-1 — called at the end of step i, so every gather ends inside a step.
    impl 'torch': torch.distributed.all_gather_into_tensor(async_op=True) (NCCL's own stream);
    impl 'sonet': sonet_allgather on a side stream ordered by events."""

    def __init__(self, world, rows_shape, dev, impl="torch"):
        self.world, self.impl, self.dev = world, impl, dev
        self.stage = [torch.empty(rows_shape, dtype=torch.float32, device=dev) for _ in range(2)]
        full = (rows_shape[0] * world,) + tuple(rows_shape[1:])
        self.out = [torch.empty(full, dtype=torch.float32, device=dev) for _ in range(2)]
        self.pending = [None, None]
        if impl == "sonet":
            self.comm = SonetComm()
            self.side = torch.cuda.Stream(device=dev)
            self.ev_fwd = [torch.cuda.Event(), torch.cuda.Event()]
            self.ev_done = [torch.cuda.Event(), torch.cuda.Event()]

    def launch(self, i, rows):
        j = i & 1
        self.stage[j].copy_(rows)
        if self.impl == "torch":
            self.pending[j] = dist.all_gather_into_tensor(self.out[j], self.stage[j], async_op=True)
        else:
            cur = torch.cuda.current_stream(self.dev)
            self.ev_fwd[j].record(cur)
            self.side.wait_event(self.ev_fwd[j])
            self.comm.all_gather(self.stage[j], self.out[j], stream=self.side)
            self.ev_done[j].record(self.side)
            self.pending[j] = self.ev_done[j]
        return self.out[j]

    def _wait(self, j):
        p = self.pending[j]
        if p is None:
            return
        if self.impl == "torch":
            p.wait()
        else:
            torch.cuda.current_stream(self.dev).wait_event(p)
        self.pending[j] = None

    def wait_prev(self, i):
        self._wait((i - 1) & 1)

    def drain(self):
        self._wait(0)
        self._wait(1)
