"""Column sharding over the GPUs of one node (one process per GPU, torch.distributed).

Every CLOUDSC2 stencil is pointwise in the column index, so rank r owns the contiguous block of
columns [start, stop) and no data-path collective exists.  The only communication of the whole
path is the all-reduce of the Taylor-test sums (SUM) and of the symmetry-test maximum (MAX), a few
hundred bytes per test, plus the broadcast of the 137 eta values derived from GLOBAL column 0
(reference physics/common/diagnostics.py:42-45).
"""
from __future__ import annotations

import contextlib
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_columns(nx_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: sizes differ by at most one column."""
    base, rem = divmod(nx_global, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


_local_only = 0


def is_distributed() -> bool:
    return _local_only == 0 and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


@contextlib.contextmanager
def local_only():
    """Inside this context the collectives below are no-ops although a process group exists: a rank runs a complete,
    un-sharded control problem next to the sharded one (bench.py's config-5 check, tests/dist_gpu_worker.py)."""
    global _local_only
    _local_only += 1
    try:
        yield
    finally:
        _local_only -= 1


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank); a no-op for single-process runs."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def allreduce_sum_(buf: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks of a small fp64 buffer (Taylor sums)."""
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def allreduce_max_(buf: torch.Tensor) -> torch.Tensor:
    """In-place MAX over ranks (symmetry-test norm3, timing)."""
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.MAX)
    return buf


def broadcast_eta(eta_field, src: int = 0) -> None:
    """Replace the local eta K-field by the one of rank `src` (which owns global column 0)."""
    if not is_distributed():
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = eta_field.buffer.to(dev)
    dist.broadcast(buf, src=src)
    eta_field.buffer.copy_(buf.cpu())
