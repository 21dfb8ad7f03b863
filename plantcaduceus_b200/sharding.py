"""Window sharding across the GPUs of one box.

Every 512-bp window is an independent unit of the scoring path (reference src/zero_shot_score.py:111-120: batches
are independent, ``shuffle=False``), so the N windows of a job are split into ``world_size`` contiguous ranges, one
process per GPU scores its range with replicated weights, and the only exchange is the gather of per-variant
outputs (4 floats per window) to rank 0.  Order is preserved by construction.  The same code runs over NCCL
(GPU tensors) and gloo (CPU tensors; used by the world-size-2 tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, end) of rank's contiguous share of n items; sizes differ by at most one, earlier ranks get the extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Joins the torchrun rendezvous if WORLD_SIZE > 1. Returns (rank, world_size)."""
    rank, _local, world = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if (device is not None and device.type == "cuda") else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Concatenates every rank's [n_local, k] rows in rank order; returns the [n_total, k] tensor on every rank
    (all_gather of padded shards: shard sizes differ by at most one row)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size()
    max_rows = -(-n_total // world)
    padded = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = []
    for r in range(world):
        s, e = shard_range(n_total, r, world)
        out.append(parts[r][: e - s])
    return torch.cat(out, dim=0)


def scatter_rows(rows: Optional[torch.Tensor], device: Optional[torch.device] = None, src: int = 0) -> torch.Tensor:
    """Rank ``src`` holds a [n, k] tensor (the parsed input); every rank gets back its contiguous ``shard_range`` share,
    so the input files are read once, not once per rank.  Other ranks pass ``None``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        meta = [(int(rows.shape[0]), tuple(rows.shape[1:]), rows.dtype)]
    dist.broadcast_object_list(meta, src=src)
    n, tail, dtype = meta[0]
    device = device if device is not None else (rows.device if rows is not None else torch.device("cpu"))
    max_rows = -(-n // world) if n else 0
    mine = torch.empty((max_rows,) + tuple(tail), dtype=dtype, device=device)
    parts = None
    if rank == src:
        parts = []
        for r in range(world):
            s, e = shard_range(n, r, world)
            p = torch.zeros((max_rows,) + tuple(tail), dtype=dtype, device=device)
            p[: e - s] = rows[s:e].to(device)
            parts.append(p)
    if max_rows:
        dist.scatter(mine, parts, src=src)
    s, e = shard_range(n, rank, world)
    return mine[: e - s]


def score_sharded(score_fn, ascii_windows: np.ndarray, device: Optional[torch.device] = None) -> np.ndarray:
    """Scores this rank's contiguous share of ``ascii_windows`` [n, L] with ``score_fn(uint8 [m, L]) -> float32 [m, 4]``
    and returns the full [n, 4] result on every rank."""
    rank, _local, world = env_world()
    n = int(ascii_windows.shape[0])
    s, e = shard_range(n, rank, world)
    local = score_fn(ascii_windows[s:e]) if e > s else np.zeros((0, 4), dtype=np.float32)
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float32))
    if device is not None and device.type == "cuda" and world > 1:
        t = t.to(device)
    full = gather_rows(t, n)
    return full.cpu().numpy()
