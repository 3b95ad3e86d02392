"""Multi-GPU sharding of the frame pipeline: one process per GPU, frames (or screen row bands) are
independent, so there is NO data-path collective -- only one gather of finished uint8 rows to assemble the output.

Plumbing is torch.distributed: NCCL over NVLink when the tensors live on a GPU, gloo for the CPU tests."""
from __future__ import annotations

import os

import numpy as np


def dist():
    import torch.distributed as d
    return d


def init_process_group(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  Returns (rank, world)."""
    import torch
    d = dist()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not d.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", rank))
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        d.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of `n_items` owned by `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_bands(height: int, world: int, align: int = 32):
    """Split buffer rows [0,height) into `world` bands on tile-row boundaries (the raster kernel owns 32-row
    tiles; an unaligned split would only duplicate work, results are identical either way)."""
    tiles = (height + align - 1) // align
    edges = [min(height, shard_range(tiles, r, world)[0] * align) for r in range(world)] + [height]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def gather_frames(local, counts, dst: int = 0):
    """Gather per-rank frame blocks (n_r, H, W, 3) uint8 to `dst` in rank order -> (sum n_r, H, W, 3) or None.

    `counts` = frames per rank (known to everybody from shard_range).  Blocks are padded to the largest count so
    a single fixed-size gather is used (NCCL gather over NVLink for CUDA tensors)."""
    import torch
    d = dist()
    if not d.is_initialized() or d.get_world_size() == 1:
        return local
    rank, world = d.get_rank(), d.get_world_size()
    cap = max(counts)
    block = local
    if local.shape[0] < cap:
        block = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        block[:local.shape[0]] = local
    recv = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    d.gather(block.contiguous(), recv, dst=dst)
    if rank != dst:
        return None
    return torch.cat([recv[r][:counts[r]] for r in range(world)], dim=0)


def assemble_bands(parts, bands, height: int):
    """Stack row bands (each a full-height frame stack with only its band's OUTPUT rows filled) into full frames.
    Output rows are flipped buffer rows: buffer band [r0,r1) lives in output rows [H-r1, H-r0)."""
    out = parts[0].clone() if hasattr(parts[0], "clone") else np.array(parts[0], copy=True)
    for part, (r0, r1) in zip(parts[1:], bands[1:]):
        out[:, height - r1:height - r0] = part[:, height - r1:height - r0]
    return out


def gather_bands(local, bands, height: int, dst: int = 0):
    """Single-frame(-stack) screen split: every rank rendered band `bands[rank]` into a full-size stack."""
    import torch
    d = dist()
    if not d.is_initialized() or d.get_world_size() == 1:
        return local
    rank, world = d.get_rank(), d.get_world_size()
    rows = max(r1 - r0 for r0, r1 in bands)
    r0, r1 = bands[rank]
    block = torch.zeros((local.shape[0], rows) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
    block[:, :r1 - r0] = local[:, height - r1:height - r0]
    recv = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    d.gather(block, recv, dst=dst)
    if rank != dst:
        return None
    out = torch.empty_like(local)
    for r, (a, b) in enumerate(bands):
        out[:, height - b:height - a] = recv[r][:, :b - a]
    return out
