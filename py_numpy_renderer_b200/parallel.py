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


class _DevView:
    """A raw device range presented through __cuda_array_interface__ (so torch / CuPy can wrap it without a copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class FrameWindow:
    """Frames of all ranks assembled on rank `dst` WITHOUT a collective: `dst` owns a device buffer of
    (slots, world, n, H, W, 3) uint8 and exports it through CUDA IPC; every other rank maps it and hands
    `block_ptr(slot)` to `render_batch(..., out=)`, so its shading kernel stores the finished pixels straight into
    `dst`'s HBM over NVLink / NVSwitch.  What remains of the gather is `fence()`: a one-element all-reduce (NCCL) or a
    host barrier (gloo) after which `dst` may read the slot.  Same node only; the handle travels through
    torch.distributed's object broadcast, whatever the backend."""

    def __init__(self, n, height, width, slots=2, dst=0):
        from . import _native
        d = dist()
        self.rank, self.world = (d.get_rank(), d.get_world_size()) if d.is_initialized() else (0, 1)
        self.dst, self.slots, self.shape = dst, slots, (n, height, width, 3)
        self.block_bytes = n * height * width * 3
        total = slots * self.world * self.block_bytes
        self.owner = self.rank == dst
        box = [None]
        if self.owner:
            self.base, handle = _native.window_create(total)
            box = [handle]
        if self.world > 1:
            d.broadcast_object_list(box, src=dst)
        if not self.owner:
            self.base = _native.window_open(box[0])
        self._token = None

    def block_ptr(self, slot, rank=None):
        """Device address this rank renders its n frames of `slot` to (an int: pass it as `out=`)."""
        rank = self.rank if rank is None else rank
        return self.base + (slot * self.world + rank) * self.block_bytes

    def block(self, slot):
        """This rank's block of `slot` as a CUDA array-interface object (n, H, W, 3): `torch.as_tensor(w.block(s),
        device='cuda').copy_(local, non_blocking=True)` pushes finished frames with the copy engines instead of the
        shading kernel's own stores (useful when many ranks would otherwise burst into `dst` at the same moment)."""
        return _DevView(self.block_ptr(slot), self.shape)

    def push(self, slot, frames, stream=None):
        """Push this rank's finished frames (a CUDA tensor (n, H, W, 3) uint8 on this rank's device) into its block of
        `slot`, asynchronously on `stream` (a torch.cuda.Stream; default: the current one).  Sparse: 32x32 tiles that hold
        one colour now and held exactly that after the previous push into the same block are not sent again
        (`b2r_window_push`, a kernel with peer stores) -- the constant background of a frame crosses NVLink once.  Falls
        back to a plain copy (copy engines) when the width is not a multiple of 32."""
        import torch
        from . import _native
        n, H, W, _ = self.shape
        stream = stream or torch.cuda.current_stream()
        if W % 32 != 0 or frames.data_ptr() % 16 != 0 or self.block_ptr(slot) % 16 != 0:
            with torch.cuda.stream(stream):
                torch.as_tensor(self.block(slot), device=frames.device).copy_(frames, non_blocking=True)
            return
        if not hasattr(self, "_push_state"):
            self._push_state = {}
        st = self._push_state.get(slot)
        if st is None:   # nothing is known about the block's content yet: every tile of the first push is sent
            st = self._push_state[slot] = torch.zeros(n * ((H + 31) // 32) * (W // 32), dtype=torch.int32, device=frames.device)
            stream.wait_stream(torch.cuda.current_stream())   # the zero fill above
        _native.window_push(frames.data_ptr(), self.block_ptr(slot), n, H, W, st.data_ptr(), stream.cuda_stream)

    def frames(self, slot):
        """On `dst`: the assembled (world * n, H, W, 3) frames of `slot` as a zero-copy CUDA array-interface object
        (torch.as_tensor(w.frames(s), device='cuda') wraps it)."""
        if not self.owner:
            return None
        n, H, W, _ = self.shape
        return _DevView(self.block_ptr(slot, 0), (self.world * n, H, W, 3))

    def fence(self, stream_sync=None, async_op=False):
        """Everything the ranks rendered into the window so far is complete and visible on `dst` once this (or the
        returned work handle) finishes.  NCCL: a one-element all-reduce ordered after the current torch stream --
        make that stream wait for the render first.  gloo: `stream_sync()` (e.g. _native.sync) then a host barrier."""
        import torch
        d = dist()
        if self.world == 1:
            if stream_sync:
                stream_sync()
            return None
        if d.get_backend() == "nccl":
            if self._token is None:
                self._token = torch.zeros(1, dtype=torch.int32, device="cuda")
            return d.all_reduce(self._token, async_op=async_op)
        if stream_sync:
            stream_sync()
        d.barrier()
        return None

    def close(self):
        from . import _native
        if self.base:
            (_native.window_destroy if self.owner else _native.window_close)(self.base)
            self.base = 0
