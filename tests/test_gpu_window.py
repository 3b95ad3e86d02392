"""Multi-GPU output window (parallel.FrameWindow, b2r_window_*): a peer process renders straight into the assembling
rank's device buffer.  Two processes share the one GPU of the test box (CUDA IPC works between processes on the same
device as well); the control plane is gloo, as in tests/test_multi_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_RANK_SCRIPT = r"""
import os, sys
import numpy as np
sys.path[:0] = [{root!r}, {root!r} + '/tests']
import torch, torch.distributed as dist
import scenes
from py_numpy_renderer_b200 import _native, parallel
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
_native.init(0)
H, W, total = 270, 480, 6
scene = scenes.c3_synthetic((H, W), tex=128)
cams = scenes.orbit_cameras(total)
dcams = scenes.orbit_cameras(total, fovy=90, near=0.05, far=20)
b, e = parallel.shard_range(total, rank, world)
win = parallel.FrameWindow(e - b, H, W, slots=2, dst=0)
for slot in (0, 1, 0):                                   # a slot is reused after a fence
    scene.render_batch(cams[b:e], debug_cameras=dcams[b:e], out=win.block_ptr(slot))
    win.fence(stream_sync=_native.sync)
    if rank == 0:
        got = torch.as_tensor(win.frames(slot), device='cuda').cpu().numpy()
        want = scene.render_batch(cams, debug_cameras=dcams)
        assert got.shape == want.shape and np.array_equal(got, want), (slot, int((got != want).sum()))
    dist.barrier()                                        # rank 0 has read the slot before anybody overwrites it
# sparse push (b2r_window_push): frames rendered locally, tiles of constant colour sent once; the orbit changes between
# pushes into the same slot, so tiles turn from background to covered and back
H2, W2 = 288, 480                                        # the push wants a width that is a multiple of 32
scene2 = scenes.c3_synthetic((H2, W2), tex=128)
win2 = parallel.FrameWindow(e - b, H2, W2, slots=2, dst=0)
local = torch.empty((e - b, H2, W2, 3), dtype=torch.uint8, device='cuda')
side = torch.cuda.Stream()
for it, slot in enumerate((0, 1, 0, 0, 1)):
    cams2 = scenes.orbit_cameras(total, start=0.9 * it, radius=3.0 + 2.0 * (it % 2))
    dcams2 = scenes.orbit_cameras(total, start=0.9 * it, radius=3.0 + 2.0 * (it % 2), fovy=90, near=0.05, far=20)
    scene2.render_batch(cams2[b:e], debug_cameras=dcams2[b:e], out=local)
    _native.sync()
    win2.push(slot, local, side)
    side.synchronize()
    win2.fence(stream_sync=_native.sync)
    if rank == 0:
        got = torch.as_tensor(win2.frames(slot), device='cuda').cpu().numpy()
        want = scene2.render_batch(cams2, debug_cameras=dcams2)
        assert got.shape == want.shape and np.array_equal(got, want), ('push', it, slot, int((got != want).sum()))
    dist.barrier()
held = win2._push_state[0].cpu().numpy()
assert (held != 0).any() and (held == 0).any()           # some tiles are remembered as constant colour, some are not
win2.close()
win.close()
dist.destroy_process_group()
print('WINDOW-OK', rank)
"""


def test_peer_renders_into_the_assembling_ranks_window():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", _RANK_SCRIPT.format(root=ROOT)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=300))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for rank, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WINDOW-OK {rank}" in out, f"rank {rank}:\n{out[-1500:]}\n{err[-3000:]}"


def test_sparse_push_state_machine_single_process():
    """b2r_window_push on one device: a tile is skipped only if it holds one colour now AND the destination is known to hold
    exactly that.  The destination is poisoned between pushes where a skip is expected, so a wrong store / a wrong skip
    both show."""
    import numpy as np
    import torch
    from py_numpy_renderer_b200 import _native
    _native.init(0)
    n, H, W = 2, 80, 96                                   # 3 x 3 tiles per view, the last tile row 16 rows high
    tiles = n * 3 * 3
    dev = torch.device("cuda:0")
    state = torch.zeros(tiles, dtype=torch.int32, device=dev)
    dst = torch.full((n, H, W, 3), 7, dtype=torch.uint8, device=dev)

    def push(src):
        _native.window_push(src.data_ptr(), dst.data_ptr(), n, H, W, state.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()

    bg = torch.tensor([10, 200, 33], dtype=torch.uint8, device=dev)
    a = bg.expand(n, H, W, 3).contiguous()
    a[0, 40:50, 40:50] = 99                                # one covered tile (view 0, tile (1, 1))
    push(a)
    assert torch.equal(dst, a)                             # first push: everything is sent
    held = state.cpu().numpy().reshape(n, 3, 3)
    assert held[0, 1, 1] == 0 and (np.delete(held.reshape(-1), 4) == (0x01000000 | 10 | 200 << 8 | 33 << 16)).all()
    dst.fill_(7)                                           # poison: only tiles that are really sent get repaired
    push(a)
    want = torch.full_like(dst, 7)
    want[0, 32:64, 32:64] = a[0, 32:64, 32:64]             # the covered tile is sent again, the constant ones are not
    assert torch.equal(dst, want)
    dst.copy_(a)
    b = a.clone()
    b[0, 40:50, 40:50] = bg                                # the covered tile turns into background ...
    b[1, 70:75, 0:5] = 1                                   # ... a background tile of the short bottom row gets covered
    push(b)
    assert torch.equal(dst, b)
    c = b.clone()
    c[1] = torch.tensor([5, 5, 5], dtype=torch.uint8, device=dev)   # another constant colour: must be sent, then remembered
    push(c)
    assert torch.equal(dst, c)
    dst[1].fill_(7)
    push(c)
    assert bool((dst[1] == 7).all()) and torch.equal(dst[0], c[0])
