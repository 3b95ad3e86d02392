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
