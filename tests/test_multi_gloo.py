"""Host-side multi-GPU logic on CPU: world_size-2 gloo job exercising the frame sharding and the final gather
(py_numpy_renderer_b200/parallel.py).  Rendering itself is replaced by a deterministic stand-in (this box has no GPU);
what is checked is the partition, the rank order of the gather and the band re-assembly."""
import os
import socket

import numpy as np
import pytest

from py_numpy_renderer_b200 import parallel

torch = pytest.importorskip("torch")


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_bands_cover_the_frame_on_tile_rows():
    for H in (1, 31, 32, 1080, 2160):
        for world in (1, 2, 4, 8):
            bands = parallel.shard_bands(H, world)
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert all(a % 32 == 0 or a == H for a, _ in bands)


def _fake_frames(first, count, H=6, W=4):
    f = np.zeros((count, H, W, 3), np.uint8)
    for k in range(count):
        f[k] = (first + k) * 3 + np.arange(3, dtype=np.uint8)
    return torch.from_numpy(f)


def _worker(rank, world, port, n_frames):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    # frame sharding + gather
    counts = [b - a for a, b in (parallel.shard_range(n_frames, q, world) for q in range(world))]
    a, b = parallel.shard_range(n_frames, rank, world)
    full = parallel.gather_frames(_fake_frames(a, b - a), counts, dst=0)
    if rank == 0:
        assert torch.equal(full, _fake_frames(0, n_frames))
    else:
        assert full is None
    # row-band sharding + gather
    H = 70
    bands = parallel.shard_bands(H, world)
    truth = torch.arange(2 * H * 5 * 3, dtype=torch.int64).reshape(2, H, 5, 3).remainder(251).to(torch.uint8)
    r0, r1 = bands[rank]
    mine = torch.zeros_like(truth)
    mine[:, H - r1:H - r0] = truth[:, H - r1:H - r0]
    whole = parallel.gather_bands(mine, bands, H, dst=0)
    if rank == 0:
        assert torch.equal(whole, truth)
    parallel.dist().barrier()
    parallel.dist().destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_world2_gloo_gather(n_frames):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_frames), nprocs=2, join=True)


def test_assemble_bands_numpy():
    H = 64
    bands = parallel.shard_bands(H, 2)
    truth = np.random.default_rng(0).integers(0, 255, (1, H, 3, 3), dtype=np.uint8)
    parts = []
    for r0, r1 in bands:
        p = np.zeros_like(truth)
        p[:, H - r1:H - r0] = truth[:, H - r1:H - r0]
        parts.append(p)
    assert np.array_equal(parallel.assemble_bands(parts, bands, H), truth)
