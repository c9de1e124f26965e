"""The N>1 host logic (partitioning + gather to rank 0) on CPU: two gloo ranks, the oracle standing
in for the GPU renderer.  What rank 0 assembles must equal the unsharded render bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

from birp_cpp_b200 import sharding  # noqa: E402


def test_block_partition_properties():
    for total in (0, 1, 7, 64, 10000):
        for world in (1, 2, 3, 8):
            counts = sharding.block_counts(total, world)
            assert sum(counts) == total and max(counts) - min(counts) <= 1
            ranges = [sharding.block_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    assert sharding.frame_block(10000, 8, 3) == (3750, 5000)
    assert sharding.row_band(1024, 8, 7) == (896, 1024)
    assert sharding.row_band(64, 3, 0) == (0, 22)
    # interleaved row bands: the ranks' bands partition the image, a short last band included
    for height, band, world in ((64, 8, 4), (70, 16, 3), (4096, 32, 8), (24, 8, 5)):
        rows = sorted(r for rank in range(world) for b in sharding.interleaved_row_bands(height, band, world, rank)
                      for r in range(*b))
        assert rows == list(range(height))
    assert sharding.interleaved_row_bands(70, 16, 3, 1) == [(16, 32), (64, 70)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import oracle as O
    from birp_cpp_b200 import sharding as sh, synthetic as S

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H = 27, 16
        # cube replicated by broadcast from rank 0
        if rank == 0:
            zyx, org, sp = S.make_cube(21, 17, 19)
            cube = torch.from_numpy(zyx)
            geom = torch.from_numpy(np.concatenate([org, sp]))
        else:
            cube = torch.empty((19, 17, 21), dtype=torch.float32)
            geom = torch.empty(6, dtype=torch.float32)
        sh.broadcast_cube(cube, geom)
        zyx, org, sp = cube.numpy(), geom.numpy()[:3], geom.numpy()[3:]

        # (a) orbit sweep: 7 frames over 2 ranks (uneven blocks)
        poses = S.orbit_poses(7) * np.float32(0.4)
        frames = O.orc_frames(poses, S.DEFAULT_AIM_X, 27.0, 16.0, W, H)
        counts = sh.block_counts(len(frames), world)
        b, e = sh.frame_block(len(frames), world, rank)
        mine = O.orc_render_cube(zyx, org, sp, True, frames[b:e], W, H, nthreads=1)
        full = sh.gather_frames(torch.from_numpy(mine), counts)

        # (a2) the same sweep interleaved over the ranks (8 frames divide evenly), gathered in 3 overlapped pieces
        poses8 = S.orbit_poses(8) * np.float32(0.4)
        frames8 = O.orc_frames(poses8, S.DEFAULT_AIM_X, 27.0, 16.0, W, H)
        idx = sh.frame_indices(8, world, rank, interleave=True)
        assert list(idx) == list(range(rank, 8, world))
        mine8 = torch.from_numpy(O.orc_render_cube(zyx, org, sp, True, frames8[idx], W, H, nthreads=1))
        fg = sh.FrameGather(8, world, rank, (H, W), torch.float32, "cpu", interleave=True, nchunks=3)
        for c in range(fg.nchunks):
            fg.post(mine8, c)
        inter = fg.finish()
        # and with uneven blocks (7 frames): falls back to contiguous blocks, single layout
        fg7 = sh.FrameGather(7, world, rank, (H, W), torch.float32, "cpu", interleave=True, nchunks=2)
        assert not fg7.interleaved
        b7, e7 = sh.frame_block(7, world, rank)
        assert list(sh.frame_indices(7, world, rank, True)) == list(range(b7, e7))
        for c in range(fg7.nchunks):
            fg7.post(torch.from_numpy(mine), c)
        blocks7 = fg7.finish()

        # (b) one image split into row bands
        bands = sh.block_counts(H, world)
        r0, r1 = sh.row_band(H, world, rank)
        img = O.orc_render_cube(zyx, org, sp, True, frames[:1], W, H, rows=(r0, r1), nthreads=1)
        whole = sh.gather_row_bands(torch.from_numpy(img[0, r0:r1].copy()), bands)
        # (c) the end-to-end layout of bench.py at N > 1: ONE host buffer on rank 0 (shared memory) that every rank
        # writes its share into -- a sweep interleaved frame by frame, then one image in interleaved row bands
        shared = sh.SharedHostResult((8, H, W), rank, world, register=False)
        for j, f in enumerate(idx):
            shared.array[f] = mine8[j].numpy()
        dist.barrier()
        shared_sweep = shared.array.copy() if rank == 0 else None
        shared.close()
        one = sh.SharedHostResult((1, H, W), rank, world, register=False)
        my_bands = sh.interleaved_row_bands(H, 4, world, rank)
        for r0b, r1b in my_bands:
            one.array[0, r0b:r1b] = O.orc_render_cube(zyx, org, sp, True, frames[:1], W, H, rows=(r0b, r1b), nthreads=1)[0, r0b:r1b]
        dist.barrier()
        shared_image = one.array.copy() if rank == 0 else None
        one.close()
        if rank == 0:
            want = O.orc_render_cube(zyx, org, sp, True, frames, W, H, nthreads=1)
            want8 = O.orc_render_cube(zyx, org, sp, True, frames8, W, H, nthreads=1)
            np.savez(out_path, full=full.numpy(), want=want, whole=whole.numpy(),
                     inter=inter.reshape(8, H, W).numpy(), want8=want8, blocks7=blocks7.numpy(),
                     shared_sweep=shared_sweep, shared_image=shared_image)
        else:
            assert full is None and whole is None and inter is None and blocks7 is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "gathered.npz"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    d = np.load(out)
    assert d["full"].shape == d["want"].shape == (7, 16, 27)
    assert np.array_equal(d["full"].view(np.uint32), d["want"].view(np.uint32))
    assert np.array_equal(d["whole"].view(np.uint32), d["want"][0].view(np.uint32))
    assert np.array_equal(d["inter"].view(np.uint32), d["want8"].view(np.uint32))
    assert np.array_equal(d["blocks7"].view(np.uint32), d["want"].view(np.uint32))
    assert np.array_equal(d["shared_sweep"].view(np.uint32), d["want8"].view(np.uint32))
    assert np.array_equal(d["shared_image"][0].view(np.uint32), d["want"][0].view(np.uint32))
    assert d["want"].max() > 0
