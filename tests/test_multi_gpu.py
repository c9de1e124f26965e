"""Two processes, two GPUs: the sweep sharded over ranks with the gather fused into the render kernel
(peer stores into rank 0's buffer through a CUDA IPC mapping).  Skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    import oracle as O
    from birp_cpp_b200 import capi, sharding as sh, synthetic as S

    torch.cuda.set_device(rank)
    capi.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        W, H, n = 27, 16, 10
        zyx, org, sp = S.make_cube(33, 29, 31)
        cube = capi.Cube(zyx, org, sp)
        poses = S.orbit_poses(n) * np.float32(0.5)
        frames = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, 27.0, 16.0, W, H)
        idx = sh.frame_indices(n, world, rank, interleave=True)
        d_frames = torch.from_numpy(np.ascontiguousarray(frames[idx]).view(np.float32).reshape(-1, 16).copy()).to(dev)
        peer = sh.PeerGather(n * H * W, rank, world, dev)
        stream = torch.cuda.current_stream().cuda_stream
        for tri in (True, False):
            capi.render_cube_device(cube, d_frames.data_ptr(), len(idx), W, H, peer.ptr, trilinear=tri, stream=stream,
                                    frame_offset=int(idx[0]), frame_stride=world)
            peer.fence()
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                got = peer.tensor((n, H, W)).cpu().numpy()
                fo = frames.view(np.float32).reshape(-1, 16)
                want = O.orc_render_cube(zyx, org, sp, tri, fo, W, H, nthreads=2)
                np.save(out_path + f".tri{int(tri)}.npy", np.stack([got, want]))
            dist.barrier()
        # one image in two row bands, both ranks writing into rank 0's single image
        one = sh.PeerGather(H * W, rank, world, dev)
        r0, r1 = sh.row_band(H, world, rank)
        d_one = torch.from_numpy(frames[:1].view(np.float32).reshape(-1, 16).copy()).to(dev)
        capi.render_cube_device(cube, d_one.data_ptr(), 1, W, H, one.ptr, rows=(r0, r1), stream=stream)
        one.fence()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            got = one.tensor((1, H, W)).cpu().numpy()
            want = O.orc_render_cube(zyx, org, sp, True, frames[:1].view(np.float32).reshape(-1, 16), W, H, nthreads=2)
            np.save(out_path + ".bands.npy", np.stack([got, want]))
        one.close()
        peer.close()
        cube.free()
    finally:
        dist.destroy_process_group()


def test_two_gpu_peer_store_gather(gpu, tmp_path):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "peer")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for tag in ("tri1", "tri0", "bands"):
        got, want = np.load(out + f".{tag}.npy")
        assert want.max() > 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), tag
