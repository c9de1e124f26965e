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


def test_one_process_multi_gpu_sweep_equals_single_gpu(gpu):
    """birp_render_*_sweep_multi: one process, one host thread per GPU, chunks of frames handed out from a shared
    counter, every GPU draining into the one host buffer; the cube replicated device to device"""
    ndev = gpu.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, str(ROOT))
    from birp_cpp_b200 import capi, synthetic as S
    W, H = S.SXI_W, S.SXI_H
    zyx, org, sp = S.make_cube(64, 64, 64)
    capi.set_device(0)
    cube0 = capi.Cube(zyx, org, sp)
    cubes = [cube0] + [cube0.replicate(d) for d in range(1, ndev)]
    assert [c.device for c in cubes] == list(range(ndev))
    assert all(c.info() == cube0.info() for c in cubes)
    poses = S.orbit_poses(1500)
    try:
        for tri, fl in ((True, 0), (False, 0), (True, capi.FAST_MARCH)):
            want_fr = np.zeros(len(poses), capi.FRAME_DTYPE)
            want = capi.render_sweep(cube0, poses, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H, trilinear=tri,
                                     extra_flags=fl, frames_out=want_fr)
            got_fr = np.zeros(len(poses), capi.FRAME_DTYPE)
            out = capi.PinnedArray((len(poses), H, W))
            out.array[:] = -1.0
            capi.render_sweep_multi(cubes, poses, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H, trilinear=tri,
                                    extra_flags=fl, frames_out=got_fr, out=out.array)
            assert np.array_equal(out.array.view(np.uint32), want.view(np.uint32)), (tri, fl)
            assert np.array_equal(got_fr.view(np.uint32), want_fr.view(np.uint32))
            out.free()
        P = capi.cmem_init()
        want = capi.render_sweep(P, poses[:600], S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
        got = capi.render_sweep_multi(P, poses[:600], S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H,
                                      devices=list(range(ndev)))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # a camera on the x axis has a NaN orientation (the reference throws std::out_of_range from .at()): the sweep
        # stops with that error, whichever device met the pose
        bad = poses.copy()
        bad[700] = (10.0, 0.0, 0.0)
        with pytest.raises(capi.BirpError) as err:
            capi.render_sweep_multi(cubes, bad, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
        assert err.value.code == capi.BIRP_ERR_RANGE
        assert capi.render_sweep_multi(cubes, poses[:130], S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H).max() > 0
    finally:
        for c in cubes:
            c.free()
        capi.set_device(0)


def test_cpp_render_sweep_and_batch_use_every_gpu(gpu, tmp_path):
    """RenderSweep / RenderBatchDirectory of the C++ classes on a multi-GPU box: same images and FITS files as with
    one device"""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, str(ROOT / "tests"))
    import mirror as M
    from birp_cpp_b200 import synthetic as S
    zyx, org, sp = S.make_cube(48, 40, 44)
    cube = M.Cube(zyx, org, sp, True)
    poses = S.orbit_poses(400)
    os.environ["BIRP_DEVICES"] = "1"
    one = M.render_sweep(cube, poses, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    os.environ.pop("BIRP_DEVICES")
    many = M.render_sweep(cube, poses, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    cm = M.Cmem()
    os.environ["BIRP_DEVICES"] = "1"
    one = M.render_sweep(cm, poses, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    os.environ.pop("BIRP_DEVICES")
    many = M.render_sweep(cm, poses, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    # batch directory: 6 small cubes over the devices
    d_in, d_one, d_many = tmp_path / "in", tmp_path / "one", tmp_path / "many"
    for d in (d_in, d_one, d_many):
        d.mkdir()
    for k in range(6):
        z, o, s = S.make_cube(21, 17, 19, seed=100 + k)
        S.write_dat(d_in / f"cube_{k}.dat", z, o, s)
    os.environ["BIRP_DEVICES"] = "1"
    n1, f1 = M.render_batch(d_in, d_one, S.DEFAULT_POS, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    os.environ.pop("BIRP_DEVICES")
    n2, f2 = M.render_batch(d_in, d_many, S.DEFAULT_POS, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    assert (n1, f1) == (n2, f2) == (6, 0)
    for k in range(6):
        assert (d_one / f"cube_{k}.dat.fits").read_bytes() == (d_many / f"cube_{k}.dat.fits").read_bytes()
