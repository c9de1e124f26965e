"""A/B timing of the march-kernel variants on one GPU (development tool, not the bench).

  python tools/ab_time.py [--workload orbit|nearest|hires256|hires512] [--flags 0,512,256,...] [--steps 5]

Times birp_render_cube_device with CUDA events (L2 flushed between launches) for every flag
combination given, and checks that all variants return bit-identical images.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def time_cmem(args, capi, S, torch, dev, stream, sptr):
    W, H = S.SXI_W, S.SXI_H
    poses = S.orbit_poses(args.frames)
    frames = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    d_frames = torch.from_numpy(frames.view(np.float32).reshape(-1, 16).copy()).to(dev)
    d_img = torch.zeros((len(poses), H, W), dtype=torch.float32, device=dev)
    P = capi.cmem_init()
    out = {"workload": "cmem", "frames": len(poses)}
    ref = None
    for fl in [int(x) for x in args.flags.split(",")]:
        mode = {capi.CMEM_FAST: True, capi.CMEM_LIBM64: "libm64", capi.CMEM_NO_FILTER: "nofilter"}.get(fl, False)
        ms = []
        for it in range(args.steps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            capi.render_cmem_device(P, d_frames.data_ptr(), len(poses), W, H, d_img.data_ptr(), stream=sptr, fast=mode)
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 1:
                ms.append(e0.elapsed_time(e1))
        img = d_img.clone()
        if ref is None:
            ref = img
        rel = ((img - ref).abs() <= 1e-5 * ref.abs().clamp_min(1e-5 * float(ref.max()))).float().mean().item()
        out[f"flags_{fl}"] = {"ms_min": round(min(ms), 3), "frames_per_s": round(len(poses) / (min(ms) * 1e-3), 1),
                              "pixels_within_1e-5_of_first": rel,
                              "bit_identical_to_first": float((img.view(torch.int32) == ref.view(torch.int32)).float().mean())}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="orbit")
    ap.add_argument("--flags", default="0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--frames", type=int, default=10000)
    args = ap.parse_args()
    import torch

    from birp_cpp_b200 import capi, synthetic as S

    capi.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    tri = args.workload != "nearest"
    if args.workload == "cmem":
        return time_cmem(args, capi, S, torch, dev, stream, sptr)
    if args.workload in ("orbit", "nearest"):
        n, W, H, fw, fh, ns = 256, S.SXI_W, S.SXI_H, S.SXI_FOV_W, S.SXI_FOV_H, 200
        poses = S.orbit_poses(args.frames)
    elif args.workload == "hires256":
        n, W, H, fw, fh, ns = 256, 1024, 1024, 36.0, 36.0, 2048
        poses = np.array([S.DEFAULT_POS], np.float32)
    else:
        n, W, H, fw, fh, ns = 512, 4096, 4096, 36.0, 36.0, 200
        poses = np.array([S.DEFAULT_POS], np.float32)
    zyx, org, sp = S.make_cube(n, n, n)
    cube = capi.Cube(zyx, org, sp)
    frames = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, fw, fh, W, H)
    d_frames = torch.from_numpy(frames.view(np.float32).reshape(-1, 16).copy()).to(dev)
    d_img = torch.zeros((len(poses), H, W), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    out = {"workload": args.workload, "lib": os.environ.get("BIRP_B200_LIB", "default"), "info": cube.info(),
           "fetch_env": os.environ.get("BIRP_FETCH")}
    for fl in [int(x) for x in args.flags.split(",")]:
        ms = []
        for it in range(args.steps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            capi.render_cube_device(cube, d_frames.data_ptr(), len(poses), W, H, d_img.data_ptr(), nsamples=ns,
                                    maxdist=44, trilinear=tri, stream=sptr, extra_flags=fl)
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                ms.append(e0.elapsed_time(e1))
        capi.check_device_status(sptr)
        img = d_img.clone()
        same = True if ref is None else bool(torch.equal(img.view(torch.int32), ref.view(torch.int32)))
        if ref is None:
            ref = img
        out[f"flags_{fl}"] = {"ms_min": round(min(ms), 4), "ms_mean": round(float(np.mean(ms)), 4),
                              "frames_per_s": round(len(poses) / (min(ms) * 1e-3), 1), "same_as_first": same}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
