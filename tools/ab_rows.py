"""A/B timing of ONE rank's share of config 4 (one 4096x4096 image of the 512^3 cube in interleaved row bands) on a
single GPU, without torch: what a rank of an N-GPU run launches, so that CTA-size / band-height variants can be
compared without an N-GPU box.  BIRP_B200_LIB selects an A/B build of the library.

  [BIRP_B200_LIB=tools/scratch/libbirp_b200_w4.so] python tools/ab_rows.py [--nphases 8] [--bands 128,32] [--steps 6]

Prints one JSON line: per band height the best / mean kernel time of phase 0's share, of the whole image, and whether
the N shares together equal the whole image bit for bit.  L2 is flushed before every timed launch (a 256 MiB memset),
as bench.py does.  Development tool, not the bench.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nphases", type=int, default=8)
    ap.add_argument("--bands", default="128,32")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--cube", type=int, default=512)
    ap.add_argument("--image", type=int, default=4096)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--nsamples", type=int, default=200)
    ap.add_argument("--warps", type=int, default=0, help="birp_set_tuning('cta_warps'): 2, 4, 8; 0 = the library's choice")
    args = ap.parse_args()
    t0 = time.perf_counter()
    from ab_quick import cudart
    from birp_cpp_b200 import capi, synthetic as S

    rt, _ = cudart()
    rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    capi.lib()
    capi.set_device(0)
    capi.set_tuning("cta_warps", args.warps)
    n, W = args.cube, args.image
    H = W
    zyx, org, sp = S.make_cube(n, n, n)
    poses = np.array([S.DEFAULT_POS], np.float32)
    frames = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, 36.0, 36.0, W, H)
    cube = capi.Cube(zyx, org, sp)
    d_frames = capi.DeviceBuffer(16)
    d_img = capi.DeviceBuffer(H * W)
    d_flush = capi.DeviceBuffer(64 << 20)  # floats: 256 MiB
    fr = np.ascontiguousarray(frames).view(np.float32).reshape(-1, 16)

    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: cudaError {rc}")

    ck(rt.cudaMemcpy(d_frames.ptr, fr.ctypes.data, fr.nbytes, 1), "H2D frames")
    e0, e1 = C.c_void_p(), C.c_void_p()
    ck(rt.cudaEventCreate(C.byref(e0)), "event")
    ck(rt.cudaEventCreate(C.byref(e1)), "event")

    def timed(launch):
        ms = []
        for it in range(args.steps + 1):
            ck(rt.cudaMemset(d_flush.ptr, it & 1, 256 << 20), "flush")
            ck(rt.cudaEventRecord(e0, None), "record")
            launch()
            ck(rt.cudaEventRecord(e1, None), "record")
            ck(rt.cudaEventSynchronize(e1), "sync")
            t = C.c_float(0)
            ck(rt.cudaEventElapsedTime(C.byref(t), e0, e1), "elapsed")
            if it >= 1:
                ms.append(t.value)
        capi.check_device_status(0)
        return {"ms_min": round(min(ms), 4), "ms_mean": round(float(np.mean(ms)), 4)}

    out = {"lib": str(capi.LIB_PATH), "cube": n, "image": W, "nphases": args.nphases, "flags": args.flags,
           "nsamples": args.nsamples, "cta_warps": args.warps}
    whole = np.empty((H, W), np.float32)
    out["whole"] = timed(lambda: capi.render_cube_device(cube, d_frames.ptr, 1, W, H, d_img.ptr, trilinear=True, stream=0,
                                                         extra_flags=args.flags, nsamples=args.nsamples))
    ck(rt.cudaMemcpy(whole.ctypes.data, d_img.ptr, whole.nbytes, 2), "D2H")
    for band in [int(b) for b in args.bands.split(",")]:
        ck(rt.cudaMemset(d_img.ptr, 0, whole.nbytes), "clear")
        r = timed(lambda: capi.render_cube_device_rows(cube, d_frames.ptr, 1, W, H, d_img.ptr, band, 0, args.nphases,
                                                       trilinear=True, stream=0, extra_flags=args.flags,
                                                       nsamples=args.nsamples))
        for ph in range(1, args.nphases):
            capi.render_cube_device_rows(cube, d_frames.ptr, 1, W, H, d_img.ptr, band, ph, args.nphases, trilinear=True,
                                         stream=0, extra_flags=args.flags, nsamples=args.nsamples)
        got = np.empty_like(whole)
        ck(rt.cudaMemcpy(got.ctypes.data, d_img.ptr, got.nbytes, 2), "D2H")
        r["shares_equal_whole"] = bool(np.array_equal(got.view(np.uint32), whole.view(np.uint32)))
        r["whole_over_share"] = round(out["whole"]["ms_min"] / r["ms_min"], 3)
        out[f"band_{band}"] = r
    out["total_s"] = round(time.perf_counter() - t0, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
