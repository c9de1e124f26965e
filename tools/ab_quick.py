"""A/B timing of march-kernel variants WITHOUT torch (seconds instead of tens of seconds of start-up): the orbit
workload through birp_render_cube_device on the default stream, CUDA events from libcudart via ctypes.

  python tools/ab_quick.py [--flags 0,65536,32768:fast_variant=0:fast_kstep=0] [--frames 10000] [--steps 4] [--dry]

A combination is BIRP_* flags, optionally followed by :knob=value settings for birp_set_tuning (reset to the defaults
before every combination).

Prints one JSON line: per flag combination the best and mean kernel time, frames/s, and whether its images are
bit-identical to those of the first combination.  No L2 flush between launches (the oct copy of the 256^3 cube is
537 MB, four times the L2).  Development tool, not the bench.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def cudart():
    cands = glob.glob(str(Path(np.__file__).resolve().parents[1] / "nvidia" / "cuda_runtime" / "lib" / "libcudart.so*"))
    cands += glob.glob("/usr/local/cuda/targets/x86_64-linux/lib/libcudart.so.*") + ["libcudart.so.12", "libcudart.so"]
    for c in cands:
        try:
            rt = C.CDLL(c)
        except OSError:
            continue
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
        rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
        rt.cudaEventSynchronize.argtypes = [C.c_void_p]
        rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        return rt, c
    raise OSError("no libcudart found")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flags", default="0,16384")
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--e2e", action="store_true", help="also time birp_render_cube_sweep (host poses -> pinned images)")
    ap.add_argument("--dry", action="store_true", help="everything that needs no GPU, then stop")
    args = ap.parse_args()
    t0 = time.perf_counter()
    from birp_cpp_b200 import capi, synthetic as S

    rt, rt_path = cudart()
    capi.lib()
    W, H = S.SXI_W, S.SXI_H
    n = args.cube
    zyx, org, sp = S.make_cube(n, n, n)
    poses = S.orbit_poses(args.frames)
    frames = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    combos = []
    for spec in args.flags.split(","):
        parts = spec.split(":")
        # a leading "n" selects nearest-neighbour sampling ("n0")
        combos.append((spec, int(parts[0].lstrip("n")), [(kv.split("=")[0], int(kv.split("=")[1])) for kv in parts[1:]]))
    flags = [c[0] for c in combos]
    out = {"workload": "orbit", "frames": len(poses), "cube": n, "cudart": rt_path,
           "host_setup_s": round(time.perf_counter() - t0, 2)}
    if args.dry:
        print(json.dumps(out))
        return

    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: cudaError {rc}")

    capi.set_device(0)
    cube = capi.Cube(zyx, org, sp)
    out["info"] = cube.info()
    d_frames = capi.DeviceBuffer(len(poses) * 16)
    d_img = capi.DeviceBuffer(len(poses) * H * W)
    fr = np.ascontiguousarray(frames).view(np.float32).reshape(-1, 16)
    ck(rt.cudaMemcpy(d_frames.ptr, fr.ctypes.data, fr.nbytes, 1), "H2D frames")
    e0, e1 = C.c_void_p(), C.c_void_p()
    ck(rt.cudaEventCreate(C.byref(e0)), "event")
    ck(rt.cudaEventCreate(C.byref(e1)), "event")
    ref = None
    img = np.empty((len(poses), H, W), np.float32)
    def tune(knobs):
        for k, v in (("fast_variant", -1), ("fast_kstep", 1), ("tile_swizzle", 0), ("tile_w_log2", -1), ("fast_coop", 0), ("cta_warps", 0)):
            capi.set_tuning(k, v)
        for k, v in knobs:
            capi.set_tuning(k, v)

    for spec, fl, knobs in combos:
        tune(knobs)
        ms = []
        for it in range(args.steps + 1):
            ck(rt.cudaEventRecord(e0, None), "record")
            capi.render_cube_device(cube, d_frames.ptr, len(poses), W, H, d_img.ptr, trilinear=not spec.startswith("n"),
                                    stream=0, extra_flags=fl)
            ck(rt.cudaEventRecord(e1, None), "record")
            ck(rt.cudaEventSynchronize(e1), "sync")
            t = C.c_float(0)
            ck(rt.cudaEventElapsedTime(C.byref(t), e0, e1), "elapsed")
            if it >= 1:
                ms.append(t.value)
        capi.check_device_status(0)
        ck(rt.cudaMemcpy(img.ctypes.data, d_img.ptr, img.nbytes, 2), "D2H images")
        same = True if ref is None else bool(np.array_equal(img.view(np.uint32), ref.view(np.uint32)))
        if ref is None:
            ref = img.copy()
        rel = float((np.abs(img.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1e-5 * np.abs(ref).max())).max())
        out[f"flags_{spec}"] = {"ms_min": round(min(ms), 4), "ms_mean": round(float(np.mean(ms)), 4),
                                "frames_per_s": round(len(poses) / (min(ms) * 1e-3), 1), "same_as_first": same,
                                "max_rel_diff_vs_first": rel, "nonzero_pixels": int(np.count_nonzero(img))}
    if args.e2e:
        # end to end through birp_render_cube_sweep: host poses in, pinned host images out (wall clock, the call
        # synchronises before it returns)
        pinned = capi.PinnedArray((len(poses), H, W))
        for spec, fl, knobs in combos:
            tune(knobs)
            secs = []
            for it in range(args.steps + 1):
                t1 = time.perf_counter()
                capi.render_sweep(cube, poses, S.DEFAULT_AIM_X, S.SXI_FOV_W, S.SXI_FOV_H, W, H, trilinear=True,
                                  extra_flags=fl, out=pinned.array)
                if it >= 1:
                    secs.append(time.perf_counter() - t1)
            out[f"e2e_flags_{spec}"] = {"ms_min": round(min(secs) * 1e3, 4), "frames_per_s": round(len(poses) / min(secs), 1),
                                      "same_as_first_device_render": bool(np.array_equal(pinned.array.view(np.uint32),
                                                                                         ref.view(np.uint32)))}
        pinned.free()
    out["total_s"] = round(time.perf_counter() - t0, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
