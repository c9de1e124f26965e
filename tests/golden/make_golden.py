"""Regenerates tests/golden/reference_renders.npz from the UNMODIFIED reference
(oracle/_ref/libbirp_ref.so, built from /root/reference by oracle/Makefile).

Run from the repo root in a container that has /root/reference:
    python tests/golden/make_golden.py
Every array in the file is an output of the reference's own Camera / DataCube / CMEM
code on the synthetic inputs of birp_cpp_b200.synthetic; nothing from this repo's
renderer or restatement is stored.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle as O  # noqa: E402
from birp_cpp_b200 import synthetic as S  # noqa: E402

# camera poses: main.cpp default; one whose rays leave the cube; one outside the cube
POSES = np.array([S.DEFAULT_POS, (30.0, -25.0, 28.0), (-14.0, 3.0, 5.0), (2.0, 9.5, -7.25)], dtype=np.float32)


def main():
    O.build()
    out = {"poses": POSES, "aim_x": np.float32(S.DEFAULT_AIM_X)}
    zyx, org, sp = S.config1_cube()
    out["cube_sha256"] = np.frombuffer(hashlib.sha256(zyx.tobytes()).digest(), dtype=np.uint8)
    neg, _, _ = S.config1_cube(negative_slab=True)
    out["cube_neg_sha256"] = np.frombuffer(hashlib.sha256(neg.tobytes()).digest(), dtype=np.uint8)
    for name, data in (("cube", zyx), ("cubeneg", neg)):
        for tri in (0, 1):
            rc = O.RefCube(data, org, sp, bool(tri))
            imgs = [O.ref_render(rc, p, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W) for p in POSES]
            out[f"{name}_tri{tri}"] = np.stack(imgs)
            rc.free()
    # a non-default march (fine step) and a ragged image size through the private members
    rc = O.RefCube(zyx, org, sp, True)
    out["cube_tri1_n517_w37_h21"] = O.ref_render(rc, POSES[0], S.DEFAULT_AIM_X, 1.0, 21.0, 37.0, nsamples=517,
                                                 maxdist=44)
    # second Render() on the same camera accumulates (src/Camera.cpp:223)
    img = O.ref_render(rc, POSES[0], S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    out["cube_tri1_twice"] = O.ref_render(rc, POSES[0], S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W,
                                          image=img.copy())
    # sampler known answers on this cube (the reference's own fixture px_uni_1000.dat is not in its repo)
    rng = np.random.default_rng(7)
    pts = np.concatenate([
        rng.uniform([-12, -42, -42], [42, 42, 42], size=(4000, 3)),
        np.array([[-10.0, -40.0, -40.0], [39.5, 39.5, 39.5], [40.0, 0, 0], [39.499996, 0, 0], [9.0, 0.0, 0.0],
                  [4.0, 9.4, 6.0], [0.25, 0.5, 0.75], [-10.0, 0.0, 0.0], [-10.000001, 0, 0]]),
    ]).astype(np.float32)
    out["sample_points"] = pts
    rc.set_trilinear(False)
    out["samples_tri0"] = O.ref_sample(rc, pts)
    rc.set_trilinear(True)
    out["samples_tri1"] = O.ref_sample(rc, pts)
    rc.free()

    # CMEM: Init() defaults, the `birp -c` parameters, and the reference test's second set
    cm_pts = np.concatenate([
        rng.uniform(-30, 30, size=(4000, 3)),
        np.array([[8, 4, 4], [4, 12, 10], [3, 1.5, 1.2], [8, 10, 10], [8, -10, -10], [8, 10, -10], [12, 12, 10],
                  [12, 10, 10], [12, -10, -10], [12, 10, -10], [0, 5, 5], [5, 0, 5], [0, 0, 9], [7, 0, 0]]),
    ]).astype(np.float32)
    out["cmem_points"] = cm_pts
    for name, kw in (("default", {}), ("cli", O.CMEM_CLI), ("test2", O.CMEM_TEST2)):
        m = O.RefCmem(False, **kw) if kw else O.RefCmem()
        out[f"cmem_{name}_constants"] = m.constants()
        out[f"cmem_{name}_samples"] = O.ref_sample(m, cm_pts)
        out[f"cmem_{name}_images"] = np.stack(
            [O.ref_render(m, p, S.DEFAULT_AIM_X, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W) for p in POSES[:2]])
        m.free()
    # per-frame setup pins
    out["orient"] = np.array([O.ref_orient(p, S.DEFAULT_AIM_X) for p in POSES], dtype=np.float32)
    out["trace_points_pose0_w5_h3_n7"] = O.ref_trace_points(POSES[0], S.DEFAULT_AIM_X, 1.0, 3.0, 5.0, 5, 3, 7, 44)
    path = Path(__file__).with_name("reference_renders.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
