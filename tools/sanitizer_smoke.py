"""Every kernel of the library once, on small inputs, for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py
"""
import sys

import numpy as np

sys.path.insert(0, '.')
import oracle as O
from birp_cpp_b200 import capi, synthetic as S

capi.set_device(0)
W, H = 27, 16
poses = np.array([S.DEFAULT_POS, (30., -25., 28.), (-14., 3., 5.), (1.5, -3., 2.)], np.float32)
fp = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, 27., 16., W, H)
fo = O.orc_frames(poses, S.DEFAULT_AIM_X, 27., 16., W, H)
zyx, org, sp = S.make_cube(41, 37, 45, negative_slab=False)
zyx0 = zyx.copy()
zyx0[10:20] = 0.0  # a non-negative cube (lerp_nonneg)
for grid in (zyx, zyx0):
    c = capi.Cube(grid, org, sp)
    for tri in (False, True):
        for fl in (0, capi.NO_LEAN_MARCH, capi.NO_OCTS, capi.WINDOW_PER_STEP, capi.NO_QUADS, capi.NO_QUADS | capi.NO_TEXTURE, capi.TEXTURE_ONLY,
                   capi.GENERAL_LERP, capi.DIV_IEEE):
            got = c.render(fp, W, H, trilinear=tri, extra_flags=fl)
            want = O.orc_render_cube(grid, org, sp, tri, fo, W, H)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (tri, fl)
    c.render(fp, W, H, nsamples=3000, trilinear=True)  # march longer than the staged ray table
    c.sample(np.random.rand(1000, 3) * 80 - 40, True)
    got = capi.render_sweep(c, poses, S.DEFAULT_AIM_X, 27., 16., W, H)
    c.free()
P = capi.cmem_init()
for mode in (False, True, "libm64", "nofilter", "notable", "libm32"):
    capi.render_cmem(P, fp, W, H, fast=mode)
    capi.sample_cmem(P, np.random.rand(1000, 3) * 20 - 10, fast=mode)
capi.selftest_cmem_math(1 << 12)
capi.selftest_cmem_filter(P, 1 << 12)
capi.selftest_cmem_table(P, 1 << 12)
capi.accumulate(np.random.rand(500, 200).astype(np.float32), 0.22)
capi.ray_directions(fp[:1], W, H)
try:
    import torch
    d_poses = torch.empty((64, 3), dtype=torch.float32, device="cuda")
    d_frames = torch.empty((64, 16), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    capi.orbit_poses_device(64, d_poses.data_ptr(), stream=s)
    capi.frame_setup_device(d_poses.data_ptr(), 64, S.DEFAULT_AIM_X, 27., 16., W, H, d_frames.data_ptr(), stream=s)
    capi.check_device_status(s)
except ImportError:
    pass
print("sanitizer script ok")
