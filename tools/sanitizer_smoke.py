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
zyx0[10:20] = 0.0  # a non-negative cube (sign-marked oct records, lerp_nonneg)
zyxn = S.make_cube(41, 37, 45, negative_slab=True)[0]  # negative voxels (NaN-marked records, marked poly cells)
for grid in (zyx, zyx0, zyxn):
    c = capi.Cube(grid, org, sp)
    for tri in (False, True):
        for fl in (0, capi.NO_LEAN_MARCH, capi.NO_OCTS, capi.WINDOW_PER_STEP, capi.NO_QUADS, capi.NO_QUADS | capi.NO_TEXTURE, capi.TEXTURE_ONLY,
                   capi.GENERAL_LERP, capi.DIV_IEEE, capi.DIV_ONE_STEP, capi.DIV_TWO_STEP):
            got = c.render(fp, W, H, trilinear=tri, extra_flags=fl)
            want = O.orc_render_cube(grid, org, sp, tri, fo, W, H)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (tri, fl)
    # the tolerance-mode march: every index-rounding variant, computed and table-read step distances, the warp-per-ray
    # experiment, lanes in 2x2 blocks
    for v in (0, 1, 2, 3, 4):
        for ks in (1, 0):
            capi.set_tuning("fast_variant", v)
            capi.set_tuning("fast_kstep", ks)
            c.render(fp, W, H, trilinear=True, extra_flags=capi.FAST_MARCH)
    capi.set_tuning("fast_variant", -1)
    capi.set_tuning("fast_kstep", 1)
    for warps in (2, 4):  # CTA sizes (8 is the default for frames of this size)
        capi.set_tuning("cta_warps", warps)
        c.render(fp, W, H, trilinear=True, extra_flags=capi.FAST_MARCH)
        c.render(fp, W, H, trilinear=True)
        c.render(fp, W, H, trilinear=False)
    capi.set_tuning("cta_warps", 0)
    for knob in ("fast_coop", "tile_swizzle"):
        capi.set_tuning(knob, 1)
        c.render(fp, W, H, trilinear=True, extra_flags=capi.FAST_MARCH)
        c.render(fp, W, H, trilinear=True)
        capi.set_tuning(knob, 0)
    c.render(fp, W, H, nsamples=3000, trilinear=True, extra_flags=capi.FAST_MARCH)
    for ph in range(3):  # interleaved row bands, strided sweep output
        capi.render_cube_rows(c, fp, W, H, 8, ph, 3)
    both = np.zeros((len(poses), H, W), np.float32)
    for r in range(2):
        capi.render_sweep(c, poses[r::2], S.DEFAULT_AIM_X, 27., 16., W, H, out=both, frame_offset=r, frame_stride=2)
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
