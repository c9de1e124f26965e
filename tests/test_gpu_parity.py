"""Parity of the CUDA path (through the C ABI) with the oracle, on a B200.

DataCube path: bit-exact (index selection, termination and accumulation order are all
reproduced).  CMEM path: SURVEY.md 8(d) tolerance -- 1e-5 per pixel, 1e-4 integrated flux,
at least 95 % of pixels (device transcendentals are not glibc's)."""
import numpy as np
import pytest

import oracle as O
from birp_cpp_b200 import capi, synthetic as S
from conftest import assert_bit_exact, assert_image_parity, pixel_fraction_within

pytestmark = pytest.mark.gpu

W, H = S.SXI_W, S.SXI_H
AIM = S.DEFAULT_AIM_X
POSES = np.array([S.DEFAULT_POS, (30.0, -25.0, 28.0), (-14.0, 3.0, 5.0), (2.0, 9.5, -7.25), (1.5, -3.0, 2.0),
                  (38.0, 38.0, 38.0), (-9.9, -39.9, 39.4)], np.float32)


def frames_pair(poses, fov_x, fov_y, w, h):
    """product frames (structured) and oracle frames (float rows); must agree bitwise"""
    fp = capi.frames_for_poses(poses, AIM, fov_x, fov_y, w, h)
    fo = O.orc_frames(poses, AIM, fov_x, fov_y, w, h)
    assert_bit_exact(fp.view(np.float32).reshape(-1, 16), fo)
    return fp, fo


@pytest.fixture(scope="module")
def dev_cube1(gpu, cube1):
    c = capi.Cube(*cube1)
    yield c
    c.free()


@pytest.fixture(scope="module")
def dev_cube1_neg(gpu, cube1_neg):
    c = capi.Cube(*cube1_neg)
    yield c
    c.free()


# ------------------------------------------------------------------ DataCube sampler
@pytest.mark.parametrize("tri", [False, True])
def test_sampler_bit_exact(dev_cube1, cube1, tri):
    zyx, org, sp = cube1
    rng = np.random.default_rng(5)
    pts = rng.uniform([-12, -42, -42], [42, 42, 42], size=(200000, 3)).astype(np.float32)
    nodes = org + sp * rng.integers(0, 100, size=(5000, 3)).astype(np.float32)
    edge = np.array([[-10.0, -40.0, -40.0], [39.5, 39.5, 39.5], [40.0, 0, 0], [39.499996, 0, 0], [-10.000001, 0, 0],
                     [np.inf, 0, 0], [-np.inf, 0, 0], [0, 1e30, 0]], np.float32)
    pts = np.concatenate([pts, nodes, edge])
    assert_bit_exact(dev_cube1.sample(pts, tri), O.orc_cube_sample(zyx, org, sp, tri, pts), f"tri={tri}")


def test_sampler_golden(dev_cube1, golden, cube1_matches_golden):
    if not cube1_matches_golden:
        pytest.skip("synthetic cube differs by an ulp on this CPU")
    for tri in (0, 1):
        assert_bit_exact(dev_cube1.sample(golden["sample_points"], bool(tri)), golden[f"samples_tri{tri}"])


def test_sampler_empty_and_nan(dev_cube1):
    assert dev_cube1.sample(np.zeros((0, 3), np.float32), True).shape == (0,)
    with pytest.raises(capi.BirpError) as e:  # the reference throws std::out_of_range here
        dev_cube1.sample([[np.nan, 0.0, 0.0]], True)
    assert e.value.code == capi.BIRP_ERR_RANGE
    # NaN in one axis but another axis out of bounds: the reference returns -1 without throwing
    assert dev_cube1.sample([[np.nan, 100.0, 0.0]], True)[0] == -1.0
    assert dev_cube1.sample([[0.0, 0.0, 0.0]], True)[0] > 0  # status word was cleared


# ------------------------------------------------------------------ DataCube renders
@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("neg", [False, True])
def test_render_config1_bit_exact(dev_cube1, dev_cube1_neg, cube1, cube1_neg, tri, neg):
    zyx, org, sp = cube1_neg if neg else cube1
    dev = dev_cube1_neg if neg else dev_cube1
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    got = dev.render(fp, W, H, trilinear=tri)
    want = O.orc_render_cube(zyx, org, sp, tri, fo, W, H)
    assert_bit_exact(got, want, f"tri={tri} neg={neg}")
    assert want[0].min() > 0 and want[2].max() == 0  # inside / outside the cube


# corner fetch (default: one 32-byte oct load, end of the in-bounds run located once per ray; two quad loads, same
# march; two quad loads with the window test in every step; tld4 + 4 loads; two tld4; 8 loads) x division x lerp flavour
# (flags 0 = the lean span march on a positive cube of at most 2^24 voxels, NO_LEAN_MARCH = the general span march)
_FETCH = [0, capi.NO_LEAN_MARCH, capi.NO_OCTS, capi.WINDOW_PER_STEP, capi.NO_QUADS, capi.TEXTURE_ONLY,
          capi.NO_QUADS | capi.NO_TEXTURE]
_DIV = [0, capi.DIV_ONE_STEP, capi.DIV_TWO_STEP, capi.DIV_IEEE]  # (cube1's spacing is 0.5: default = the multiply alone)
_LERP = [0, capi.GENERAL_LERP]


@pytest.mark.parametrize("tri", [False, True])
@pytest.mark.parametrize("neg", [False, True])
@pytest.mark.parametrize("extra", [f | d | l for f in _FETCH for d in _DIV for l in _LERP])
def test_every_kernel_variant_bit_exact(dev_cube1, dev_cube1_neg, cube1, cube1_neg, tri, neg, extra):
    """all template variants of the DataCube march kernel (corner fetch x division x lerp flavour) agree with
    the oracle"""
    zyx, org, sp = cube1_neg if neg else cube1
    dev = dev_cube1_neg if neg else dev_cube1
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    got = dev.render(fp, W, H, trilinear=tri, extra_flags=extra)
    assert_bit_exact(got, O.orc_render_cube(zyx, org, sp, tri, fo, W, H), f"flags={extra}")


def test_upload_analysis(dev_cube1, dev_cube1_neg, gpu):
    a, b = dev_cube1.info(), dev_cube1_neg.info()
    assert a["positive"] and not b["positive"]
    assert a["texture"] and b["texture"], "layered-texture copy missing or failed its gather self-check"
    assert a["quads"] and b["quads"], "quad copy missing or failed its self-check"
    assert a["octs"] and b["octs"], "oct copy missing or failed its self-check"
    assert a["divmode"] == 3 == b["divmode"], "spacing 0.5: the multiply by the exact reciprocal is the division"
    print("division variant proven exact for spacing 0.5:", a["divmode"])
    # awkward spacings: whatever variant is chosen, the renders must stay bit-exact
    rng = np.random.default_rng(9)
    for sp in (0.3137255, 1.0 / 3.0, 0.1, 0.7, 1.9999999):
        zyx = rng.uniform(1e-7, 1e-5, size=(24, 20, 28)).astype(np.float32)
        zyx[3, 4, 5] = 0.0  # one zero voxel: not "positive" any more
        org = np.array([-3.3, -2.9, -3.1], np.float32)
        spacing = np.array([sp, sp * 1.25, sp * 0.8], np.float32)
        c = capi.Cube(zyx, org, spacing)
        try:
            assert not c.info()["positive"]
            fp, fo = frames_pair(np.array([(0.5, 0.25, 1.0), (1.0, -1.0, 0.5)], np.float32), 36.0, 36.0, 48, 48)
            for tri in (False, True):
                got = c.render(fp, 48, 48, nsamples=300, maxdist=12, trilinear=tri)
                want = O.orc_render_cube(zyx, org, spacing, tri, fo, 48, 48, nsamples=300, maxdist=12)
                assert_bit_exact(got, want, f"spacing {sp} tri={tri} divmode={c.info()['divmode']}")
        finally:
            c.free()


def test_rays_in_a_grid_plane_take_the_cold_path(gpu):
    """hand-made frames whose rays run exactly inside the first grid plane (ingress 0) or within
    1e-35 of it: below the 2^-100 window of the fast division, so the generic sampler must take
    over -- and still match the oracle bit for bit"""
    rng = np.random.default_rng(12)
    zyx = rng.uniform(1e-7, 1e-5, size=(40, 30, 20)).astype(np.float32)
    org = np.array([1.0, -2.0, 0.0], np.float32)
    sp = np.array([0.3137255, 0.25, 0.5], np.float32)
    c = capi.Cube(zyx, org, sp)
    try:
        frames = np.zeros(3, capi.FRAME_DTYPE)
        for n, px in enumerate((1.0, np.float32(1.0) + np.float32(1e-35), 1.0000001)):
            frames[n]["pos"] = (px, -1.0, 0.125)
            frames[n]["rinv"] = (1, 0, 0, 0, 1, 0, 0, 0, -1)  # camera -z maps to world +z
            frames[n]["tan_half_x"] = 0.0  # every ray has direction x == 0 exactly
            frames[n]["tan_half_y"] = 0.2
            frames[n]["aspect"] = 1.0
        fo = frames.view(np.float32).reshape(-1, 16)
        for tri in (False, True):
            got = c.render(frames, 8, 8, nsamples=64, maxdist=30, trilinear=tri)
            want = O.orc_render_cube(zyx, org, sp, tri, fo, 8, 8, nsamples=64, maxdist=30)
            assert_bit_exact(got, want, f"tri={tri}")
            assert want.min() > 0
    finally:
        c.free()


def test_first_cells_and_near_one_weights(gpu):
    """cameras inside and next to the first cell of every axis, looking both ways: those samples are outside
    the hot window (index < 1, where the interpolation weight can exceed 1 - 2^-23 and std::lerp's clamp
    is live) and must come out of the literal sampler bit for bit; rays cross the window edge repeatedly"""
    rng = np.random.default_rng(33)
    zyx = rng.uniform(1e-7, 1e-5, size=(24, 26, 28)).astype(np.float32)
    org = np.array([-2.0, 1.0, 0.5], np.float32)
    sp = np.array([0.3137255, 0.25, 0.4], np.float32)
    poses = []
    for fx, fy, fz in ((0.5, 0.5, 0.5), (0.999999, 3.2, 4.1), (5.3, 0.99999994, 2.2), (6.1, 4.4, 0.9999999),
                       (1.0000001, 1.0000001, 1.0000001), (0.25, 12.0, 0.75), (20.0, 0.3, 11.0)):
        poses.append((org[0] + fx * sp[0], org[1] + fy * sp[1], org[2] + fz * sp[2]))
    poses = np.array(poses, np.float32)
    c = capi.Cube(zyx, org, sp)
    try:
        assert c.info()["positive"]
        for aim in (3.0, -30.0):
            fp = capi.frames_for_poses(poses, aim, 60.0, 60.0, 40, 40)
            fo = O.orc_frames(poses, aim, 60.0, 60.0, 40, 40)
            for fl in _FETCH:
                got = c.render(fp, 40, 40, nsamples=600, maxdist=12, trilinear=True, extra_flags=fl)
                want = O.orc_render_cube(zyx, org, sp, True, fo, 40, 40, nsamples=600, maxdist=12)
                assert_bit_exact(got, want, f"aim={aim} flags={fl}")
            assert want.max() > 0
    finally:
        c.free()


def test_cube_with_zero_regions_stays_bit_exact(gpu):
    """emissivity grids of real runs are zero inside the magnetosphere: zero samples must stay on the hot path
    (their scaled value is exactly zero) and the std::lerp sign branch fires at every zero corner"""
    zyx, org, sp = S.make_cube(64, 64, 64)
    zyx = zyx.copy()
    x = org[0] + sp[0] * np.arange(64)
    r = np.sqrt(x[None, None, :] ** 2 + x[None, :, None] ** 2 + x[:, None, None] ** 2)
    zyx[r < 12.0] = 0.0
    zyx[np.random.default_rng(8).random(zyx.shape) < 0.08] = 0.0  # and isolated zero voxels: every corner pattern
    c = capi.Cube(zyx, org, sp)
    try:
        assert not c.info()["positive"] and c.info()["nonnegative"]
        poses = np.array([S.DEFAULT_POS, (20.0, -15.0, 3.0), (0.5, 0.25, 1.0)], np.float32)
        fp, fo = frames_pair(poses, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
        for tri in (False, True):
            want = O.orc_render_cube(zyx, org, sp, tri, fo, W, H)
            for fl in (_FETCH + [capi.GENERAL_LERP] if tri else [0]):
                assert_bit_exact(c.render(fp, W, H, trilinear=tri, extra_flags=fl), want, f"tri={tri} flags={fl}")
            assert want.max() > 0
    finally:
        c.free()
    # a slab of NEGATIVE zeros takes the cube out of the non-negative class: full std::lerp, same parity
    zyx[40:44, :, :] = -0.0
    c = capi.Cube(zyx, org, sp)
    try:
        assert not c.info()["nonnegative"]
        for tri in (False, True):
            want = O.orc_render_cube(zyx, org, sp, tri, fo, W, H)
            assert_bit_exact(c.render(fp, W, H, trilinear=tri), want, f"negative zeros, tri={tri}")
    finally:
        c.free()


def test_oct_record_fetch_on_a_small_grid(gpu):
    """the default corner fetch (one 32-byte load from the oct copy, every record verified against the linear grid at
    upload) on a small non-negative grid seen from inside: cameras in the first and in the last cells of every
    axis, where the records' clamped neighbours sit; the quad and the 4-byte forms must agree with it"""
    rng = np.random.default_rng(5)
    zyx = rng.uniform(1e-7, 1e-5, size=(24, 26, 28)).astype(np.float32)
    zyx[rng.random(zyx.shape) < 0.1] = 0.0
    org = np.array([-2.0, 1.0, 0.5], np.float32)
    sp = np.array([0.3137255, 0.25, 0.4], np.float32)
    poses = np.array([(org[0] + fx * sp[0], org[1] + fy * sp[1], org[2] + fz * sp[2])
                      for fx, fy, fz in ((0.5, 0.5, 0.5), (26.5, 24.5, 22.5), (13.0, 0.99999994, 11.0), (5.3, 12.2, 22.9))],
                     np.float32)
    c = capi.Cube(zyx, org, sp)
    try:
        assert c.info()["octs"] and c.info()["quads"] and c.info()["nonnegative"] and not c.info()["positive"]
        for aim in (3.0, -30.0):
            fq = capi.frames_for_poses(poses, aim, 60.0, 60.0, 40, 40)
            fr = O.orc_frames(poses, aim, 60.0, 60.0, 40, 40)
            want = O.orc_render_cube(zyx, org, sp, True, fr, 40, 40, nsamples=600, maxdist=12)
            for fl in _FETCH + [capi.GENERAL_LERP]:
                got = c.render(fq, 40, 40, nsamples=600, maxdist=12, trilinear=True, extra_flags=fl)
                assert_bit_exact(got, want, f"small grid, aim={aim} flags={fl}")
            assert want.max() > 0
    finally:
        c.free()


def test_cube_beyond_texture_limits_uses_global_loads(gpu):
    """more z slices than a layered texture can hold (2048): no texture copy, the global-load kernel runs"""
    rng = np.random.default_rng(21)
    zyx = rng.uniform(1e-7, 1e-5, size=(2100, 6, 7)).astype(np.float32)
    org = np.array([-1.0, -1.0, -10.0], np.float32)
    sp = np.array([0.5, 0.5, 0.01], np.float32)
    c = capi.Cube(zyx, org, sp)
    try:
        assert not c.info()["texture"] and c.info()["positive"] and c.info()["quads"] and c.info()["octs"]
        fp, fo = frames_pair(np.array([(0.2, 0.3, 5.0), (0.5, 0.1, -3.0)], np.float32), 36.0, 36.0, 24, 24)
        want = O.orc_render_cube(zyx, org, sp, True, fo, 24, 24, nsamples=400, maxdist=30)
        assert want.max() > 0
        for fl in (0, capi.NO_OCTS, capi.NO_QUADS):  # one oct load / two quad loads / eight global loads
            got = c.render(fp, 24, 24, nsamples=400, maxdist=30, trilinear=True, extra_flags=fl)
            assert_bit_exact(got, want, f"flags={fl}")
    finally:
        c.free()


@pytest.fixture(scope="module")
def orbit_sweep(gpu):
    """the headline workload and its CPU restatement (all host cores, ~20-60 s), shared by the exact and the
    tolerance-mode test"""
    import os
    zyx, org, sp = S.make_cube(256, 256, 256)
    poses = S.orbit_poses(10000)
    fp = capi.frames_for_poses(poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    fo = fp.view(np.float32).reshape(-1, 16)
    want, valid = O.orc_render_cube(zyx, org, sp, True, fo, W, H, nthreads=len(os.sched_getaffinity(0)),
                                    return_valid=True)
    return zyx, org, sp, fp, want, valid


def test_whole_orbit_sweep_bit_exact(orbit_sweep):
    """the headline workload itself, all of it: 10 000 SXI frames (69 M pixels, 1.4e10 samples) through a
    256^3 cube, every pixel compared bit for bit with the CPU restatement; default march (lean) and the general one"""
    zyx, org, sp, fp, want, valid = orbit_sweep
    c = capi.Cube(zyx, org, sp)
    try:
        got = c.render(fp, W, H, trilinear=True)
        general = c.render(fp[::10], W, H, trilinear=True, extra_flags=capi.NO_LEAN_MARCH)
    finally:
        c.free()
    assert valid > 1.3e10
    assert_bit_exact(got, want, "orbit sweep")
    assert_bit_exact(general, want[::10], "orbit sweep, general march, every 10th frame")
    assert (want.reshape(10000, -1).max(axis=1) > 0).mean() > 0.99  # the orbit stays inside the cube


def test_bad_spacing_rejected(gpu):
    zyx = np.ones((4, 4, 4), np.float32)
    for sp in ((0.0, 1, 1), (1, -1.0, 1), (1, 1, np.nan), (np.inf, 1, 1)):
        with pytest.raises(capi.BirpError) as e:
            capi.Cube(zyx, (0, 0, 0), sp)
        assert e.value.code == capi.BIRP_ERR_ARG


@pytest.mark.parametrize("name", ["cube", "cubeneg"])
@pytest.mark.parametrize("tri", [0, 1])
def test_render_golden_reference_images(dev_cube1, dev_cube1_neg, golden, cube1_matches_golden, name, tri):
    dev = dev_cube1 if name == "cube" else dev_cube1_neg
    fp = capi.frames_for_poses(golden["poses"], float(golden["aim_x"]), S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    got = dev.render(fp, W, H, trilinear=bool(tri))
    want = golden[f"{name}_tri{tri}"]
    if cube1_matches_golden:
        assert_bit_exact(got, want, "vs images rendered by the unmodified reference")
    else:
        for g, w in zip(got, want):
            assert_image_parity(g, w)


@pytest.mark.parametrize("w,h,n,maxdist", [(37, 21, 517, 44), (1, 1, 200, 44), (8, 4, 1, 44), (9, 5, 2, 10),
                                             (131, 3, 64, 60), (3, 131, 33, 44)])
def test_render_ragged_shapes(dev_cube1, cube1, w, h, n, maxdist):
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES[:3], 37.0, 21.0, w, h)
    for tri in (False, True):
        got = dev_cube1.render(fp, w, h, nsamples=n, maxdist=maxdist, trilinear=tri)
        want = O.orc_render_cube(zyx, org, sp, tri, fo, w, h, nsamples=n, maxdist=maxdist)
        assert_bit_exact(got, want, f"{w}x{h} n={n}")


def test_render_march_longer_than_the_staged_ray_table(dev_cube1, cube1):
    """more samples than one shared-memory chunk of the ray table (2048): 5000-sample march"""
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES[:2], 9.0, 7.0, 18, 14)
    for tri in (False, True):
        got = dev_cube1.render(fp, 18, 14, nsamples=5000, maxdist=44, trilinear=tri)
        want = O.orc_render_cube(zyx, org, sp, tri, fo, 18, 14, nsamples=5000, maxdist=44)
        assert_bit_exact(got, want, f"tri={tri}")


def test_render_golden_nondefault_march(dev_cube1, golden, cube1_matches_golden):
    if not cube1_matches_golden:
        pytest.skip("synthetic cube differs by an ulp on this CPU")
    fp = capi.frames_for_poses(golden["poses"][:1], float(golden["aim_x"]), 37.0, 21.0, 37, 21)
    assert_bit_exact(dev_cube1.render(fp, 37, 21, nsamples=517)[0], golden["cube_tri1_n517_w37_h21"])


def test_render_twice_accumulates_like_reference(dev_cube1, golden, cube1_matches_golden, cube1):
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES[:1], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    once = dev_cube1.render(fp, W, H)
    twice = dev_cube1.render(fp, W, H, images=once.copy())
    want = O.orc_render_cube(zyx, org, sp, True, fo, W, H, images=O.orc_render_cube(zyx, org, sp, True, fo, W, H))
    assert_bit_exact(twice, want)
    if cube1_matches_golden:
        assert_bit_exact(twice[0], golden["cube_tri1_twice"])


def test_render_row_bands_tile_the_image(dev_cube1):
    """row bands are the unit of multi-GPU image sharding: they tile the full image exactly"""
    fp, _ = frames_pair(POSES[:2], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    full = dev_cube1.render(fp, W, H)
    tiled = np.zeros_like(full)
    for r0, r1 in ((0, 9), (9, 10), (10, 41), (41, 64)):
        band = dev_cube1.render(fp, W, H, rows=(r0, r1))
        assert band[:, :r0].max(initial=0) == 0 and band[:, r1:].max(initial=0) == 0
        tiled[:, r0:r1] = band[:, r0:r1]
    assert_bit_exact(tiled, full)
    # with BIRP_ACCUMULATE only the band's rows are read and written back
    canvas = np.full_like(full, 3.0)
    dev_cube1.render(fp, W, H, rows=(9, 10), images=canvas)
    assert (canvas[:, :9] == 3.0).all() and (canvas[:, 10:] == 3.0).all() and (canvas[:, 9] > 3.0).all()


@pytest.mark.parametrize("h,band,nph", [(64, 8, 4), (70, 16, 3), (64, 32, 8), (24, 8, 5)])
def test_render_interleaved_row_bands_tile_the_image(dev_cube1, h, band, nph):
    """interleaved row bands (one large image over several GPUs, birp_render_cube_rows): the phases tile the image
    exactly, a phase touches only its own rows, also when the last band is short or a phase owns no band at all"""
    fp, _ = frames_pair(POSES[:2], S.SXI_FOV_W, float(h) / 4, W, h)
    full = dev_cube1.render(fp, W, h)
    tiled = np.full_like(full, -7.0)
    for ph in range(nph):
        one = np.full_like(full, -7.0)
        capi.render_cube_rows(dev_cube1, fp, W, h, band, ph, nph, out=one)
        own = ((np.arange(h) // band) % nph) == ph
        assert (one[:, ~own] == -7.0).all(), "a phase wrote rows of another phase"
        assert_bit_exact(one[:, own], full[:, own], f"phase {ph}")
        capi.render_cube_rows(dev_cube1, fp, W, h, band, ph, nph, out=tiled)
    assert_bit_exact(tiled, full, "all phases")


def test_sweep_with_strided_output_interleaves_into_one_buffer(dev_cube1):
    """birp_render_cube_sweep_strided: two calls over the even and the odd frames of a sweep (what two processes
    would do) fill one buffer that equals the single sweep"""
    poses = S.orbit_poses(601)
    whole = capi.render_sweep(dev_cube1, poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    both = np.full_like(whole, -1.0)
    for r in range(2):
        capi.render_sweep(dev_cube1, poses[r::2], AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, out=both, frame_offset=r,
                          frame_stride=2)
    assert_bit_exact(both, whole)
    P = capi.cmem_init()
    whole = capi.render_sweep(P, poses[:100], AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    both = np.full_like(whole, -1.0)
    for r in range(4):
        capi.render_sweep(P, poses[:100][r::4], AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, out=both, frame_offset=r,
                          frame_stride=4)
    assert_bit_exact(both, whole, "cmem, stride 4")


def test_render_many_frames_chunked_pipeline(dev_cube1, cube1):
    """more frames than one staging chunk of a larger image, odd count"""
    zyx, org, sp = cube1
    poses = S.orbit_poses(37)
    fp, fo = frames_pair(poses, 36.0, 36.0, 512, 512)  # 1 MiB frames -> 16 per chunk -> 3 chunks
    got = dev_cube1.render(fp, 512, 512, nsamples=40, trilinear=True)
    want = O.orc_render_cube(zyx, org, sp, True, fo, 512, 512, nsamples=40)
    assert_bit_exact(got, want)


def test_render_nan_pose_reports_range_error(dev_cube1):
    fr = capi.frame_setup((10.0, 0.0, 0.0), AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    assert np.isnan(fr["rinv"]).any()
    with pytest.raises(capi.BirpError) as e:
        dev_cube1.render(fr, W, H)
    assert e.value.code == capi.BIRP_ERR_RANGE
    ok = dev_cube1.render(capi.frame_setup(S.DEFAULT_POS, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H), W, H)
    assert ok.min() > 0


def test_render_zero_frames(dev_cube1):
    out = dev_cube1.render(np.zeros(0, capi.FRAME_DTYPE), W, H)
    assert out.shape == (0, H, W)


def test_render_linearity_in_the_cube(gpu, cube1):
    """scaling the emissivity by a power of two scales every pixel exactly"""
    zyx, org, sp = cube1
    fp, _ = frames_pair(POSES[:2], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    a = capi.Cube(zyx, org, sp)
    b = capi.Cube(zyx * np.float32(4.0), org, sp)
    try:
        assert_bit_exact(b.render(fp, W, H), a.render(fp, W, H) * np.float32(4.0))
    finally:
        a.free()
        b.free()


def test_render_constant_cube_counts_samples(gpu):
    """a constant field integrates to (live samples) x constant: checks the termination prefix"""
    zyx = np.full((33, 33, 33), 2.0 ** -10, np.float32)
    org, sp = np.array([-8, -8, -8], np.float32), np.array([0.5, 0.5, 0.5], np.float32)
    c = capi.Cube(zyx, org, sp)
    try:
        fp, fo = frames_pair(np.array([(1.0, 2.0, 3.0)], np.float32), 36.0, 36.0, 64, 64)
        for tri in (False, True):
            got = c.render(fp, 64, 64, trilinear=tri)
            want, valid = O.orc_render_cube(zyx, org, sp, tri, fo, 64, 64, return_valid=True)
            assert_bit_exact(got, want)
            assert 0 < valid < 64 * 64 * 200  # every ray leaves this small cube
    finally:
        c.free()


def test_sweep_call_equals_frames_call(gpu, dev_cube1):
    """birp_render_*_sweep (poses in, chunked: host setup of one chunk while the previous one renders) returns
    the frame records and the images of the two-step path, bit for bit; 1000 frames make four chunks"""
    poses = S.orbit_poses(1000)
    fp = capi.frames_for_poses(poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    for tri in (True, False):
        fr = np.zeros(len(poses), capi.FRAME_DTYPE)
        got = capi.render_sweep(dev_cube1, poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, trilinear=tri, frames_out=fr)
        assert_bit_exact(fr.view(np.float32), fp.view(np.float32), "frame records")
        assert_bit_exact(got, dev_cube1.render(fp, W, H, trilinear=tri), f"tri={tri}")
    P = capi.cmem_init()
    got = capi.render_sweep(P, poses[:300], AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, rows=(5, 40))
    assert_bit_exact(got, capi.render_cmem(P, fp[:300], W, H, rows=(5, 40)), "cmem sweep, row band, no frames_out")
    assert capi.render_sweep(dev_cube1, np.zeros((0, 3), np.float32), AIM, 27.0, 16.0, W, H).shape == (0, H, W)



# ------------------------------------------------------------------ tolerance-mode march (BIRP_FAST_MARCH)
FAST = capi.FAST_MARCH
FAST_MIN_FRACTION = 0.999  # of the pixels within 1e-5 (SURVEY 8(d) asks for 95 %); flux within 1e-4


def assert_fast_parity(got, want, what):
    """SURVEY 8(d) tolerance at 99.9 %, and the rays end where the reference ends them: a pixel that is exactly zero
    in the reference (no live sample) is exactly zero here and vice versa"""
    frac = assert_image_parity(got, want, what, min_fraction=FAST_MIN_FRACTION)
    assert np.array_equal(np.asarray(got) == 0, np.asarray(want) == 0), f"{what}: zero pixels differ"
    return frac


@pytest.fixture
def tuning(gpu):
    """birp_set_tuning with the defaults restored afterwards"""
    yield capi.set_tuning
    for k, v in (("fast_variant", -1), ("fast_kstep", 1), ("tile_swizzle", 0), ("tile_w_log2", -1), ("fast_coop", 0),
                 ("cta_warps", 0)):
        capi.set_tuning(k, v)


@pytest.mark.parametrize("warps", [2, 4, 8])
def test_cta_size_does_not_change_a_bit(dev_cube1, cube1, tuning, warps):
    """the launch picks 8 warps per CTA for many small frames and 4 for large images (pick_cta_warps): every size
    renders the same bits -- trilinear, nearest, tolerance mode, interleaved row bands, and CMEM"""
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    tuning("cta_warps", 0)
    auto = {tri: dev_cube1.render(fp, W, H, trilinear=tri) for tri in (True, False)}
    auto_fast = dev_cube1.render(fp, W, H, trilinear=True, extra_flags=FAST)
    P = capi.cmem_init()
    auto_cmem = capi.render_cmem(P, fp, W, H)
    tuning("cta_warps", warps)
    for tri in (True, False):
        got = dev_cube1.render(fp, W, H, trilinear=tri)
        assert_bit_exact(got, O.orc_render_cube(zyx, org, sp, tri, fo, W, H))
        assert_bit_exact(got, auto[tri])
    assert_bit_exact(dev_cube1.render(fp, W, H, trilinear=True, extra_flags=FAST), auto_fast)
    assert_bit_exact(capi.render_cmem(P, fp, W, H), auto_cmem)
    bands = np.full_like(auto[True], -7.0)
    for ph in range(3):
        capi.render_cube_rows(dev_cube1, fp, W, H, 8, ph, 3, out=bands)
    assert_bit_exact(bands, auto[True])


def test_large_image_takes_the_small_cta_and_stays_bit_exact(gpu, tuning):
    """a 1024 x 576 image is above the frame size where the launch switches to 4 warps per CTA: compared with the oracle
    and with the forced 8-warp launch"""
    zyx, org, sp = S.make_cube(48, 40, 44)
    poses = np.array([S.DEFAULT_POS], np.float32)
    Wl, Hl = 1024, 576
    fp, fo = frames_pair(poses, 36.0, 20.25, Wl, Hl)
    dev = capi.Cube(zyx, org, sp)
    try:
        got = dev.render(fp, Wl, Hl, trilinear=True)
        assert_bit_exact(got, O.orc_render_cube(zyx, org, sp, True, fo, Wl, Hl))
        tuning("cta_warps", 8)
        assert_bit_exact(dev.render(fp, Wl, Hl, trilinear=True), got)
    finally:
        dev.free()


@pytest.mark.parametrize("variant", [-1, 0, 1, 2, 3, 4])
@pytest.mark.parametrize("kstep", [1, 0])
@pytest.mark.parametrize("neg", [False, True])
def test_fast_march_config1_every_variant(dev_cube1, dev_cube1_neg, cube1, cube1_neg, tuning, variant, kstep, neg):
    """tolerance mode on the reference test cube's shape, every rounding variant of the cell index, with computed
    and with table-read step distances; the cube with a negative slab has marked cells (rays end at the reference's
    step)"""
    zyx, org, sp = cube1_neg if neg else cube1
    dev = dev_cube1_neg if neg else dev_cube1
    tuning("fast_variant", variant)
    tuning("fast_kstep", kstep)
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    got = dev.render(fp, W, H, trilinear=True, extra_flags=FAST)
    want = O.orc_render_cube(zyx, org, sp, True, fo, W, H)
    assert_fast_parity(got, want, f"variant {variant} kstep {kstep} neg {neg}")
    assert not np.array_equal(got, want), "the tolerance-mode kernel did not run (images are bit-identical)"


def test_fast_march_flag_leaves_nearest_neighbour_exact(dev_cube1, cube1):
    """nearest-neighbour sampling has no tolerance mode (a sample next to a cell boundary would pick another voxel):
    the flag is ignored, the render stays bit-exact"""
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    assert_bit_exact(dev_cube1.render(fp, W, H, trilinear=False, extra_flags=FAST),
                     O.orc_render_cube(zyx, org, sp, False, fo, W, H))


def test_fast_march_live_samples_equal_the_reference(dev_cube1_neg, cube1_neg):
    """termination is decided by the reference's arithmetic: the number of accumulated samples is the oracle's"""
    torch = pytest.importorskip("torch")
    zyx, org, sp = cube1_neg
    fp, fo = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    _, valid = O.orc_render_cube(zyx, org, sp, True, fo, W, H, return_valid=True)
    d_frames = torch.from_numpy(fp.view(np.float32).reshape(-1, 16).copy()).cuda()
    d_images = torch.zeros((len(fp), H, W), device="cuda")
    d_valid = torch.zeros(1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    capi.render_cube_device(dev_cube1_neg, d_frames.data_ptr(), len(fp), W, H, d_images.data_ptr(),
                            d_valid=d_valid.data_ptr(), stream=stream, extra_flags=FAST)
    capi.check_device_status(stream)
    assert int(d_valid.item()) == valid


def test_fast_march_zero_regions_negatives_and_nonfinite(gpu):
    """a non-negative cube with zero regions (no marked cells), then 30 % negative voxels (marked cells everywhere),
    then NaN / inf voxels: tolerance, termination, and NaN / inf pixels exactly where the reference has them"""
    zyx, org, sp = S.make_cube(64, 64, 64)
    zyx = zyx.copy()
    x = org[0] + sp[0] * np.arange(64)
    r = np.sqrt(x[None, None, :] ** 2 + x[None, :, None] ** 2 + x[:, None, None] ** 2)
    zyx[r < 12.0] = 0.0
    rng = np.random.default_rng(8)
    zyx[rng.random(zyx.shape) < 0.08] = 0.0
    poses = np.array([S.DEFAULT_POS, (20.0, -15.0, 3.0), (0.5, 0.25, 1.0)], np.float32)
    fp, fo = frames_pair(poses, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    neg = zyx.copy()
    neg[rng.random(zyx.shape) < 0.3] *= -1.0
    bad = zyx.copy()
    bad[10, 20, 30] = np.nan
    bad[40, 30, 20] = np.inf
    for name, grid in (("zeros", zyx), ("negatives", neg), ("nan/inf", bad)):
        c = capi.Cube(grid, org, sp)
        try:
            got = c.render(fp, W, H, trilinear=True, extra_flags=FAST)
            want = O.orc_render_cube(grid, org, sp, True, fo, W, H)
            assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want)), name
            fin = np.isfinite(want)
            assert pixel_fraction_within(got[fin], want[fin]) >= FAST_MIN_FRACTION, name
            assert np.array_equal(got[fin] == 0, want[fin] == 0), name
            assert want[fin].max() > 0
        finally:
            c.free()


@pytest.mark.parametrize("w,h,n,maxdist", [(37, 21, 517, 44), (1, 1, 200, 44), (8, 4, 1, 44), (9, 5, 2, 10),
                                             (131, 3, 64, 60), (18, 14, 5000, 44)])
def test_fast_march_ragged_shapes_and_long_marches(dev_cube1, cube1, tuning, w, h, n, maxdist):
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES[:3], 37.0, 21.0, w, h)
    want = O.orc_render_cube(zyx, org, sp, True, fo, w, h, nsamples=n, maxdist=maxdist)
    for kstep in (1, 0):
        tuning("fast_kstep", kstep)
        got = dev_cube1.render(fp, w, h, nsamples=n, maxdist=maxdist, trilinear=True, extra_flags=FAST)
        assert_fast_parity(got, want, f"{w}x{h} n={n} kstep={kstep}")


def test_fast_march_row_bands_accumulate_and_sweep(dev_cube1, cube1):
    zyx, org, sp = cube1
    fp, fo = frames_pair(POSES[:2], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    full = dev_cube1.render(fp, W, H, extra_flags=FAST)
    tiled = np.zeros_like(full)
    for r0, r1 in ((0, 9), (9, 10), (10, 41), (41, 64)):
        dev_cube1.render(fp, W, H, rows=(r0, r1), out=tiled, extra_flags=FAST)
    assert_bit_exact(tiled, full, "row bands of the tolerance-mode march tile the image exactly")
    twice = dev_cube1.render(fp, W, H, images=full.copy(), extra_flags=FAST)
    want = O.orc_render_cube(zyx, org, sp, True, fo, W, H, images=O.orc_render_cube(zyx, org, sp, True, fo, W, H))
    assert_fast_parity(twice, want, "accumulate")
    poses = S.orbit_poses(300)
    got = capi.render_sweep(dev_cube1, poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, extra_flags=FAST)
    fq = capi.frames_for_poses(poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    assert_bit_exact(got, dev_cube1.render(fq, W, H, extra_flags=FAST), "sweep call = frames call")


def test_fast_march_whole_orbit_sweep(orbit_sweep, tuning):
    """the headline workload in tolerance mode, all 69 M pixels against the CPU restatement"""
    zyx, org, sp, fp, want, valid = orbit_sweep
    c = capi.Cube(zyx, org, sp)
    try:
        for swz in (0, 1):
            tuning("tile_swizzle", swz)
            got = c.render(fp, W, H, trilinear=True, extra_flags=FAST)
            frac = assert_fast_parity(got, want, f"orbit sweep, tolerance mode, swizzle {swz}")
            rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-5 * np.abs(want).max())
            print(f"tolerance mode, whole sweep: {frac:.5%} of pixels within 1e-5, max relative difference {rel.max():.3g}")
            assert rel.max() < 1e-5
            tuning("tile_swizzle", 0)
        # the exact march with swizzled lanes is the exact march
        tuning("tile_swizzle", 1)
        assert_bit_exact(c.render(fp[::20], W, H, trilinear=True), want[::20], "exact march, swizzled lanes")
    finally:
        c.free()


def test_fast_march_config3_and_config4(gpu):
    """BASELINE configs[2] and [3] in tolerance mode: 256^3 / 1024^2 (200 and 2048 samples) and 512^3 (integer cell
    index: more than 2^24 voxels) / 1024^2, bands against the oracle"""
    for n, nsamples, rows in ((256, 200, (0, 1024)), (256, 2048, (500, 532)), (512, 200, (500, 564))):
        zyx, org, sp = S.make_cube(n, n, n)
        c = capi.Cube(zyx, org, sp)
        try:
            fp, fo = frames_pair(np.array([S.DEFAULT_POS], np.float32), 36.0, 36.0, 1024, 1024)
            got = c.render(fp, 1024, 1024, nsamples=nsamples, trilinear=True, extra_flags=FAST)
            want = O.orc_render_cube(zyx, org, sp, True, fo, 1024, 1024, nsamples=nsamples, rows=rows, nthreads=16)
            assert_fast_parity(got[:, rows[0]:rows[1]], want[:, rows[0]:rows[1]], f"{n}^3, {nsamples} samples")
        finally:
            c.free()


# ------------------------------------------------------------------ device-side sweep setup
def test_device_orbit_poses_and_frame_setup(gpu, dev_cube1, cube1):
    """poses generated and turned into frame records on the device: against the host path (glibc, bit-identical
    to the reference).  The device evaluates every libm call in fp64 and rounds once, i.e. correctly rounded
    but for ~1e-8 of the arguments; glibc 2.39's fp32 atan2f / acosf / asinf are not (34 % / 3.4 % / 3.1 % of
    their results are the OTHER neighbouring float, sinf / cosf about 1 %), so 4 frames in 10 differ from the
    host's in the last bit of some intermediate, which the `2 pi - atan2f(..)` style cancellations of
    src/Camera.cpp:99-104 turn into up to ~2e-6 in a rotation entry.  The bounds below are that measured
    spread with a factor 2-3; the host path stays the bit-exact one (and the default of the C++ classes)."""
    import torch

    n = 10000
    dev = torch.device("cuda", 0)
    d_poses = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d_frames = torch.zeros((n, 16), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    capi.orbit_poses_device(n, d_poses.data_ptr(), stream=s)
    capi.frame_setup_device(d_poses.data_ptr(), n, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, d_frames.data_ptr(), stream=s)
    capi.check_device_status(s)
    poses_dev = d_poses.cpu().numpy()
    poses = S.orbit_poses(n)
    same_pose = (poses_dev.view(np.uint32) == poses.view(np.uint32)).all(axis=1).mean()
    assert same_pose >= 0.999 and np.abs(poses_dev - poses).max() <= 2e-6, (same_pose, np.abs(poses_dev - poses).max())

    # frame records from the SAME poses: host (glibc) vs device
    d_poses.copy_(torch.from_numpy(poses))
    capi.frame_setup_device(d_poses.data_ptr(), n, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H, d_frames.data_ptr(), stream=s)
    capi.check_device_status(s)
    got = d_frames.cpu().numpy()
    want = capi.frames_for_poses(poses, AIM, S.SXI_FOV_W, S.SXI_FOV_H, W, H).view(np.float32).reshape(-1, 16)
    identical = (got.view(np.uint32) == want.view(np.uint32)).all(axis=1).mean()
    print(f"device frame setup: {identical:.4%} of {n} frame records bit-identical to the host's")
    assert np.array_equal(got[:, :3], want[:, :3]) and np.array_equal(got[:, 12:15], want[:, 12:15])
    rdiff = np.abs(got[:, 3:12] - want[:, 3:12]).max()
    adiff = np.abs(got[:, 15] - want[:, 15]).max()
    print(f"  max |rotation_inverse entry difference| {rdiff:.3g}, max |sun angle difference| {adiff:.3g} deg")
    assert rdiff <= 6e-6      # rotation_inverse entries (|r| <= 1); measured 2.2e-6
    assert adiff <= 3e-4      # sun angle, degrees (ulp at 360 is 3e-5)
    assert identical >= 0.4   # measured 58 % (same figure as an fp64-then-round evaluation on the host)

    # and what it means for the images: every 50th frame, device-made records vs host-made records
    idx = np.arange(0, n, 50)
    zyx, org, sp = cube1
    fr_dev = np.zeros(len(idx), capi.FRAME_DTYPE)
    fr_dev.view(np.float32).reshape(-1, 16)[:] = got[idx]
    fr_host = np.zeros(len(idx), capi.FRAME_DTYPE)
    fr_host.view(np.float32).reshape(-1, 16)[:] = want[idx]
    a = dev_cube1.render(fr_dev, W, H, trilinear=True)
    b = dev_cube1.render(fr_host, W, H, trilinear=True)
    frac = pixel_fraction_within(a, b)
    print(f"images from device-made frame records: {frac:.4%} of pixels within 1e-5 of the host-made ones")
    flux = abs(a.sum(dtype=np.float64) - b.sum(dtype=np.float64)) / b.sum(dtype=np.float64)
    print(f"  integrated flux differs by {flux:.3g} (relative)")
    assert frac >= 0.95 and flux <= 1e-4  # >= 95 % of pixels within 1e-5, flux within 1e-4


# ------------------------------------------------------------------ full-size configs (BASELINE.json configs[2], [3])
@pytest.mark.parametrize("n", [256])
def test_config3_256cube_1024_image(gpu, n):
    zyx, org, sp = S.make_cube(n, n, n)
    c = capi.Cube(zyx, org, sp)
    try:
        fp, fo = frames_pair(np.array([S.DEFAULT_POS], np.float32), 36.0, 36.0, 1024, 1024)
        got = c.render(fp, 1024, 1024, nsamples=200, trilinear=True)
        want = O.orc_render_cube(zyx, org, sp, True, fo, 1024, 1024, nsamples=200, nthreads=16)
        assert_bit_exact(got, want, "256^3, 1024x1024, 200 samples")
        assert c.info()["texture"]
        assert c.info()["quads"]
        for fl in (capi.NO_LEAN_MARCH, capi.NO_TEXTURE, capi.TEXTURE_ONLY, capi.NO_QUADS, capi.QUADS_ONLY,
                   capi.NO_QUADS | capi.NO_TEXTURE):
            assert_bit_exact(c.render(fp, 1024, 1024, nsamples=200, trilinear=True, extra_flags=fl), want,
                             f"corner-fetch variant {fl}")
        # fine step: a row band against the oracle, the whole image through band-tiling consistency
        fine = c.render(fp, 1024, 1024, nsamples=2048, trilinear=True)
        want_band = O.orc_render_cube(zyx, org, sp, True, fo, 1024, 1024, nsamples=2048, rows=(500, 532), nthreads=16)
        assert_bit_exact(fine[:, 500:532], want_band[:, 500:532], "fine step band")
        band = c.render(fp, 1024, 1024, nsamples=2048, trilinear=True, rows=(100, 228))
        assert_bit_exact(band[:, 100:228], fine[:, 100:228])
        # the fine march converges to the coarse one (same integral, step-size independent scale)
        rel = abs(fine.sum(dtype=np.float64) - got.sum(dtype=np.float64)) / got.sum(dtype=np.float64)
        assert rel < 5e-3
    finally:
        c.free()


def test_config4_512cube_image_in_row_bands(gpu):
    """BASELINE.json configs[3]: the 512^3 cube (512 MiB; 2 GiB as quads), a 1024x1024 image cut into the eight row
    bands the 8-GPU run gives its ranks.  Two bands against the oracle bit for bit, and the eight bands rendered
    one by one tile the single-call image exactly (the decomposition changes nothing)"""
    zyx, org, sp = S.make_cube(512, 512, 512)
    c = capi.Cube(zyx, org, sp)
    try:
        assert c.info()["quads"] and c.info()["positive"]
        fp, fo = frames_pair(np.array([S.DEFAULT_POS], np.float32), 36.0, 36.0, 1024, 1024)
        full = c.render(fp, 1024, 1024, trilinear=True)
        tiled = np.zeros_like(full)
        for b in range(8):
            c.render(fp, 1024, 1024, trilinear=True, rows=(128 * b, 128 * (b + 1)), out=tiled)
        assert_bit_exact(tiled, full, "eight row bands vs one call")
        for rows in ((0, 24), (500, 524)):
            want = O.orc_render_cube(zyx, org, sp, True, fo, 1024, 1024, rows=rows, nthreads=16)
            assert_bit_exact(full[:, rows[0]:rows[1]], want[:, rows[0]:rows[1]], f"rows {rows}")
        near = c.render(fp, 1024, 1024, trilinear=False, rows=(500, 524))
        want = O.orc_render_cube(zyx, org, sp, False, fo, 1024, 1024, rows=(500, 524), nthreads=16)
        assert_bit_exact(near[:, 500:524], want[:, 500:524], "nearest neighbour band")
        assert full.min() > 0
    finally:
        c.free()


# ------------------------------------------------------------------ CMEM
KA_DEFAULT = [((8, 4, 4), 0.0), ((4, 12, 10), 4.46853346e-06), ((3, 1.5, 1.2), 0.0), ((8, 10, 10), 8.15947317e-07),
              ((8, -10, -10), 8.15947317e-07), ((8, 10, -10), 8.15947317e-07)]
KA_TEST2 = [((8, 4, 4), 0.0), ((12, 12, 10), 4.57952495e-07), ((3, 1.5, 1.2), 0.0), ((12, 10, 10), 5.48567925e-07),
            ((12, -10, -10), 5.48567925e-07), ((12, 10, -10), 5.48567925e-07)]


def test_cmem_math_accuracy(gpu):
    """the library's own fp64 exp2 / log2 / pow / exp / acos / angle bundle stay within 2^-44 of the CUDA fp64
    library, so rounding them once to fp32 is a correctly rounded fp32 result almost everywhere"""
    err = capi.selftest_cmem_math(1 << 23)
    print({k: f"2^{np.log2(max(v, 1e-300)):.1f}" for k, v in err.items()})
    for name, v in err.items():
        assert v < 2.0 ** -44, (name, v)


def test_cmem_division_by_ten_is_the_ieee_division_for_every_float(gpu):
    """r / 10 (src/Cmem.cpp:158,165) is computed without a division where r is an ordinary number: all 2^32 bit patterns
    on the device"""
    kernel_form, bare_form = capi.selftest_div10()
    assert kernel_form == 0
    assert bare_form == 419431  # the subnormal quotients of r in [2^-126, 2^-125), which the range test excludes


CMEM_KW = [{}, O.CMEM_CLI, O.CMEM_TEST2, dict(dipole=0.2, b=(1, -2, -6), v=(450, 10, -20))]


@pytest.mark.parametrize("kw", CMEM_KW)
def test_cmem_filter_margin(gpu, kw):
    """the fp32 radius estimates of the region filter stay within an eighth of the margin the filter allows them
    (2e-3 R + 2e-3 RE), over 10^8 sample positions (incl. positions squeezed towards the degenerate planes and axes);
    the library itself refuses the filter for a parameter set that uses more than a quarter"""
    out = capi.selftest_cmem_filter(capi.cmem_init(**kw), 100_000_000)
    print(out)
    assert out["margin_used_mp"] <= 0.125 and out["margin_used_bs"] <= 0.125, out
    assert out["trusted_fraction"] > 0.9


@pytest.mark.parametrize("kw", CMEM_KW + [dict(dipole=0.5, p0=2.0, bs=20.0, A1=1e-5, ay_bs=1.5, az_bs=0.4),
                                          dict(dipole=-0.4, p3=1, bs=6.0, alpha=0.5, beta=3.0)])
def test_cmem_filter_is_invisible(gpu, kw):
    """filtered and unfiltered evaluation return the same bits: random positions, and 4000 rays sampled 2500
    times each so that every boundary is crossed at a resolution of 0.015 RE"""
    P = capi.cmem_init(**kw)
    rng = np.random.default_rng(5)
    pts = [rng.uniform(-35, 35, size=(4_000_000, 3))]
    d = rng.normal(size=(4000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:200, 2] *= 1e-4   # rays in (almost) the z = 0 plane
    d[200:400, 1] *= 1e-4
    t = np.linspace(2.0, 40.0, 2500)
    pts.append((d[:, None, :] * t[None, :, None]).reshape(-1, 3))
    pts = np.concatenate(pts).astype(np.float32)
    a = capi.sample_cmem(P, pts)
    b = capi.sample_cmem(P, pts, fast="nofilter")
    assert_bit_exact(a, b, "region table / filter changed a sample")
    assert_bit_exact(capi.sample_cmem(P, pts, fast="notable"), b, "region filter (without the table) changed a sample")


@pytest.mark.parametrize("kw", CMEM_KW + [dict(dipole=0.5, p0=2.0, bs=20.0, A1=1e-5, ay_bs=1.5, az_bs=0.4),
                                          dict(dipole=-0.4, p3=1, bs=6.0, alpha=0.5, beta=3.0)])
def test_cmem_region_table_never_disagrees(gpu, kw):
    """the region table (per bin of cos(theta): bounds of both radii over every phi) decides nine samples in ten with
    one 16-byte load -- and never differently from the exact evaluation: 10^8 pseudo-random positions"""
    out = capi.selftest_cmem_table(capi.cmem_init(**kw), 100_000_000)
    print(out)
    assert out["differing_samples"] == 0
    if out["table_in_use"]:
        assert out["decided_fraction"] > 0.5


@pytest.mark.parametrize("fast", [False, True, "libm64", "libm32"])
@pytest.mark.parametrize("kw,answers", [({}, KA_DEFAULT), (O.CMEM_TEST2, KA_TEST2)])
def test_cmem_known_answers_on_device(gpu, kw, answers, fast):
    """the reference's own CMEM tests (tests/CMEM_test.cpp), evaluated by the device sampler"""
    P = capi.cmem_init(**kw)
    got = capi.sample_cmem(P, [p for p, _ in answers], fast=fast)
    want = np.array([v for _, v in answers], np.float32)
    np.testing.assert_allclose(got, want, rtol=(3e-6 if fast in (True, "libm32") else 3e-7), atol=0)
    assert got[3] == got[4] == got[5]  # mirror symmetry in y and z
    if fast not in (True, "libm32"):
        assert_bit_exact(got, want, "precise mode reproduces the known answers bitwise")


@pytest.mark.parametrize("fast", [False, True, "libm64", "libm32"])
@pytest.mark.parametrize("kw", [{}, O.CMEM_CLI, O.CMEM_TEST2, dict(dipole=0.2, b=(1, -2, -6), v=(450, 10, -20))])
def test_cmem_samples_vs_oracle(gpu, kw, fast):
    P = capi.cmem_init(**kw)
    Po = O.orc_cmem_init(**kw)
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-30, 30, size=(200000, 3)), [[0, 5, 5], [5, 0, 5], [0, 0, 9], [7, 0, 0]]])
    got = capi.sample_cmem(P, pts, fast=fast)
    want = O.orc_cmem_sample(Po, pts)
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
    rel[(got == 0) & (want == 0)] = 0
    # region flips at the layer boundaries change a sample discontinuously; they must be rare
    assert np.mean(rel <= 1e-5) > 0.9995, np.mean(rel <= 1e-5)
    if fast not in (True, "libm32"):
        assert np.mean(got.view(np.uint32) == want.view(np.uint32)) > 0.90
    if fast is False:
        assert_bit_exact(got, capi.sample_cmem(P, pts, fast="nofilter"), "region filter changed a sample")
        # the two fp64-then-round evaluations (own blocks / CUDA library) must agree almost everywhere
        other = capi.sample_cmem(P, pts, fast="libm64")
        same = np.mean(got.view(np.uint32) == other.view(np.uint32))
        print(f"own fp64 blocks vs CUDA fp64 library: {same:.6%} of samples bit-identical")
        assert same > 0.9999


def test_cmem_radii_that_are_not_ordinary_numbers(gpu):
    """zero, tiny, huge, infinite and NaN radii take the literal b1 / b2 (out of line in the kernel): same values as the
    oracle and as the CUDA-library evaluation, with and without the region filter"""
    big = np.float32(3e38)
    pts = np.array([(0, 0, 0), (1e-35, 0, 0), (1e-32, 1e-33, -1e-33), (7e-31, 0, 0), (9e-31, 0, 0), (1e-20, 1e-21, 0),
                    (1e29, 0, 0), (0, 1e30, 1e29), (2e30, 0, 0), (-3e30, 1e30, 0), (1e19, 1e19, 1e19), (big, 0, 0),
                    (big, big, 0), (np.inf, 0, 0), (0, -np.inf, 1), (np.nan, 1, 1), (5, np.nan, 0), (12, 3, np.nan),
                    (1.5e19, 0, 0), (0, 0, 1.9e19)], np.float32)

    def same(a, b):
        return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))

    for kw in ({}, O.CMEM_TEST2):
        P = capi.cmem_init(**kw)
        with np.errstate(all="ignore"):
            want = O.orc_cmem_sample(O.orc_cmem_init(**kw), pts)
        for mode in (False, "nofilter", "notable", "libm64"):
            got = capi.sample_cmem(P, pts, fast=mode)
            assert same(got, want), (mode, got, want)


@pytest.mark.parametrize("B", [0.05, 2.4e-9, 2.46e-9, 1e-6])
def test_cmem_exponential_factor_of_b1(gpu, B):
    """CMEM::Init leaves B = |b| * 1e-9, for which expf(-B (theta/2)^4) is 1.0f for every theta and the kernel skips it.
    A record with another B (poked in: Init cannot produce one) still goes through the exponential"""
    P = capi.cmem_init()
    P["B"] = np.float32(B)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-25, 25, size=(200000, 3)).astype(np.float32)
    want = O.orc_cmem_sample(P.view(np.float32).copy(), pts)
    got = capi.sample_cmem(P, pts)
    assert_bit_exact(got, capi.sample_cmem(P, pts, fast="nofilter"), "region filter")
    other = capi.sample_cmem(P, pts, fast="libm64")  # the literal evaluation with the CUDA fp64 library
    assert np.mean(got.view(np.uint32) == other.view(np.uint32)) > 0.9999
    assert np.mean((got.view(np.uint32) == want.view(np.uint32)) | ((got == 0) & (want == 0))) > 0.9
    fp, fo = frames_pair(POSES[:2], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    img = capi.render_cmem(P, fp, W, H)
    ref = O.orc_render_cmem(P.view(np.float32).copy(), fo, W, H)
    for g, w in zip(img, ref):
        assert assert_image_parity(g, w, f"cmem B={B}") >= 0.99


@pytest.mark.parametrize("fast", [False, True, "libm64", "libm32"])
@pytest.mark.parametrize("name,kw", [("default", {}), ("cli", O.CMEM_CLI), ("test2", O.CMEM_TEST2)])
def test_cmem_render_vs_oracle_and_golden(gpu, golden, name, kw, fast):
    P = capi.cmem_init(**kw)
    fp, fo = frames_pair(golden["poses"], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    got = capi.render_cmem(P, fp, W, H, fast=fast)
    want = O.orc_render_cmem(O.orc_cmem_init(**kw), fo, W, H)
    for g, w in zip(got, want):
        assert assert_image_parity(g, w, f"cmem {name}") >= 0.99
    for g, w in zip(got[:2], golden[f"cmem_{name}_images"]):
        assert assert_image_parity(g, w, f"cmem {name} vs reference-rendered fixture") >= 0.99


def test_cmem_render_modes_are_bit_identical(gpu):
    """the default CMEM march (region filter) against the unfiltered evaluation and the CUDA-library one: same
    bits in every pixel; odd image sizes leave partially filled warps"""
    P = capi.cmem_init()
    poses = np.concatenate([S.orbit_poses(40), [S.DEFAULT_POS, (30.0, -25.0, 28.0), (-15.0, 3.0, 2.0)]]).astype(np.float32)
    for (w, h, fw, fh) in ((W, H, S.SXI_FOV_W, S.SXI_FOV_H), (37, 21, 40.0, 25.0)):
        fp = capi.frames_for_poses(poses, AIM, fw, fh, w, h)
        a = capi.render_cmem(P, fp, w, h)
        assert_bit_exact(a, capi.render_cmem(P, fp, w, h, fast="nofilter"), "filtered vs unfiltered")
        assert_bit_exact(a, capi.render_cmem(P, fp, w, h, fast="libm64"), "own fp64 blocks vs CUDA fp64 library")
        assert a.max() > 0


def test_cmem_row_bands_and_accumulate(gpu):
    P = capi.cmem_init()
    fp, _ = frames_pair(POSES[:2], S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    full = capi.render_cmem(P, fp, W, H)
    band = capi.render_cmem(P, fp, W, H, rows=(17, 40))
    assert_bit_exact(band[:, 17:40], full[:, 17:40])
    assert band[:, :17].max() == 0 and band[:, 40:].max() == 0
    twice = capi.render_cmem(P, fp, W, H, images=full.copy())
    np.testing.assert_allclose(twice, 2 * full, rtol=1e-5)


# ------------------------------------------------------------------ K4: user-defined Space subclasses
def test_accumulate_matches_oracle(gpu):
    rng = np.random.default_rng(2)
    s = rng.uniform(0, 1e-5, size=(6912, 200)).astype(np.float32)
    s[::7, 50] = -1.0     # rays that die early
    s[3, 0] = -1.0        # dies at the first sample
    s[5, 10] = np.nan     # NaN is not < 0: it accumulates
    _, w = capi.ray_table(200, 44)
    got = capi.accumulate(s, w)
    want = O.orc_accumulate(s, w)
    assert_bit_exact(got, want)
    assert got[3] == 0 and np.isnan(got[5])
    # CubeFake of the reference tests (tests/Camera_test.cpp:25-29): constant 1.f
    ones = np.ones((144 * 144, 200), np.float32)
    assert_bit_exact(capi.accumulate(ones, w), O.orc_accumulate(ones, w))
    img0 = rng.uniform(0, 1, size=6912).astype(np.float32)
    assert_bit_exact(capi.accumulate(s, w, image=img0.copy()), O.orc_accumulate(s, w, image=img0.copy()))


def test_scale_chain_exhaustive_slice(gpu):
    """the fp64 scale (src/Camera.cpp:201): fast path + exact fallback must equal the plain chain
    for every fp32 input; checked on a dense slice of bit patterns incl. subnormals, inf, NaN"""
    _, w = capi.ray_table(200, 44)
    pats = np.concatenate([
        np.arange(0, 1 << 32, 4099, dtype=np.uint64),
        np.arange(0x00000000, 0x00002000, dtype=np.uint64),  # subnormals
        np.arange(0x7F7FF000, 0x7F802000, dtype=np.uint64),  # around +inf
        np.arange(0x3A000000, 0x3A100000, dtype=np.uint64),  # a dense run of ordinary values
    ]).astype(np.uint32)
    s = pats.view(np.float32).reshape(-1, 1)
    got = capi.accumulate(s, w)
    want = O.orc_accumulate(s, w)
    assert_bit_exact(got, want)


# ------------------------------------------------------------------ device-buffer entry points (what bench.py times)
def test_device_entry_points_match_host_entry_points(dev_cube1):
    torch = pytest.importorskip("torch")
    fp, _ = frames_pair(POSES, S.SXI_FOV_W, S.SXI_FOV_H, W, H)
    want = dev_cube1.render(fp, W, H)
    d_frames = torch.from_numpy(fp.view(np.float32).reshape(-1, 16).copy()).cuda()
    d_images = torch.full((len(fp), H, W), -1.0, device="cuda")
    d_valid = torch.zeros(1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    capi.render_cube_device(dev_cube1, d_frames.data_ptr(), len(fp), W, H, d_images.data_ptr(),
                            d_valid=d_valid.data_ptr(), stream=stream)
    capi.check_device_status(stream)
    assert_bit_exact(d_images.cpu().numpy(), want)
    assert 0 < int(d_valid.item()) <= len(fp) * W * H * 200
    P = capi.cmem_init()
    capi.render_cmem_device(P, d_frames.data_ptr(), len(fp), W, H, d_images.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    assert_bit_exact(d_images.cpu().numpy(), capi.render_cmem(P, fp, W, H))
    assert capi.launch_count() > 0
