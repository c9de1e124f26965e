"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
its host-side setup math is bit-identical to the oracle, and compute entry points fail loudly
(not silently, not through a CPU path) when no GPU is present."""
import ctypes as C
import re

import numpy as np
import pytest

import oracle as O
from birp_cpp_b200 import capi, synthetic as S
from conftest import assert_bit_exact


def test_library_exports_every_declared_symbol():
    header = capi.HEADER_PATH.read_text()
    declared = set(re.findall(r"\b(birp_[a-z_0-9]+)\s*\(", header))
    declared -= {"birp_cube", "birp_frame", "birp_cmem_params"}
    assert len(declared) >= 20
    lib = capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/birp_b200.h but not exported"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert b"sm_100a" in lib.birp_version()


def test_struct_layouts_match_header():
    assert capi.FRAME_DTYPE.itemsize == 64 and capi.CMEM_DTYPE.itemsize == 108
    assert O.orc().orc_frame_floats() * 4 == capi.FRAME_DTYPE.itemsize
    assert O.orc().orc_cmem_param_floats() * 4 == capi.CMEM_DTYPE.itemsize


def test_frame_setup_matches_oracle():
    rng = np.random.default_rng(0)
    poses = np.concatenate([rng.uniform(-25, 25, size=(300, 3)), S.orbit_poses(200), [S.DEFAULT_POS]])
    for fov_x, fov_y, W, H in ((27.0, 16.0, 108, 64), (36.0, 36.0, 1024, 1024), (9.0, 7.0, 9, 7)):
        got = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, fov_x, fov_y, W, H)
        want = O.orc_frames(poses, S.DEFAULT_AIM_X, fov_x, fov_y, W, H)
        assert_bit_exact(got.view(np.float32).reshape(-1, 16), want, "birp_frame_setup")


def test_frame_setup_with_the_whole_aim_vector():
    """birp_frame_setup_aim: the record of a camera aimed at (x, y, z) is the restatement's bit for bit, its Sun roll
    angle the reference's own Orient() (which reads all three aim components, src/Camera.cpp:323-327); aim y = z = 0
    is birp_frame_setup"""
    rng = np.random.default_rng(3)
    for t in range(400):
        pos = rng.uniform(-30, 30, 3).astype(np.float32)
        aim = rng.uniform(-10, 10, 3).astype(np.float32)
        if t % 4 == 0:
            aim[1:] = 0
        fr = capi.frame_setup_aim(pos, aim, 27.0, 16.0, 108, 64).view(np.float32)
        assert_bit_exact(fr, O.orc_frame_aim(pos, aim, 27.0, 16.0, 108, 64), f"pos {pos} aim {aim}")
        if O.have_ref() and t < 100:
            assert_bit_exact(fr[15:16], np.array([O.ref_orient_aim(pos, aim)]), "Orient()")
        if t % 4 == 0:
            assert_bit_exact(fr, capi.frame_setup(pos, float(aim[0]), 27.0, 16.0, 108, 64).view(np.float32))


def test_frame_setup_batch_equals_single_calls():
    poses = S.orbit_poses(1000)
    batch = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, 27.0, 16.0, 108, 64)
    one_thread = capi.frames_for_poses(poses, S.DEFAULT_AIM_X, 27.0, 16.0, 108, 64, nthreads=1)
    singles = np.concatenate([capi.frame_setup(p, S.DEFAULT_AIM_X, 27.0, 16.0, 108, 64) for p in poses[:50]])
    assert batch.tobytes() == one_thread.tobytes()
    assert batch[:50].tobytes() == singles.tobytes()
    assert len(capi.frames_for_poses(np.zeros((0, 3), np.float32), 1.0, 27.0, 16.0, 108, 64)) == 0


def test_frame_setup_golden(golden):
    got = capi.frames_for_poses(golden["poses"], float(golden["aim_x"]), 27.0, 16.0, 108, 64)
    assert_bit_exact(got["sun_angle"], golden["orient"])


def test_frame_setup_on_x_axis_is_nan_like_reference():
    # camera on the x axis sunward of the aim point: the reference produces NaN angles, then throws from .at() during the
    # march; the frame record carries the NaNs and the render call reports BIRP_ERR_RANGE (GPU test)
    fr = capi.frame_setup((10.0, 0.0, 0.0), S.DEFAULT_AIM_X, 27.0, 16.0, 108, 64)
    want = O.orc_frame((10.0, 0.0, 0.0), S.DEFAULT_AIM_X, 27.0, 16.0, 108, 64)
    assert np.isnan(want[3:12]).any()
    assert np.array_equal(np.isnan(fr.view(np.float32).reshape(16)), np.isnan(want))


def test_orbit_poses_equal_the_numpy_generator():
    """birp_orbit_poses (C, fp64 Newton on Kepler's equation) == synthetic.orbit_poses (numpy), bit for bit"""
    for n in (1, 7, 1000, 10000):
        got, want = capi.orbit_poses(n), S.orbit_poses(n)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), n
    other = capi.orbit_poses(64, perigee=2.5, apogee=12.0, inclination_deg=40.0)
    assert np.array_equal(other.view(np.uint32), S.orbit_poses(64, 2.5, 12.0, 40.0).view(np.uint32))
    with pytest.raises(capi.BirpError):
        capi.orbit_poses(4, perigee=5.0, apogee=1.0)


def test_ray_table_matches_oracle():
    for n, m in ((200, 44), (2048, 44), (517, 31), (1, 44), (2, 5)):
        d, w = capi.ray_table(n, m)
        do, wo = O.orc_ray_table(n, m)
        assert_bit_exact(d, do)
        assert_bit_exact(w, wo)


@pytest.mark.parametrize("kw", [{}, O.CMEM_CLI, O.CMEM_TEST2, dict(dipole=0.2, b=(1, -2, -6), v=(450, 10, -20)),
                                dict(density=7.5, ay_bs=0.8, az_bs=0.7, bs=11.0, A1=1e-5, A2=2e-6, p0=0.9)])
def test_cmem_init_matches_oracle(kw):
    got = capi.cmem_init(**kw)
    assert_bit_exact(got.view(np.float32), O.orc_cmem_init(**kw), "birp_cmem_init")


def test_cmem_cli_parameters(golden):
    assert_bit_exact(capi.cmem_init_cli().view(np.float32), golden["cmem_cli_constants"])
    assert_bit_exact(capi.cmem_init().view(np.float32), golden["cmem_default_constants"])


def test_argument_errors():
    lib = capi.lib()
    assert lib.birp_ray_table(0, 44, None, None) == capi.BIRP_ERR_ARG
    assert lib.birp_frame_setup(None, 1.0, 1.0, 1.0, 4, 4, None) == capi.BIRP_ERR_ARG
    assert b"bad" in lib.birp_last_error()
    out = C.c_void_p()
    assert lib.birp_cube_upload(None, 4, 4, 4, None, None, C.byref(out)) == capi.BIRP_ERR_ARG


def test_compute_fails_loudly_without_gpu():
    try:
        n = capi.device_count()
    except capi.BirpError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    zyx = np.ones((4, 4, 4), np.float32)
    with pytest.raises(capi.BirpError) as e:
        capi.Cube(zyx, (0, 0, 0), (1, 1, 1))
    assert e.value.code == capi.BIRP_ERR_CUDA
    with pytest.raises(capi.BirpError):
        capi.sample_cmem(capi.cmem_init(), [[8, 4, 4]])
    with pytest.raises(capi.BirpError):
        capi.accumulate(np.ones((2, 3), np.float32), 0.22)


def test_synthetic_inputs():
    zyx, org, sp = S.make_cube(17, 19, 23)
    assert zyx.shape == (23, 19, 17) and zyx.dtype == np.float32 and zyx.min() > 0
    neg, _, _ = S.make_cube(33, 9, 9, negative_slab=True)
    assert (neg < 0).any()
    p = S.orbit_poses(1000)
    r = np.linalg.norm(p, axis=1)
    assert p.dtype == np.float32 and 1.7 < r.min() < 1.9 and 19.9 < r.max() < 20.1
    assert not np.any((p[:, 1] == 0) & (p[:, 2] == 0))
    assert p[np.argmax(r), 2] > 18.0  # apogee over the north pole
