"""The C++ classes with the reference's public API (Space / DataCube / CMEM / Camera), rendering on
the GPU, against the unmodified reference classes driven the same way; and the reference's own CLI
(src/main.cpp, unmodified) linked against those classes, against the same CLI linked against the
reference's classes (BASELINE.json configs[0] and configs[1])."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mirror as M
import oracle as O
from birp_cpp_b200 import synthetic as S
from conftest import assert_bit_exact, assert_image_parity

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
REF_DIR = O.HERE / "_ref"
AIM = S.DEFAULT_AIM_X


@needs_ref
@pytest.mark.parametrize("tri", [False, True])
def test_camera_datacube_matches_reference_classes(gpu, cube1, tri):
    zyx, org, sp = cube1
    ref, mir = O.RefCube(zyx, org, sp, tri), M.Cube(zyx, org, sp, tri)
    for pos in (S.DEFAULT_POS, (30.0, -25.0, 28.0), (-14.0, 3.0, 5.0)):
        want = O.ref_render(ref, pos, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        got = M.render(mir, pos, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        assert_bit_exact(got, want, f"pose {pos}")
    # the main.cpp default camera: 36 x 36 deg at 0.25 deg/px = 144 x 144
    want = O.ref_render(ref, S.DEFAULT_POS, AIM, 0.25, 36.0, 36.0)
    assert want.shape == (144, 144)
    assert_bit_exact(M.render(mir, S.DEFAULT_POS, AIM, 0.25, 36.0, 36.0), want)
    # Render() twice on one Camera accumulates (src/Camera.cpp:223)
    once = O.ref_render(ref, S.DEFAULT_POS, AIM, 1.0, 16.0, 27.0)
    twice = O.ref_render(ref, S.DEFAULT_POS, AIM, 1.0, 16.0, 27.0, image=once.copy())
    assert_bit_exact(M.render(mir, S.DEFAULT_POS, AIM, 1.0, 16.0, 27.0, times=2), twice)
    # GetSample through the class (device sampler), incl. the out-of-bounds -1
    pts = np.array([[9.0, 0.0, 0.0], [4.0, 9.4, 6.0], [100.0, 0, 0], [-10.0, -40.0, -40.0]], np.float32)
    assert_bit_exact(M.sample(mir, pts), O.ref_sample(ref, pts))
    with pytest.raises(IndexError):
        M.sample(mir, [[np.nan, 0.0, 0.0]])
    mir.free()


@needs_ref
def test_camera_aim_y_and_z_reach_orient_like_the_reference(gpu, cube1):
    """Camera::SetAim(x, y, z): Orient() builds its look direction from all three components (src/Camera.cpp:323-327),
    so a non-zero aim y / z changes the roll of the image (the frame setup used to drop them)"""
    zyx, org, sp = cube1
    ref, mir = O.RefCube(zyx, org, sp, True), M.Cube(zyx, org, sp, True)
    base = O.ref_render(ref, S.DEFAULT_POS, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    for aim in ((AIM, 3.0, -2.0), (AIM, 0.0, 5.0), (2.0, -4.0, 0.0)):
        want = O.ref_render_aim(ref, S.DEFAULT_POS, aim, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        got = M.render_aim(mir, S.DEFAULT_POS, aim, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        assert_bit_exact(got, want, f"aim {aim}")
        assert not np.array_equal(want, base), "the aim's y / z components must matter"
    ref.free(), mir.free()


@needs_ref
def test_render_sweep_equals_one_reference_camera_per_position(gpu, cube1):
    """RenderSweep (C++ helper over birp_render_*_sweep) against the unmodified reference classes driven with
    one Camera per position: bit-identical images for the DataCube, tolerance for CMEM"""
    zyx, org, sp = cube1
    positions = np.concatenate([S.orbit_poses(6), [S.DEFAULT_POS, (30.0, -25.0, 28.0)]]).astype(np.float32)
    for tri in (True, False):
        ref, mir = O.RefCube(zyx, org, sp, tri), M.Cube(zyx, org, sp, tri)
        got = M.render_sweep(mir, positions, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        assert got.shape == (len(positions), S.SXI_H, S.SXI_W)
        for k, pos in enumerate(positions):
            assert_bit_exact(got[k], O.ref_render(ref, pos, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W), f"pose {k}")
        mir.free()
    ref, mir = O.RefCmem(), M.Cmem()
    got = M.render_sweep(mir, positions[:3], AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    for k in range(3):
        want = O.ref_render(ref, positions[k], AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
        assert assert_image_parity(got[k], want, f"cmem pose {k}") >= 0.99
    mir.free()
    with pytest.raises(Exception):  # a user-defined Space cannot be swept
        M.render_sweep(M.Fake(1.0), positions[:1], AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)


@needs_ref
def test_camera_nondefault_march(gpu, cube1):
    zyx, org, sp = cube1
    ref, mir = O.RefCube(zyx, org, sp, True), M.Cube(zyx, org, sp, True)
    want = O.ref_render(ref, S.DEFAULT_POS, AIM, 1.0, 21.0, 37.0, nsamples=517, maxdist=31)
    assert_bit_exact(M.render(mir, S.DEFAULT_POS, AIM, 1.0, 21.0, 37.0, nsamples=517, maxdist=31), want)
    mir.free()


@needs_ref
@pytest.mark.parametrize("kw", [{}, O.CMEM_CLI, O.CMEM_TEST2])
def test_camera_cmem_matches_reference_classes(gpu, kw):
    ref = O.RefCmem(False, **kw) if kw else O.RefCmem()
    mir = M.Cmem(**kw)
    want = O.ref_render(ref, S.DEFAULT_POS, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    got = M.render(mir, S.DEFAULT_POS, AIM, S.SXI_PX, S.SXI_FOV_H, S.SXI_FOV_W)
    assert assert_image_parity(got, want, "CMEM through Camera") >= 0.99
    # the reference's own known answers through CMEM::GetSample (tests/CMEM_test.cpp)
    pts = np.array([[8, 4, 4], [4, 12, 10], [3, 1.5, 1.2], [8, 10, 10], [8, -10, -10], [8, 10, -10]], np.float32)
    np.testing.assert_allclose(M.sample(mir, pts), O.ref_sample(ref, pts), rtol=3e-7, atol=0)
    mir.free()


@needs_ref
def test_camera_user_defined_space(gpu):
    """the reference tests' CubeFake (tests/Camera_test.cpp:25-49): any Space subclass still works,
    sees the same GetSample calls, and produces the same image"""
    fake = M.Fake(1.0, -1.0)
    img = M.render(fake, (1.0, 1.0, 1.0), 1.0, 0.25, 36.0, 36.0)
    assert img.shape == (144, 144) and fake.calls == 144 * 144 * 200
    # constant-ish field: compare with the restatement's accumulate on the same samples
    fr = O.orc_frame((1.0, 1.0, 1.0), 1.0, 36.0, 36.0, 144, 144)
    pts = O.orc_trace_points(fr, 144, 144, 200, 44)
    samples = (1.0 + np.float32(0.001) * pts[..., 0]).astype(np.float32).reshape(-1, 200)
    _, w = O.orc_ray_table(200, 44)
    assert_bit_exact(img.reshape(-1), O.orc_accumulate(samples, w))
    # a Space that turns negative beyond 6 RE: rays stop there, and so do the GetSample calls
    fake2 = M.Fake(1.0, 6.0)
    img2 = M.render(fake2, (1.0, 1.0, 1.0), 1.0, 1.0, 16.0, 27.0)
    assert 0 < fake2.calls < 16 * 27 * 200
    fr = O.orc_frame((1.0, 1.0, 1.0), 1.0, 27.0, 16.0, 27, 16)
    pts = O.orc_trace_points(fr, 27, 16, 200, 44)
    r2 = (pts.astype(np.float32) ** 2).sum(-1, dtype=np.float32)
    val = (1.0 + np.float32(0.001) * pts[..., 0]).astype(np.float32)
    # (the fake evaluates x*x + y*y + z*z in fp32 left to right)
    r2 = (pts[..., 0] * pts[..., 0] + pts[..., 1] * pts[..., 1]) + pts[..., 2] * pts[..., 2]
    samples = np.where(r2 > np.float32(36.0), np.float32(-1.0), val).astype(np.float32).reshape(-1, 200)
    assert_bit_exact(img2.reshape(-1), O.orc_accumulate(samples, w))
    first_neg = np.argmax(samples < 0, axis=1)
    first_neg[~(samples < 0).any(axis=1)] = 199
    assert fake2.calls == int((first_neg + 1).sum())
    fake.free()
    fake2.free()


def test_todat_and_tofits(gpu, tmp_path, cube1):
    zyx, org, sp = cube1
    mir = M.Cube(zyx, org, sp, True)
    img = M.render(mir, S.DEFAULT_POS, AIM, 1.0, 16.0, 27.0, dat=tmp_path / "out", fits=tmp_path / "out")
    rows = (tmp_path / "out.dat").read_text().splitlines()
    assert len(rows) == 16 and all(r.endswith(",") for r in rows)
    vals = np.array([[float(v) for v in r.split(",")[:-1]] for r in rows], np.float32)
    np.testing.assert_allclose(vals, img, atol=5e-7 * np.abs(img).max() + 1e-6)
    hdr, data = M.read_fits(tmp_path / "out.fits")
    assert_bit_exact(data, img)
    assert (tmp_path / "out.fits").stat().st_size % 2880 == 0
    assert hdr["BITPIX"] == "-32" and hdr["NAXIS1"] == "27" and hdr["NAXIS2"] == "16"
    assert float(hdr["CRVAL1"]) == -13.5 and float(hdr["CDELT1"]) == 1.0 and float(hdr["CRVAL2"]) == -8.0
    assert np.float32(hdr["POS_X"]) == np.float32(S.DEFAULT_POS[0]) and np.float32(hdr["AIM_X"]) == np.float32(AIM)
    mir.free()


# ------------------------------------------------------------------ the reference's own test files, unmodified
@pytest.mark.skipif(not (REF_DIR / "reference_tests_b200").exists(), reason="reference test build absent")
def test_reference_test_files_pass_against_our_classes(gpu):
    """tests/CMEM_test.cpp + tests/Space_test.cpp of the reference, compiled unchanged against
    birp_cpp_b200's classes (oracle/Makefile: reference_tests_b200): CMEM::GetSample runs on the GPU and
    the 12 known answers are compared with EXPECT_EQ, i.e. bit-exactly"""
    out = subprocess.run([str(REF_DIR / "reference_tests_b200")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "3 tests ran, 0 failed" in out.stdout
    assert "[       OK ] CMEM.BasicSamples" in out.stdout and "[       OK ] CMEM.InitParamsSamples" in out.stdout


# ------------------------------------------------------------------ the reference CLI, unmodified
def _run(exe, args, cwd):
    out = subprocess.run([str(exe)] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


@needs_ref
@pytest.mark.skipif(not (REF_DIR / "birp_main_b200").exists() or not (REF_DIR / "birp_main_ref").exists(),
                    reason="reference CLI builds absent")
@pytest.mark.parametrize("tri", [False, True])
def test_reference_cli_datacube_drop_in(gpu, tmp_path, tri):
    """BASELINE.json configs[0]: `birp -i cube.dat [-t] -h 16 -w 27 -s 0.25`, the reference main.cpp
    linked against our classes vs linked against its own"""
    zyx, org, sp = S.make_cube(41, 65, 65, lo=(-10.0, -40.0, -40.0), hi=(40.0, 40.0, 40.0))
    dat = tmp_path / "syn_41x65x65.dat"
    S.write_dat(dat, zyx, org, sp)
    for sub in ("ref", "b200"):
        (tmp_path / sub).mkdir()
    args = ["-i", str(dat), "-h", "16", "-w", "27", "-s", "0.25"] + (["-t"] if tri else [])
    _run(REF_DIR / "birp_main_ref", args + ["-o", str(tmp_path / "ref") + "/"], tmp_path)
    out = _run(REF_DIR / "birp_main_b200", args + ["-o", str(tmp_path / "b200") + "/"], tmp_path)
    assert "Completed rendering" in out
    h_ref, img_ref = M.read_fits(tmp_path / "ref" / (dat.name + ".fits"))
    h_new, img_new = M.read_fits(tmp_path / "b200" / (dat.name + ".fits"))
    assert img_ref.shape == (64, 108) and img_ref.max() > 0
    assert_bit_exact(img_new, img_ref, "CLI render")
    assert h_new == h_ref


@needs_ref
@pytest.mark.skipif(not (REF_DIR / "birp_main_b200").exists() or not (REF_DIR / "birp_main_ref").exists(),
                    reason="reference CLI builds absent")
def test_reference_cli_batch_directory_drop_in(gpu, tmp_path):
    """the reference's batch mode (src/main.cpp:242-264): `birp -i dir/` renders every cube of a
    directory with one pose -- one upload + one render per cube through our classes"""
    cubes = tmp_path / "cubes"
    cubes.mkdir()
    for n, seed in enumerate((11, 12, 13)):
        zyx, org, sp = S.make_cube(31, 41, 37, lo=(-10.0, -40.0, -40.0), hi=(40.0, 40.0, 40.0), seed=seed)
        S.write_dat(cubes / f"px_uni_{n:04d}.dat", zyx, org, sp)
    for sub in ("ref", "b200"):
        (tmp_path / sub).mkdir()
    args = ["-i", str(cubes) + "/", "-t", "-h", "8", "-w", "12", "-s", "0.25"]
    _run(REF_DIR / "birp_main_ref", args + ["-o", str(tmp_path / "ref") + "/"], tmp_path)
    _run(REF_DIR / "birp_main_b200", args + ["-o", str(tmp_path / "b200") + "/"], tmp_path)
    names = sorted(p.name for p in (tmp_path / "ref").glob("*.fits"))
    assert names == [f"px_uni_{n:04d}.dat.fits" for n in range(3)]
    images = []
    for name in names:
        h_ref, img_ref = M.read_fits(tmp_path / "ref" / name)
        h_new, img_new = M.read_fits(tmp_path / "b200" / name)
        assert_bit_exact(img_new, img_ref, name)
        assert h_new == h_ref
        images.append(img_ref)
    assert not np.array_equal(images[0], images[1])  # different cubes really were rendered


def test_pipelined_batch_directory_equals_cube_by_cube(gpu, tmp_path):
    """RenderBatchDirectory (parse-ahead threads) writes the same files as rendering cube by cube"""
    cubes = tmp_path / "cubes"
    cubes.mkdir()
    data = []
    for n, seed in enumerate((21, 22, 23, 24, 25)):
        zyx, org, sp = S.make_cube(25, 33, 29, lo=(-10.0, -40.0, -40.0), hi=(40.0, 40.0, 40.0), seed=seed)
        S.write_dat(cubes / f"px_uni_{n:04d}.dat", zyx, org, sp)
        data.append((zyx, org, sp))
    (cubes / "broken.dat").write_text("not a cube\n")
    out = tmp_path / "out"
    out.mkdir()
    ok, failed = M.render_batch(cubes, out, S.DEFAULT_POS, AIM, 0.25, 8.0, 12.0, trilinear=True, threads=3)
    assert (ok, failed) == (5, 1)
    for n in range(len(data)):
        hdr, img = M.read_fits(out / f"px_uni_{n:04d}.dat.fits")
        cube = M.Cube(path=cubes / f"px_uni_{n:04d}.dat", trilinear=True)  # (spacing comes from the file's axes)
        assert_bit_exact(img, M.render(cube, S.DEFAULT_POS, AIM, 0.25, 8.0, 12.0), f"cube {n}")
        cube.free()
    assert not (out / "broken.dat.fits").exists()


@needs_ref
@pytest.mark.skipif(not (REF_DIR / "birp_main_b200").exists() or not (REF_DIR / "birp_main_ref").exists(),
                    reason="reference CLI builds absent")
def test_reference_cli_cmem_drop_in(gpu, tmp_path):
    """BASELINE.json configs[1]: `birp -c -h 16 -w 27` (CLI parameters p0=1, alpha=beta=0)"""
    for sub in ("ref", "b200"):
        (tmp_path / sub).mkdir()
    args = ["-c", "-h", "16", "-w", "27", "-s", "0.25"]
    _run(REF_DIR / "birp_main_ref", args + ["-o", str(tmp_path / "ref") + "/"], tmp_path)
    _run(REF_DIR / "birp_main_b200", args + ["-o", str(tmp_path / "b200") + "/"], tmp_path)
    h_ref, img_ref = M.read_fits(tmp_path / "ref" / "cmembirp.fits")
    h_new, img_new = M.read_fits(tmp_path / "b200" / "cmembirp.fits")
    assert img_ref.shape == (64, 108) and img_ref.max() > 0
    assert assert_image_parity(img_new, img_ref, "CLI CMEM render") >= 0.99
    assert h_new == h_ref
    # the CLI path really uses the alpha = beta = 0 parameter set
    fr = O.orc_frame(S.DEFAULT_POS, AIM, 27.0, 16.0, 108, 64)
    assert_bit_exact(O.orc_render_cmem(O.orc_cmem_init(**O.CMEM_CLI), fr[None], 108, 64)[0], img_ref)
