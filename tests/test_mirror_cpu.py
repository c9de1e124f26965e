"""Host-only behaviour of the C++ classes that mirror the reference API: the ASCII cube loader
against the reference's own loader, Helper's matrix routines, CMEM constants.  No GPU."""
import numpy as np
import pytest

import mirror as M
import oracle as O
from birp_cpp_b200 import synthetic as S
from conftest import assert_bit_exact

needs_ref = pytest.mark.skipif(not (O.have_ref() or O.REFERENCE_ROOT.exists()), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def small_dat(tmp_path_factory):
    zyx, org, sp = S.make_cube(13, 7, 9, lo=(-3.0, -1.5, -2.0), hi=(3.0, 1.5, 2.0))
    path = tmp_path_factory.mktemp("dat") / "syn_13x7x9.dat"
    S.write_dat(path, zyx, org, sp, per_line=5)
    return path, zyx, org, sp


@needs_ref
def test_loader_matches_reference_loader(small_dat):
    path, zyx, org, sp = small_dat
    ref = O.RefCube(path=path)
    mir = M.Cube(path=path)
    rd, rg = ref.info()
    md, mg = mir.info()
    assert list(md[:3]) == list(rd[:3]) == [13, 7, 9]
    assert_bit_exact(mg, rg, "origin / spacing")
    assert_bit_exact(mir.export(), ref.export(), "slices")
    # the text round-trips the synthetic values exactly (repr of float32)
    assert_bit_exact(mir.export(), zyx)
    assert_bit_exact(mg[:3], org)


@needs_ref
def test_loader_blank_lines_between_slices(tmp_path, small_dat):
    path, zyx, org, sp = small_dat
    lines = path.read_text().splitlines()
    # re-emit with a blank line after every slice, as exported MHD cubes have
    nx, ny, nz = 13, 7, 9
    head = lines[:10]
    rest = lines[10:]
    ncoord = -(-nx // 5) + -(-ny // 5) + -(-nz // 5)
    coords, data = rest[:ncoord], rest[ncoord:]
    per_thread = -(-nx // 5)
    out = head + coords
    for k in range(nz):
        out += data[k * ny * per_thread:(k + 1) * ny * per_thread] + [""]
    p2 = tmp_path / "blank.dat"
    p2.write_text("\n".join(out) + "\n")
    ref, mir = O.RefCube(path=p2), M.Cube(path=p2)
    assert list(mir.info()[0][:3]) == list(ref.info()[0][:3])
    # the reference leaves one extra, empty trailing slice in this layout; so do we
    assert mir.info()[0][3] == ref.info()[0][3] == nz + 1
    assert_bit_exact(mir.export(), ref.export())


@needs_ref
@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_loader_worker_threads_do_not_change_the_result(tmp_path, monkeypatch, threads):
    """the value section is converted on worker threads over line-aligned pieces; whatever the number of pieces
    (more pieces than lines included), the reference's line-by-line state machine sees the same lines"""
    zyx, org, sp = S.make_cube(37, 23, 19, lo=(-3.0, -1.5, -2.0), hi=(3.0, 1.5, 2.0))
    path = tmp_path / "syn_37x23x19.dat"
    S.write_dat(path, zyx, org, sp, per_line=7)
    text = path.read_text().splitlines()
    # ragged extras the reference tolerates: a blank line in the middle of the values, no newline at the very end
    text.insert(len(text) // 2, "")
    path.write_text("\n".join(text))
    monkeypatch.setenv("BIRP_LOAD_THREADS", threads)
    ref, mir = O.RefCube(path=path), M.Cube(path=path)
    assert list(mir.info()[0]) == list(ref.info()[0])
    assert_bit_exact(mir.info()[1], ref.info()[1], "origin / spacing")
    assert_bit_exact(mir.export(), ref.export(), "slices")


def test_loader_errors(tmp_path):
    with pytest.raises(ValueError):
        M.Cube(path=tmp_path / "missing.dat")
    bad = tmp_path / "bad.dat"
    bad.write_text("meta\n" + "#\n" * 8 + "0 3 3\n1 2 3\n")
    with pytest.raises(ValueError):
        M.Cube(path=bad)


def test_helper_inverse_matches_frame_setup():
    # Helper::GetInverse on a rotation reproduces what birp_frame_setup stores
    rng = np.random.default_rng(4)
    for _ in range(20):
        a, b, c = rng.uniform(0, 6.28, 3)
        rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
        rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]], np.float32)
        m = (rz @ rx).astype(np.float32)
        inv = M.helper_inverse3(m)
        np.testing.assert_allclose(inv @ m, np.eye(3), atol=2e-6)
    with pytest.raises(RuntimeError, match="Determinant is 0"):
        M.helper_inverse3(np.zeros(9))
