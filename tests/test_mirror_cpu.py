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


@needs_ref
@pytest.mark.parametrize("flavour", ["crlf", "trailing-space", "tabs", "cr-only-separators"])
def test_loader_whitespace_like_the_reference(tmp_path, small_dat, flavour):
    """the reference runs Helper::trim (isspace at both ends) over every VALUE line before splitting it at ' ':
    CRLF files, lines ending in spaces / tabs and "\r"-only separator lines between slices load there (ADVICE r1:
    they used to abort our loader)"""
    path, zyx, org, sp = small_dat
    lines = path.read_text().splitlines()
    nx, ny, nz = 13, 7, 9
    ncoord = -(-nx // 5) + -(-ny // 5) + -(-nz // 5)
    head, coords, data = lines[:10], lines[10:10 + ncoord], lines[10 + ncoord:]
    if flavour == "crlf":
        text = "\r\n".join(head + coords + data) + "\r\n"
    elif flavour == "trailing-space":
        text = "\n".join(head + coords + [d + "   " for d in data]) + "\n"
    elif flavour == "tabs":
        text = "\n".join(head + coords + ["\t" + d + " \t" for d in data]) + "\n"
    else:
        per_thread = -(-nx // 5)
        out = []
        for k in range(nz):
            out += data[k * ny * per_thread:(k + 1) * ny * per_thread] + ["\r"]
        text = "\n".join(head + coords + out) + "\n"
    p2 = tmp_path / f"{flavour}.dat"
    p2.write_bytes(text.encode())
    ref, mir = O.RefCube(path=p2), M.Cube(path=p2)
    assert list(mir.info()[0]) == list(ref.info()[0])
    assert_bit_exact(mir.info()[1], ref.info()[1], "origin / spacing")
    assert_bit_exact(mir.export(), ref.export(), flavour)
    assert_bit_exact(mir.export(), zyx)


@needs_ref
def test_loader_irregular_files_like_the_reference(tmp_path, small_dat):
    """files the flat fast path must NOT take: a value line that overshoots its row (the reference never completes
    that row), and a file that stops short -- the nested vectors come out as the reference leaves them"""
    path, zyx, org, sp = small_dat
    lines = path.read_text().splitlines()
    for name, edit in (("overshoot", lambda L: L[:40] + [L[40] + " 1.0 2.0"] + L[41:]), ("short", lambda L: L[:-30])):
        p2 = tmp_path / f"{name}.dat"
        p2.write_text("\n".join(edit(lines)) + "\n")
        ref, mir = O.RefCube(path=p2), M.Cube(path=p2)
        assert not mir.state()[0], "an irregular file must be loaded into the nested vectors"
        assert list(mir.info()[0]) == list(ref.info()[0]), name


def test_loader_keeps_one_flat_copy_until_slices_are_touched(small_dat):
    """Load() parses straight into the flat buffer; the nested vectors appear when somebody touches `slices`, and a
    write through them is what the next upload sees"""
    path, zyx, org, sp = small_dat
    c = M.Cube(path=path)
    assert c.state() == (True, False)
    dims, _ = c.info()  # (reads slices.size(): a const access keeps the flat buffer)
    assert dims[3] == 9 and c.state()[0]
    assert_bit_exact(c.export(), zyx)  # non-const access: the nested vectors are the truth now
    assert not c.state()[0]
    c.free()


def test_loader_side_car_cache(tmp_path, small_dat):
    """a side-car is written only on request, used whenever it is valid, ignored when the ASCII file changed, and
    can live in a directory of its own"""
    import os
    import shutil
    path, zyx, org, sp = small_dat
    src = tmp_path / "cube.dat"
    shutil.copy(path, src)
    try:
        M.cache_policy(False)
        a = M.Cube(path=src)
        assert not (tmp_path / "cube.dat.birpcache").exists() and a.state() == (True, False)
        M.cache_policy(True)
        b = M.Cube(path=src)
        assert (tmp_path / "cube.dat.birpcache").exists() and not b.state()[1]
        M.cache_policy(False)  # reading needs no switch
        c = M.Cube(path=src)
        assert c.state() == (True, True)
        assert list(c.info()[0]) == list(a.info()[0])
        assert_bit_exact(c.info()[1], a.info()[1])
        assert_bit_exact(c.export(), zyx)
        # the ASCII file changes (same size, new mtime): the side-car is stale and ignored
        text = src.read_text()
        src.write_text(text)
        os.utime(src, ns=(1_000_000_000, 1_000_000_000))
        d = M.Cube(path=src)
        assert d.state() == (True, False)
        # a truncated side-car is ignored, not trusted
        M.cache_policy(True)
        M.Cube(path=src)
        cache = tmp_path / "cube.dat.birpcache"
        cache.write_bytes(cache.read_bytes()[:-8])
        M.cache_policy(False)
        assert M.Cube(path=src).state() == (True, False)
        # side-cars in their own directory
        cdir = tmp_path / "caches"
        cdir.mkdir()
        cache.unlink()
        M.cache_policy(True, cdir)
        M.Cube(path=src)
        assert not cache.exists() and len(list(cdir.glob("cube.dat.*.birpcache"))) == 1
        M.cache_policy(False, cdir)
        e = M.Cube(path=src)
        assert e.state() == (True, True)
        assert_bit_exact(e.export(), zyx)
    finally:
        M.cache_policy(False, None)


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
