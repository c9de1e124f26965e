"""ctypes access to tests/cpp/libmirror_harness.so: the product's C++ classes (same public API as
the reference's) driven exactly like oracle/ref_harness.cpp drives the reference's."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "cpp" / "libmirror_harness.so"
_vp = C.c_void_p
_lib = None

KIND_CUBE, KIND_CMEM, KIND_FAKE = 0, 1, 2
ERR_KINDS = {1: ValueError, 2: IndexError, 3: RuntimeError, 4: Exception}


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", str(HERE.parent / "birp_cpp_b200")], check=True, capture_output=True)
        subprocess.run(["make", "-C", str(HERE / "cpp")], check=True, capture_output=True)
        L = C.CDLL(str(SO))
        L.mir_last_error.restype = C.c_char_p
        L.mir_last_error_kind.restype = C.c_int
        L.mir_cube_create.restype = _vp
        L.mir_cube_create.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]
        L.mir_cube_load.restype = _vp
        L.mir_cube_load.argtypes = [C.c_char_p, C.c_int]
        L.mir_cube_set_trilinear.argtypes = [_vp, C.c_int]
        L.mir_cube_info.argtypes = [_vp, _vp, _vp]
        L.mir_cube_export.restype = C.c_int
        L.mir_cube_export.argtypes = [_vp, _vp]
        L.mir_cube_free.argtypes = [_vp]
        L.mir_cube_state.argtypes = [_vp, _vp]
        L.mir_cube_poke.restype = C.c_int
        L.mir_cube_poke.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_float]
        L.mir_cache_policy.argtypes = [C.c_int, C.c_char_p]
        L.mir_cube_samples.restype = C.c_int
        L.mir_cube_samples.argtypes = [_vp, _vp, C.c_long, _vp]
        L.mir_cmem_create_default.restype = _vp
        L.mir_cmem_create.restype = _vp
        L.mir_cmem_create.argtypes = [C.c_int, _vp, C.c_int, _vp, C.c_float, C.c_float, C.c_int, C.c_int,
                                      C.c_int] + [C.c_float] * 9
        L.mir_cmem_free.argtypes = [_vp]
        L.mir_fake_create.restype = _vp
        L.mir_fake_create.argtypes = [C.c_float, C.c_float]
        L.mir_fake_calls.restype = C.c_long
        L.mir_fake_calls.argtypes = [_vp]
        L.mir_fake_free.argtypes = [_vp]
        L.mir_space_sample.restype = C.c_int
        L.mir_space_sample.argtypes = [_vp, C.c_int, _vp, C.c_long, _vp]
        L.mir_render.restype = C.c_int
        L.mir_render.argtypes = [_vp, C.c_int, C.c_float, C.c_float, C.c_float, _vp, C.c_float, C.c_int, C.c_int,
                                 C.c_int, _vp, _vp, C.c_char_p, C.c_char_p]
        L.mir_render3.restype = C.c_int
        L.mir_render3.argtypes = [_vp, C.c_int, C.c_float, C.c_float, C.c_float, _vp, _vp, C.c_int, C.c_int,
                                  C.c_int, _vp, _vp, C.c_char_p, C.c_char_p]
        L.mir_fits_of_image.restype = C.c_int
        L.mir_fits_of_image.argtypes = [C.c_float, C.c_float, C.c_float, _vp, _vp, _vp, _vp, C.c_char_p]
        L.mir_render_batch.restype = C.c_int
        L.mir_render_batch.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_float, _vp, C.c_float, C.c_int,
                                       C.c_int, _vp]
        L.mir_render_sweep.restype = C.c_int
        L.mir_render_sweep.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp, _vp]
        L.mir_helper_inverse3.restype = C.c_int
        L.mir_helper_inverse3.argtypes = [_vp, _vp]
        _lib = L
    return _lib


def _raise():
    kind = lib().mir_last_error_kind()
    raise ERR_KINDS.get(kind, Exception)(lib().mir_last_error().decode())


def _p(a):
    return a.ctypes.data_as(_vp)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


class Space:
    kind = KIND_CUBE
    h = None


class Cube(Space):
    kind = KIND_CUBE

    def __init__(self, zyx=None, origin=None, spacing=None, trilinear=False, path=None):
        if path is not None:
            self.h = lib().mir_cube_load(str(path).encode(), int(trilinear))
            if not self.h:
                _raise()
        else:
            zyx = _f32(zyx)
            nz, ny, nx = zyx.shape
            o, s = _f32(origin, (3,)), _f32(spacing, (3,))
            self.h = lib().mir_cube_create(_p(zyx), nx, ny, nz, _p(o), _p(s), int(trilinear))

    def set_trilinear(self, on):
        lib().mir_cube_set_trilinear(self.h, int(on))

    def info(self):
        dims = np.zeros(4, np.int32)
        geom = np.zeros(6, np.float32)
        lib().mir_cube_info(self.h, _p(dims), _p(geom))
        return dims, geom

    def export(self):
        dims, _ = self.info()
        out = np.zeros((dims[2], dims[1], dims[0]), np.float32)
        if lib().mir_cube_export(self.h, _p(out)):
            _raise()
        return out

    def state(self):
        """(nested slices not built since Load, Load served by a side-car)"""
        st = np.zeros(2, np.int32)
        lib().mir_cube_state(self.h, _p(st))
        return bool(st[0]), bool(st[1])

    def poke(self, k, j, i, v):
        if lib().mir_cube_poke(self.h, k, j, i, v):
            _raise()

    def samples(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.float32)
        if lib().mir_cube_samples(self.h, _p(xyz), len(xyz), _p(out)):
            _raise()
        return out

    def free(self):
        if self.h:
            lib().mir_cube_free(self.h)
            self.h = None


class Cmem(Space):
    kind = KIND_CMEM

    def __init__(self, **kw):
        if not kw:
            self.h = lib().mir_cmem_create_default()
        else:
            import oracle as O
            a = dict(O.CMEM_DEFAULTS)
            a.update(kw)
            hv, va, hb, ba, *rest = O._cmem_args(**a)
            self.h = lib().mir_cmem_create(hv, _p(va), hb, _p(ba), *rest)

    def free(self):
        if self.h:
            lib().mir_cmem_free(self.h)
            self.h = None


class Fake(Space):
    kind = KIND_FAKE

    def __init__(self, value=1.0, negative_beyond=-1.0):
        self.h = lib().mir_fake_create(value, negative_beyond)

    @property
    def calls(self):
        return lib().mir_fake_calls(self.h)

    def free(self):
        if self.h:
            lib().mir_fake_free(self.h)
            self.h = None


def cache_policy(writes: bool, directory=None):
    """DataCube::SetCacheWrites / SetCacheDirectory"""
    lib().mir_cache_policy(int(writes), str(directory).encode() if directory else None)


def sample(space, xyz):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), np.float32)
    if lib().mir_space_sample(space.h, space.kind, _p(xyz), len(xyz), _p(out)):
        _raise()
    return out


def render(space, pos, aim_x, px_deg, fov_h, fov_w, nsamples=0, maxdist=0, times=1, dat=None, fits=None):
    wh = np.zeros(2, np.int32)
    p = _f32(pos, (3,))
    if lib().mir_render(space.h, space.kind, px_deg, fov_h, fov_w, _p(p), aim_x, nsamples, maxdist, times, None,
                        _p(wh), None, None):
        _raise()
    img = np.zeros((int(wh[1]), int(wh[0])), np.float32)
    if lib().mir_render(space.h, space.kind, px_deg, fov_h, fov_w, _p(p), aim_x, nsamples, maxdist, times, _p(img),
                        _p(wh), str(dat).encode() if dat else None, str(fits).encode() if fits else None):
        _raise()
    return img


def render_aim(space, pos, aim, px_deg, fov_h, fov_w, nsamples=0, maxdist=0):
    """Camera(space, ...).SetPosition(pos).SetAim(aim x, y, z).Render()"""
    wh = np.zeros(2, np.int32)
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    if lib().mir_render3(space.h, space.kind, px_deg, fov_h, fov_w, _p(p), _p(a), nsamples, maxdist, 1, None, _p(wh),
                         None, None):
        _raise()
    img = np.zeros((int(wh[1]), int(wh[0])), np.float32)
    if lib().mir_render3(space.h, space.kind, px_deg, fov_h, fov_w, _p(p), _p(a), nsamples, maxdist, 1, _p(img), _p(wh),
                         None, None):
        _raise()
    return img


def fits_of_image(path, image, pos, aim, px_deg, fov_h, fov_w):
    """Camera::ToFITS of a hand-filled image (no render, no GPU); writes <path>.fits; image None: (W, H) only"""
    wh = np.zeros(2, np.int32)
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    if image is None:
        if lib().mir_fits_of_image(px_deg, fov_h, fov_w, _p(p), _p(a), None, _p(wh), None):
            _raise()
        return int(wh[0]), int(wh[1])
    img = _f32(image)
    if lib().mir_fits_of_image(px_deg, fov_h, fov_w, _p(p), _p(a), _p(img), _p(wh), str(path).encode()):
        _raise()
    return int(wh[0]), int(wh[1])


def render_sweep(space, positions, aim_x, px_deg, fov_h, fov_w):
    """RenderSweep: n positions -> [n][H][W] images"""
    pos = _f32(positions).reshape(-1, 3)
    wh = np.zeros(2, np.int32)
    if lib().mir_render_sweep(space.h, space.kind, _p(pos), 0, aim_x, px_deg, fov_h, fov_w, None, _p(wh)):
        _raise()
    img = np.zeros((len(pos), int(wh[1]), int(wh[0])), np.float32)
    if lib().mir_render_sweep(space.h, space.kind, _p(pos), len(pos), aim_x, px_deg, fov_h, fov_w, _p(img), _p(wh)):
        _raise()
    return img


def render_batch(in_dir, out_dir, pos, aim_x, px_deg, fov_h, fov_w, trilinear=True, threads=0):
    p = _f32(pos, (3,))
    failed = np.zeros(1, np.int32)
    ok = lib().mir_render_batch(str(in_dir).encode(), str(out_dir).encode(), px_deg, fov_h, fov_w, _p(p), aim_x,
                                int(trilinear), threads, _p(failed))
    if ok < 0:
        _raise()
    return ok, int(failed[0])


def helper_inverse3(m):
    m = _f32(m, (9,))
    out = np.zeros(9, np.float32)
    if lib().mir_helper_inverse3(_p(m), _p(out)):
        _raise()
    return out.reshape(3, 3)


def read_fits(path):
    """minimal reader for the single-image files Camera::ToFITS writes: (header dict, float32 image)"""
    raw = Path(path).read_bytes()
    hdr = {}
    off = 0
    done = False
    while not done:
        block = raw[off:off + 2880]
        off += 2880
        for c in range(0, 2880, 80):
            card = block[c:c + 80].decode("ascii")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                val = card[10:].split("/")[0].strip()
                hdr[key] = val
    w, h = int(hdr["NAXIS1"]), int(hdr["NAXIS2"])
    data = np.frombuffer(raw[off:off + 4 * w * h], dtype=">f4").astype(np.float32).reshape(h, w)
    return hdr, data
