"""TEST INFRASTRUCTURE -- ctypes access to the two CPU checkers.

  ref()    oracle/_ref/libbirp_ref.so   the UNMODIFIED reference sources + ref_harness.cpp
  orc()    oracle/libbirp_oracle.so     the CPU restatement (birp_oracle.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  Nothing in birp_cpp_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libbirp_ref.so"
ORC_SO = HERE / "libbirp_oracle.so"
REFERENCE_ROOT = Path("/root/reference")

_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p
FRAME_FLOATS = 16
CMEM_FLOATS = 27


def build(verbose: bool = False) -> None:
    """Compile the restatement, and oracle/_ref when the reference tree is present."""
    out = subprocess.run(["make", "-C", str(HERE), "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


_orc = None
_ref = None


def have_ref() -> bool:
    return REF_SO.exists()


def orc() -> C.CDLL:
    global _orc
    if _orc is None:
        if not ORC_SO.exists():
            build()
        L = C.CDLL(str(ORC_SO))
        L.orc_frame_setup.restype = C.c_int
        L.orc_frame_setup.argtypes = [_vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp]
        L.orc_orient.restype = C.c_float
        L.orc_orient.argtypes = [_vp, C.c_float]
        L.orc_ray_table.restype = None
        L.orc_ray_table.argtypes = [C.c_int, C.c_int, _vp, _vp]
        L.orc_trace_points.restype = None
        L.orc_trace_points.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
        L.orc_render_cube.restype = C.c_int
        L.orc_render_cube.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp] + [C.c_int] * 7 + [
            _vp, C.c_int, _vp]
        L.orc_render_cmem.restype = C.c_int
        L.orc_render_cmem.argtypes = [_vp, _vp] + [C.c_int] * 7 + [_vp, C.c_int, _vp]
        L.orc_cube_sample.restype = C.c_int
        L.orc_cube_sample.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.c_long, _vp]
        L.orc_cmem_init.restype = None
        L.orc_cmem_init.argtypes = [C.c_int, _vp, C.c_int, _vp, C.c_float, C.c_float, C.c_int, C.c_int,
                                    C.c_int] + [C.c_float] * 9 + [_vp]
        L.orc_cmem_sample.restype = None
        L.orc_cmem_sample.argtypes = [_vp, _vp, C.c_long, _vp]
        L.orc_accumulate.restype = None
        L.orc_accumulate.argtypes = [_vp, C.c_long, C.c_int, C.c_float, _vp]
        _orc = L
    return _orc


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            if REFERENCE_ROOT.exists():
                build()
            else:
                raise FileNotFoundError(f"{REF_SO} missing and {REFERENCE_ROOT} absent")
        L = C.CDLL(str(REF_SO))
        L.ref_last_error.restype = C.c_char_p
        L.ref_cube_create.restype = _vp
        L.ref_cube_create.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]
        L.ref_cube_load.restype = _vp
        L.ref_cube_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_cube_set_trilinear.restype = None
        L.ref_cube_set_trilinear.argtypes = [_vp, C.c_int]
        L.ref_cube_info.restype = None
        L.ref_cube_info.argtypes = [_vp, _vp, _vp]
        L.ref_cube_export.restype = C.c_int
        L.ref_cube_export.argtypes = [_vp, _vp]
        L.ref_cube_free.restype = None
        L.ref_cube_free.argtypes = [_vp]
        L.ref_cmem_create_default.restype = _vp
        L.ref_cmem_create_default.argtypes = []
        L.ref_cmem_create.restype = _vp
        L.ref_cmem_create.argtypes = [C.c_int, _vp, C.c_int, _vp, C.c_float, C.c_float, C.c_int, C.c_int,
                                      C.c_int] + [C.c_float] * 9
        L.ref_cmem_constants.restype = None
        L.ref_cmem_constants.argtypes = [_vp, _vp]
        L.ref_cmem_free.restype = None
        L.ref_cmem_free.argtypes = [_vp]
        L.ref_space_sample.restype = C.c_int
        L.ref_space_sample.argtypes = [_vp, C.c_int, _vp, C.c_long, _vp]
        L.ref_render.restype = C.c_int
        L.ref_render.argtypes = [_vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_float,
                                 C.c_int, C.c_int, _vp, _vp, _vp]
        L.ref_orient.restype = C.c_int
        L.ref_orient.argtypes = [_vp, C.c_float, _vp]
        L.ref_trace_points.restype = C.c_int
        L.ref_trace_points.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_float, C.c_int,
                                       C.c_int, _vp]
        _ref = L
    return _ref


# ======================================================================== restatement wrappers
def orc_frame(pos, aim_x, fov_x, fov_y, W, H) -> np.ndarray:
    fr = np.zeros(FRAME_FLOATS, dtype=np.float32)
    p = _f32(pos, (3,))
    rc = orc().orc_frame_setup(_p(p), aim_x, fov_x, fov_y, W, H, _p(fr))
    if rc:
        raise RuntimeError("Determinant is 0")
    return fr


def orc_frames(poses, aim_x, fov_x, fov_y, W, H) -> np.ndarray:
    poses = _f32(poses).reshape(-1, 3)
    return np.stack([orc_frame(p, aim_x, fov_x, fov_y, W, H) for p in poses])


def orc_orient(pos, aim_x) -> np.float32:
    p = _f32(pos, (3,))
    return np.float32(orc().orc_orient(_p(p), aim_x))


def orc_ray_table(nsamples, maxdist):
    d = np.zeros(nsamples, dtype=np.float32)
    w = np.zeros(1, dtype=np.float32)
    orc().orc_ray_table(nsamples, maxdist, _p(d), _p(w))
    return d, w[0]


def orc_trace_points(frame, W, H, nsamples, maxdist) -> np.ndarray:
    out = np.zeros((H, W, nsamples, 3), dtype=np.float32)
    fr = _f32(frame, (FRAME_FLOATS,))
    orc().orc_trace_points(_p(fr), W, H, nsamples, maxdist, _p(out))
    return out


def orc_render_cube(zyx, origin, spacing, trilinear, frames, W, H, nsamples=200, maxdist=44, rows=(0, -1),
                    images=None, nthreads=8, return_valid=False):
    zyx = _f32(zyx)
    nz, ny, nx = zyx.shape
    frames = _f32(frames).reshape(-1, FRAME_FLOATS)
    if images is None:
        images = np.zeros((len(frames), H, W), dtype=np.float32)
    valid = C.c_ulonglong(0)
    o, s = _f32(origin, (3,)), _f32(spacing, (3,))
    rc = orc().orc_render_cube(_p(zyx), nx, ny, nz, _p(o), _p(s), int(trilinear), _p(frames), len(frames), W, H,
                               rows[0], rows[1], nsamples, maxdist, _p(images), nthreads, C.byref(valid))
    if rc:
        raise IndexError("reference would throw std::out_of_range (NaN sample coordinate)")
    return (images, valid.value) if return_valid else images


def orc_render_cmem(params, frames, W, H, nsamples=200, maxdist=44, rows=(0, -1), images=None, nthreads=8,
                    return_valid=False):
    params = _f32(params, (CMEM_FLOATS,))
    frames = _f32(frames).reshape(-1, FRAME_FLOATS)
    if images is None:
        images = np.zeros((len(frames), H, W), dtype=np.float32)
    valid = C.c_ulonglong(0)
    orc().orc_render_cmem(_p(params), _p(frames), len(frames), W, H, rows[0], rows[1], nsamples, maxdist,
                          _p(images), nthreads, C.byref(valid))
    return (images, valid.value) if return_valid else images


def orc_cube_sample(zyx, origin, spacing, trilinear, xyz) -> np.ndarray:
    zyx = _f32(zyx)
    nz, ny, nx = zyx.shape
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), dtype=np.float32)
    o, s = _f32(origin, (3,)), _f32(spacing, (3,))
    rc = orc().orc_cube_sample(_p(zyx), nx, ny, nz, _p(o), _p(s), int(trilinear), _p(xyz), len(xyz), _p(out))
    if rc:
        raise IndexError("reference would throw std::out_of_range (NaN sample coordinate)")
    return out


def _cmem_args(v, b, dipole, p0, p1, p2, p3, B, alpha, beta, bs, A1, A2, ay_bs, az_bs, density):
    va = np.array(v if v is not None else (0, 0, 0), dtype=np.int32)
    ba = np.array(b if b is not None else (0, 0, 0), dtype=np.int32)
    return (int(v is not None), va, int(b is not None), ba, dipole, p0, int(p1), int(p2), int(p3), B, alpha, beta, bs,
            A1, A2, ay_bs, az_bs, density)


CMEM_DEFAULTS = dict(v=None, b=None, dipole=0.0, p0=0.786300004, p1=1, p2=3, p3=4, B=2.0, alpha=2.5, beta=-1.6,
                     bs=12.6400003, A1=7.2e-06, A2=3.5000000000000004e-06, ay_bs=0.0, az_bs=0.0, density=0.0)
# what `birp -c` passes (reference src/main.cpp:66-78,294)
CMEM_CLI = dict(v=None, b=None, dipole=0.0, p0=1.0, p1=1, p2=3, p3=4, B=0.0, alpha=0.0, beta=0.0, bs=0.0, A1=0.0,
                A2=0.0, ay_bs=0.0, az_bs=0.0, density=0.0)
# what tests/CMEM_test.cpp:48-63 actually passes (its positional arguments are shifted by one)
CMEM_TEST2 = dict(CMEM_DEFAULTS, v=(400, 0, 0), b=(0, 0, 4), dipole=0.0, p0=1.0, p1=4, p2=6, p3=3, B=3.6, alpha=-1.0)


def orc_cmem_init(**kw) -> np.ndarray:
    a = dict(CMEM_DEFAULTS)
    a.update(kw)
    hv, va, hb, ba, *rest = _cmem_args(**a)
    out = np.zeros(CMEM_FLOATS, dtype=np.float32)
    orc().orc_cmem_init(hv, _p(va), hb, _p(ba), *rest, _p(out))
    return out


def orc_cmem_sample(params, xyz) -> np.ndarray:
    params = _f32(params, (CMEM_FLOATS,))
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), dtype=np.float32)
    orc().orc_cmem_sample(_p(params), _p(xyz), len(xyz), _p(out))
    return out


def orc_accumulate(samples, ray_width, image=None) -> np.ndarray:
    samples = _f32(samples)
    nrays, nsamples = samples.shape
    if image is None:
        image = np.zeros(nrays, dtype=np.float32)
    orc().orc_accumulate(_p(samples), nrays, nsamples, float(ray_width), _p(image))
    return image


# ======================================================================== unmodified-reference wrappers
class RefCube:
    def __init__(self, zyx=None, origin=None, spacing=None, trilinear=False, path=None):
        if path is not None:
            self.h = ref().ref_cube_load(str(path).encode(), int(trilinear))
            if not self.h:
                raise RuntimeError(ref().ref_last_error().decode())
        else:
            zyx = _f32(zyx)
            nz, ny, nx = zyx.shape
            o, s = _f32(origin, (3,)), _f32(spacing, (3,))
            self.h = ref().ref_cube_create(_p(zyx), nx, ny, nz, _p(o), _p(s), int(trilinear))
        self.is_cmem = 0

    def set_trilinear(self, on: bool):
        ref().ref_cube_set_trilinear(self.h, int(on))

    def info(self):
        dims = np.zeros(4, dtype=np.int32)
        geom = np.zeros(6, dtype=np.float32)
        ref().ref_cube_info(self.h, _p(dims), _p(geom))
        return dims, geom

    def export(self) -> np.ndarray:
        dims, _ = self.info()
        out = np.zeros((dims[2], dims[1], dims[0]), dtype=np.float32)
        if ref().ref_cube_export(self.h, _p(out)):
            raise RuntimeError(ref().ref_last_error().decode())
        return out

    def free(self):
        if self.h:
            ref().ref_cube_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class RefCmem:
    def __init__(self, default_init: bool = True, **kw):
        if default_init and not kw:
            self.h = ref().ref_cmem_create_default()
        else:
            a = dict(CMEM_DEFAULTS)
            a.update(kw)
            hv, va, hb, ba, *rest = _cmem_args(**a)
            self.h = ref().ref_cmem_create(hv, _p(va), hb, _p(ba), *rest)
        self.is_cmem = 1

    def constants(self) -> np.ndarray:
        out = np.zeros(CMEM_FLOATS, dtype=np.float32)
        ref().ref_cmem_constants(self.h, _p(out))
        return out

    def free(self):
        if self.h:
            ref().ref_cmem_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ref_sample(space, xyz) -> np.ndarray:
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), dtype=np.float32)
    if ref().ref_space_sample(space.h, space.is_cmem, _p(xyz), len(xyz), _p(out)):
        raise IndexError(ref().ref_last_error().decode())
    return out


def ref_render(space, pos, aim_x, px_deg, fov_h, fov_w, W=0, H=0, nsamples=200, maxdist=44, image=None,
               return_seconds=False):
    """One frame through the unmodified reference Camera.  W=H=0 keeps round(fov/px)."""
    wh = np.zeros(2, dtype=np.int32)
    p = _f32(pos, (3,))
    if ref().ref_render(space.h, space.is_cmem, px_deg, fov_h, fov_w, W, H, _p(p), aim_x, nsamples, maxdist, None,
                        _p(wh), None):
        raise RuntimeError(ref().ref_last_error().decode())
    w, h = int(wh[0]), int(wh[1])
    if image is None:
        image = np.zeros((h, w), dtype=np.float32)
    secs = C.c_double(0)
    if ref().ref_render(space.h, space.is_cmem, px_deg, fov_h, fov_w, W, H, _p(p), aim_x, nsamples, maxdist,
                        _p(image), _p(wh), C.byref(secs)):
        raise RuntimeError(ref().ref_last_error().decode())
    return (image, secs.value) if return_seconds else image


def ref_render_aim(space, pos, aim, px_deg, fov_h, fov_w, W=0, H=0, nsamples=200, maxdist=44):
    """One frame through the unmodified reference Camera aimed at the point `aim` (SetAim(x, y, z))."""
    L = ref()
    L.ref_render3.restype = C.c_int
    L.ref_render3.argtypes = [_vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, _vp, C.c_int,
                              C.c_int, _vp, _vp, _vp]
    wh = np.zeros(2, dtype=np.int32)
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    if L.ref_render3(space.h, space.is_cmem, px_deg, fov_h, fov_w, W, H, _p(p), _p(a), nsamples, maxdist, None, _p(wh),
                     None):
        raise RuntimeError(L.ref_last_error().decode())
    image = np.zeros((int(wh[1]), int(wh[0])), dtype=np.float32)
    if L.ref_render3(space.h, space.is_cmem, px_deg, fov_h, fov_w, W, H, _p(p), _p(a), nsamples, maxdist, _p(image),
                     _p(wh), None):
        raise RuntimeError(L.ref_last_error().decode())
    return image


def ref_orient_aim(pos, aim) -> np.float32:
    L = ref()
    L.ref_orient3.restype = C.c_int
    L.ref_orient3.argtypes = [_vp, _vp, _vp]
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    out = np.zeros(1, dtype=np.float32)
    if L.ref_orient3(_p(p), _p(a), _p(out)):
        raise RuntimeError(L.ref_last_error().decode())
    return out[0]


def orc_frame_aim(pos, aim, fov_x, fov_y, W, H) -> np.ndarray:
    """the restatement's frame record (16 floats) for a camera aimed at the point `aim`"""
    L = orc()
    L.orc_frame_setup3.restype = C.c_int
    L.orc_frame_setup3.argtypes = [_vp, _vp, C.c_float, C.c_float, C.c_int, C.c_int, _vp]
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    out = np.zeros(FRAME_FLOATS, dtype=np.float32)
    if L.orc_frame_setup3(_p(p), _p(a), fov_x, fov_y, W, H, _p(out)):
        raise ZeroDivisionError("Determinant is 0")
    return out


def ref_orient(pos, aim_x) -> np.float32:
    p = _f32(pos, (3,))
    out = np.zeros(1, dtype=np.float32)
    if ref().ref_orient(_p(p), aim_x, _p(out)):
        raise RuntimeError(ref().ref_last_error().decode())
    return out[0]


def ref_trace_points(pos, aim_x, px_deg, fov_h, fov_w, W, H, nsamples, maxdist) -> np.ndarray:
    out = np.zeros((H, W, nsamples, 3), dtype=np.float32)
    p = _f32(pos, (3,))
    if ref().ref_trace_points(px_deg, fov_h, fov_w, W, H, _p(p), aim_x, nsamples, maxdist, _p(out)):
        raise RuntimeError(ref().ref_last_error().decode())
    return out
