"""ctypes binding of libbirp_b200.so (include/birp_b200.h).

This is the same C ABI the reference's C++ code would link against;
Python is only used here to drive it from the tests and from bench.py.  There is no
fallback: if the shared library is missing or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("BIRP_B200_LIB", _HERE / "libbirp_b200.so"))  # override: A/B builds only
HEADER_PATH = _HERE.parent / "include" / "birp_b200.h"

BIRP_OK, BIRP_ERR_CUDA, BIRP_ERR_ARG, BIRP_ERR_SINGULAR, BIRP_ERR_RANGE, BIRP_ERR_NOMEM = range(6)
NEAREST, TRILINEAR, ACCUMULATE, CMEM_FAST = 0, 1, 2, 4
DIV_TWO_STEP, DIV_IEEE, GENERAL_LERP, NO_TEXTURE, TEXTURE_ONLY, NO_QUADS, QUADS_ONLY, CMEM_LIBM64, CMEM_NO_FILTER = 8, 16, 32, 64, 128, 256, 512, 1024, 2048
WINDOW_PER_STEP, NO_OCTS, LEAN_MARCH, FAST_MARCH, NO_LEAN_MARCH, CMEM_NO_TABLE, CMEM_LIBM32, DIV_ONE_STEP = 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288

FRAME_DTYPE = np.dtype(
    [("pos", "<f4", (3,)), ("rinv", "<f4", (9,)), ("tan_half_x", "<f4"), ("tan_half_y", "<f4"),
     ("aspect", "<f4"), ("sun_angle", "<f4")]
)
assert FRAME_DTYPE.itemsize == 64

CMEM_FIELDS = [
    "p0", "p1", "p2", "p3", "bs", "A1", "A2", "B", "alpha", "beta", "ay_bs", "az_bs", "r0_lin", "charlie",
    "dn", "ds", "theta_n", "theta_s", "a5", "a21", "beta_c0", "beta_c1", "beta_c2", "beta_c3",
    "pressure_dynamic", "pressure_magnetic", "density",
]
CMEM_DTYPE = np.dtype([(n, "<f4") for n in CMEM_FIELDS])
assert CMEM_DTYPE.itemsize == 27 * 4


class BirpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"birp_b200 error {code}: {message}")
        self.code = code


_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/birp_b200.h declares
SIGNATURES = {
    "birp_version": (C.c_char_p, []),
    "birp_last_error": (C.c_char_p, []),
    "birp_device_count": (C.c_int, [_i32p]),
    "birp_set_device": (C.c_int, [C.c_int]),
    "birp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "birp_host_free": (C.c_int, [_vp]),
    "birp_host_register": (C.c_int, [_vp, C.c_size_t]),
    "birp_host_unregister": (C.c_int, [_vp]),
    "birp_render_cube_rows": (C.c_int, [_vp, _vp] + [C.c_int] * 9 + [_vp]),
    "birp_render_cube_device_rows": (C.c_int, [_vp, _vp] + [C.c_int] * 9 + [_vp, C.c_longlong, C.c_longlong, _vp, _vp]),
    "birp_frame_setup": (C.c_int, [_f32p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp]),
    "birp_frame_setup_aim": (C.c_int, [_f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int, _vp]),
    "birp_frame_setup_batch": (C.c_int, [_vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp,
                                         C.c_int]),
    "birp_frame_setup_device": (C.c_int, [_vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, _vp]),
    "birp_orbit_poses": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, _vp]),
    "birp_orbit_poses_device": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, _vp, _vp]),
    "birp_ray_table": (C.c_int, [C.c_int, C.c_int, _f32p, _f32p]),
    "birp_cmem_init": (C.c_int, [C.c_int, _i32p, C.c_int, _i32p, C.c_float, C.c_float, C.c_int, C.c_int,
                                 C.c_int] + [C.c_float] * 9 + [_vp]),
    "birp_cube_upload": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.POINTER(_vp)]),
    "birp_cube_upload_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _vp, C.POINTER(_vp)]),
    "birp_cube_free": (None, [_vp]),
    "birp_cube_info": (C.c_int, [_vp, _i32p]),
    "birp_cube_has_oct_records": (C.c_int, [_vp, _i32p]),
    "birp_render_cube": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp]),
    "birp_render_cmem": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp]),
    "birp_render_cube_sweep": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float] + [C.c_int] * 7 + [_vp, _vp]),
    "birp_render_cmem_sweep": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float] + [C.c_int] * 7 + [_vp, _vp]),
    "birp_render_cube_sweep_strided": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float] + [C.c_int] * 7 +
                                       [_vp, _vp, C.c_longlong, C.c_longlong]),
    "birp_render_cmem_sweep_strided": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float] + [C.c_int] * 7 +
                                       [_vp, _vp, C.c_longlong, C.c_longlong]),
    "birp_render_cube_sweep_multi": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float] +
                                     [C.c_int] * 7 + [_vp, _vp]),
    "birp_render_cmem_sweep_multi": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_float, C.c_float, C.c_float] +
                                     [C.c_int] * 7 + [_vp, _vp]),
    "birp_cube_replicate": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "birp_cube_device": (C.c_int, [_vp, _i32p]),
    "birp_accumulate": (C.c_int, [_vp, C.c_long, C.c_int, C.c_float, C.c_int, _vp]),
    "birp_ray_directions": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "birp_sample_cube": (C.c_int, [_vp, _vp, C.c_long, C.c_int, _vp]),
    "birp_sample_cmem": (C.c_int, [_vp, _vp, C.c_long, C.c_int, _vp]),
    "birp_render_cube_device": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp, _vp, _vp]),
    "birp_render_cmem_device": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp, _vp, _vp]),
    "birp_render_cube_device_strided": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp, C.c_longlong, C.c_longlong, _vp, _vp]),
    "birp_render_cmem_device_strided": (C.c_int, [_vp, _vp] + [C.c_int] * 8 + [_vp, C.c_longlong, C.c_longlong, _vp, _vp]),
    "birp_device_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "birp_device_free": (C.c_int, [_vp]),
    "birp_ipc_export": (C.c_int, [_vp, _vp]),
    "birp_ipc_import": (C.c_int, [_vp, C.POINTER(_vp)]),
    "birp_ipc_close": (C.c_int, [_vp]),
    "birp_peer_signal": (C.c_int, [_vp, C.c_uint, _vp]),
    "birp_peer_wait": (C.c_int, [_vp, C.c_int, C.c_uint, _vp]),
    "birp_check_device_status": (C.c_int, [_vp]),
    "birp_selftest_cmem_math": (C.c_int, [C.c_long, _vp]),
    "birp_selftest_div10": (C.c_int, [_vp]),
    "birp_selftest_cmem_filter": (C.c_int, [_vp, C.c_long, _vp]),
    "birp_selftest_cmem_table": (C.c_int, [_vp, C.c_long, _vp]),
    "birp_launch_count": (C.c_ulonglong, []),
    "birp_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
}


def lib() -> C.CDLL:
    """Load libbirp_b200.so (built in-tree by `make -C birp_cpp_b200/csrc` / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {_HERE / 'csrc'}` "
                "(birp_cpp_b200 has no CPU or pure-Python path)"
            )
        handle = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(rc: int) -> None:
    if rc != BIRP_OK:
        raise BirpError(rc, lib().birp_last_error().decode(errors="replace"))


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


# ------------------------------------------------------------------ host-side setup
def frame_setup(pos, aim_x: float, fov_x_deg: float, fov_y_deg: float, W: int, H: int) -> np.ndarray:
    """One birp_frame record (FRAME_DTYPE scalar array of shape (1,))."""
    out = np.zeros(1, dtype=FRAME_DTYPE)
    p = _f32(pos, (3,))
    _check(lib().birp_frame_setup(p.ctypes.data_as(_f32p), aim_x, fov_x_deg, fov_y_deg, W, H, _ptr(out)))
    return out


def frame_setup_aim(pos, aim, fov_x_deg: float, fov_y_deg: float, W: int, H: int) -> np.ndarray:
    """One birp_frame record for a camera aimed at the point `aim` (all three components, Camera::SetAim)."""
    out = np.zeros(1, dtype=FRAME_DTYPE)
    p, a = _f32(pos, (3,)), _f32(aim, (3,))
    _check(lib().birp_frame_setup_aim(p.ctypes.data_as(_f32p), a.ctypes.data_as(_f32p), fov_x_deg, fov_y_deg, W, H,
                                      _ptr(out)))
    return out


def frames_for_poses(poses, aim_x: float, fov_x_deg: float, fov_y_deg: float, W: int, H: int, out=None,
                     nthreads: int = 0) -> np.ndarray:
    """birp_frame records of a sweep (birp_frame_setup_batch); `out`: optional preallocated FRAME_DTYPE array"""
    poses = _f32(poses).reshape(-1, 3)
    if out is None:
        out = np.zeros(len(poses), dtype=FRAME_DTYPE)
    assert out.dtype == FRAME_DTYPE and len(out) == len(poses) and out.flags.c_contiguous
    _check(lib().birp_frame_setup_batch(_ptr(poses), len(poses), aim_x, fov_x_deg, fov_y_deg, W, H, _ptr(out),
                                        nthreads))
    return out


def frame_setup_device(d_poses: int, n: int, aim_x: float, fov_x_deg: float, fov_y_deg: float, W: int, H: int,
                       d_frames: int, stream: int = 0) -> None:
    """poses already on the device -> frame records on the device (enqueued on `stream`)"""
    _check(lib().birp_frame_setup_device(_vp(d_poses), n, aim_x, fov_x_deg, fov_y_deg, W, H, _vp(d_frames),
                                         _vp(stream)))


def orbit_poses(n: int, perigee: float = 1.78, apogee: float = 20.0, inclination_deg: float = 73.0) -> np.ndarray:
    out = np.empty((n, 3), dtype=np.float32)
    _check(lib().birp_orbit_poses(n, perigee, apogee, inclination_deg, _ptr(out)))
    return out


def orbit_poses_device(n: int, d_poses: int, perigee: float = 1.78, apogee: float = 20.0,
                       inclination_deg: float = 73.0, stream: int = 0) -> None:
    _check(lib().birp_orbit_poses_device(n, perigee, apogee, inclination_deg, _vp(d_poses), _vp(stream)))


def ray_table(nsamples: int, maxdist: int):
    dist = np.zeros(nsamples, dtype=np.float32)
    w = C.c_float(0)
    _check(lib().birp_ray_table(nsamples, maxdist, dist.ctypes.data_as(_f32p), C.byref(w)))
    return dist, np.float32(w.value)


def cmem_init(v=None, b=None, dipole=0.0, p0=0.786300004, p1=1, p2=3, p3=4, B=2.0, alpha=2.5, beta=-1.6,
              bs=12.6400003, A1=7.2e-06, A2=3.5000000000000004e-06, ay_bs=0.0, az_bs=0.0, density=0.0) -> np.ndarray:
    """CMEM::Init(true, ...) with the reference's default arguments (include/Cmem.h:21-39)."""
    out = np.zeros(1, dtype=CMEM_DTYPE)
    va = (C.c_int * 3)(*(v if v is not None else (0, 0, 0)))
    ba = (C.c_int * 3)(*(b if b is not None else (0, 0, 0)))
    _check(lib().birp_cmem_init(int(v is not None), va, int(b is not None), ba, dipole, p0, int(p1), int(p2),
                                int(p3), B, alpha, beta, bs, A1, A2, ay_bs, az_bs, density, _ptr(out)))
    return out


def cmem_init_cli() -> np.ndarray:
    """The parameters `birp -c` ends up with (reference src/main.cpp:66-78,294): every CMEM flag
    left at NULL/0 and p = {1,1,3,4}."""
    return cmem_init(v=None, b=None, dipole=0.0, p0=1.0, p1=1, p2=3, p3=4, B=0.0, alpha=0.0, beta=0.0, bs=0.0,
                     A1=0.0, A2=0.0, ay_bs=0.0, az_bs=0.0, density=0.0)


# ------------------------------------------------------------------ pinned host memory
class PinnedArray:
    """float32 / structured numpy array backed by cudaHostAlloc memory."""

    def __init__(self, shape, dtype=np.float32):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._p = _vp()
        _check(lib().birp_host_alloc(max(n, 1), C.byref(self._p)))
        buf = (C.c_char * max(n, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().birp_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------ cube
class Cube:
    """Device-resident DataCube grid."""

    def __init__(self, zyx: np.ndarray, origin, spacing):
        zyx = _f32(zyx)
        assert zyx.ndim == 3, "cube must be [nz][ny][nx]"
        self.nz, self.ny, self.nx = zyx.shape
        self.origin = _f32(origin, (3,))
        self.spacing = _f32(spacing, (3,))
        self._h = _vp()
        _check(lib().birp_cube_upload(_ptr(zyx), self.nx, self.ny, self.nz, self.origin.ctypes.data_as(_f32p),
                                      self.spacing.ctypes.data_as(_f32p), C.byref(self._h)))

    @classmethod
    def from_device(cls, d_ptr: int, nx: int, ny: int, nz: int, origin, spacing, stream: int = 0) -> "Cube":
        self = cls.__new__(cls)
        self.nx, self.ny, self.nz = nx, ny, nz
        self.origin = _f32(origin, (3,))
        self.spacing = _f32(spacing, (3,))
        self._h = _vp()
        _check(lib().birp_cube_upload_device(_vp(d_ptr), nx, ny, nz, self.origin.ctypes.data_as(_f32p),
                                             self.spacing.ctypes.data_as(_f32p), _vp(stream), C.byref(self._h)))
        return self

    @property
    def handle(self):
        return self._h

    def replicate(self, device: int) -> "Cube":
        """the same grid on another GPU of the box (birp_cube_replicate)"""
        other = Cube.__new__(Cube)
        other.nx, other.ny, other.nz = self.nx, self.ny, self.nz
        other.origin, other.spacing = self.origin, self.spacing
        other._h = _vp()
        _check(lib().birp_cube_replicate(self._h, int(device), C.byref(other._h)))
        return other

    @property
    def device(self) -> int:
        d = (C.c_int * 1)()
        _check(lib().birp_cube_device(self._h, d))
        return int(d[0])

    def free(self):
        if self._h:
            lib().birp_cube_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def info(self) -> dict:
        out = (C.c_int * 4)()
        _check(lib().birp_cube_info(self._h, out))
        oct_ = (C.c_int * 1)()
        _check(lib().birp_cube_has_oct_records(self._h, oct_))
        return {"divmode": int(out[0]), "positive": out[1] == 1, "nonnegative": out[1] in (1, 2),
                "texture": bool(out[2]), "quads": bool(out[3]), "octs": bool(oct_[0])}

    def sample(self, xyz, trilinear: bool) -> np.ndarray:
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.empty(len(xyz), dtype=np.float32)
        _check(lib().birp_sample_cube(self._h, _ptr(xyz), len(xyz), TRILINEAR if trilinear else NEAREST, _ptr(out)))
        return out

    def render(self, frames: np.ndarray, W: int, H: int, nsamples: int = 200, maxdist: int = 44,
               trilinear: bool = True, rows=(0, 0), images: np.ndarray | None = None,
               out: np.ndarray | None = None, extra_flags: int = 0) -> np.ndarray:
        """Host-buffer render.  `images` given => accumulated into (BIRP_ACCUMULATE);
        `out` given => overwritten (rows outside the band untouched); neither => new zeros."""
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        flags = (TRILINEAR if trilinear else NEAREST) | extra_flags
        if images is not None:
            flags |= ACCUMULATE
        elif out is not None:
            images = out
        else:
            images = np.zeros((len(frames), H, W), dtype=np.float32)
        assert images.dtype == np.float32 and images.flags.c_contiguous
        assert images.size == len(frames) * H * W
        _check(lib().birp_render_cube(self._h, _ptr(frames), len(frames), W, H, rows[0], rows[1], nsamples, maxdist,
                                      flags, _ptr(images)))
        return images


def render_cube_rows(cube: "Cube", frames: np.ndarray, W: int, H: int, band_rows: int, phase: int, nphases: int, *,
                     nsamples: int = 200, maxdist: int = 44, trilinear: bool = True, out: np.ndarray | None = None,
                     extra_flags: int = 0) -> np.ndarray:
    """interleaved row bands of every frame (host buffers): bands with index = phase (mod nphases) are rendered"""
    frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
    images = out if out is not None else np.zeros((len(frames), H, W), dtype=np.float32)
    assert images.dtype == np.float32 and images.flags.c_contiguous and images.size == len(frames) * H * W
    flags = (TRILINEAR if trilinear else NEAREST) | extra_flags
    _check(lib().birp_render_cube_rows(cube.handle, _ptr(frames), len(frames), W, H, band_rows, phase, nphases, nsamples,
                                       maxdist, flags, _ptr(images)))
    return images


def render_cube_device_rows(cube: "Cube", d_frames: int, nframes: int, W: int, H: int, d_images: int, band_rows: int,
                            phase: int, nphases: int, *, nsamples=200, maxdist=44, trilinear=True, d_valid: int = 0,
                            stream: int = 0, extra_flags: int = 0, frame_offset: int = 0, frame_stride: int = 1) -> None:
    flags = (TRILINEAR if trilinear else NEAREST) | extra_flags
    _check(lib().birp_render_cube_device_rows(cube.handle, _vp(d_frames), nframes, W, H, band_rows, phase, nphases,
                                              nsamples, maxdist, flags, _vp(d_images), frame_offset, frame_stride,
                                              _vp(d_valid), _vp(stream)))


def host_register(array: np.ndarray) -> None:
    _check(lib().birp_host_register(_ptr(array), array.nbytes))


def host_unregister(array: np.ndarray) -> None:
    _check(lib().birp_host_unregister(_ptr(array)))


def render_sweep(field, poses, aim_x: float, fov_x_deg: float, fov_y_deg: float, W: int, H: int, *, nsamples: int = 200,
                 maxdist: int = 44, trilinear: bool = True, fast=False, rows=(0, 0), extra_flags: int = 0,
                 frames_out: np.ndarray | None = None, out: np.ndarray | None = None, frame_offset: int = 0,
                 frame_stride: int = 1) -> np.ndarray:
    """poses (host) -> images (host) in one call; `field` is a Cube or a CMEM parameter record.  frame_offset /
    frame_stride: frame f of the call lands in slot offset + f * stride of `out` (which then holds the whole sweep)"""
    poses = _f32(poses).reshape(-1, 3)
    n = len(poses)
    images = out if out is not None else np.zeros((n, H, W), dtype=np.float32)
    assert images.dtype == np.float32 and images.flags.c_contiguous
    assert images.size >= (frame_offset + max(n - 1, 0) * frame_stride + (1 if n else 0)) * H * W
    fr = _vp(0)
    if frames_out is not None:
        assert frames_out.dtype == FRAME_DTYPE and len(frames_out) == n and frames_out.flags.c_contiguous
        fr = _ptr(frames_out)
    if isinstance(field, Cube):
        flags = (TRILINEAR if trilinear else NEAREST) | extra_flags
        _check(lib().birp_render_cube_sweep_strided(field.handle, _ptr(poses), n, aim_x, fov_x_deg, fov_y_deg, W, H,
                                                    rows[0], rows[1], nsamples, maxdist, flags, fr, _ptr(images),
                                                    frame_offset, frame_stride))
    else:
        params = np.ascontiguousarray(field, dtype=CMEM_DTYPE)
        _check(lib().birp_render_cmem_sweep_strided(_ptr(params), _ptr(poses), n, aim_x, fov_x_deg, fov_y_deg, W, H,
                                                    rows[0], rows[1], nsamples, maxdist,
                                                    _cmem_mode(fast) | extra_flags, fr, _ptr(images), frame_offset,
                                                    frame_stride))
    return images


def render_sweep_multi(fields, poses, aim_x: float, fov_x_deg: float, fov_y_deg: float, W: int, H: int, *,
                       nsamples: int = 200, maxdist: int = 44, trilinear: bool = True, fast=False, rows=(0, 0),
                       extra_flags: int = 0, frames_out: np.ndarray | None = None, out: np.ndarray | None = None,
                       devices=None) -> np.ndarray:
    """one sweep over several GPUs from this process: `fields` is a list of Cubes (one per device) or a CMEM parameter
    record with `devices`"""
    poses = _f32(poses).reshape(-1, 3)
    n = len(poses)
    images = out if out is not None else np.zeros((n, H, W), dtype=np.float32)
    assert images.dtype == np.float32 and images.flags.c_contiguous and images.size == n * H * W
    fr = _vp(0)
    if frames_out is not None:
        assert frames_out.dtype == FRAME_DTYPE and len(frames_out) == n and frames_out.flags.c_contiguous
        fr = _ptr(frames_out)
    if isinstance(fields, (list, tuple)):
        handles = (_vp * len(fields))(*[f.handle for f in fields])
        flags = (TRILINEAR if trilinear else NEAREST) | extra_flags
        _check(lib().birp_render_cube_sweep_multi(handles, len(fields), _ptr(poses), n, aim_x, fov_x_deg, fov_y_deg, W, H,
                                                  rows[0], rows[1], nsamples, maxdist, flags, fr, _ptr(images)))
    else:
        params = np.ascontiguousarray(fields, dtype=CMEM_DTYPE)
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        _check(lib().birp_render_cmem_sweep_multi(_ptr(params), _ptr(devs), len(devs), _ptr(poses), n, aim_x, fov_x_deg,
                                                  fov_y_deg, W, H, rows[0], rows[1], nsamples, maxdist,
                                                  _cmem_mode(fast) | extra_flags, fr, _ptr(images)))
    return images


def _cmem_mode(fast) -> int:
    """fast: False = default (own fp64 blocks behind the region table and filter), True = tolerance mode (fp32 values
    behind the same decisions), "libm64" = CUDA fp64 library, "nofilter" = own fp64 blocks for every sample,
    "notable" = default without the region table, "libm32" = CUDA fp32 library for every sample"""
    if isinstance(fast, int) and not isinstance(fast, bool):
        return fast  # raw BIRP_CMEM_* flags
    if fast == "libm64":
        return CMEM_LIBM64
    if fast == "nofilter":
        return CMEM_NO_FILTER
    if fast == "notable":
        return CMEM_NO_TABLE
    if fast == "libm32":
        return CMEM_LIBM32
    return CMEM_FAST if fast else 0


def render_cmem(params: np.ndarray, frames: np.ndarray, W: int, H: int, nsamples: int = 200, maxdist: int = 44,
                fast: bool = False, rows=(0, 0), images: np.ndarray | None = None,
                out: np.ndarray | None = None) -> np.ndarray:
    frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
    params = np.ascontiguousarray(params, dtype=CMEM_DTYPE)
    flags = _cmem_mode(fast)
    if images is not None:
        flags |= ACCUMULATE
    elif out is not None:
        images = out
    else:
        images = np.zeros((len(frames), H, W), dtype=np.float32)
    assert images.dtype == np.float32 and images.flags.c_contiguous
    assert images.size == len(frames) * H * W
    _check(lib().birp_render_cmem(_ptr(params), _ptr(frames), len(frames), W, H, rows[0], rows[1], nsamples,
                                  maxdist, flags, _ptr(images)))
    return images


def sample_cmem(params: np.ndarray, xyz, fast: bool = False) -> np.ndarray:
    params = np.ascontiguousarray(params, dtype=CMEM_DTYPE)
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty(len(xyz), dtype=np.float32)
    _check(lib().birp_sample_cmem(_ptr(params), _ptr(xyz), len(xyz), _cmem_mode(fast), _ptr(out)))
    return out


def ray_directions(frame: np.ndarray, W: int, H: int) -> np.ndarray:
    frame = np.ascontiguousarray(frame, dtype=FRAME_DTYPE).reshape(-1)[:1]
    out = np.empty((H, W, 3), dtype=np.float32)
    _check(lib().birp_ray_directions(_ptr(frame), W, H, _ptr(out)))
    return out


def accumulate(samples, ray_width: float, image: np.ndarray | None = None) -> np.ndarray:
    samples = _f32(samples)
    assert samples.ndim == 2
    nrays, nsamples = samples.shape
    flags = 0
    if image is None:
        image = np.zeros(nrays, dtype=np.float32)
    else:
        assert image.dtype == np.float32 and image.flags.c_contiguous and image.size == nrays
        flags = ACCUMULATE
    _check(lib().birp_accumulate(_ptr(samples), nrays, nsamples, float(ray_width), flags, _ptr(image)))
    return image


# ------------------------------------------------------------------ device-buffer entry points
def render_cube_device(cube: Cube, d_frames: int, nframes: int, W: int, H: int, d_images: int, *, nsamples=200,
                       maxdist=44, trilinear=True, rows=(0, 0), accumulate_into=False, d_valid: int = 0,
                       stream: int = 0, extra_flags: int = 0, frame_offset: int = 0, frame_stride: int = 1) -> None:
    flags = (TRILINEAR if trilinear else NEAREST) | (ACCUMULATE if accumulate_into else 0) | extra_flags
    _check(lib().birp_render_cube_device_strided(cube.handle, _vp(d_frames), nframes, W, H, rows[0], rows[1],
                                                 nsamples, maxdist, flags, _vp(d_images), frame_offset, frame_stride,
                                                 _vp(d_valid), _vp(stream)))


def render_cmem_device(params: np.ndarray, d_frames: int, nframes: int, W: int, H: int, d_images: int, *,
                       nsamples=200, maxdist=44, fast=False, rows=(0, 0), accumulate_into=False, d_valid: int = 0,
                       stream: int = 0, frame_offset: int = 0, frame_stride: int = 1) -> None:
    params = np.ascontiguousarray(params, dtype=CMEM_DTYPE)
    flags = _cmem_mode(fast) | (ACCUMULATE if accumulate_into else 0)
    _check(lib().birp_render_cmem_device_strided(_ptr(params), _vp(d_frames), nframes, W, H, rows[0], rows[1],
                                                 nsamples, maxdist, flags, _vp(d_images), frame_offset, frame_stride,
                                                 _vp(d_valid), _vp(stream)))


class DeviceBuffer:
    """cudaMalloc'ed float32 buffer that other processes of the box can map (CUDA IPC)"""

    def __init__(self, nfloats: int):
        self.nfloats = int(nfloats)
        self._p = _vp()
        self._imported = False
        _check(lib().birp_device_alloc(self.nfloats * 4, C.byref(self._p)))

    @classmethod
    def from_handle(cls, handle: bytes, nfloats: int) -> "DeviceBuffer":
        self = cls.__new__(cls)
        self.nfloats = int(nfloats)
        self._p = _vp()
        self._imported = True
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _check(lib().birp_ipc_import(buf, C.byref(self._p)))
        return self

    @property
    def ptr(self) -> int:
        return int(self._p.value)

    def export_handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        _check(lib().birp_ipc_export(self._p, buf))
        return bytes(buf)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nfloats,), "typestr": "<f4", "data": (self.ptr, False), "version": 2}

    def free(self):
        if self._p:
            (lib().birp_ipc_close if self._imported else lib().birp_device_free)(self._p)
            self._p = _vp()


def peer_signal(d_flag: int, value: int, stream: int = 0) -> None:
    _check(lib().birp_peer_signal(_vp(d_flag), int(value) & 0xffffffff, _vp(stream)))


def peer_wait(d_flags: int, n: int, value: int, stream: int = 0) -> None:
    _check(lib().birp_peer_wait(_vp(d_flags), int(n), int(value) & 0xffffffff, _vp(stream)))


def check_device_status(stream: int = 0) -> None:
    _check(lib().birp_check_device_status(_vp(stream)))


def selftest_cmem_math(n: int = 1 << 22) -> dict:
    out = np.zeros(6, dtype=np.float64)
    _check(lib().birp_selftest_cmem_math(n, _ptr(out)))
    return dict(zip(("exp2", "log2", "pow", "exp", "acos", "angle"), out.tolist()))


def selftest_div10() -> tuple[int, int]:
    """(patterns where the kernel's r / 10 differs from the IEEE division, patterns where the bare multiplication form would)"""
    out = np.zeros(2, dtype=np.uint64)
    _check(lib().birp_selftest_div10(_ptr(out)))
    return int(out[0]), int(out[1])


def selftest_cmem_filter(params: np.ndarray, n: int = 1 << 24) -> dict:
    params = np.ascontiguousarray(params, dtype=CMEM_DTYPE)
    out = np.zeros(3, dtype=np.float64)
    _check(lib().birp_selftest_cmem_filter(_ptr(params), n, _ptr(out)))
    return dict(zip(("margin_used_mp", "margin_used_bs", "trusted_fraction"), out.tolist()))


def selftest_cmem_table(params: np.ndarray, n: int = 1 << 24) -> dict:
    params = np.ascontiguousarray(params, dtype=CMEM_DTYPE)
    out = np.zeros(3, dtype=np.float64)
    _check(lib().birp_selftest_cmem_table(_ptr(params), n, _ptr(out)))
    return dict(zip(("differing_samples", "decided_fraction", "table_in_use"), out.tolist()))


def launch_count() -> int:
    return int(lib().birp_launch_count())


def set_tuning(name: str, value: int) -> None:
    """diagnostic launch-shape knobs (birp_set_tuning): tile_w_log2, tile_swizzle, fast_variant, fast_kstep"""
    _check(lib().birp_set_tuning(name.encode(), int(value)))


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().birp_device_count(C.byref(n)))
    return n.value


def set_device(ordinal: int) -> None:
    _check(lib().birp_set_device(ordinal))
