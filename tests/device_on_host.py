"""ctypes access to tests/cpp/libdevice_on_host.so: the DEVICE code of the march kernels
(birp_cpp_b200/csrc/birp_device.cuh) compiled for the host, so that the source the GPU runs can be checked
against the oracle without a GPU.  Test infrastructure; see tests/cpp/device_on_host.cpp."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "cpp" / "libdevice_on_host.so"
_vp = C.c_void_p
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", str(HERE.parent / "birp_cpp_b200" / "csrc")], check=True, capture_output=True)
        subprocess.run(["make", "-C", str(HERE / "cpp")], check=True, capture_output=True)
        L = C.CDLL(str(SO))
        L.doh_cube_create.restype = _vp
        L.doh_cube_create.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]
        L.doh_cube_free.argtypes = [_vp]
        L.doh_cube_class.argtypes = [_vp]
        L.doh_cube_window.argtypes = [_vp, _vp, _vp]
        L.doh_verify_div.argtypes = [C.c_float, C.c_int, C.c_uint, _vp]
        L.doh_sample_cube.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_long, _vp, _vp, _vp]
        L.doh_render_cube.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, _vp, _vp]
        L.doh_render_cube_fast.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           _vp, _vp]
        L.doh_scale.argtypes = [_vp, C.c_long, C.c_float, C.c_int, _vp, _vp, _vp, _vp]
        L.doh_lerps.argtypes = [_vp, _vp, _vp, C.c_long, _vp, _vp, _vp]
        L.doh_scale_exhaustive.argtypes = [C.c_int, C.c_uint, C.c_int, _vp]
        L.doh_sample_cmem.argtypes = [_vp, _vp, C.c_long, C.c_int, _vp]
        L.doh_div10_check.argtypes = [C.c_uint, C.c_uint, C.c_uint, _vp]
        L.doh_cmem_filter_decides.argtypes = [_vp, _vp, C.c_long, _vp]
        L.doh_cmem_table.argtypes = [_vp, _vp]
        L.doh_sample_cmem_table.argtypes = [_vp, _vp, _vp, C.c_long, _vp, _vp]
        L.doh_render_cmem.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
        _lib = L
    return _lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return a.ctypes.data_as(_vp)


def verify_div(spacing: float, n: int, stride: int = 1, with_multiply: bool = False):
    """(one-step division exact, two-step division exact[, the multiply by the reciprocal alone exact]) over every
    `stride`-th ingress of the hot window"""
    bad = (C.c_int * 3)()
    lib().doh_verify_div(float(spacing), int(n), int(stride), bad)
    return (bad[0] == 0, bad[1] == 0, bad[2] == 0) if with_multiply else (bad[0] == 0, bad[1] == 0)


class HostCube:
    """the device view of a cube (CubeView) with its quad / oct copies, built on the host"""

    def __init__(self, zyx, origin, spacing, divmode=2):
        self.zyx = _f32(zyx)
        nz, ny, nx = self.zyx.shape
        self._o, self._s = _f32(origin, (3,)), _f32(spacing, (3,))
        self._h = lib().doh_cube_create(_p(self.zyx), nx, ny, nz, _p(self._o), _p(self._s), int(divmode))

    def free(self):
        if self._h:
            lib().doh_cube_free(self._h)
            self._h = None

    def __del__(self):
        self.free()

    @property
    def lerp_class(self) -> int:
        return lib().doh_cube_class(self._h)

    def window(self):
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib().doh_cube_window(self._h, _p(lo), _p(hi))
        return lo, hi

    def sample(self, xyz, trilinear=True, lerp=-1, fetch=0):
        """(cube_sample, cube_sample_hot, hot mask, status)"""
        xyz = _f32(xyz).reshape(-1, 3)
        n = len(xyz)
        literal, hot = np.zeros(n, np.float32), np.zeros(n, np.float32)
        mask = np.zeros(n, np.uint8)
        st = lib().doh_sample_cube(self._h, int(trilinear), int(lerp), int(fetch), _p(xyz), n, _p(literal), _p(hot),
                                   _p(mask))
        return literal, hot, mask.astype(bool), st

    def render(self, frames, W, H, trilinear=True, lerp=-1, divmode=-1, fetch=5, span=True, nsamples=200, maxdist=44,
               accumulate_into=None):
        """march_cube_kernel on the host; frames: birp_frame records (structured array or (n, 16) floats);
        span: False / 0 the window test in every step, True / 1 run end by bisection, 2 the lean form of it"""
        fr = np.ascontiguousarray(frames).view(np.float32).reshape(-1, 16)
        img = np.zeros((len(fr), H, W), np.float32) if accumulate_into is None else _f32(accumulate_into).copy()
        valid = C.c_ulonglong(0)
        st = lib().doh_render_cube(self._h, int(trilinear), int(lerp), int(divmode), int(fetch), int(span), _p(fr),
                                   len(fr), W, H, nsamples, maxdist, int(accumulate_into is not None), _p(img),
                                   C.byref(valid))
        if st < 0:
            raise ValueError("unknown kernel variant")
        return img, int(valid.value), st


def render_fast(cube: HostCube, frames, W, H, variant=2, signs=-1, kstep=False, nsamples=200, maxdist=44,
                accumulate_into=None):
    """march_cube_fast_kernel (BIRP_FAST_MARCH, tolerance mode) on the host; signs < 0: as the kernel chooses"""
    fr = np.ascontiguousarray(frames).view(np.float32).reshape(-1, 16)
    img = np.zeros((len(fr), H, W), np.float32) if accumulate_into is None else _f32(accumulate_into).copy()
    valid = C.c_ulonglong(0)
    st = lib().doh_render_cube_fast(cube._h, int(variant), int(signs), int(kstep), _p(fr), len(fr), W, H, nsamples, maxdist,
                                    int(accumulate_into is not None), _p(img), C.byref(valid))
    if st < 0:
        raise ValueError("unknown variant")
    return img, int(valid.value), st


def scale(s, ray_width, maybe_zero=False):
    """(scale_fast value, its exactness claim, the reference's three-step chain, scale_sample)"""
    s = _f32(s).ravel()
    n = len(s)
    y, chain, full = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    exact = np.zeros(n, np.uint8)
    lib().doh_scale(_p(s), n, float(ray_width), int(maybe_zero), _p(y), _p(exact), _p(chain), _p(full))
    return y, exact.astype(bool), chain, full


def scale_exhaustive(maybe_zero=False, stride=1, nthreads=0):
    """scale_fast against the reference's chain over every stride-th fp32 product:
    (accepted, accepted but different, accepted but negative / NaN, in range but sent to the cold path,
    in range and different whatever the flag says)"""
    import os
    out = (C.c_ulonglong * 5)()
    lib().doh_scale_exhaustive(int(maybe_zero), int(stride), int(nthreads or len(os.sched_getaffinity(0))), out)
    return tuple(int(x) for x in out)


def lerps(a, b, t):
    """(std::lerp transcription, positive-cube form, non-negative-cube form)"""
    a, b, t = _f32(a).ravel(), _f32(b).ravel(), _f32(t).ravel()
    n = len(a)
    ref, pos, nn = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib().doh_lerps(_p(a), _p(b), _p(t), n, _p(ref), _p(pos), _p(nn))
    return ref, pos, nn


def sample_cmem(params, xyz, mode=3):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), np.float32)
    params = np.ascontiguousarray(params)
    lib().doh_sample_cmem(_p(params), _p(xyz), len(xyz), int(mode), _p(out))
    return out


def div10_check(first_bits: int, last_bits: int, stride: int = 1):
    """(mismatches of the kernel's r / 10, mismatches of the bare multiply-and-correct form) over a range of fp32 bit patterns"""
    bad = np.zeros(2, np.uint64)
    lib().doh_div10_check(first_bits, last_bits, stride, _p(bad))
    return int(bad[0]), int(bad[1])


def cmem_filter_decides(params, xyz):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), np.uint8)
    params = np.ascontiguousarray(params)
    lib().doh_cmem_filter_decides(_p(params), _p(xyz), len(xyz), _p(out))
    return out.astype(bool)


def cmem_table(params):
    """the region table of a parameter set (cmem_table_cell over all cells, on the host): (256 * 64, 4) floats"""
    params = np.ascontiguousarray(params)
    out = np.zeros((256 * 64, 4), np.float32)
    lib().doh_cmem_table(_p(params), _p(out))
    return out


def sample_cmem_table(params, table, xyz):
    """(CMEM samples with the region table in front of the filter, which samples the table decided)"""
    xyz = _f32(xyz).reshape(-1, 3)
    params = np.ascontiguousarray(params)
    table = np.ascontiguousarray(table, dtype=np.float32)
    out = np.zeros(len(xyz), np.float32)
    dec = np.zeros(len(xyz), np.uint8)
    lib().doh_sample_cmem_table(_p(params), _p(table), _p(xyz), len(xyz), _p(out), _p(dec))
    return out, dec.astype(bool)


def render_cmem(params, frames, W, H, mode=3, nsamples=200, maxdist=44):
    """march_kernel<CmemField<mode>> on the host"""
    fr = np.ascontiguousarray(frames).view(np.float32).reshape(-1, 16)
    img = np.zeros((len(fr), H, W), np.float32)
    params = np.ascontiguousarray(params)
    lib().doh_render_cmem(_p(params), int(mode), _p(fr), len(fr), W, H, nsamples, maxdist, _p(img))
    return img
