import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden" / "reference_renders.npz"

# per-pixel |gpu - ref| <= 1e-5*|ref| with a floor of 1e-5*max(image);
# integrated flux within 1e-4 relative; pass = at least 95 % of the pixels inside.
PIXEL_RTOL = 1e-5
FLUX_RTOL = 1e-4
MIN_PIXEL_FRACTION = 0.95


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_exact(got, want, what=""):
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = bits(got) != bits(want)
    # +0 and -0 compare equal as pixels; a NaN must be matched by a NaN (payload bits are not
    # preserved by the GPU's arithmetic and carry no meaning here)
    bad &= ~((got == 0) & (want == 0))
    bad &= ~(np.isnan(got) & np.isnan(want))
    n = int(bad.sum())
    if n:
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {n} of {got.size} values differ bitwise, first at {idx.tolist()}: "
                             f"{got[bad][:5]} vs {want[bad][:5]}")


def pixel_fraction_within(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    floor = PIXEL_RTOL * np.abs(want).max() if want.size else 0.0
    ok = np.abs(got - want) <= np.maximum(PIXEL_RTOL * np.abs(want), floor)
    return float(ok.mean()) if ok.size else 1.0


def assert_image_parity(got, want, what="", min_fraction=MIN_PIXEL_FRACTION):
    frac = pixel_fraction_within(got, want)
    fg, fw = float(np.sum(got, dtype=np.float64)), float(np.sum(want, dtype=np.float64))
    assert frac >= min_fraction, f"{what}: only {frac:.4%} of pixels within {PIXEL_RTOL} of the reference"
    assert abs(fg - fw) <= FLUX_RTOL * abs(fw) + 1e-30, f"{what}: flux {fg} vs {fw}"
    return frac


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def cube1():
    from birp_cpp_b200 import synthetic as S
    return S.config1_cube()


@pytest.fixture(scope="session")
def cube1_neg():
    from birp_cpp_b200 import synthetic as S
    return S.config1_cube(negative_slab=True)


@pytest.fixture(scope="session")
def cube1_matches_golden(golden, cube1, cube1_neg):
    """The fixtures were rendered from these exact cube bytes (numpy's exp could differ by an ulp on
    another CPU; when it does the image fixtures are compared at tolerance instead of bitwise)."""
    a = hashlib.sha256(cube1[0].tobytes()).digest() == golden["cube_sha256"].tobytes()
    b = hashlib.sha256(cube1_neg[0].tobytes()).digest() == golden["cube_neg_sha256"].tobytes()
    return a and b


def _cuda_ok():
    try:
        from birp_cpp_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The GPU tests must run the CUDA extension: no skip, no fallback."""
    from birp_cpp_b200 import capi
    n = capi.device_count()
    assert n > 0, "no CUDA device visible: -m gpu tests need the B200 box"
    capi.set_device(0)
    return capi
