"""The claim behind march_steps_span (birp_cpp_b200/csrc/birp_kernels.cu), checked on the CPU in the same fp32
arithmetic: along a monotone ray table every ingress RN(RN(RN(d * t) + p) - o) is a monotone function of the
step, so the steps whose three ingresses lie inside the hot window [lo, l) form ONE contiguous run, and the
bisection the kernel uses (same midpoint formula, same window test) finds the end of that run exactly.

numpy's float32 operations are IEEE round-to-nearest without contraction, i.e. the kernel's __fmul_rn /
__fadd_rn / __fsub_rn."""
import numpy as np
import pytest

from birp_cpp_b200 import capi

f32 = np.float32


def ray_table(nsamples, maxdist):
    """the library's own host table (Camera::GenerateRayDistWidth)"""
    dist, _ = capi.ray_table(nsamples, maxdist)
    return np.asarray(dist, f32)


def ingress(d, p, o, t):
    return ((d * t).astype(f32) + p).astype(f32) - o


def hot_mask(dirs, pos, org, lo, hi, t):
    """hot[ray, step]: all three ingresses inside [lo, hi)"""
    m = np.ones((len(dirs), len(t)), bool)
    for a in range(3):
        xi = ingress(dirs[:, a:a + 1], f32(pos[a]), f32(org[a]), t[None, :]).astype(f32)
        m &= (xi >= f32(lo[a])) & (xi < f32(hi[a]))
    return m


def bisect_run_end(hot_row, start, end):
    """march_steps_span's search in units of steps: first step in (start, end] that is not hot"""
    lo, hi = start + 1, end
    evaluations = 0
    while lo < hi:
        mid = lo + ((hi - lo) >> 1)
        assert start < mid < end
        evaluations += 1
        if hot_row[mid]:
            lo = mid + 1
        else:
            hi = mid
    return lo, evaluations


@pytest.mark.parametrize("nsamples,maxdist", [(200, 44), (517, 44), (2, 10), (2048, 100), (200, -44), (64, 0)])
def test_hot_samples_of_a_ray_form_one_run_and_bisection_finds_its_end(nsamples, maxdist):
    rng = np.random.default_rng(nsamples * 131 + maxdist)
    t = ray_table(nsamples, maxdist)
    assert np.all(np.diff(t) >= 0) or np.all(np.diff(t) <= 0)
    nrays = 4000
    dirs = rng.normal(size=(nrays, 3)).astype(f32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(f32)
    dirs[:50, 0] = 0  # rays inside a grid plane, axis-aligned rays, signed zeros
    dirs[50:100, 1] = -0.0
    dirs[100:120] = np.array([1, 0, 0], f32)
    dirs[120:140] = np.array([0, 0, -1], f32)
    for trial in range(6):
        org = rng.uniform(-40, -5, 3).astype(f32)
        size = rng.uniform(20, 80, 3).astype(f32)
        lo = rng.uniform(1e-3, 1.0, 3).astype(f32)  # "first cell" edge of the window
        hi = (size - rng.uniform(0, 1, 3)).astype(f32)
        pos = (org + rng.uniform(-0.2, 1.2, 3).astype(f32) * size).astype(f32)  # inside, on the edge, outside
        hot = hot_mask(dirs, pos, org, lo, hi, t)
        # (1) one contiguous run per ray
        changes = np.diff(hot.astype(np.int8), axis=1)
        assert np.all((changes == 1).sum(1) + hot[:, 0] <= 1), "a ray re-entered the hot window"
        # (2) the bisection started at any hot step returns the first cold step after it
        for r in rng.choice(nrays, 200, replace=False):
            steps = np.flatnonzero(hot[r])
            if len(steps) == 0:
                continue
            for start in {int(steps[0]), int(steps[len(steps) // 2]), int(steps[-1])}:
                want = int(steps[-1]) + 1
                got, evals = bisect_run_end(hot[r], start, nsamples)
                assert got == want
                assert evals <= max(1, int(np.ceil(np.log2(nsamples))) + 1)


def test_ray_tables_of_the_library_are_monotone():
    """RayTable::monotone is computed where the table is made; host_ray_table = maxdist * ((i + 0.5f) / n)"""
    for n, m in [(1, 44), (2, 1), (200, 44), (600, 12), (4097, 1000), (200, -44), (200, 0)]:
        d, w = capi.ray_table(n, m)
        df = np.diff(np.asarray(d, f32))
        assert np.all(df >= 0) or np.all(df <= 0)
