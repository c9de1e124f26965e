#!/usr/bin/env python
"""Generate the polynomial coefficients of birp_cpp_b200/csrc/birp_cmem_math.cuh (fp64 asin core).

    python tools/gen_cmem_coeffs.py            # prints the coefficient table and the measured error

asin(t) = t + t^3 * R(t^2) for |t| <= 1/2: R is interpolated at Chebyshev nodes of z = t^2 in
[0, 1/4] with mpmath (50 digits), converted to monomial coefficients and rounded to fp64; the
error of the ROUNDED polynomial, evaluated in fp64 Horner form, is then measured against mpmath.
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def R(z):
    z = mp.mpf(z)
    if z == 0:
        return mp.mpf(1) / 6
    t = mp.sqrt(z)
    return (mp.asin(t) / t - 1) / z


def cheb_fit(f, a, b, n):
    """monomial coefficients (in z) of the degree-n interpolant of f at the Chebyshev nodes of [a, b]"""
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
    xs = [(a + b) / 2 + (b - a) / 2 * x for x in nodes]
    V = mp.matrix(n + 1, n + 1)
    y = mp.matrix(n + 1, 1)
    for i, x in enumerate(xs):
        for j in range(n + 1):
            V[i, j] = mp.mpf(x) ** j
        y[i] = f(x)
    c = mp.lu_solve(V, y)
    return [c[j] for j in range(n + 1)]


def horner64(c, z):
    p = np.float64(c[-1])
    for k in range(len(c) - 2, -1, -1):
        p = p * z + np.float64(c[k])
    return p


def main():
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    c = cheb_fit(R, mp.mpf(0), mp.mpf(1) / 4, deg)
    c64 = [float(x) for x in c]
    worst = 0.0
    for t in np.linspace(0.0, 0.5, 20001):
        z = np.float64(t) * np.float64(t)
        got = np.float64(t) + np.float64(t) * z * horner64(c64, z)
        want = mp.asin(mp.mpf(float(t)))
        err = abs(mp.mpf(float(got)) - want)
        worst = max(worst, float(err))
    print(f"// asin(t) = t + t^3 R(t^2), |t| <= 1/2: degree {deg} in z = t^2, max abs error {worst:.3g} (fp64 Horner)")
    for k, x in enumerate(c64):
        print(f"    {x.hex()},  // z^{k}  {x!r}")


if __name__ == "__main__":
    main()
