"""Fits and checks the polynomial behind gelu_erf_poly_f (magicdance_b200/csrc/common.cuh): erf(z) ~ z P(z^2) on
|z| <= zmax, clamped beyond.  CPU only (numpy + scipy).  Prints the coefficients (highest power last) and the
maximum absolute error of erf and of GELU when the polynomial is evaluated in float32 Horner arithmetic.

    python scripts/fit_erf_poly.py [--zmax 3.0] [--deg 8]
"""
import argparse

import numpy as np
from scipy.special import erf


def fit(zmax, deg, n=400):
    k = np.arange(n)
    u = (np.cos(np.pi * (k + 0.5) / n) + 1) / 2 * zmax ** 2          # Chebyshev nodes in u = z^2
    z = np.sqrt(u)
    f = np.where(z > 0, erf(z) / np.maximum(z, 1e-30), 2 / np.sqrt(np.pi))
    a = np.vander(u, deg + 1, increasing=True) * z[:, None]           # weight z: absolute error of z * P
    coef, *_ = np.linalg.lstsq(a, f * z, rcond=None)
    return coef


def erf_poly_f32(z, coef, zmax):
    z = np.clip(z.astype(np.float32), np.float32(-zmax), np.float32(zmax))
    u = (z * z).astype(np.float32)
    p = np.full_like(u, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        p = (p * u + np.float32(c)).astype(np.float32)
    return (p * z).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zmax", type=float, default=3.0)
    ap.add_argument("--deg", type=int, default=8)
    args = ap.parse_args()
    coef = fit(args.zmax, args.deg)
    x = np.linspace(-12, 12, 800001)
    z = x * 0.70710678118654752
    e = erf_poly_f32(z, coef, args.zmax).astype(np.float64)
    gelu = 0.5 * x * (1 + e)
    ref = 0.5 * x * (1 + erf(z))
    print("coefficients (z^0 ... z^%d of P(z^2)):" % (2 * args.deg), [float(c) for c in coef])
    print(f"max |erf error| {np.abs(e - erf(z)).max():.3e}   max |gelu error| {np.abs(gelu - ref).max():.3e}   "
          f"max |gelu error| for |x| <= 4.2: {np.abs(gelu - ref)[np.abs(x) <= 4.2].max():.3e}")


if __name__ == "__main__":
    main()
