"""Derive the fp32 polynomial coefficients used by vg_renderer_b200/csrc/vgb_math.h.

bx (the reference's math dependency) is not vendored and not pinned, so this
build defines its own explicit fp32 routines, shared bit-for-bit by the oracle's bx shim and
the CUDA kernels.  This script is the provenance of the constants: near-minimax fits
(least squares on Chebyshev nodes, float64) rounded to float32, with the achieved error
printed.  Run:  python tools/fit_math.py
"""
import numpy as np

def cheb_nodes(a, b, n):
    k = np.arange(n)
    x = np.cos((2 * k + 1) * np.pi / (2 * n))
    return 0.5 * (a + b) + 0.5 * (b - a) * x

def fit(fn_resid, lo, hi, deg, n=4000, weight=None):
    """fit resid(z) ~ sum c_k z^k on [lo,hi] (z is already the squared variable)."""
    z = cheb_nodes(lo, hi, n)
    y = fn_resid(z)
    A = np.vander(z, deg + 1, increasing=True)
    w = np.ones_like(z) if weight is None else weight(z)
    c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
    return c.astype(np.float32)

def show(name, c):
    print(name, "= {", ", ".join(float(v).hex() + "f" for v in c), "}")
    print("   ", [float(v) for v in c])

# sin(r) = r + r^3 * S(r^2),  |r| <= pi/4
from math import factorial
def sin_resid(z):   # (sin r - r)/r^3 by Taylor series in z=r^2 (no cancellation)
    return sum((-1) ** k * z ** (k - 1) / factorial(2 * k + 1) for k in range(1, 12))
def cos_resid(z):   # (cos r - 1 + r^2/2)/r^4
    return sum((-1) ** k * z ** (k - 2) / factorial(2 * k) for k in range(2, 13))
LO = 4e-4           # sqrt(LO)=0.02: below this the polynomial is dominated by its leading terms anyway
S = fit(sin_resid, 0.0, (np.pi / 4) ** 2 * 1.02, 2)
show("SIN", S)
# cos(r) = 1 - r^2/2 + r^4 * C(r^2)
C = fit(cos_resid, 0.0, (np.pi / 4) ** 2 * 1.02, 2)
show("COS", C)
# atan(t) = t + t^3 * A(t^2), |t| <= tan(pi/8)
T8 = np.tan(np.pi / 8)
A = fit(lambda z: (np.arctan(np.sqrt(z)) - np.sqrt(z)) / np.sqrt(z) ** 3, LO, T8 ** 2 * 1.02, 4)
show("ATAN", A)
# asin(x) = x + x^3 * R(x^2), |x| <= 0.5
R = fit(lambda z: (np.arcsin(np.sqrt(z)) - np.sqrt(z)) / np.sqrt(z) ** 3, LO, 0.25 * 1.02, 4)
show("ASIN", R)

# ---- round-1 revision: FMA-evaluated variants (every step is one IEEE fma, identical on CPU
# (fmaf) and GPU (__fmaf_rn)).  atan: one odd polynomial over the whole [0,1] octant so that
# atan2 needs a single division and no branch.
A1 = fit(lambda z: (np.arctan(np.sqrt(z)) - np.sqrt(z)) / np.sqrt(z) ** 3, LO, 1.0, 8, n=20000)
show("ATAN01", A1)
t = np.linspace(0, 1, 2000001)
z = t * t
q = np.zeros_like(t)
for c in A1[::-1].astype(np.float64):
    q = q * z + c
print("   atan01 max abs err (float64 eval):", np.max(np.abs(t + t ** 3 * q - np.arctan(t))))
