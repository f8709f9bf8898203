"""CPU experiment: outer iterations when the dense level-1 inverse of the geometric hierarchy (n = 8192) and its right-hand
side are rounded to FP8 e4m3 (per-row / per-column power-of-two scales), BF16 or kept exact.  Two-grid cycle with the even-odd
post-smoother of degree 16 in S; FGMRES to 1e-12."""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import exp_geometric_aggregation as geo
from exp_geometric_aggregation import A, gmres_poly_omega, probes, smoother_product_form, fgmres
L = 128; V = L * L; n = 2 * V
tv = geo.g['tv0']
P = geo.geo_P(tv, 4, 4, 4); R = P.conj().T.tocsr(); A1 = (R @ A @ P).tocsc(); lu = spla.splu(A1)
t0 = time.time(); n1 = A1.shape[0]
Minv = np.empty((n1, n1), dtype=np.complex128)
for c0 in range(0, n1, 1024):
    E = np.zeros((n1, 1024), dtype=np.complex128); E[c0:c0 + 1024] = np.eye(1024)
    Minv[:, c0:c0 + 1024] = lu.solve(E)
print('dense inverse %.0f s' % (time.time() - t0), flush=True)
def rnd_fmt(x, mant, emin, emax_val):
    """round real array to a float format with `mant` mantissa bits, min normal exponent emin, saturating at emax_val"""
    ax = np.abs(x); out = np.zeros_like(x)
    nz = ax > 0
    e = np.floor(np.log2(ax[nz])); e = np.maximum(e, emin)
    q = 2.0 ** (e - mant)
    out[nz] = np.sign(x[nz]) * np.minimum(np.round(ax[nz] / q) * q, emax_val)
    return out
def fp8(x): return rnd_fmt(x, 3, -6, 448.0)
def bf16(x): return rnd_fmt(x, 7, -126, 3.3e38)
def scaled(f, z, axis):
    """apply real format f to re/im of z after a power-of-two scale per row (axis=1) / column (axis=0) mapping the max to ~256"""
    m = np.maximum(np.abs(z.real).max(axis=axis, keepdims=True), np.abs(z.imag).max(axis=axis, keepdims=True))
    sc = 2.0 ** np.floor(np.log2(256.0 / np.maximum(m, 1e-300)))
    return (f(z.real * sc) + 1j * f(z.imag * sc)) / sc
s_, x_, t_ = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing='ij')
par = ((x_ + t_) % 2).ravel(); ie = np.where(par == 0)[0]; io = np.where(par == 1)[0]
Ac = A.tocsr(); Heo = Ac[ie][:, io]; Hoe = Ac[io][:, ie]; c = Ac.diagonal()[0].real
def S(v): return c * v - Heo @ (Hoe @ v) / c
rv = np.random.RandomState(7); b0 = rv.standard_normal(len(ie)) + 1j * rv.standard_normal(len(ie))
nu, p0 = smoother_product_form(gmres_poly_omega(S, b0, 16))
def smooth(r):
    y = r[ie] - Heo @ r[io] / c
    for v in nu: y = y - v * S(y)
    xe = p0 * y; xo = (r[io] - Hoe @ xe) / c
    out = np.zeros_like(r); out[ie] = xe; out[io] = xo
    return out
def make(Mm, frhs):
    def M(b):
        rc = R @ b
        x = P @ (Mm @ frhs(rc))
        return x + smooth(b - A @ x)
    return M
ident = lambda v: v
variants = [('exact inverse, exact rhs', Minv, ident),
            ('BF16 inverse, BF16 rhs (the device today)', scaled(bf16, Minv, 1), lambda v: scaled(bf16, v[:, None], 0)[:, 0]),
            ('FP8 e4m3 inverse (row scales), BF16 rhs', scaled(fp8, Minv, 1), lambda v: scaled(bf16, v[:, None], 0)[:, 0]),
            ('FP8 e4m3 inverse (row scales), FP8 rhs', scaled(fp8, Minv, 1), lambda v: scaled(fp8, v[:, None], 0)[:, 0])]
for name, Mm, fr in variants:
    print('%-48s outer iterations %s' % (name, [fgmres(make(Mm, fr), b) for b in probes[:2]]), flush=True)
