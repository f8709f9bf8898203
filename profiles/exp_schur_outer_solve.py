"""CPU experiment: outer FGMRES on the even-odd Schur complement S x_e = b^_e, preconditioned by the even part of the two-grid
cycle applied to (v_e, 0), against the outer FGMRES on A itself with the same cycle (geometric 4x4 hierarchy, exact coarse solve,
even-odd post-smoother of degree 16 in S).  Counts outer iterations to 1e-12 (relative to ||b||)."""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import exp_geometric_aggregation as geo
from exp_geometric_aggregation import A, gmres_poly_omega, probes, smoother_product_form
L = 128; V = L * L; n = 2 * V
tv = geo.g['tv0']
P = geo.geo_P(tv, 4, 4, 4); R = P.conj().T.tocsr(); A1 = (R @ A @ P).tocsc(); lu = spla.splu(A1)
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
def M(b):
    x = P @ lu.solve(R @ b)
    return x + smooth(b - A @ x)
def fgmres(op, prec, b, normb, tol=1e-12, maxit=60):
    beta = np.linalg.norm(b); Vs = [b / beta]; H = np.zeros((maxit + 1, maxit), dtype=complex)
    for j in range(maxit):
        z = prec(Vs[j]); w = op(z)
        for _ in range(2):
            for i in range(j + 1):
                h = np.vdot(Vs[i], w); H[i, j] += h; w = w - h * Vs[i]
        H[j + 1, j] = np.linalg.norm(w); Vs.append(w / H[j + 1, j])
        e1 = np.zeros(j + 2, dtype=complex); e1[0] = beta
        y = np.linalg.lstsq(H[:j + 2, :j + 1], e1, rcond=None)[0]
        if np.linalg.norm(e1 - H[:j + 2, :j + 1] @ y) < tol * normb: return j + 1
    return maxit
def Me(ve):
    v = np.zeros(n, dtype=complex); v[ie] = ve
    return M(v)[ie]
for b in probes[:2]:
    it_full = fgmres(lambda v: A @ v, M, b, np.linalg.norm(b))
    bh = b[ie] - Heo @ b[io] / c
    it_schur = fgmres(S, Me, bh, np.linalg.norm(b))
    print('outer iterations: full system', it_full, ' Schur complement system (half-length Krylov vectors)', it_schur, flush=True)
