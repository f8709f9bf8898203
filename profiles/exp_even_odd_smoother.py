import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import exp_geometric_aggregation as geo
from exp_geometric_aggregation import A, gmres_poly_omega, fgmres, probes, smoother_product_form
L = 128; V = L * L; n = 2 * V
tv = geo.g['tv0']
P = geo.geo_P(tv, 4, 4, 4); R = P.conj().T.tocsr(); A1 = (R @ A @ P).tocsc(); lu = spla.splu(A1)
s_, x_, t_ = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing='ij')
par = ((x_ + t_) % 2).ravel()
ie = np.where(par == 0)[0]; io = np.where(par == 1)[0]
Acsr = A.tocsr()
Aee = Acsr[ie][:, ie]; Aeo = Acsr[ie][:, io]; Aoe = Acsr[io][:, ie]; Aoo = Acsr[io][:, io]
c = Aee.diagonal()[0]
assert abs(Aee - c * sp.identity(len(ie))).max() < 1e-12 and abs(Aoo - c * sp.identity(len(io))).max() < 1e-12
def Ahat(v): return c * v - Aeo @ (Aoe @ v) / c
rv = np.random.RandomState(7); b0 = rv.standard_normal(len(ie)) + 1j * rv.standard_normal(len(ie))
def test(dh):
    nu, p0 = smoother_product_form(gmres_poly_omega(Ahat, b0, dh))
    def S(r):
        re = r[ie] - Aeo @ r[io] / c
        y = re
        for v in nu: y = y - v * Ahat(y)
        xe = p0 * y
        xo = (r[io] - Aoe @ xe) / c
        out = np.zeros_like(r); out[ie] = xe; out[io] = xo
        return out
    def M(b):
        x = P @ lu.solve(R @ b); r = b - A @ x
        return x + S(r)
    its = [fgmres(M, b) for b in probes[:1]]
    print('even-odd Schur polynomial degree', dh, '(operator applications ~', dh + 1, ') iters', its, flush=True)
for dh in (8, 12, 16, 18, 20, 24):
    test(dh)
