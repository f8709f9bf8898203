"""CPU experiment for the next round: outer iterations of the geometric two-grid cycle (4x4-site spin-split aggregates, exact
coarse solve, even-odd post-smoother) when the preconditioner hierarchy uses MORE test vectors than the estimator's (the
nv smallest eigenvectors of A, scipy eigs as multigrid.py:174).  Coarse size = 2048 nv.   python profiles/exp_more_test_vectors.py"""
import sys, os, time, numpy as np, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import exp_geometric_aggregation as geo
from exp_geometric_aggregation import A, gmres_poly_omega, probes, smoother_product_form, fgmres
L = 128; V = L * L; n = 2 * V
t0 = time.time()
w, tv8 = spla.eigs(A.tocsc(), k=8, which='LM', tol=1e-9, maxiter=1000000, sigma=0.0)
order = np.argsort(np.abs(w)); tv8 = tv8[:, order]
print('eigs k=8: %.0f s' % (time.time() - t0), np.round(w[order], 4), flush=True)
s_, x_, t_ = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing='ij')
par = ((x_ + t_) % 2).ravel(); ie = np.where(par == 0)[0]; io = np.where(par == 1)[0]
Ac = A.tocsr(); Heo = Ac[ie][:, io]; Hoe = Ac[io][:, ie]; c = Ac.diagonal()[0].real
def S(v): return c * v - Heo @ (Hoe @ v) / c
rv = np.random.RandomState(7); b0 = rv.standard_normal(len(ie)) + 1j * rv.standard_normal(len(ie))
polys = {d: smoother_product_form(gmres_poly_omega(S, b0, d)) for d in (4, 8, 12, 16)}
for nv in (4, 6, 8):
    P = geo.geo_P(tv8, 4, 4, nv); R = P.conj().T.tocsr(); lu = spla.splu((R @ A @ P).tocsc())
    for d, (nu, p0) in polys.items():
        def M(b):
            x = P @ lu.solve(R @ b); r = b - A @ x
            y = r[ie] - Heo @ r[io] / c
            for v in nu: y = y - v * S(y)
            xe = p0 * y; xo = (r[io] - Hoe @ xe) / c
            x[ie] += xe; x[io] += xo
            return x
        print('nv', nv, 'coarse size', P.shape[1], 'even-odd degree', d, 'outer iterations', [fgmres(M, b) for b in probes[:1]], flush=True)
