"""CPU experiment (round 2): mixed-precision outer solve on the even-odd Schur complement.  Outer loop = iterative refinement
in complex128 (true residual r = b^ - S x, x += dx); inner solve = FGMRES(m) whose Krylov basis V, preconditioned vectors Z
and work vector are STORED in complex64 (coefficients accumulated in FP64), restarted every m steps.  Gram-Schmidt, the basis
normalisation and the solution update then move half the bytes of the complex128 Schur solve and a quarter of the full-lattice
one.  Counts the total number of preconditioner applications to a true relative residual of 1e-12 for several m."""
import sys, numpy as np, scipy.sparse.linalg as spla
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
    xe = p0 * y
    return xe
def Me(ve):
    v = np.zeros(n, dtype=complex); v[ie] = ve
    x = P @ lu.solve(R @ v)
    return x[ie] + smooth(v - A @ x)
c64 = lambda a: a.astype(np.complex64).astype(np.complex128)

def ir_solve(b, m, store=c64, tol=1e-12, inner_drop=1e-5, reorth=False):
    normb = np.linalg.norm(b)
    bh = b[ie] - Heo @ b[io] / c
    x = np.zeros_like(bh); r = bh.copy(); napply = 0; cycles = 0
    while np.linalg.norm(r) >= tol * normb and napply < 60:
        beta = np.linalg.norm(r); Vs = [store(r / beta)]; Zs = []; H = np.zeros((m + 1, m), dtype=complex)
        target = max(tol * normb / beta * 0.5, inner_drop)
        for j in range(m):
            z = store(Me(Vs[j])); Zs.append(z); napply += 1
            w = store(S(z))
            for _ in range(2 if reorth else 1):
                hs = [np.vdot(Vs[i], w) for i in range(j + 1)]
                for i in range(j + 1): w = w - hs[i] * Vs[i]
                H[:j + 1, j] += hs
                w = store(w)
            H[j + 1, j] = np.linalg.norm(w); Vs.append(store(w / H[j + 1, j]))
            e1 = np.zeros(j + 2, dtype=complex); e1[0] = 1.0
            y = np.linalg.lstsq(H[:j + 2, :j + 1], e1, rcond=None)[0]
            if np.linalg.norm(e1 - H[:j + 2, :j + 1] @ y) < target: break
        x = x + beta * sum(yi * zi for yi, zi in zip(y, Zs))
        r = bh - S(x); cycles += 1
    return napply, cycles, np.linalg.norm(r) / normb

for b in probes[:2]:
    print('complex128 storage, no restart:', ir_solve(b, 40, store=lambda a: a, inner_drop=0.0), flush=True)
    for m in (3, 4, 5, 6, 8):
        for drop in (1e-5, 1e-6):
            print('complex64 storage, m = %d, inner target %.0e:' % (m, drop), ir_solve(b, m, inner_drop=drop), flush=True)
