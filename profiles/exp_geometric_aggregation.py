"""CPU experiment (dev tool, uses the oracle for the operator): outer FGMRES iterations of the exact two-grid method on
Schwinger 128^2 with geometric aggregates (bx x bt sites, split by spin) for several smoother degrees.  Also the shared
helpers of the other exp_*.py scripts.   python profiles/exp_geometric_aggregation.py"""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refport
from deflatedmlmc_schwinger_b200.multigrid import leja_order, smoother_product_form
A = refport.load_matrix('schwinger128', -0.1320).tocsr()
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'schwinger128.npz'))
n = A.shape[0]; L = 128; V = L * L
def geo_P(tv, bx, bt, nv):
    tv = tv[:, :nv]
    s, x, t = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing='ij')
    ag = (x // bx) * (L // bt) + t // bt
    cb = (2 * ag + s).ravel()           # coarse block of each row
    ncb = cb.max() + 1
    Pv = np.zeros((n, nv), dtype=complex)
    # orthonormalise per block
    order = np.argsort(cb, kind='stable'); m = n // ncb
    rows = order.reshape(ncb, m)
    for b in range(ncb):
        Q, _ = np.linalg.qr(tv[rows[b]])
        Pv[rows[b]] = Q
    indptr = np.arange(n + 1) * nv
    idx = (cb[:, None] * nv + np.arange(nv)[None, :]).ravel()
    return sp.csr_matrix((Pv.ravel(), idx, indptr), shape=(n, ncb * nv))
def gmres_poly_omega(apply_A, b, degree):
    n = b.shape[0]
    Vv = np.zeros((degree + 1, n), dtype=np.complex128); H = np.zeros((degree + 1, degree), dtype=np.complex128)
    Vv[0] = b / np.linalg.norm(b)
    for j in range(degree):
        w = apply_A(Vv[j])
        for _ in range(2):
            hh = np.conj(Vv[:j + 1]) @ w; H[:j + 1, j] += hh; w = w - hh @ Vv[:j + 1]
        H[j + 1, j] = np.linalg.norm(w); Vv[j + 1] = w / H[j + 1, j]
    Hm = H[:degree, :degree]; em = np.zeros(degree); em[-1] = 1.0
    f = np.linalg.solve(Hm.conj().T, em)
    theta = np.linalg.eigvals(Hm + (abs(H[degree, degree - 1]) ** 2) * np.outer(f, em))
    return 1.0 / np.array(leja_order(theta))
def fgmres(M, b, tol=1e-12, maxit=80):
    beta = np.linalg.norm(b); Vs = [b / beta]; H = np.zeros((maxit + 1, maxit), dtype=complex)
    for j in range(maxit):
        z = M(Vs[j]); w = A @ z
        for _ in range(2):
            for i in range(j + 1):
                h = np.vdot(Vs[i], w); H[i, j] += h; w = w - h * Vs[i]
        H[j + 1, j] = np.linalg.norm(w); Vs.append(w / H[j + 1, j])
        e1 = np.zeros(j + 2, dtype=complex); e1[0] = beta
        y = np.linalg.lstsq(H[:j + 2, :j + 1], e1, rcond=None)[0]
        rn = np.linalg.norm(e1 - H[:j + 2, :j + 1] @ y)
        if rn < tol * beta: return j + 1
    return maxit
rs = np.random.RandomState(123456)
probes = [(2.0 * rs.randint(0, 2, n) - 1).astype(complex) for _ in range(2)]
rv = np.random.RandomState(7); b0 = rv.standard_normal(n) + 1j * rv.standard_normal(n)
def run(P, label, degrees):
    R = P.conj().T.tocsr(); A1 = (R @ A @ P).tocsc(); lu = spla.splu(A1)
    print(label, 'n1', A1.shape[0], 'nnz/row', A1.nnz / A1.shape[0], flush=True)
    for d in degrees:
        omega = gmres_poly_omega(lambda v: A @ v, b0, d)
        nu, p0 = smoother_product_form(omega)
        def M(b):
            x = P @ lu.solve(R @ b); y = b - A @ x
            for v in nu: y = y - v * (A @ y)
            return x + p0 * y
        its = [fgmres(M, b) for b in probes]
        print('   post-smoother degree', d, 'iters', its, 'cost(step units, overhead 70)', its[0] * (d + 70), flush=True)
if __name__ == '__main__':
    tv = g['tv0']
    run(geo_P(tv, 4, 4, 4), 'geometric 4x4, 4 tv', (4, 8, 12, 16, 24, 32))
    run(geo_P(tv, 2, 2, 4), 'geometric 2x2, 4 tv', (4, 8, 16))
    run(geo_P(tv, 4, 4, 2), 'geometric 4x4, 2 tv', (8, 16, 32))
