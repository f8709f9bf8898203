"""TEST INFRASTRUCTURE ONLY (oracle shim) -- restatement of pyamg.krylov.fgmres.

Call site in the reference: multigrid.py:362
    fgmres(lop1, b, tol=tol, M=lop2, callback=callback, maxiter=maxiter)

Semantics restated from pyamg's published algorithm (pyamg >= 4, version unpinned
by the reference, source not available in this image -> "parity unpinned"):
  * x0 = 0, restart=None  ->  ONE cycle of at most min(maxiter, n) Krylov vectors
  * right preconditioning, flexible: z_j = M v_j (all z_j stored), w = A z_j
  * orthogonalisation: pyamg uses Householder reflectors; any backward-stable
    Arnoldi spans the same spaces, here modified Gram-Schmidt with one
    re-orthogonalisation pass
  * Givens rotations keep the least-squares residual norm; stop as soon as
    ||r|| < tol * ||b||   (||b|| = 0 is replaced by 1)
  * x = sum_j y_j z_j ;  callback(x_k) once per inner iteration (the reference only
    counts the calls, multigrid.py:349-352)
"""
import numpy as np


def fgmres(A, b, x0=None, tol=1e-5, restart=None, maxiter=None, M=None,
           callback=None, residuals=None):
    b = np.asarray(b).reshape(-1)
    n = b.shape[0]
    dtype = np.result_type(b.dtype, np.complex128) if np.iscomplexobj(b) else np.float64
    x = np.zeros(n, dtype=dtype) if x0 is None else np.array(x0, dtype=dtype).reshape(-1)
    if maxiter is None:
        maxiter = min(n, 40)
    max_inner = min(maxiter, n)

    normb = np.linalg.norm(b)
    if normb == 0.0:
        normb = 1.0
    r = b - A.matvec(x) if x0 is not None else b.astype(dtype, copy=True)
    normr = np.linalg.norm(r)
    if residuals is not None:
        residuals[:] = [normr]
    if normr < tol * normb:
        return x, 0

    V = np.zeros((max_inner + 1, n), dtype=dtype)
    Z = []
    H = np.zeros((max_inner + 1, max_inner), dtype=dtype)
    cs = np.zeros(max_inner, dtype=dtype)
    sn = np.zeros(max_inner, dtype=dtype)
    g = np.zeros(max_inner + 1, dtype=dtype)
    g[0] = normr
    V[0] = r / normr

    niter = 0
    for j in range(max_inner):
        z = M.matvec(V[j]) if M is not None else V[j].copy()
        z = np.asarray(z).reshape(-1)
        Z.append(z)
        w = np.asarray(A.matvec(z)).reshape(-1).astype(dtype)
        for _pass in range(2):
            for i in range(j + 1):
                hij = np.vdot(V[i], w)
                H[i, j] += hij
                w = w - hij * V[i]
        hn = np.linalg.norm(w)
        H[j + 1, j] = hn
        if hn != 0.0:
            V[j + 1] = w / hn
        # apply the previous rotations to the new column
        for i in range(j):
            t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
            H[i + 1, j] = -np.conj(sn[i]) * H[i, j] + cs[i] * H[i + 1, j]
            H[i, j] = t
        # new rotation annihilating H[j+1, j]
        a_, b_ = H[j, j], H[j + 1, j]
        den = np.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
        if den == 0.0:
            cs[j], sn[j] = 1.0, 0.0
        else:
            cs[j] = abs(a_) / den if a_ != 0 else 0.0
            sn[j] = (a_ / abs(a_)) * np.conj(b_) / den if a_ != 0 else 1.0
        H[j, j] = cs[j] * a_ + sn[j] * b_
        H[j + 1, j] = 0.0
        g[j + 1] = -np.conj(sn[j]) * g[j]
        g[j] = cs[j] * g[j]
        niter += 1
        normr = abs(g[j + 1])
        if residuals is not None:
            residuals.append(normr)
        if callback is not None:
            callback(x)          # the reference only counts calls
        if normr < tol * normb:
            break

    k = niter
    y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k > 0 else np.zeros(0, dtype=dtype)
    for i in range(k):
        x = x + y[i] * Z[i]
    return x, (0 if normr < tol * normb else niter)
