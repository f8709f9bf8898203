"""Block eigensolvers of the set-up phase, written against two callables so that the heavy work is the batched device
solver / SpMM (and so that the CPU tests can drive the same code with scipy):

  smallest_eigenpairs(apply_A, apply_Ainv, X0, k, tol)
      the k eigenpairs of smallest modulus of a NON-Hermitian operator -- the multigrid test vectors of
      multigrid.py:174 (`eigs(Al, k, which='LM', sigma=0.0)`): block Arnoldi on A^{-1} (shift-invert at 0 exactly as
      ARPACK does it there), every block step is ONE batched solve of p columns, Rayleigh-Ritz on the block Hessenberg
      matrix, explicit residuals ||A x - theta x|| <= tol ||x||, restart from the Ritz vectors.  Conjugate pairs cut by k
      are broken deterministically (Im theta > 0 is kept), which scipy's random start vector does not do.

  largest_hermitian_eigenpairs(apply_Op, X0, k, tol)
      the k eigenpairs of largest modulus of a Hermitian operator -- the deflation vectors of utils.py:137-158
      (`eigsh(Q, k, which='LM', sigma=0)` = largest of Q^{-1}; `eigsh(LinearOperator(diff_op_Q), k, which='LM')`):
      block Lanczos with full re-orthogonalisation (thick restart), every block step ONE batched operator application
      of p columns instead of eigsh's one vector at a time.

Tensors are torch (CUDA in the product, CPU in the tests); the small dense algebra (p x p, (m p) x (m p)) is numpy on
the host.  Nothing here knows about lattices or hierarchies."""
import numpy as np


def _orth(torch, W, basis=None, passes=2):
    """W [n, p] orthonormalised against the orthonormal blocks in `basis` and in itself (block classical Gram-Schmidt,
    `passes` times, then Householder QR).  Returns (Q, coefficient blocks per basis block summed over passes, R)."""
    coefs = None
    if basis:
        B = torch.cat(basis, dim=1)
        coefs = torch.zeros((B.shape[1], W.shape[1]), dtype=W.dtype, device=W.device)
        for _ in range(passes):
            c = B.conj().T @ W
            W = W - B @ c
            coefs += c
    Q, R = torch.linalg.qr(W)
    return Q, coefs, R


def _sort_smallest(theta, rel=1e-5):
    """order by modulus ascending; members of a conjugate pair (equal modulus): Im > 0 first"""
    theta = np.asarray(theta)
    key = np.abs(theta)
    order = np.argsort(key, kind='stable')
    out = list(order)
    i = 0
    while i + 1 < len(out):
        a, b = out[i], out[i + 1]
        if abs(key[a] - key[b]) <= rel * max(key[a], key[b]) and theta[a].imag < theta[b].imag:
            out[i], out[i + 1] = b, a
        i += 1
    return np.array(out)


def smallest_eigenpairs(apply_A, apply_Ainv, X0, k, tol=1e-9, max_blocks=10, max_restarts=12, min_blocks=1, verbose=False):
    """X0: torch [n, p] start block (p >= k; extra columns are guard vectors).  min_blocks: block steps taken before the
    convergence test (a warm start from an invariant subspace passes it at once; a few steps let the guard vectors
    show whether a smaller eigenvalue exists).  Returns (theta[k] numpy complex, X [n, k] torch with unit columns,
    residuals[k], info)."""
    import torch
    p = X0.shape[1]
    assert p >= k
    X = X0
    solves = 0
    info = {"restarts": 0, "block_solves": 0}
    done = False
    for restart in range(max_restarts):
        Q, _, _ = _orth(torch, X)
        V = [Q]
        Hrows = []                       # block columns of the (m+1)p x mp block Hessenberg matrix of A^{-1}
        for j in range(max_blocks):
            W = apply_Ainv(V[j])
            solves += 1
            Qn, coefs, R = _orth(torch, W, V)
            Hrows.append(torch.cat([coefs, R], dim=0))          # [(j+2) p, p]
            V.append(Qn)
            m = j + 1
            # Rayleigh-Ritz for A^{-1} on span(V_1..V_m)
            H = torch.zeros(((m + 1) * p, m * p), dtype=W.dtype, device=W.device)
            for jj, c in enumerate(Hrows):
                H[:c.shape[0], jj * p:(jj + 1) * p] = c
            Hh = H.cpu().numpy()
            mu, S = np.linalg.eig(Hh[:m * p])
            with np.errstate(divide='ignore'):
                theta = 1.0 / mu
            order = _sort_smallest(theta)[:p]
            # Arnoldi residual of A^{-1}, || H_{m+1,m} E_m^T s || / |mu|: a cheap convergence monitor
            est = np.array([np.linalg.norm(Hh[m * p:, (m - 1) * p:] @ S[(m - 1) * p:, i]) * abs(theta[i]) for i in order[:k]])
            if verbose:
                print("  restart %d block %d: theta %s est %s" % (restart, m, np.round(theta[order[:k]], 8), est), flush=True)
            if (est.max() < tol and m >= min_blocks) or m == max_blocks:
                Vall = torch.cat(V[:m], dim=1)
                Sx = torch.from_numpy(np.ascontiguousarray(S[:, order])).to(Vall.dtype).to(Vall.device)
                X = Vall @ Sx
                X = X / torch.linalg.vector_norm(X, dim=0, keepdim=True)
                AX = apply_A(X[:, :k])
                th = (X[:, :k].conj() * AX).sum(dim=0)
                res = torch.linalg.vector_norm(AX - X[:, :k] * th[None, :], dim=0).cpu().numpy()
                if verbose:
                    print("  restart %d block %d: explicit residuals %s" % (restart, m, res), flush=True)
                if res.max() <= tol:
                    done = True
                    break
        info["restarts"] = restart
        info["block_solves"] = solves
        if done:
            break
    info["converged"] = done
    return th.cpu().numpy(), X[:, :k].contiguous(), res, info


def largest_hermitian_eigenpairs(apply_Op, X0, k, tol=1e-9, max_blocks=12, max_restarts=30, verbose=False):
    """k eigenpairs of largest modulus of a Hermitian operator.  X0 [n, p], p >= 1 (block size).  Returns
    (lam[k] numpy float, sorted as eigsh returns them: algebraically ascending; X [n, k]; residuals; info)."""
    import torch
    p = X0.shape[1]
    X = X0
    info = {"restarts": 0, "block_applies": 0}
    applies = 0
    keep = None
    for restart in range(max_restarts):
        V = []
        AV = []
        if keep is None:
            Q, _, _ = _orth(torch, X)
        else:
            Q, _, _ = _orth(torch, torch.cat([keep, X], dim=1))     # thick restart: Ritz vectors + the new direction block
        V.append(Q)
        for j in range(max_blocks):
            W = apply_Op(V[j])
            applies += 1
            AV.append(W)
            Vall = torch.cat(V, dim=1)
            AVall = torch.cat(AV, dim=1)
            T = (Vall.conj().T @ AVall).cpu().numpy()
            T = 0.5 * (T + T.conj().T)
            lam, S = np.linalg.eigh(T)
            order = np.argsort(-np.abs(lam), kind='stable')[:k]
            Sx = torch.from_numpy(np.ascontiguousarray(S[:, order])).to(Vall.dtype).to(Vall.device)
            Xr = Vall @ Sx
            Rr = AVall @ Sx - Xr * torch.from_numpy(lam[order]).to(Xr.device).to(Xr.dtype)[None, :]
            res = torch.linalg.vector_norm(Rr, dim=0).cpu().numpy() / np.maximum(np.abs(lam[order]), 1e-300)
            if verbose:
                print("  restart %d block %d: lam %s res %s" % (restart, j + 1, lam[order], res), flush=True)
            if len(order) == k and res.max() <= tol:
                break
            if j + 1 == max_blocks:
                break
            Qn, _, _ = _orth(torch, W, V)
            V.append(Qn)
        info["restarts"] = restart
        info["block_applies"] = applies
        if len(order) == k and res.max() <= tol:
            break
        # thick restart: keep the k best Ritz vectors (+ a few more), continue with the residual directions of the worst
        kk = min(Vall.shape[1], k + p)
        order2 = np.argsort(-np.abs(lam), kind='stable')[:kk]
        S2 = torch.from_numpy(np.ascontiguousarray(S[:, order2])).to(Vall.dtype).to(Vall.device)
        keep = Vall @ S2
        X = (AVall @ S2 - keep * torch.from_numpy(lam[order2]).to(keep.device).to(keep.dtype)[None, :])[:, :p]
    asc = np.argsort(lam[order], kind='stable')
    info["converged"] = bool(res.max() <= tol)
    return lam[order][asc], Xr[:, torch.from_numpy(asc).to(Xr.device)].contiguous(), res[asc], info
