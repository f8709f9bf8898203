"""Gauge-link <-> sparse-matrix converters and the synthetic U(1) generator.

Stencil identity (verified bit-exact on both reference matrices):
    row i = s*V + x*LT + t,  V = LX*LT,  s in {0,1} spin (slowest), t fastest
    S = 4 I - sum_mu [ (1-sigma_mu) U_mu(x) d_{x+mu,y} + (1+sigma_mu) U_mu(x-mu)^* d_{x-mu,y} ]
    mu=1 <-> t with sigma_1 = [[0,1],[1,0]],  mu=2 <-> x with sigma_2 = [[0,-i],[i,0]]
This is the input format of the level-0 kernel (2 complex links per site instead of 9
coefficients per row), cf. matrix.py:14-31 of the reference for the matrix it replaces.
"""
import numpy as np
import scipy.sparse as sp


class NotAStencil(Exception):
    """the matrix is not a 2-D Wilson-Dirac stencil on the given lattice in the reference's index layout"""


def wilson_matrix(links, mass=0.0):
    """Assemble A = S + m I (CSC, complex128) from links[2][LX][LT]."""
    links = np.asarray(links, dtype=np.complex128)
    Ut, Ux = links[0], links[1]
    LX, LT = Ut.shape
    V = LX * LT
    site = np.arange(V).reshape(LX, LT)
    tp, tm = np.roll(site, -1, axis=1), np.roll(site, 1, axis=1)
    xp, xm = np.roll(site, -1, axis=0), np.roll(site, 1, axis=0)
    Utb = np.conj(np.roll(Ut, 1, axis=1))
    Uxb = np.conj(np.roll(Ux, 1, axis=0))
    s = site.ravel()
    rows, cols, vals = [], [], []

    def add(sr, sc, col_site, v):
        rows.append(sr * V + s); cols.append(sc * V + col_site.ravel()); vals.append(np.asarray(v).ravel())

    d = np.full(V, 4.0 + mass, dtype=np.complex128)
    add(0, 0, site, d); add(1, 1, site, d)
    # forward t: -(1-s1) Ut = [[-U, U], [U, -U]]
    add(0, 0, tp, -Ut); add(0, 1, tp, Ut); add(1, 0, tp, Ut); add(1, 1, tp, -Ut)
    # backward t: -(1+s1) Ut(x-t)^*
    add(0, 0, tm, -Utb); add(0, 1, tm, -Utb); add(1, 0, tm, -Utb); add(1, 1, tm, -Utb)
    # forward x: -(1-s2) Ux = [[-U, -iU], [iU, -U]]
    add(0, 0, xp, -Ux); add(0, 1, xp, -1j * Ux); add(1, 0, xp, 1j * Ux); add(1, 1, xp, -Ux)
    # backward x: -(1+s2) Ux(x-x)^* = [[-U*, iU*], [-iU*, -U*]]
    add(0, 0, xm, -Uxb); add(0, 1, xm, 1j * Uxb); add(1, 0, xm, -1j * Uxb); add(1, 1, xm, -Uxb)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(2 * V, 2 * V)).tocsc()
    A.sum_duplicates()
    return A


def links_from_matrix(A, LX, LT):
    """Inverse of wilson_matrix: returns (links[2][LX][LT], diag) or raises if A is not a
    2-D Wilson-Dirac stencil on an LX x LT lattice in the reference's index layout."""
    A = sp.csr_matrix(A)
    V = LX * LT
    if A.shape != (2 * V, 2 * V):
        raise NotAStencil("matrix size does not match the lattice dimensions")
    site = np.arange(V).reshape(LX, LT)
    tp = np.roll(site, -1, axis=1).ravel()
    xp = np.roll(site, -1, axis=0).ravel()
    s = site.ravel()
    Ut = -np.asarray(A[s, tp]).ravel().reshape(LX, LT)
    Ux = -np.asarray(A[s, xp]).ravel().reshape(LX, LT)
    diag = A.diagonal()
    if not np.all(diag == diag[0]):
        raise NotAStencil("matrix diagonal is not constant: not a Wilson-Dirac stencil")
    links = np.stack([Ut, Ux])
    B = wilson_matrix(links, 0.0) + (diag[0] - 4.0) * sp.identity(2 * V, dtype=np.complex128, format="csc")
    D = (B - sp.csc_matrix(A))
    if D.nnz and np.abs(D.data).max() > 1e-13:
        raise NotAStencil("matrix is not a 2-D Wilson-Dirac stencil in the expected layout")
    return links, diag[0]


def random_u1_links(L, seed, sigma=0.204, LT=None):
    """Synthetic random-U(1) configuration: theta_mu(x) ~ N(0, sigma^2) i.i.d., U = exp(i theta).
    sigma = 0.204 gives <Re plaq> = exp(-2 sigma^2) ~ 0.92 like schwinger128.mat."""
    LT = L if LT is None else LT
    rng = np.random.default_rng(seed)
    theta = rng.normal(0.0, sigma, size=(2, L, LT))
    return np.exp(1j * theta)


def unpack_bf16_vectors(a):
    """uint16 [n][c][2] (real, imaginary parts as BF16 bit patterns: the top 16 bits of a float32) -> complex128 [n][c].
    The compact storage of the synthetic-lattice test vectors (tests/golden/synthetic256.npz)."""
    f = (np.asarray(a).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    return f[..., 0] + 1j * f[..., 1]
