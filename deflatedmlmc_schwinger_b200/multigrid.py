"""Drop-in for the reference's multigrid.py: class MG with the same attributes and methods
(setup / solve / one_mg_step / diff_op / diff_op_Q / matvec), whose solve path runs on the
GPU through libdmlmc_sm100.so (batched FGMRES + V-cycle, see csrc/dmlmc.cu).

The set-up (multigrid.py:100-344 of the reference) runs on the device as well since round 2: the
test-vector eigensolve of :174 (block Arnoldi on the batched solver, two-stage bootstrap on large
lattices), the per-aggregate classical Gram-Schmidt of :232-259 (closed-form index maps of
:192-227, bit-exact; never the dense n_l x n_{l+1} array of :200), the Galerkin products R A P of
:276 straight into the padded block-sparse layout of the coarse-level kernels, the dense coarsest
inverse of :342-344, the smoother polynomials.  The host keeps scipy / numpy copies of P, R, A_l
and coarsest_inv under the reference's attribute names, assembled from the device's numbers;
params['host_galerkin' | 'host_prolongator' | 'host_coarsest_inverse' | 'host_eigensolver' |
'host_smoother_setup'] restore the host computations one by one.

The smoother is NOT the reference's lgmres(maxiter=2) (multigrid.py:393-394): FGMRES is
flexible and parity is on the converged solution, so the V-cycle uses a
reduction-free fixed polynomial of A (Leja-ordered harmonic-Ritz roots of a degree-`smoother_degree`
GMRES polynomial computed at setup), applied in product form, one fused operator+update kernel per factor.
"""
import sys

import numpy as np
from scipy.sparse import csr_matrix, identity, diags
from scipy.sparse.linalg import eigs

from .utils import CustomTimer
from . import lattice
from . import _lib


def _same_on_all_ranks(arr, device=None):
    """rank 0's copy of a set-up array on every rank of an initialised process group (the ranks run the same deterministic
    eigensolver; the broadcast makes identical hierarchies a fact instead of an expectation)"""
    try:
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return arr
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
        t = torch.view_as_real(t).contiguous()
        dist.broadcast(t, src=0)
        return torch.view_as_complex(t).cpu().numpy()
    except ImportError:
        return arr


def _warn(msg):
    """to stderr, not through `warnings`: loadMatrix silences that module globally, as the reference does (matrix.py:16)"""
    print("[deflatedmlmc_schwinger_b200] WARNING: " + msg, file=sys.stderr, flush=True)


class LevelML:
    """multigrid.py:26-37"""
    R = 0
    P = 0
    A = 0
    Q = 0
    Pperm = 0
    perm_shift = 0
    Bblock_perm = 0
    g3 = 0


class SimpleML:
    """multigrid.py:39-48"""
    def __init__(self):
        self.levels = []

    def __str__(self):
        out = ""
        for idx, level in enumerate(self.levels[:-1]):
            out += "Level: " + str(idx) + "\n"
            out += "\tsize(R) = " + str(level.R.shape) + "\n"
            out += "\tsize(P) = " + str(level.P.shape) + "\n"
            out += "\tsize(A) = " + str(level.A.shape) + "\n"
        return out


# ------------------------------------------------------------------------------------------
# host helpers of the setup

def aggregation_maps(n, aggr_size, dofi, nvec):
    """Closed-form structure of P_l (multigrid.py:203-227): for fine row r returns
    (half[r], first_col[r]) with its nvec entries in columns first_col[r] + [0, nvec)."""
    r = np.arange(n)
    j = r // aggr_size
    q = (r % aggr_size) % dofi
    half = (q >= dofi // 2).astype(np.int64)
    return half, (2 * j + half) * nvec


def build_prolongator_values(eig_vecs, aggr_size, dofi, nvec):
    """Per-(aggregate, half) classical Gram-Schmidt (multigrid.py:232-259): the projections of column k
    on the already orthonormalised columns w<k are all taken from the unmodified column, subtracted in
    order, then the column is normalised.  The inner products run over the full aggregate column
    (the other half's rows are zero) with np.vdot, as in the reference, so that the values are
    bit-identical to the reference's for the same test vectors.  Returns pvals[n][nvec]."""
    n = eig_vecs.shape[0]
    na = n // aggr_size
    h = dofi // 2
    nw = aggr_size // dofi
    rel = [(np.arange(nw)[:, None] * dofi + np.arange(h)[None, :]).ravel() + half * h for half in (0, 1)]
    pv = np.zeros((n, nvec), dtype=np.complex128)
    ev = np.asarray(eig_vecs)
    sqrt = np.sqrt
    for j in range(na):
        base = j * aggr_size
        blk = np.zeros((aggr_size, 2 * nvec), dtype=np.complex128)
        blk[rel[0], :nvec] = ev[base + rel[0], :nvec]
        blk[rel[1], nvec:] = ev[base + rel[1], :nvec]
        for off in (0, nvec):
            for k in range(nvec):
                col = blk[:, off + k]
                rs = [np.vdot(blk[:, off + w], col) for w in range(k)]
                for w in range(k):
                    col -= rs[w] * blk[:, off + w]
                col /= sqrt(np.vdot(col, col).real)
        pv[base + rel[0], :] = blk[rel[0], :nvec]
        pv[base + rel[1], :] = blk[rel[1], nvec:]
    return pv


def prolongator_csr(pvals, aggr_size, dofi, nvec):
    n = pvals.shape[0]
    _, first = aggregation_maps(n, aggr_size, dofi, nvec)
    indices = (first[:, None] + np.arange(nvec)[None, :]).ravel()
    indptr = np.arange(n + 1) * nvec
    n_c = (n // aggr_size) * 2 * nvec
    return csr_matrix((pvals.ravel(), indices, indptr), shape=(n, n_c))


# ---- geometric aggregates of the preconditioner hierarchy ---------------------------------------
# The estimator's levels are the reference's (aggregates = runs of consecutive rows, multigrid.py:203-227).  The
# V-cycle that PRECONDITIONS the level-0 solve is free to use any hierarchy (FGMRES is flexible, parity is on the
# converged solution): with aggregates that are bx x bt blocks of lattice sites, split by spin, and the SAME test
# vectors, the two-grid method needs 8 outer iterations with a degree-32 smoother where the reference's
# aggregation needs 16 with degree 80 (128^2, tol 1e-12; CPU experiment profiles/exp_geometric_aggregation.py).

def geometric_blocks_level0(LX, LT, bx, bt):
    """coarse block of every level-0 row i = s*V + x*LT + t: ((x/bx)*(LT/bt) + t/bt)*2 + s"""
    s, x, t = np.meshgrid(np.arange(2), np.arange(LX), np.arange(LT), indexing='ij')
    return (((x // bx) * (LT // bt) + t // bt) * 2 + s).ravel().astype(np.int32)


def geometric_blocks_coarse(LXc, LTc, nv, bx, bt):
    """coarse block of every row ((X*LTc + T)*2 + half)*nv + v of a geometric coarse level"""
    X, T, h, v = np.meshgrid(np.arange(LXc), np.arange(LTc), np.arange(2), np.arange(nv), indexing='ij')
    return (((X // bx) * (LTc // bt) + T // bt) * 2 + h).ravel().astype(np.int32)


def geometric_blocks_level1(LX, LT, a_sites, nv1, bx=2):
    """coarse block of every row of the ESTIMATOR's level-1 operator (reference aggregation, multigrid.py:203-227, dofi = 2).
    Row r = (2 j + half) nv1 + v; strip j = s V/a + x LT/a + q is the run of a = a_sites consecutive rows (spin s, lattice
    column x, t in [q a, (q+1) a)), half = odd/even t.  Blocks: bx neighbouring strips in x, both halves, split by the spin s:
    ((x / bx) (LT / a) + q) 2 + s.  Returns (cblk, (LX / bx, LT / a))."""
    V = LX * LT
    nq = LT // a_sites
    n1 = (2 * V // a_sites) * 2 * nv1
    r = np.arange(n1)
    j = r // (2 * nv1)
    s = j // (V // a_sites)
    x = (j % (V // a_sites)) // nq
    q = j % nq
    return (((x // bx) * nq + q) * 2 + s).astype(np.int32), (LX // bx, nq)


def block_orthonormal_values(vecs, cblk, nvec):
    """pvals[n][nvec]: the first nvec columns of `vecs` orthonormalised within every coarse block (batched QR)"""
    vecs = np.asarray(vecs)[:, :nvec]
    n = vecs.shape[0]
    nb = int(cblk.max()) + 1
    m = n // nb
    order = np.argsort(cblk, kind='stable').reshape(nb, m)
    if m < nvec:
        raise Exception("geometric aggregation: aggregates smaller than the number of test vectors")
    Q, _ = np.linalg.qr(vecs[order])               # [nb, m, nvec]
    pv = np.zeros((n, nvec), dtype=np.complex128)
    pv[order.ravel()] = Q.reshape(nb * m, nvec)
    return pv


def prolongator_csr_indexed(pvals, cblk):
    n, nvec = pvals.shape
    indices = (cblk.astype(np.int64)[:, None] * nvec + np.arange(nvec)[None, :]).ravel()
    return csr_matrix((pvals.ravel(), indices, np.arange(n + 1) * nvec), shape=(n, (int(cblk.max()) + 1) * nvec))


def even_odd_schur(A0, LX, LT):
    """S = c - H_eo H_oe / c of the level-0 operator A = c I + H (row i = s*V + x*LT + t, parity (x + t) & 1): the even-odd
    Schur complement whose inverse the even-odd smoother approximates.  Returns (S as csr on the even sites, c)."""
    A0 = csr_matrix(A0)
    s, x, t = np.meshgrid(np.arange(2), np.arange(LX), np.arange(LT), indexing='ij')
    par = ((x + t) & 1).ravel()
    ie, io = np.where(par == 0)[0], np.where(par == 1)[0]
    c = A0.diagonal()[0]
    Aee = A0[ie][:, ie]
    if abs(Aee - c * identity(len(ie), format='csr')).max() > 1e-13 or abs(c.imag) > 0 or LX % 2 or LT % 2:
        raise Exception("even-odd smoother: the operator does not have the form c I + H with H odd-even")
    return (c * identity(len(ie), format='csr') - (A0[ie][:, io] @ A0[io][:, ie]) / c).tocsr(), c.real


def harmonic_ritz_inv_roots(A, degree, seed=7):
    """Inverse roots 1/theta_i of the degree-`degree` GMRES residual polynomial of A for a fixed
    random vector (harmonic Ritz values of an Arnoldi run), in Leja order for stability."""
    n = A.shape[0]
    degree = int(min(degree, n - 1))
    rs = np.random.RandomState(seed)
    b = rs.standard_normal(n) + 1j * rs.standard_normal(n)
    b /= np.linalg.norm(b)
    V = np.zeros((degree + 1, n), dtype=np.complex128)
    Vc = np.zeros((degree + 1, n), dtype=np.complex128)      # conj(V), kept so that no conjugated copy is made per step
    H = np.zeros((degree + 1, degree), dtype=np.complex128)
    V[0] = b
    Vc[0] = np.conj(b)
    for j in range(degree):
        w = A @ V[j]
        for _ in range(2):
            hh = Vc[:j + 1] @ w
            H[:j + 1, j] += hh
            w = w - hh @ V[:j + 1]
        H[j + 1, j] = np.linalg.norm(w)
        V[j + 1] = w / H[j + 1, j]
        Vc[j + 1] = np.conj(V[j + 1])
    return _harmonic_ritz_from_hessenberg(H, degree)


def _harmonic_ritz_from_hessenberg(H, degree):
    Hm = H[:degree, :degree]
    em = np.zeros(degree)
    em[-1] = 1.0
    f = np.linalg.solve(Hm.conj().T, em)
    theta = np.linalg.eigvals(Hm + (abs(H[degree, degree - 1]) ** 2) * np.outer(f, em))
    out = leja_order(theta)
    return 1.0 / np.array(out, dtype=np.complex128)


def harmonic_ritz_inv_roots_device(apply, n, degree, device, seed=7, support=None):
    """The same Arnoldi run with the operator applied on the device (`apply`: torch complex128 [n, 1] -> [n, 1], the level's
    SpMM kernel) and the basis kept there; the start vector is the host version's, so both give the same polynomial up to
    rounding.  The (degree + 1) x degree Hessenberg matrix comes back once at the end.  support: the operator acts on the
    rows `support` of the n-vectors only (the even sites of the even-odd Schur complement); the host version's start vector
    of that length is scattered there."""
    import torch
    m = n if support is None else int(support.shape[0])
    degree = int(min(degree, m - 1))
    rs = np.random.RandomState(seed)
    b = rs.standard_normal(m) + 1j * rs.standard_normal(m)
    b /= np.linalg.norm(b)
    V = torch.zeros((degree + 1, n), dtype=torch.complex128, device=device)
    H = torch.zeros((degree + 1, degree), dtype=torch.complex128, device=device)
    if support is None:
        V[0] = torch.from_numpy(b).to(device)
    else:
        V[0, support] = torch.from_numpy(b).to(device)
    for j in range(degree):
        w = apply(V[j].reshape(n, 1)).reshape(n)
        for _ in range(2):
            hh = torch.mv(V[:j + 1], w.conj_physical()).conj_physical()      # V^* w without a conjugated copy of the basis
            H[:j + 1, j] += hh
            w = w - hh @ V[:j + 1]
        nw = torch.linalg.vector_norm(w)
        H[j + 1, j] = nw
        V[j + 1] = w / nw
    return _harmonic_ritz_from_hessenberg(H.cpu().numpy(), degree)


class SmootherPolynomialError(Exception):
    """the product form of a smoother polynomial cannot be evaluated accurately at this degree / in this precision"""


def leja_order(points):
    """Leja ordering of complex points: start from the largest modulus, then repeatedly take the point
    that maximises the product of distances to the points already chosen (running log-products, O(d^2))."""
    pts = np.asarray(points, dtype=np.complex128).reshape(-1)
    m = pts.shape[0]
    left = np.ones(m, dtype=bool)
    first = int(np.argmax(np.abs(pts)))
    order = [first]
    left[first] = False
    logp = np.log(np.abs(pts - pts[first]) + 1e-300)
    for _ in range(m - 1):
        cand = np.where(left, logp, -np.inf)
        nxt = int(np.argmax(cand))
        order.append(nxt)
        left[nxt] = False
        logp = logp + np.log(np.abs(pts - pts[nxt]) + 1e-300)
    return pts[order]


def smoother_product_form(omega):
    """The smoother polynomial p(z) = sum_i omega_i prod_{j<i} (1 - omega_j z) (what d Richardson steps
    r <- r - omega_i A r, e <- e + omega_i r accumulate; 1 - z p(z) is the GMRES residual polynomial)
    rewritten as  p(z) = p0 * prod_{i<d-1} (1 - nu_i z):  each factor is then ONE kernel that reads a
    vector and writes a vector -- no accumulator to read and write, half the memory traffic.
    The roots 1/nu_i of p are the eigenvalues of the comrade matrix of p in the basis
    N_i(z) = prod_{j<i} (1 - omega_j z)  (z N_i = (N_i - N_{i+1}) / omega_i), Leja-ordered.
    Returns (nu[d-1], p0)."""
    w = np.asarray(omega, dtype=np.complex128).reshape(-1)
    d = w.shape[0]
    p0 = complex(np.sum(w))
    m = d - 1
    if m == 0:
        return np.zeros(0, dtype=np.complex128), p0
    C = np.zeros((m, m), dtype=np.complex128)
    for i in range(m):
        C[i, i] = 1.0 / w[i]
        if i + 1 < m:
            C[i, i + 1] = -1.0 / w[i]
    C[m - 1, :] += w[:m] / (w[m - 1] * w[m])
    mu = leja_order(np.linalg.eigvals(C))
    nu = 1.0 / mu
    # self-check on the spectrum-side sample points 1/omega_j (the harmonic Ritz values), where p = 1/z
    z = 1.0 / w
    ref = 1.0 / z
    val = p0 * np.prod(1.0 - np.outer(z, nu), axis=1)
    err = np.max(np.abs(val - ref) / np.abs(ref))
    if not err < 1e-6:
        raise SmootherPolynomialError("smoother polynomial: product form is inaccurate (%.2e); lower the degree" % err)
    return nu, p0


def _bf16_round(z):
    """complex64 array with real and imaginary parts rounded to BF16 (round to nearest even), as the device stores
    the smoother's intermediate vectors"""
    z = np.ascontiguousarray(z, dtype=np.complex64)
    u = z.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.complex64)


def smoother_storage_error(A, omega, nu, p0, storage):
    """Relative error of the product-form smoother  p0 prod (I - nu_i A) b  evaluated like the device does
    (complex64 arithmetic, intermediate vectors stored as `storage` = 'bf16' or 'f32') against the Richardson form
    in complex128, for one random vector.  High degrees on small levels can be unstable in 16-bit storage; the setup
    uses this to pick the storage (or lower the degree) per level."""
    n = A.shape[0]
    rs = np.random.RandomState(11)
    b = rs.standard_normal(n) + 1j * rs.standard_normal(n)
    b /= np.linalg.norm(b)
    r = b.copy()
    e = np.zeros_like(b)
    for i, wi in enumerate(omega):
        e = e + wi * r
        if i < len(omega) - 1:
            r = r - wi * (A @ r)
    A32 = A.astype(np.complex64)
    y = (64.0 * b).astype(np.complex64)
    for i, v in enumerate(nu):
        y = (y - np.complex64(v) * (A32 @ y)).astype(np.complex64)
        if storage == 'bf16' and i < len(nu) - 1:
            y = _bf16_round(y)
        if not np.all(np.isfinite(y)):
            return np.inf
    y = (p0 / 64.0) * y.astype(np.complex128)
    return float(np.linalg.norm(y - e) / np.linalg.norm(e))


def smoother_storage_error_device(dev, level, nu, p0, storage):
    """The same measurement with the device's own smoother kernels: p(A) b in the complex64 cycle's arithmetic with the
    intermediates stored as `storage` against the complex128 evaluation of the same product form (whose agreement with the
    Richardson form smoother_product_form has already checked)."""
    import torch
    n = dev.sizes[level]
    rs = np.random.RandomState(11)
    b = rs.standard_normal(n) + 1j * rs.standard_normal(n)
    b /= np.linalg.norm(b)
    B = torch.from_numpy(b.reshape(n, 1).repeat(2, axis=1)).to(dev.device).contiguous()      # two columns: the BF16 kernels pack pairs
    dev.set_smoother(level, nu, p0, storage16=(storage == 'bf16'))
    E = dev.smooth(level, B.to(torch.complex64).contiguous()).to(torch.complex128)
    Eref = dev.smooth(level, B)
    if not bool(torch.isfinite(E.real).all()) or not bool(torch.isfinite(E.imag).all()):
        return np.inf
    return float((torch.linalg.vector_norm(E - Eref) / torch.linalg.vector_norm(Eref)).item())


def even_odd_schur_device(dev, LX, LT, c):
    """S = c - H_eo H_oe / c of the stencil on level 0 of `dev` as a callable on FULL-lattice column batches that vanish on
    the odd sites (two applications of the level's SpMM kernel and two masks), for complex128 and complex64 input.  Returns
    (apply_S, even_rows): even_rows = the rows of the even sites in increasing order, the ordering of even_odd_schur's S."""
    import torch
    s, x, t = np.meshgrid(np.arange(2), np.arange(LX), np.arange(LT), indexing='ij')
    par = ((x + t) & 1).ravel()
    pe = torch.from_numpy((par == 0).astype(np.float64).reshape(-1, 1)).to(dev.device)
    po = 1.0 - pe
    masks = {torch.complex128: (pe, po), torch.complex64: (pe.to(torch.float32), po.to(torch.float32))}
    even_rows = torch.from_numpy(np.where(par == 0)[0]).to(dev.device)

    def apply_S(X):
        me, mo = masks[X.dtype]
        W = dev.spmm(0, X.contiguous()) * mo                 # H_oe x_e on the odd sites
        Z = dev.spmm(0, W.contiguous()) * me                 # H_eo H_oe x_e on the even sites
        return c * X - Z / c
    return apply_S, even_rows


def smoother_storage_error_op(apply, n, support, omega, nu, p0, device):
    """smoother_storage_error for an operator given as a device callable (complex128 and complex64 column batches [n, 1]
    supported on the rows `support`): the product form evaluated in complex64 with BF16-stored intermediates against the
    Richardson form in complex128, same start vector as the host version."""
    import torch
    m = int(support.shape[0])
    rs = np.random.RandomState(11)
    b = rs.standard_normal(m) + 1j * rs.standard_normal(m)
    b /= np.linalg.norm(b)
    B = torch.zeros((n, 1), dtype=torch.complex128, device=device)
    B[support, 0] = torch.from_numpy(b).to(device)
    r = B.clone()
    e = torch.zeros_like(B)
    for i, wi in enumerate(omega):
        e = e + complex(wi) * r
        if i < len(omega) - 1:
            r = r - complex(wi) * apply(r)
    y = (64.0 * B).to(torch.complex64)
    for i, v in enumerate(nu):
        y = y - complex(np.complex64(v)) * apply(y)
        if i < len(nu) - 1:
            y = torch.view_as_complex(torch.view_as_real(y).to(torch.bfloat16).to(torch.float32))
        if not bool(torch.isfinite(torch.view_as_real(y)).all()):
            return np.inf
    y = (complex(p0) / 64.0) * y.to(torch.complex128)
    return float((torch.linalg.vector_norm(y - e) / torch.linalg.vector_norm(e)).item())


def bsr_padded(A, bs):
    """scipy matrix -> (colidx[nb][bpr] with -1 padding, vals[nb][bpr][bs][bs])."""
    B = csr_matrix(A).tobsr(blocksize=(bs, bs))
    B.sort_indices()
    nb = B.shape[0] // bs
    counts = np.diff(B.indptr)
    bpr = int(counts.max())
    col = -np.ones((nb, bpr), dtype=np.int32)
    vals = np.zeros((nb, bpr, bs, bs), dtype=np.complex128)
    rowid = np.repeat(np.arange(nb), counts)
    pos = np.arange(B.indices.shape[0]) - np.repeat(B.indptr[:-1], counts)
    col[rowid, pos] = B.indices
    vals[rowid, pos] = B.data
    return col, vals


def csr_from_padded_bsr(col, vals):
    """(colidx[nb][bpr] with -1 padding, vals[nb][bpr][bs][bs]) -> scipy CSR (the inverse of bsr_padded)"""
    from scipy.sparse import bsr_matrix
    col = np.asarray(col)
    vals = np.asarray(vals)
    nb, bs = col.shape[0], vals.shape[2]
    mask = col >= 0
    indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int64)
    M = bsr_matrix((vals[mask], col[mask].astype(np.int64), indptr), shape=(nb * bs, nb * bs)).tocsr()
    M.eliminate_zeros()
    return M


def ell_padded(M):
    """scipy matrix -> (cols[n][w] with -1 padding, vals[n][w])."""
    M = csr_matrix(M)
    M.sort_indices()
    n = M.shape[0]
    counts = np.diff(M.indptr)
    w = int(max(counts.max(), 1))
    cols = -np.ones((n, w), dtype=np.int32)
    vals = np.zeros((n, w), dtype=np.complex128)
    rowid = np.repeat(np.arange(n), counts)
    pos = np.arange(M.indices.shape[0]) - np.repeat(M.indptr[:-1], counts)
    cols[rowid, pos] = M.indices
    vals[rowid, pos] = M.data
    return cols, vals


# ------------------------------------------------------------------------------------------

def _close_hierarchies(pm):
    """free the device side of a (temporary) hierarchy and of the hierarchies that precondition it"""
    for sub in [pm.precond_mg, pm.precond_mg1] + list(getattr(pm, "precond_mg_coarse", {}).values()) + [pm]:
        if sub is not None and getattr(sub, "dev", None) is not None:
            sub.dev.close()
            sub.dev = None


class MG:
    """Same public surface as the reference's MG (multigrid.py:56-557)."""

    def __init__(self, A, smooth_iters=2, smoother_degree=80, restart=40, inner_precision="c64",
                 device=None, dense_coarse_threshold=8192, pre_smooth=False, aggregation="reference",
                 geometric_precond=True, precond_degree=36, precond_blocks=(4, 4), precond_coarse_degree=None,
                 level0_block=1, precond_eo_degree=16, eo_degree=None):
        self.level_nr = 0
        self.ml = []
        self.A = A
        self.x = []
        self.num_iters = 0
        self.total_levels = 0
        self.coarsest_iters = 0
        self.coarsest_iters_tot = 0
        self.coarsest_iters_avg = 0
        self.nr_calls = 0
        self.smooth_iters = smooth_iters
        self.coarsest_lev_iters = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
        self.level_for_diff_op = 0
        self.solve_tol = 1.0e-1
        self.coarsest_inv = []
        self.timer = CustomTimer()
        self.skip_level = False
        # B200 specifics
        self.smoother_degree = smoother_degree
        self.restart = restart
        self.inner_precision = inner_precision
        self.device = device
        self.dense_coarse_threshold = dense_coarse_threshold
        self.pre_smooth = bool(pre_smooth)   # False: V-cycle = coarse correction + polynomial post-smoother
        self.dense_level = None              # level at which the V-cycle bottoms out with a dense inverse
        self.dev = None                      # _lib.Hierarchy
        self.test_vectors = []
        self.level_shapes = []
        # "reference": the aggregates of multigrid.py:203-227 (the estimator's levels).  "geometric": blocks of lattice
        # sites split by spin -- only for the hierarchy that preconditions the level-0 solve (self.precond_mg), which
        # the reference-aggregation instance builds for itself when geometric_precond is on and level 0 is a stencil.
        self.aggregation = aggregation
        self.geometric_precond = bool(geometric_precond) and aggregation == "reference"
        self.precond_degree = precond_degree
        self.precond_blocks = tuple(precond_blocks)
        self.precond_coarse_degree = 16 if precond_coarse_degree is None else precond_coarse_degree
        self.level0_block = int(level0_block)    # BSR block size of level 0 when it is not a Wilson stencil
        # even-odd (Schur complement) form of the level-0 post-smoother: degree of its polynomial on THIS hierarchy
        # (eo_degree) and on the geometric preconditioner hierarchy of level 0 (precond_eo_degree); None / 0 = not used
        self.eo_degree = eo_degree
        self.precond_eo_degree = precond_eo_degree
        self.eo_poly = None
        self.precond_mg = None                   # geometric hierarchy preconditioning the level-0 solve
        self.precond_mg1 = None                  # ... and the level-1 solve (lattices whose level 1 has no dense inverse)

    def set_option(self, name, value):
        """kernel / cycle option on this hierarchy and on its preconditioner hierarchy"""
        for pm in [self.precond_mg, self.precond_mg1] + list(getattr(self, "precond_mg_coarse", {}).values()):
            if pm is not None:
                pm.dev.set_option(name, value)
        self.dev.set_option(name, value)        # (also drops this hierarchy's CUDA graphs, which embed the preconditioner's kernels)

    # ---- multigrid.py:100-344 ---------------------------------------------------------------
    def setup(self, dof=[2, 8, 8], aggrs=[2 * 2, 2 * 2], max_levels=3, dim=2, acc_eigvs='low',
              sys_type='schwinger', params=None, test_vectors=None):
        if params is None:
            params = {}
        use_permuted = bool(params.get('use_permuted', False))
        tv_type = params.get('test_vectors_type', 'EVs')
        if tv_type != "EVs":
            raise Exception("only test_vectors_type='EVs' is supported (LSVs/RSVs are disabled in the reference set)")
        if test_vectors is None:
            test_vectors = params.get('test_vectors', None)

        Al = self.A.copy()
        ml = SimpleML()
        ml.levels.append(LevelML())
        ml.levels[0].A = Al.copy()
        self.test_vectors = []
        self._transfer_meta = []
        self._setup_args = dict(dof=list(dof), aggrs=list(aggrs), max_levels=max_levels, acc_eigvs=acc_eigvs)
        geometric = self.aggregation == "geometric"
        if geometric:
            use_permuted = False
            dims = params.get('latt_dims', None)
            if dims is None:
                Ls = int(round(np.sqrt(Al.shape[0] / 2)))
                dims = [Ls, Ls]
            geo = [dims[1] if len(dims) > 1 else dims[0], dims[0]]       # (LX, LT) of the current level's site lattice
            # level 0 of this hierarchy is not the spin-major site lattice (it is a coarse operator of the estimator):
            # the caller passes the blocks of its rows and the lattice of the resulting coarse blocks
            first = params.get('geometric_first', None)

        # The hierarchy is built ON the device (round 2): the level's operator is uploaded, the transfer operator's values are
        # orthonormalised there, and the Galerkin product R A P (multigrid.py:276) is formed there straight in the padded
        # block-sparse layout of the coarse-level kernels (dmlmc_galerkin); params['host_galerkin'] restores scipy's R*A*P.
        gdev = None
        self._gdev = None
        if not params.get('host_galerkin', False):
            gdev = _lib.Hierarchy(max_levels, self.device)
            self._set_level0_operator(gdev, Al, params)

        for i in range(max_levels - 1):
            dofi = dof[i] if i == 0 else int(dof[i] / 2)
            dofip1 = int(dof[i + 1] / 2)
            n = Al.shape[0]
            diag_g3 = np.ones(n, dtype=Al.dtype)
            diag_g3[int(n / 2):] = -1.0
            ml.levels[i].g3 = diags([diag_g3], [0])

            if use_permuted and i == 0:
                nt = params['latt_dims'][0]
                mat_disp = nt * 2 * params['x_displacement']
                ml.levels[0].perm_shift = mat_disp
                ml.levels[0].Pperm = diags([np.ones(n - mat_disp), np.ones(mat_disp)],
                                           [-mat_disp, n - mat_disp]).transpose()
                ml.levels[0].Bblock_perm = identity(n, dtype=Al.dtype)

            if acc_eigvs == 'low':
                tolx, ncvx = 1.0e-3, dofip1 + 2
            elif acc_eigvs == 'high':
                tolx, ncvx = 1.0e-9, None
            else:
                raise Exception("<accuracy_mg_eigvs> does not have a possible value.")

            if test_vectors is not None and i < len(test_vectors) and test_vectors[i] is not None:
                eig_vecs = np.asarray(test_vectors[i])
            elif geometric and i > 0:
                # the fine test vectors lie in range(P), so their restrictions are eigenvectors of R A P for the
                # same eigenvalues: no eigensolve on the coarse levels of the preconditioner hierarchy
                eig_vecs = ml.levels[i - 1].R @ self.test_vectors[i - 1]
            elif params.get('host_eigensolver', False):
                _, eig_vecs = eigs(Al, k=dofip1, which='LM', tol=tolx, maxiter=1000000, sigma=0.0, ncv=ncvx)
            else:
                hint = None
                if i > 0 and self.test_vectors[i - 1].shape[1] >= dofip1:
                    # A fine eigenvector lies in range(P), so its restriction is an eigenvector of R A P for the same
                    # eigenvalue: the start block of the coarse eigensolve (which then only has to confirm it)
                    hint = ml.levels[i - 1].R @ self.test_vectors[i - 1][:, :dofip1]
                eig_vecs = self.device_test_vectors(Al, dofip1, tolx, params, i, hint, gdev=gdev)
                eig_vecs = _same_on_all_ranks(eig_vecs, self.device)
            self.test_vectors.append(eig_vecs)

            if geometric:
                bx, bt = self.precond_blocks if i == 0 else (2, 2)
                dofip1 = min(dofip1, eig_vecs.shape[1])
                if i == 0 and first is not None:
                    cblk = np.asarray(first['cblk'], dtype=np.int32)
                    geo = [first['coarse_dims'][0] * bx, first['coarse_dims'][1] * bt]
                elif geo[0] % bx or geo[1] % bt:
                    raise Exception("geometric aggregation: lattice %dx%d is not divisible into %dx%d blocks" % (geo[0], geo[1], bx, bt))
                elif i == 0:
                    cblk = geometric_blocks_level0(geo[0], geo[1], bx, bt)
                else:
                    cblk = geometric_blocks_coarse(geo[0], geo[1], n // (2 * geo[0] * geo[1]), bx, bt)
                if cblk.shape[0] != n:
                    raise Exception("geometric aggregation: level size does not match the lattice")
                m_rows = n // (int(cblk.max()) + 1)
                if gdev is None or params.get('host_prolongator', False) or m_rows * dofip1 * 16 > 48 * 1024:
                    pvals = block_orthonormal_values(eig_vecs, cblk, dofip1)
                else:           # classical Gram-Schmidt with re-orthogonalisation, all blocks at once on the device
                    if m_rows < dofip1:
                        raise Exception("geometric aggregation: aggregates smaller than the number of test vectors")
                    pvals = gdev.block_orthonormal_values(eig_vecs, cblk, dofip1)[0].cpu().numpy()
                Pl = prolongator_csr_indexed(pvals, cblk)
                self._transfer_meta.append(("indexed", cblk, dofip1, pvals))
                geo = [geo[0] // bx, geo[1] // bt]
                ml.levels[i].P = Pl
                Rl = Pl.conjugate().transpose().tocsr()
                ml.levels[i].R = Rl
                if gdev is not None:
                    gdev.set_transfer_indexed(i, n, dofip1, pvals, cblk)
                    Al = self._device_galerkin(gdev, i, dofip1, Pl.shape[1])
                else:
                    Al = (Rl * Al * Pl).tocsr()
                ml.levels.append(LevelML())
                ml.levels[i + 1].A = Al.copy()
                continue

            aggr_size = aggrs[i] * dofi if i == 0 else aggrs[i] * dofi * 2
            if dofi < 2 or dofi % 2 or aggr_size % dofi or n % aggr_size or dofip1 < 1:
                raise Exception("inconsistent dof/aggrs for level " + str(i))
            if params.get('host_prolongator', False) or (aggr_size // 2) * dofip1 * 16 > 48 * 1024:
                pvals = build_prolongator_values(eig_vecs, aggr_size, dofi, dofip1)      # bit-identical to the reference
            else:
                # the same classical Gram-Schmidt, all aggregates at once on the device (values agree to ~1e-15)
                if gdev is None and getattr(self, "_setup_dev", None) is None:
                    self._setup_dev = _lib.Hierarchy(1, self.device)
                pvals = (gdev or self._setup_dev).prolongator_values(eig_vecs, aggr_size, dofi, dofip1).cpu().numpy()
            Pl = prolongator_csr(pvals, aggr_size, dofi, dofip1)
            self._transfer_meta.append((aggr_size, dofi, dofip1, pvals))
            ml.levels[i].P = Pl
            Rl = Pl.conjugate().transpose().tocsr()
            ml.levels[i].R = Rl
            if gdev is not None:
                gdev.set_transfer(i, n, aggr_size, dofi, dofip1, pvals)
                Al = self._device_galerkin(gdev, i, dofip1, Pl.shape[1])
            else:
                Al = (Rl * Al * Pl).tocsr()
            ml.levels.append(LevelML())
            ml.levels[i + 1].A = Al.copy()

            if use_permuted:
                mat_disp = int((ml.levels[i].perm_shift / (dof[i] * aggrs[i])) * dof[i + 1])
                ml.levels[i + 1].perm_shift = mat_disp
                nc = Pl.shape[1]
                ml.levels[i + 1].Pperm = diags([np.ones(nc - mat_disp), np.ones(mat_disp)],
                                               [-mat_disp, nc - mat_disp]).transpose()
                Bl = ml.levels[i].Pperm.transpose().conjugate() * (Pl * ml.levels[i + 1].Pperm)
                Bl = (Rl * ml.levels[i].Bblock_perm) * Bl
                ml.levels[i + 1].Bblock_perm = Bl

        if getattr(self, "_setup_dev", None) is not None:
            self._setup_dev.close()
            self._setup_dev = None
        self.ml = ml
        n_last = ml.levels[-1].A.shape[0]
        if gdev is not None and n_last <= 8192 and not params.get('host_coarsest_inverse', False):
            # multigrid.py:342-344 on the device (Gauss-Jordan with partial pivoting, dmlmc_dense_inverse)
            self.coarsest_inv = gdev.dense_inverse(np.asarray(ml.levels[-1].A.todense())).cpu().numpy()
        else:
            self.coarsest_inv = np.linalg.inv(np.asarray(ml.levels[-1].A.todense()))
        self.level_shapes = [l.A.shape[0] for l in ml.levels]
        self._gdev = gdev
        self._upload(params, use_permuted)

    def _device_galerkin(self, gdev, i, nvec, n_coarse):
        """A_{i+1} = R_i A_i P_i on the device (installed on level i + 1 of gdev); the scipy copy that the host-side consumers
        of ml.levels[i + 1].A read (the reference's attribute) is assembled from the same numbers."""
        col, vals = gdev.galerkin(i, nvec, n_coarse)
        self._bsr_on_device = getattr(self, "_bsr_on_device", {})
        self._bsr_on_device[i + 1] = (int(col.shape[1]), nvec)
        return csr_from_padded_bsr(col.cpu().numpy(), vals.cpu().numpy())

    def _links_of(self, dims):
        """lattice.links_from_matrix of the level-0 matrix self.A, once per matrix object (the check rebuilds the matrix from
        the links on the host; the hierarchies that precondition this one are built on the same object and share the answer)"""
        cache = getattr(self, "_links_cache", None)
        if cache is not None and cache[0] is self.A:
            if isinstance(cache[1], Exception):
                raise cache[1]
            return cache[1]
        try:
            out = lattice.links_from_matrix(self.A, dims[1] if len(dims) > 1 else dims[0], dims[0])
        except lattice.NotAStencil as why:
            self._links_cache = (self.A, why)
            raise
        self._links_cache = (self.A, out)
        return out

    def _set_level0_operator(self, dev, A0, params):
        """level 0 on the device: link form if A is a Wilson-Dirac stencil, else generic padded rows (block size level0_block)"""
        n0 = A0.shape[0]
        dims = params.get('latt_dims', None)
        if dims is None:
            L = int(round(np.sqrt(n0 / 2)))
            dims = [L, L]
        try:
            links, diag = self._links_of(dims)
        except lattice.NotAStencil as why:
            links = None
            if params.get('geometric_first') is None:
                # the caller's own matrix: say that the link-form kernels are not in use (orders of magnitude slower)
                _warn("level 0 is not a Wilson-Dirac stencil on the %s lattice (%s): generic block-sparse kernels are used"
                              % (dims, why))
        if links is not None:
            dev.set_stencil(0, links, diag)
            self.level0_format = "stencil"
            self._level0_diag = diag
        else:
            col, vals = bsr_padded(A0, self.level0_block)
            dev.set_bsr(0, n0, self.level0_block, col, vals)
            self.level0_format = "bsr%d" % self.level0_block
        return dims

    def device_test_vectors(self, Al, nvec, tol, params, level, hint=None, gdev=None):
        """multigrid.py:174 (`eigs(Al, k=nvec, which='LM', sigma=0.0, tol)`) on the device: block Arnoldi on A_l^{-1}
        (eigensolve.smallest_eigenpairs), every block step one batched FGMRES solve of p columns.  No hierarchy exists yet
        when the test vectors of a level are wanted, so the solver is the bootstrap one: FGMRES preconditioned by the
        level's own smoother polynomial (option precond_smoother_only) -- or, on large levels, by a geometric hierarchy built
        from rough vectors (two-stage bootstrap, below).  Deterministic: seeded start block, conjugate pairs cut by nvec keep
        the member with Im > 0 (scipy's eigs starts from a random vector and may return either).  gdev: the hierarchy under
        construction, whose level `level` already holds A_l (else a temporary one is made from the scipy matrix).
        Returns eig_vecs[n][nvec] (numpy); self.test_vector_info[level] holds eigenvalues, residuals and solve counts."""
        import torch
        from . import eigensolve
        n = Al.shape[0]
        dims = params.get('latt_dims', None)
        if dims is None:
            Ls = int(round(np.sqrt(self.A.shape[0] / 2)))
            dims = [Ls, Ls]
        own = gdev is None
        fmt = None
        if own:
            dev, lvl = _lib.Hierarchy(2, self.device), 0
            if level == 0 and self.aggregation == "reference":
                try:
                    links, diag = self._links_of(dims)
                    dev.set_stencil(0, links, diag)
                    fmt = "stencil"
                except lattice.NotAStencil:
                    fmt = None
            if fmt is None:
                bs = 1
                if level > 0 and self._transfer_meta and self._transfer_meta[level - 1][0] != "indexed":
                    bs = self._transfer_meta[level - 1][2]
                col, vals = bsr_padded(Al, bs)
                dev.set_bsr(0, n, bs, col, vals)
        else:
            dev, lvl = gdev, level
            if level == 0 and getattr(self, "level0_format", None) == "stencil":
                fmt = "stencil"
        deg = int(params.get('bootstrap_degree', 80))
        while True:
            try:
                if n >= 4096:
                    om = harmonic_ritz_inv_roots_device(lambda X: dev.spmm(lvl, X.contiguous()), n, deg, dev.device)
                else:
                    om = harmonic_ritz_inv_roots(csr_matrix(Al), deg)
                nu, p0 = smoother_product_form(om)
                break
            except SmootherPolynomialError:
                if deg <= 4:
                    raise
                deg = max(4, (3 * deg) // 4)
        dev.set_smoother(lvl, nu, p0, storage16=False)
        dev.set_inner_precision(_lib.C128)
        dev.set_option("precond_smoother_only", 1)
        p = int(params.get('eigensolver_block', max(4 * nvec, 16)))
        p = max(nvec, min(p, n // 4))
        gen = torch.Generator(device=dev.device).manual_seed(20240531 + level)      # (Philox: the same block on every rank)
        X0 = torch.complex(torch.randn(n, p, dtype=torch.float64, generator=gen, device=dev.device),
                           torch.randn(n, p, dtype=torch.float64, generator=gen, device=dev.device))
        if hint is not None:
            X0[:, :hint.shape[1]] = torch.from_numpy(np.ascontiguousarray(hint)).to(dev.device)
        # Two-stage bootstrap (large levels): FGMRES preconditioned by a polynomial alone needs ~200 iterations per solve
        # there.  Level 0 -- stage 1: the start block is smoothed, X <- orth(p(A) X), a few times (approximate inverse
        # iteration with no solves) and its Ritz vectors of smallest modulus serve as ROUGH test vectors of a geometric
        # hierarchy; stage 2: the block Arnoldi run below is preconditioned by that hierarchy's V-cycle.  Coarse levels -- the
        # restricted fine test vectors (`hint`: eigenvectors of R A P up to the accuracy of the fine ones) give the level's
        # geometric hierarchy directly; it is kept for the level's solves of the sampling phase (_build_level1_preconditioner).
        # The eigenvectors returned meet the same residual test either way.
        pm0 = None
        big = n >= int(params.get('two_stage_min_n', 100000)) and self.aggregation == "reference" and self.geometric_precond
        if big and level == 0 and fmt == "stencil" and hint is None and self._geometric_precond_possible(params):
            X = X0
            for _ in range(int(params.get('two_stage_smoothing_passes', 4))):
                X, _ = torch.linalg.qr(dev.smooth(lvl, X.contiguous()))
            Hs = (X.conj().T @ dev.spmm(lvl, X.contiguous())).cpu().numpy()
            th, S = np.linalg.eig(Hs)
            keep = np.argsort(np.abs(th), kind='stable')
            X0 = X @ torch.from_numpy(np.ascontiguousarray(S[:, keep])).to(X.device)
            X0 = X0 / torch.linalg.vector_norm(X0, dim=0, keepdim=True)
            rough = X0[:, :nvec].cpu().numpy()
            p2 = dict(params)
            p2['two_stage_min_n'] = 1 << 62
            pm0 = self._make_geometric_hierarchy(p2, rough)
        elif big and level >= 1 and hint is not None and hint.shape[1] >= nvec and \
                getattr(self, "level0_format", None) == "stencil" and level <= int(params.get('geometric_coarse_levels', 2)):
            LX, LT = (dims[1] if len(dims) > 1 else dims[0]), dims[0]
            pm0 = self._coarse_geometric_hierarchy(level, Al, np.ascontiguousarray(hint[:, :nvec]), params, LX, LT)
        if pm0 is not None:
            dev.set_option("precond_smoother_only", 0)
            dev.set_inner_precision(_lib.C64 if self.inner_precision == "c64" else _lib.C128)
            dev.set_preconditioner(lvl, pm0.dev, 0)
        maxit = n if n < 4000 else 4000
        restart = min(self.restart, maxit)
        solve_tol = min(1e-4, max(1e-12, 1e-2 * tol))      # the Arnoldi relation needs the solves two digits below the target
        stats = {"solves": 0, "iters": 0}

        def apply_Ainv(V):
            X, it, rr = dev.fgmres(lvl, V.contiguous(), solve_tol, restart=restart, maxiter=maxit)
            stats["solves"] += 1
            stats["iters"] += int(it.max())
            return X

        def apply_A(V):
            return dev.spmm(lvl, V.contiguous())
        theta, X, res, info = eigensolve.smallest_eigenpairs(apply_A, apply_Ainv, X0, nvec, tol=max(tol, 1e-11),
                                                             max_blocks=int(params.get('eigensolver_blocks', 16)),
                                                             min_blocks=3 if hint is not None else 1)
        if not info["converged"]:
            raise Exception("test-vector eigensolver did not converge on level %d (residuals %s)" % (level, res))
        info.update({"theta": theta, "residuals": res, "fgmres_iters": stats["iters"], "block": p, "bootstrap_degree": deg,
                     "two_stage": pm0 is not None})
        self.test_vector_info = getattr(self, "test_vector_info", {})
        self.test_vector_info[level] = info
        out = X.cpu().numpy()
        dev.release_workspace()
        dev.set_option("precond_smoother_only", 0)
        if pm0 is not None:
            dev.set_preconditioner(lvl, None)
            if level >= 1:
                self._early_coarse_pm = getattr(self, "_early_coarse_pm", {})
                self._early_coarse_pm[level] = pm0           # reused by _build_level1_preconditioner
            else:
                _close_hierarchies(pm0)
        if own:
            dev.close()
        return out

    def _upload(self, params, use_permuted):
        """Re-lay out the hierarchy for the device kernels and copy it to the GPU once."""
        lv = self.ml.levels
        nl = len(lv)
        A0 = lv[0].A
        n0 = A0.shape[0]
        on_dev = getattr(self, "_gdev", None) is not None          # operators and transfers are already there (setup)
        if on_dev:
            dev = self._gdev
            self._gdev = None
            dims = params.get('latt_dims', None)
            if dims is None:
                L = int(round(np.sqrt(n0 / 2)))
                dims = [L, L]
            diag = getattr(self, "_level0_diag", None)
        else:
            dev = _lib.Hierarchy(nl, self.device)
            dims = self._set_level0_operator(dev, A0, params)
            diag = getattr(self, "_level0_diag", None)
        for i in range(0 if not on_dev else nl - 1, nl - 1):
            if self._transfer_meta[i][0] == "indexed":
                _, cblk, nvec, pvals = self._transfer_meta[i]
                dev.set_transfer_indexed(i, lv[i].A.shape[0], nvec, pvals, cblk)
            else:
                aggr_size, dofi, nvec, pvals = self._transfer_meta[i]
                dev.set_transfer(i, lv[i].A.shape[0], aggr_size, dofi, nvec, pvals)
            if i + 1 < nl - 1:
                col, vals = bsr_padded(lv[i + 1].A, nvec)
                dev.set_bsr(i + 1, lv[i + 1].A.shape[0], nvec, col, vals)
        dev.set_coarsest_inverse(self.coarsest_inv)
        self.smoother_polys = []
        self.smoother_degrees_used = []
        self.smoother_storage = []
        for i in range(nl - 1):
            d = self.level_degree(i)
            Ai = csr_matrix(lv[i].A)
            while True:
                # the product form is checked against the Richardson form (exact arithmetic) and its evaluation in the
                # device's precisions against complex128 (a high degree on a small level can be unstable with BF16-
                # stored intermediates): fall back to FP32 storage, then to a lower degree
                try:
                    on_device = Ai.shape[0] >= 4096 and not params.get('host_smoother_setup', False)
                    if on_device:       # the level's operator is already on the device: Arnoldi and the storage check there
                        omega = harmonic_ritz_inv_roots_device(lambda X, _i=i: dev.spmm(_i, X.contiguous()), Ai.shape[0], d, dev.device)
                    else:
                        omega = harmonic_ritz_inv_roots(Ai, d)
                    nu, p0 = smoother_product_form(omega)
                    storage = None
                    for st in ('bf16', 'f32'):
                        err = smoother_storage_error_device(dev, i, nu, p0, st) if on_device else \
                            smoother_storage_error(Ai, omega, nu, p0, st)
                        if err < 0.15:
                            storage = st
                            break
                    if storage is None:
                        raise SmootherPolynomialError("smoother polynomial unstable in complex64")
                    break
                except SmootherPolynomialError:
                    if d <= 4:
                        raise
                    d = max(4, (3 * d) // 4)
            self.smoother_polys.append((nu, p0))
            self.smoother_degrees_used.append(d)
            self.smoother_storage.append(storage)
            dev.set_smoother(i, nu, p0, storage16=(storage == 'bf16'))
            if i == 0 and self.eo_degree and self.level0_format == "stencil" and storage == 'bf16':
                # polynomial in the even-odd Schur complement: degree d there ~ degree 2d in A at the cost of d applications
                LXs, LTs = (dims[1] if len(dims) > 1 else dims[0]), dims[0]
                self.eo_poly = None
                if LXs % 2 or LTs % 2 or abs(complex(diag).imag) > 0:
                    _warn("even-odd smoother not used: odd lattice extent or complex mass")
                else:
                    eo_on_device = n0 >= 4096 and not params.get('host_smoother_setup', False)
                    if eo_on_device:    # Arnoldi and storage check of the Schur complement through the device SpMM
                        apply_S, even_rows = even_odd_schur_device(dev, LXs, LTs, float(complex(diag).real))
                    else:
                        S, _c = even_odd_schur(lv[0].A, LXs, LTs)
                    de = int(self.eo_degree)
                    while de >= 2:
                        om = harmonic_ritz_inv_roots_device(apply_S, n0, de, dev.device, support=even_rows) if eo_on_device \
                            else harmonic_ritz_inv_roots(S, de)
                        try:
                            nue, p0e = smoother_product_form(om)
                        except SmootherPolynomialError:
                            de = (3 * de) // 4
                            continue
                        err_eo = smoother_storage_error_op(apply_S, n0, even_rows, om, nue, p0e, dev.device) if eo_on_device \
                            else smoother_storage_error(S, om, nue, p0e, 'bf16')
                        if err_eo < 0.15:
                            dev.set_smoother_eo(0, nue, p0e)
                            self.eo_poly = (nue, p0e)
                            break
                        de = (3 * de) // 4
                    if self.eo_poly is None:
                        _warn("even-odd smoother not used: no stable polynomial of degree <= %d" % int(self.eo_degree))
        if use_permuted:
            for i in range(nl):
                if i == 0:
                    dev.set_perm(0, lv[0].perm_shift)
                else:
                    cols, vals = ell_padded(lv[i].Bblock_perm)
                    dev.set_perm(i, lv[i].perm_shift, cols, vals)
        dev.set_inner_precision(_lib.C64 if self.inner_precision == "c64" else _lib.C128)
        dev.set_option("pre_smooth", 1 if self.pre_smooth else 0)
        self.dev = dev
        # V-cycle bottom: every intermediate level small enough gets a dense inverse, computed with the
        # batched device solver itself (A_l X = I), coarser levels first so that finer ones already use them.
        # Up to n = 4096 the inverse is kept in all precisions; larger ones only as the BF16 tensor-core
        # operand of the complex64 V-cycle (dmlmc_set_dense_inverse_device).
        self.dense_level = nl - 1
        self.dense_levels = {nl - 1: "host"}
        # params['skip_unused_inverses'] (the drivers set it): level 1 is neither sampled (mlmc_levels_to_skip = [1]) nor part of
        # the V-cycle that preconditions level 0 (the geometric hierarchy does that), so its dense inverse -- n_1 device solves
        # and n_1^2 BF16 numbers -- is not built; a solve on level 1, should one be asked for, cycles down to level 2 instead
        unused1 = bool(params.get('skip_unused_inverses', False)) and self.geometric_precond and self.level0_format == "stencil" \
            and list(params.get('mlmc_levels_to_skip', [])) == [1] and nl >= 4 and self._geometric_precond_possible(params)
        self.level1_unused = unused1
        for i in range(nl - 2, 0, -1):
            n_i = lv[i].A.shape[0]
            if n_i > self.dense_coarse_threshold or n_i % 8 or (i == 1 and unused1):
                break
            small = n_i <= 4096
            Minv = self._device_inverse(i, 1e-13 if small else 1e-6)     # the BF16 operand keeps 3 digits
            dev.set_dense_inverse_device(i, Minv, full=small)
            self.dense_levels[i] = "host" if small else "tensor"
            self.dense_level = i
            del Minv
        dev.release_workspace()
        import torch
        free_b, total_b = torch.cuda.mem_get_info(dev.device)
        if free_b < 0.5 * total_b:       # the library allocates its operators with cudaMalloc, outside torch's cache
            torch.cuda.empty_cache()
        if self.geometric_precond and self.level0_format == "stencil":
            self._build_geometric_preconditioner(params)

    def _geometric_precond_possible(self, params):
        dims = params.get('latt_dims', None)
        n0 = self.A.shape[0]
        if dims is None:
            Ls = int(round(np.sqrt(n0 / 2)))
            dims = [Ls, Ls]
        LX, LT = (dims[1] if len(dims) > 1 else dims[0]), dims[0]
        bx, bt = self.precond_blocks
        return not (LX % bx or LT % bt or 2 * LX * LT != n0 or (bx * bt) < self._setup_args['dof'][1] // 2)

    def _build_geometric_preconditioner(self, params):
        """A second hierarchy on geometric aggregates (precond_blocks sites at level 0, then 2 x 2, split by spin; same
        number of test vectors per level, level-0 test vectors shared with the estimator's hierarchy) whose V-cycle
        preconditions the level-0 FGMRES.  Skipped (the estimator's own hierarchy preconditions) when the lattice does
        not divide into the blocks."""
        pm = self._make_geometric_hierarchy(params, self.test_vectors[0])
        if pm is None:
            return
        dims = params.get('latt_dims', None)
        if dims is None:
            Ls = int(round(np.sqrt(self.A.shape[0] / 2)))
            dims = [Ls, Ls]
        LX, LT = (dims[1] if len(dims) > 1 else dims[0]), dims[0]
        self.precond_mg = pm
        self.dev.set_preconditioner(0, pm.dev, 0)
        self._build_level1_preconditioner(params, LX, LT)

    def _make_geometric_hierarchy(self, params, tv0):
        """the geometric hierarchy of level 0 built from the level-0 vectors tv0 (None when the lattice does not divide)"""
        dims = params.get('latt_dims', None)
        n0 = self.A.shape[0]
        if dims is None:
            Ls = int(round(np.sqrt(n0 / 2)))
            dims = [Ls, Ls]
        LX, LT = (dims[1] if len(dims) > 1 else dims[0]), dims[0]
        bx, bt = self.precond_blocks
        sa = self._setup_args
        if not self._geometric_precond_possible(params):
            return None
        # levels: blocks of bx x bt sites, then 2 x 2, until the level is small enough for a host-side dense inverse
        nv = [int(d // 2) for d in sa['dof'][1:]]
        gx, gt = LX // bx, LT // bt
        levels = 2
        while True:
            nvl = nv[min(levels - 2, len(nv) - 1)]
            if 2 * gx * gt * nvl <= 1024 or gx % 2 or gt % 2 or gx < 4 or gt < 4:
                break
            gx, gt = gx // 2, gt // 2
            levels += 1
        dof = [2] + [2 * nv[min(j, len(nv) - 1)] for j in range(levels - 1)]
        degs = [self.precond_degree] + [self.precond_coarse_degree] * (levels - 1)
        pm = MG(self.A, smoother_degree=degs, restart=self.restart, inner_precision=self.inner_precision,
                device=self.device, dense_coarse_threshold=self.dense_coarse_threshold, pre_smooth=self.pre_smooth,
                aggregation="geometric", precond_blocks=self.precond_blocks, eo_degree=self.precond_eo_degree)
        pm._links_cache = getattr(self, "_links_cache", None)
        p2 = dict(params)
        p2['use_permuted'] = False
        p2['latt_dims'] = [LT, LX]
        p2.pop('geometric_first', None)
        pm.setup(dof=dof, aggrs=[bx * bt] + [4] * (levels - 2), max_levels=levels, acc_eigvs=sa['acc_eigvs'],
                 params=p2, test_vectors=[tv0])
        return pm

    def _build_level1_preconditioner(self, params, LX, LT):
        """Geometric preconditioner hierarchies for the estimator's COARSE-level solves on lattices whose level l >= 1 is too
        large for a dense inverse.  A row of A_l is (strip group j, half, v): j = s V/a_l + x LT/a_l + q is the run of a_l
        consecutive t-sites (spin s, lattice column x, t in [q a_l, (q+1) a_l)) that the aggregates of multigrid.py:203-227
        have merged up to level l (a_1 = aggr_size of level 0, a_{l+1} = a_l times the strips per level-l aggregate).
        Geometric blocks: 2 neighbouring groups in x, both halves, split by the spin s; then 2 x 2.  CPU experiment at 128^2
        (exact two-grid method on A_1): 11 outer iterations at degree 32 against 33 (18 at degree 80) with the estimator's own
        level-2 aggregates.  Level 1 -> self.precond_mg1, deeper levels -> self.precond_mg_coarse[l].  A hierarchy that the
        set-up already built for the level's eigensolve (device_test_vectors, from the restricted fine test vectors) is used
        as it is."""
        lv = self.ml.levels
        self.precond_mg_coarse = {}
        early = getattr(self, "_early_coarse_pm", {})
        self._early_coarse_pm = {}
        try:
            if len(lv) < 3 or self._transfer_meta[0][0] == "indexed" or getattr(self, "level1_unused", False) and len(lv) < 4:
                return
            max_level = int(params.get('geometric_coarse_levels', 2))
            for l in range(1, len(lv) - 1):
                if l in self.dense_levels or l > max_level:
                    return
                if l == 1 and getattr(self, "level1_unused", False):
                    continue
                pm = early.pop(l, None)
                if pm is None:
                    pm = self._coarse_geometric_hierarchy(l, lv[l].A, self.test_vectors[l], params, LX, LT)
                if pm is None:
                    return
                if l == 1:
                    self.precond_mg1 = pm
                else:
                    self.precond_mg_coarse[l] = pm
                self.dev.set_preconditioner(l, pm.dev, 0)
        finally:
            for pm in early.values():          # built for an eigensolve, not needed by the sampling phase
                _close_hierarchies(pm)

    def _strip_sites(self, l):
        """a_l of _build_level1_preconditioner: t-sites per strip group of the rows of level l (None: other structure)"""
        if self._transfer_meta[0][0] == "indexed":
            return None
        aggr0, dofi0, _, _ = self._transfer_meta[0]
        if dofi0 != 2:
            return None
        a = aggr0                               # rows of one spin component per strip (dofi = 2: rows = sites of one spin)
        for m in range(2, l + 1):
            meta = self._transfer_meta[m - 1]
            if meta[0] == "indexed" or meta[0] % (2 * self._transfer_meta[m - 2][2]):
                return None
            a = a * (meta[0] // (2 * self._transfer_meta[m - 2][2]))
        return a

    def _coarse_geometric_hierarchy(self, l, A_l, tv_l, params, LX, LT):
        """the geometric hierarchy of the estimator's level l >= 1 (operator A_l, its test vectors tv_l); None if the level's
        rows do not have the strip structure"""
        sa = self._setup_args
        a_sites = self._strip_sites(l)
        if a_sites is None or LX % 2:
            return None
        nv_l = self._transfer_meta[l - 1][2]
        V = LX * LT
        n_l = A_l.shape[0]
        if LT % a_sites or n_l != (2 * V // a_sites) * 2 * nv_l or (LT // a_sites) < 1:
            return None
        cblk, (gx, gt) = geometric_blocks_level1(LX, LT, a_sites, nv_l, 2)
        nq = gt
        nvs = [int(d // 2) for d in sa['dof'][l + 1:]] or [nv_l]
        levels = 2
        while True:
            nvl = nvs[min(levels - 2, len(nvs) - 1)]
            if 2 * gx * gt * nvl <= 1024 or gx % 2 or gt % 2 or gx < 4 or gt < 4:
                break
            gx, gt = gx // 2, gt // 2
            levels += 1
        dof = [2 * nv_l] + [2 * nvs[min(jj, len(nvs) - 1)] for jj in range(levels - 1)]
        degs = [self.precond_degree] + [self.precond_coarse_degree] * (levels - 1)
        pm = MG(A_l, smoother_degree=degs, restart=self.restart, inner_precision=self.inner_precision,
                device=self.device, dense_coarse_threshold=self.dense_coarse_threshold, pre_smooth=self.pre_smooth,
                aggregation="geometric", precond_blocks=(1, 1), level0_block=nv_l)
        p2 = dict(params)
        p2['use_permuted'] = False
        p2['latt_dims'] = [LT, LX]
        p2['geometric_first'] = {'cblk': cblk, 'coarse_dims': (LX // 2, nq)}
        pm.setup(dof=dof, aggrs=[4] * (levels - 1), max_levels=levels, acc_eigvs=sa['acc_eigvs'], params=p2,
                 test_vectors=[tv_l])
        return pm

    def _device_inverse(self, level, tol, batch=1024):
        """A_level^{-1} as a torch complex128 CUDA tensor [n, n], solved in column batches on the device"""
        import torch
        dev = self.dev
        n = self.ml.levels[level].A.shape[0]
        Minv = torch.empty((n, n), dtype=torch.complex128, device=dev.device)
        kb = min(n, batch)
        for c0 in range(0, n, kb):
            w = min(kb, n - c0)
            B = torch.zeros((n, w), dtype=torch.complex128, device=dev.device)
            B[c0:c0 + w, :] = torch.eye(w, dtype=torch.complex128, device=dev.device)
            X, _, relres = dev.fgmres(level, B, tol, restart=min(self.restart, n), maxiter=n)
            if not np.all(relres < 10 * tol):
                raise Exception("dense coarse inverse: the device solve did not converge")
            Minv[:, c0:c0 + w] = X
            del B, X
        return Minv

    def level_degree(self, i):
        """smoother polynomial degree on level i (smoother_degree may be an int or a per-level list)"""
        d = self.smoother_degree
        if isinstance(d, (list, tuple)):
            return int(d[min(i, len(d) - 1)])
        return int(d)

    # ---- device-side batched API ---------------------------------------------------------------
    def _to_dev(self, v):
        import torch
        a = np.asarray(v, dtype=np.complex128)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev.device)

    def solve_batch(self, level, B, tol, maxiter=None):
        """B: torch complex128 CUDA tensor [n_level, k] -> (X, iters[k], relres[k])."""
        n = self.level_shapes[level]
        if maxiter is None:
            maxiter = n if n < 1000 else 1000                       # multigrid.py:354-357
        return self.dev.fgmres(level, B, tol, restart=min(self.restart, maxiter), maxiter=maxiter)

    # ---- multigrid.py:347-366 --------------------------------------------------------------------
    def solve(self, A, b, tol):
        B = self._to_dev(b)
        X, iters, _ = self.solve_batch(self.level_nr, B, tol)
        x = X.cpu().numpy()
        self.x = x.reshape(-1) if np.asarray(b).ndim == 1 else x
        self.num_iters = int(iters.max())

    # ---- multigrid.py:369-447 --------------------------------------------------------------------
    def one_mg_step(self, b):
        B = self._to_dev(b)
        X = self.dev.vcycle(self.level_nr, B)
        self.coarsest_lev_iters[self.level_nr] += 1
        self.nr_calls += 1
        x = X.cpu().numpy()
        return x.reshape(-1) if np.asarray(b).ndim == 1 else x

    def __str__(self):
        str_out = ""
        str_out += "\nMultilevel information:\n"
        for idx, level in enumerate(self.ml.levels):
            str_out += "Level: " + str(idx) + "\n"
            if idx < (len(self.ml.levels) - 1): str_out += "\tsize(R) = " + str(level.R.shape) + "\n"
            if idx < (len(self.ml.levels) - 1): str_out += "\tsize(P) = " + str(level.P.shape) + "\n"
            str_out += "\tsize(A) = " + str(level.A.shape) + "\n"
        return str_out

    # ---- multigrid.py:461-549 (the input vector is NOT modified, unlike :465-466) -----------------
    def diff_op_Q(self, v):
        vx = np.array(v, dtype=np.complex128, copy=True).reshape(-1)
        h = int(vx.shape[0] / 2)
        vx[h:] = -vx[h:]
        return self.diff_op(vx)

    def diff_op(self, v):
        """( Af^{-1} - P Ac^{-1} R ) v at tolerance self.solve_tol, on the device."""
        V = self._to_dev(np.asarray(v).reshape(-1))
        return self.diff_op_batch(V).cpu().numpy().reshape(-1)

    def diff_op_Q_batch(self, V):
        """diff_op_Q on a block: V torch complex128 CUDA tensor [n_l, p] -> [n_l, p]"""
        h = V.shape[0] // 2
        Vx = V.clone()
        Vx[h:] = -Vx[h:]
        return self.diff_op_batch(Vx)

    def diff_op_batch(self, V):
        """( Af^{-1} - P Ac^{-1} R ) V for a block of columns (torch CUDA tensor in and out)"""
        l = self.level_for_diff_op
        nl = len(self.ml.levels)
        skip = self.skip_level and l == 0
        lc = l + 2 if skip else l + 1
        V = V.contiguous()
        self.level_nr = l
        T1, _, _ = self.solve_batch(l, V, self.solve_tol)
        Vc = self.dev.restrict(l, V)
        if skip:
            Vc = self.dev.restrict(l + 1, Vc)
        if lc == nl - 1:
            T2 = self.dev.coarsest_apply(Vc)
        else:
            self.level_nr = lc
            T2, _, _ = self.solve_batch(lc, Vc, self.solve_tol)
        if skip:
            Tm = self.dev.torch.zeros((self.level_shapes[l + 1], V.shape[1]), dtype=T2.dtype, device=T2.device)
            self.dev.prolong_add(l + 1, T2, Tm)
            T2 = Tm
        W = self.dev.torch.zeros_like(T1)
        self.dev.prolong_add(l, T2, W)
        return T1 - W

    # ---- multigrid.py:552-557 -----------------------------------------------------------------------
    def matvec(self, x):
        X = self._to_dev(x)
        Y = self.dev.spmm(self.level_nr, X)
        y = Y.cpu().numpy()
        return y.reshape(-1) if np.asarray(x).ndim == 1 else y
