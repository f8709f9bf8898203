"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference algorithm.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker / the timed CPU baseline.
The product (deflatedmlmc_schwinger_b200/) never imports it.

Restates, function by function (file:line into /root/reference):
  load_matrix            matrix.py:14-31
  MGPort.setup           multigrid.py:100-344   (index maps :192-227, CGS :232-259,
                                                 R=P^H :267-274, RAP :276, perm :142-155,320-331,
                                                 coarsest inverse :342-344)
  MGPort.solve           multigrid.py:347-366   (pyamg fgmres restated in
                                                 oracle/shims/pyamg/krylov.py)
  MGPort.one_mg_step     multigrid.py:369-447   (scipy lgmres smoother, 2 cycles)
  MGPort.diff_op(_Q)     multigrid.py:461-549
  deflation_pre_computations   utils.py:130-201 (defl_type "exact" only)
  one_defl_hutch_step    utils.py:207-361
  hutchinson / mlmc      stoch_trace.py:33-179 / 185-471

Third-party arithmetic (absent from /root/reference, versions unpinned by it):
scipy.sparse kernels, scipy.sparse.linalg.{lgmres,eigs,eigsh} (called exactly as the
reference calls them -- the same scipy is installed on the GPU box), numpy, and
pyamg.krylov.fgmres (restated).  The reference holds NO tests or golden vectors for
this path except the exact-trace comment gateway.py:100-104, so parity is pinned by
(a) that number, (b) running the unmodified reference modules in the authoring
container (oracle/ref_shim.py) and comparing with this port on identical test
vectors and probe streams (oracle/make_golden.py writes tests/golden/*), and is
otherwise "parity unpinned" for the pyamg restatement.

Differences from the reference, all deliberate and result-neutral:
  * P is assembled per aggregate (same floating-point operations in the same order,
    bit-identical values) instead of through the dense n_l x n_{l+1} array of
    multigrid.py:200, which cannot be allocated beyond 128^2.
  * the probe stream comes from an explicit numpy RandomState (legacy MT19937, same
    words as the reference's global np.random usage).
  * test vectors / deflation vectors can be injected (hierarchies are only
    comparable with identical vectors).
  * no printing.
"""
import os
from math import sqrt

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator, eigs, eigsh, lgmres

from oracle.shims.pyamg.krylov import fgmres

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "deflatedmlmc_schwinger_b200", "data")


# ----------------------------------------------------------------------------------
# operator (matrix.py:14-31 + the stencil identity of SURVEY.md section 0)

def wilson_from_links(links):
    """S = 4 I - sum_mu [(1-sigma_mu) U_mu(x) d_{x+mu,y} + (1+sigma_mu) U_mu(x-mu)^* d_{x-mu,y}]
    row index i = s*V + x*L + t; mu=1 <-> t (sigma_1), mu=2 <-> x (sigma_2)."""
    Ut, Ux = links[0], links[1]
    L = Ut.shape[0]
    V = L * L
    X, T = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")

    def idx(s, x, t):
        return (s * V + (x % L) * L + (t % L)).ravel()

    rows, cols, vals = [], [], []
    s1 = np.array([[0, 1], [1, 0]], dtype=complex)
    s2 = np.array([[0, -1j], [1j, 0]], dtype=complex)
    one = np.eye(2, dtype=complex)
    Utb = np.conj(np.roll(Ut, 1, axis=1))     # U_t(x, t-1)^*
    Uxb = np.conj(np.roll(Ux, 1, axis=0))     # U_x(x-1, t)^*
    for s in (0, 1):
        rows.append(idx(s, X, T)); cols.append(idx(s, X, T)); vals.append(np.full(V, 4.0 + 0j))
        for sp_ in (0, 1):
            for (dx, dt, sig, sgn, U) in ((0, 1, s1, -1, Ut), (0, -1, s1, +1, Utb),
                                          (1, 0, s2, -1, Ux), (-1, 0, s2, +1, Uxb)):
                c = -(one[s, sp_] + sgn * sig[s, sp_])
                if c == 0:
                    continue
                rows.append(idx(s, X, T)); cols.append(idx(sp_, X + dx, T + dt))
                vals.append((c * U).ravel())
    S = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(2 * V, 2 * V)).tocsc()
    S.sum_duplicates()
    return S


def synthetic_links(L, seed, sigma=0.204):
    """theta_mu(x) ~ N(0, sigma^2) i.i.d. from numpy default_rng(seed), U = exp(i theta);
    links[mu][x][t], mu = 0 <-> t, 1 <-> x (the generator the synthetic-lattice parity fixtures are built from)"""
    rng = np.random.default_rng(seed)
    return np.exp(1j * rng.normal(0.0, sigma, size=(2, L, L)))


def load_matrix(matrix_name, mass):
    """matrix.py:14-31: A = S + m I.  The 16^2 file stores gamma3*S and the reference
    flips it back (matrix.py:25-27); the link fixture already describes S itself."""
    base = matrix_name[:-4] if matrix_name.endswith(".mat") else matrix_name
    links = np.load(os.path.join(_DATA, base + "_links.npy"))
    S = wilson_from_links(links)
    return (S + mass * sp.identity(S.shape[0], dtype=S.dtype, format="csc")).tocsc()


# ----------------------------------------------------------------------------------

class Level:
    """multigrid.py:26-37"""
    def __init__(self):
        self.R = self.P = self.A = 0
        self.Pperm = 0
        self.perm_shift = 0
        self.Bblock_perm = 0
        self.g3 = 0
        self.test_vectors = None


class MGPort:
    def __init__(self, A, smooth_iters=2):
        self.A = A
        self.level_nr = 0
        self.levels = []
        self.x = None
        self.num_iters = 0
        self.total_levels = 0
        self.smooth_iters = smooth_iters
        self.level_for_diff_op = 0
        self.solve_tol = 1.0e-1
        self.coarsest_inv = None
        self.skip_level = False
        self.coarsest_lev_iters = [0] * 10
        self.nr_spmv = 0

    # -- multigrid.py:100-344 -------------------------------------------------------
    def setup(self, dof, aggrs, max_levels, acc_eigvs, params, test_vectors=None):
        Al = self.A.copy()
        self.levels = [Level()]
        self.levels[0].A = Al.copy()
        for i in range(max_levels - 1):
            dofi = dof[i] if i == 0 else int(dof[i] / 2)
            dofip1 = int(dof[i + 1] / 2)
            n = Al.shape[0]
            diag_g3 = np.ones(n, dtype=Al.dtype)
            diag_g3[int(n / 2):] = -1.0
            self.levels[i].g3 = sp.diags([diag_g3], [0])

            if params["use_permuted"] and i == 0:
                nt = params["latt_dims"][0]
                mat_disp = nt * 2 * params["x_displacement"]
                self.levels[0].perm_shift = mat_disp
                self.levels[0].Pperm = sp.diags([np.ones(n - mat_disp), np.ones(mat_disp)],
                                                [-mat_disp, n - mat_disp]).transpose()
                self.levels[0].Bblock_perm = sp.identity(n, dtype=Al.dtype)

            if acc_eigvs == "low":
                tolx, ncvx = 1.0e-3, dofip1 + 2
            elif acc_eigvs == "high":
                tolx, ncvx = 1.0e-9, None
            else:
                raise Exception("<accuracy_mg_eigvs> does not have a possible value.")
            if params["test_vectors_type"] != "EVs":
                raise Exception("oracle port: only test_vectors_type='EVs' is restated")
            if test_vectors is not None:
                eig_vecs = np.asarray(test_vectors[i])
            else:
                _, eig_vecs = eigs(Al, k=dofip1, which="LM", tol=tolx, maxiter=1000000,
                                   sigma=0.0, ncv=ncvx)
            self.levels[i].test_vectors = eig_vecs

            aggr_size = aggrs[i] * dofi if i == 0 else aggrs[i] * dofi * 2
            nr_aggrs = int(n / aggr_size)
            h = int(dofi / 2)
            nw = int(int(aggr_size / 2) / (dofi / 2))
            # rows (relative to the aggregate) of the two "spin" halves, multigrid.py:207-227
            rel0 = np.array([w * dofi + z for w in range(nw) for z in range(h)])
            rel1 = rel0 + h
            P_rows, P_cols, P_vals = [], [], []
            for j in range(nr_aggrs):
                blk = np.zeros((aggr_size, 2 * dofip1), dtype=Al.dtype)
                base = j * aggr_size
                blk[rel0, :dofip1] = eig_vecs[base + rel0, :dofip1]
                blk[rel1, dofip1:] = eig_vecs[base + rel1, :dofip1]
                # classical Gram-Schmidt on full-aggregate columns, multigrid.py:232-259
                for off in (0, dofip1):
                    for k in range(dofip1):
                        rs = [np.vdot(blk[:, off + w], blk[:, off + k]) for w in range(k)]
                        for w in range(k):
                            blk[:, off + k] -= rs[w] * blk[:, off + w]
                        nrm = np.vdot(blk[:, off + k], blk[:, off + k])
                        blk[:, off + k] /= sqrt(nrm.real)
                for half, rel in ((0, rel0), (1, rel1)):
                    for k in range(dofip1):
                        P_rows.append(base + rel)
                        P_cols.append(np.full(rel.shape[0], j * dofip1 * 2 + half * dofip1 + k))
                        P_vals.append(blk[rel, half * dofip1 + k])
            Pl = sp.csr_matrix((np.concatenate(P_vals),
                                (np.concatenate(P_rows), np.concatenate(P_cols))),
                               shape=(n, nr_aggrs * dofip1 * 2), dtype=Al.dtype)
            self.levels[i].P = Pl
            Rl = Pl.conjugate().transpose().tocsr()
            self.levels[i].R = Rl
            Al = (Rl * Al * Pl)
            self.levels.append(Level())
            self.levels[i + 1].A = Al.copy()

            if params["use_permuted"]:
                mat_disp = int((self.levels[i].perm_shift / (dof[i] * aggrs[i])) * dof[i + 1])
                self.levels[i + 1].perm_shift = mat_disp
                nc = Pl.shape[1]
                self.levels[i + 1].Pperm = sp.diags([np.ones(nc - mat_disp), np.ones(mat_disp)],
                                                    [-mat_disp, nc - mat_disp]).transpose()
                Bl = self.levels[i].Pperm.transpose().conjugate() * (Pl * self.levels[i + 1].Pperm)
                Bl = (Rl * self.levels[i].Bblock_perm) * Bl
                self.levels[i + 1].Bblock_perm = Bl
        self.coarsest_inv = np.linalg.inv(np.asarray(self.levels[-1].A.todense()))
        self.total_levels = len(self.levels)

    # -- multigrid.py:552-557 -------------------------------------------------------
    def matvec(self, x):
        self.nr_spmv += 1
        return self.A * x

    # -- multigrid.py:347-366 -------------------------------------------------------
    def solve(self, A, b, tol):
        cnt = [0]

        def callback(xk):
            cnt[0] += 1
        maxiter = A.shape[0] if A.shape[0] < 1000 else 1000
        self.A = self.levels[self.level_nr].A
        lop1 = LinearOperator(A.shape, matvec=self.matvec, dtype=A.dtype)
        lop2 = LinearOperator(A.shape, matvec=self.one_mg_step, dtype=A.dtype)
        self.x, _ = fgmres(lop1, b, tol=tol, M=lop2, callback=callback, maxiter=maxiter)
        self.num_iters = cnt[0]

    # -- multigrid.py:369-447 -------------------------------------------------------
    def one_mg_step(self, b):
        lv = self.levels
        l0 = self.level_nr
        level_id = self.total_levels - l0
        dt = lv[l0].A.dtype
        rs = [np.zeros(lv[i].A.shape[0], dtype=dt) for i in range(l0, self.total_levels)]
        bs = [np.zeros(lv[i].A.shape[0], dtype=dt) for i in range(l0, self.total_levels)]
        xs = [np.zeros(lv[i].A.shape[0], dtype=dt) for i in range(l0, self.total_levels)]
        bs[0][:] = np.asarray(b).reshape(-1)
        i = -1
        for i in range(level_id - 1):
            Ai = lv[i + l0].A
            rs[i] = bs[i] - Ai * xs[i]
            self.A = Ai
            lop = LinearOperator(Ai.shape, matvec=self.matvec, dtype=dt)
            e, _ = lgmres(lop, rs[i], rtol=1.0e-20, maxiter=self.smooth_iters)
            self.A = lv[l0].A
            xs[i] += e
            rs[i] = bs[i] - Ai * xs[i]
            bs[i + 1] = lv[i + l0].R * rs[i]
        i += 1
        xs[i] = np.asarray(np.dot(self.coarsest_inv, bs[i])).reshape(-1)
        self.coarsest_lev_iters[l0] += 1
        for i in range(level_id - 2, -1, -1):
            Ai = lv[i + l0].A
            xs[i] += lv[i + l0].P * xs[i + 1]
            rs[i] = bs[i] - Ai * xs[i]
            self.A = Ai
            lop = LinearOperator(Ai.shape, matvec=self.matvec, dtype=dt)
            e, _ = lgmres(lop, rs[i], rtol=1.0e-20, maxiter=self.smooth_iters)
            self.A = lv[l0].A
            xs[i] += e
        return xs[0]

    # -- multigrid.py:461-549 (without the in-place mutation quirk of :465-466) ------
    def diff_op_Q(self, v):
        vx = np.array(v, copy=True).reshape(-1)
        h = int(vx.shape[0] / 2)
        vx[h:] = -vx[h:]
        return self.diff_op(vx)

    def diff_op(self, v):
        l = self.level_for_diff_op
        vx = np.asarray(v).reshape(-1)
        lv = self.levels
        skip = self.skip_level and l == 0
        lc = l + 2 if skip else l + 1
        Af, Ac = lv[l].A, lv[lc].A
        vc = lv[l + 1].R * (lv[l].R * vx) if skip else lv[l].R * vx
        self.level_nr = l
        self.solve(Af, vx, self.solve_tol)
        t1 = self.x
        if lc == len(lv) - 1:
            t2 = np.asarray(np.dot(self.coarsest_inv, vc)).reshape(-1)
        else:
            self.level_nr = lc
            self.solve(Ac, vc, self.solve_tol)
            t2 = self.x
        return t1 - (lv[l].P * (lv[l + 1].P * t2) if skip else lv[l].P * t2)


# ----------------------------------------------------------------------------------
# utils.py:130-201

def deflation_pre_computations(A, nr_deflat_vctrs, tolx, method, params, mg, lop=None,
                               level_nr=0, eigpairs=None):
    """eigpairs=(Sy, Vx) injects the eigensolver output (utils.py:140,143)."""
    if nr_deflat_vctrs > 0:
        if eigpairs is not None:
            Sy, Vx = np.array(eigpairs[0]), np.array(eigpairs[1])
        elif method == "hutchinson":
            Q = mg.levels[0].g3 * A
            Sy, Vx = eigsh(Q, k=nr_deflat_vctrs, which="LM", tol=tolx, sigma=0.0)
        else:
            mg.solve_tol = params["diff_lev_op_tol"]
            Sy, Vx = eigsh(lop, k=nr_deflat_vctrs, which="LM", tol=tolx)
        sgnS = np.where(Sy > 0, 1.0, -1.0)
        Sy = Sy * sgnS
        Ux = Vx * sgnS[None, :]
        if method == "hutchinson":
            Ux = mg.levels[0].g3 * Ux
            if params["use_permuted"]:
                Ux = mg.levels[0].Pperm * Ux
        else:
            Vx = mg.levels[level_nr].g3 * Vx
        mg.solve_tol = params["function_params"]["tol"]
        # NB utils.py:173,176: `*` between ndarrays is element-wise, so only the diagonal
        # of Ux^H Vx contributes to the trace
        d = np.einsum("ij,ij->j", np.conj(Ux), Vx)
        if method == "hutchinson":
            tr1 = np.sum(d / Sy)
        else:
            if params["defl_type"] != "exact":
                raise Exception("oracle port: only defl_type='exact' is restated")
            tr1 = np.sum(d * Sy)
    else:
        tr1, Vx, Ux = 0.0, None, None
    if method == "hutchinson":
        return (Ux, tr1)
    return (Vx, Ux, tr1)


# ----------------------------------------------------------------------------------
# utils.py:207-361

def rademacher(rs, n, dtype=np.complex128):
    """utils.py:213-216: randint(2,size=n)*2-1 -> one MT19937 word per element, LSB."""
    x = rs.randint(2, size=n)
    x *= 2
    x -= 1
    return x.astype(dtype)


def one_defl_hutch_step(Af, Ac, mg, params, method, nr_deflat_vctrs, Vx, Ux, rs, i=0,
                        iters=None, trace=None):
    """Returns (e, itrs).  `trace`, when a dict, receives the intermediates (x0, z, e1, e2)."""
    tol = params["function_params"]["tol"]
    lv = mg.levels
    if method == "hutchinson":
        x = rademacher(rs, Af.shape[0], Af.dtype)
        x_def = x - np.dot(Vx, np.dot(Vx.transpose().conjugate(), x)) if nr_deflat_vctrs > 0 else x
        mg.level_nr = 0
        rhs = lv[0].Pperm.transpose() * x_def if params["use_permuted"] else x_def
        mg.solve(Af, rhs, tol)
        z = mg.x
        e = np.vdot(x, z)
        if trace is not None:
            trace.update(x0=x, z=z, e=e, iters=mg.num_iters)
        return e, mg.num_iters

    x0 = rademacher(rs, Af.shape[0], Af.dtype)
    x_def = x0 - np.dot(Vx, np.dot(Vx.transpose().conjugate(), x0)) if nr_deflat_vctrs > 0 else x0
    mg.level_nr = i
    if params["use_permuted"]:
        x_perm = lv[i].Pperm.transpose() * x_def
        x_def = lv[i].Bblock_perm * x_perm
    mg.solve(Af, x_def, tol)
    z = mg.x
    it1 = mg.num_iters
    skip = mg.skip_level and i == 0
    lc = i + 2 if skip else i + 1
    xc = lv[i + 1].R * (lv[i].R * x_def) if skip else lv[i].R * x_def
    if lc == len(lv) - 1:
        y = np.asarray(np.dot(mg.coarsest_inv, xc)).reshape(-1)
        it2 = 1
    else:
        mg.level_nr = lc
        mg.solve(Ac, xc, tol)
        y = mg.x
        it2 = mg.num_iters
    if iters is not None:
        iters[i] += it1
        iters[lc] += it2
    e1 = np.vdot(x0, z)
    w = lv[i].P * (lv[i + 1].P * y) if skip else lv[i].P * y
    e2 = np.vdot(x0, w)
    if trace is not None:
        trace.update(x0=x0, z=z, w=w, e1=e1, e2=e2, e=e1 - e2, iters=(it1, it2))
    return e1 - e2, 0


# ----------------------------------------------------------------------------------
# stoch_trace.py

def _stats(ests, j):
    """stoch_trace.py:143-147 / 394-398"""
    avg = np.sum(ests[0:(j + 1)]) / (j + 1)
    dev = sqrt(np.sum(np.square(np.abs(ests[0:(j + 1)] - avg))) / (j + 1))
    return avg, dev, dev / sqrt(j + 1)


def _setup_mg(A, params, test_vectors):
    mg = MGPort(A)
    mg.setup(dof=params["dof"], aggrs=params["aggrs"], max_levels=params["max_nr_levels"],
             acc_eigvs=params["accuracy_mg_eigvs"], params=params, test_vectors=test_vectors)
    if len(mg.levels) < 3:
        raise Exception("Use three or more levels.")
    return mg


def hutchinson(A, params, test_vectors=None, defl_eigpairs=None, max_samples=None, log=None):
    """stoch_trace.py:33-179.  max_samples bounds the sampling loop (bench samples)."""
    mg = _setup_mg(A, params, test_vectors)
    nd = params["nr_deflat_vctrs"]
    Vx, tr1 = deflation_pre_computations(A, nd, params["defl_eigvs_tol_Hutch"], "hutchinson",
                                         params, mg, eigpairs=defl_eigpairs)
    rs = np.random.RandomState(123456)                     # stoch_trace.py:103
    ests = np.zeros(5, dtype=A.dtype)
    for i in range(5):
        ests[i], _ = one_defl_hutch_step(A, None, mg, params, "hutchinson", nd, Vx, None, rs)
    rough_trace = np.sum(ests) / 5 + tr1
    rough_tol = abs(params["tol"] * rough_trace)
    nmax = params["max_nr_ests"] if max_samples is None else max_samples
    ests = np.zeros(nmax, dtype=A.dtype)
    function_iters = 0
    for i in range(nmax):
        ests[i], itrs = one_defl_hutch_step(A, None, mg, params, "hutchinson", nd, Vx, None, rs)
        function_iters += itrs
        avg, dev, err = _stats(ests, i)
        if i >= 5 and err < rough_tol:
            break
    if log is not None:
        log.update(ests=ests[:i + 1].copy(), rough_trace=rough_trace, tr1=tr1)
    return {"trace": avg + tr1, "std_dev": dev, "nr_ests": i, "function_iters": function_iters}


def mlmc(A, params, test_vectors=None, defl_eigpairs=None, mlmc_eigpairs=None,
         max_samples=None, log=None):
    """stoch_trace.py:185-471 (complexity bookkeeping omitted: not a parity target)."""
    if len(params["mlmc_levels_to_skip"]) > 1:
        raise Exception("Only allowed to skip one level for now")
    skip_level = len(params["mlmc_levels_to_skip"]) == 1
    if skip_level and params["mlmc_levels_to_skip"][0] != 1:
        raise Exception("Only allowed to skip the second level for now")
    mg = _setup_mg(A, params, test_vectors)
    nl = len(mg.levels)
    mg.skip_level = skip_level
    ndv = params["mlmc_deflat_vctrs"]
    Vxs, Uxs, tr1s = [], [], []
    for ix in range(nl - 1):
        if skip_level and ix == 1:
            Vxs.append([]); Uxs.append([]); tr1s.append(0.0)
            continue
        mg.level_for_diff_op = ix
        lop = LinearOperator(mg.levels[ix].A.shape, matvec=mg.diff_op_Q, dtype=A.dtype)
        ep = None if mlmc_eigpairs is None else mlmc_eigpairs[ix]
        Vx, Ux, tr1 = deflation_pre_computations(A, ndv[ix], params["defl_eigvs_tol_MLMC"], "mlmc",
                                                 params, mg, lop, level_nr=ix, eigpairs=ep)
        Vxs.append(Vx); Uxs.append(Ux); tr1s.append(tr1)
    nd = params["nr_deflat_vctrs"]
    Vx, tr1 = deflation_pre_computations(A, nd, params["defl_eigvs_tol_Hutch"], "hutchinson",
                                         params, mg, eigpairs=defl_eigpairs)
    rs = np.random.RandomState(123456)                     # stoch_trace.py:288
    ests = np.zeros(5, dtype=A.dtype)
    for i in range(5):
        ests[i], _ = one_defl_hutch_step(A, None, mg, params, "hutchinson", nd, Vx, None, rs)
    rough_trace = np.sum(ests) / 5 + tr1
    out = {"nr_levels": nl, "trace": 0.0, "results": [
        {"function_iters": 0, "nr_ests": 0, "ests_avg": 0.0, "ests_dev": 0.0} for _ in range(nl)]}
    if nl == 3:
        f0, f1 = 0.8, 0.2
    else:
        f0, f1 = 0.45, 0.45
    if skip_level:
        f0 = f0 + f1
    iters = [0] * nl
    per_level = {}
    nmax = params["max_nr_ests"] if max_samples is None else max_samples
    for i in range(nl - 1):
        if skip_level and i == 1:
            continue
        if i == 0:
            fct = sqrt(f0)
        elif i == 1:
            fct = sqrt(f1)
        elif skip_level:
            fct = sqrt(1.0 - f0) / sqrt(nl - 3)
        else:
            fct = sqrt(1.0 - f0 - f1) / sqrt(nl - 3)
        level_tol = abs(params["tol"] * rough_trace * fct)
        Af = mg.levels[i].A
        Ac = mg.levels[i + 2].A if (skip_level and i == 0) else mg.levels[i + 1].A
        ests = np.zeros(nmax, dtype=Af.dtype)
        for j in range(nmax):
            ests[j], _ = one_defl_hutch_step(Af, Ac, mg, params, "mlmc", ndv[i], Vxs[i], Uxs[i],
                                             rs, i, iters)
            avg, dev, err = _stats(ests, j)
            if j >= 5 and err < level_tol:
                break
        out["results"][i]["nr_ests"] += j
        out["results"][i]["ests_avg"] = avg + tr1s[i]
        out["results"][i]["ests_dev"] = dev
        per_level[i] = ests[:j + 1].copy()
    if not params["coarsest_level_directly"]:
        raise Exception("Stochastic coarsest-level computation is disabled at the moment.")
    out["results"][nl - 1]["nr_ests"] += 1
    crst = mg.coarsest_inv
    if params["use_permuted"]:
        crst = mg.levels[nl - 1].Pperm.transpose().conjugate() * (crst * mg.levels[nl - 1].Bblock_perm)
    out["results"][nl - 1]["ests_avg"] = np.trace(crst)
    for i in range(nl):
        out["results"][i]["function_iters"] = iters[i]
        out["trace"] += out["results"][i]["ests_avg"]
    if log is not None:
        log.update(per_level=per_level, rough_trace=rough_trace, tr1=tr1)
    return out


def trace_params(params, example):
    """utils.py:73-125 (same whitelist copy)."""
    tp = {"function_params": {"tol": params["function_tol"]}, "tol": params["trace_tol"],
          "max_nr_ests": 100000}
    keys = ["max_nr_levels", "nr_deflat_vctrs", "defl_eigvs_tol_Hutch", "accuracy_mg_eigvs", "aggrs",
            "dof", "use_permuted", "latt_dims", "x_displacement", "check_quality_MG",
            "test_vectors_type", "defl_type"]
    if example == "mlmc":
        keys += ["mlmc_deflat_vctrs", "defl_eigvs_tol_MLMC", "diff_lev_op_tol",
                 "coarsest_level_directly", "mlmc_levels_to_skip"]
    for k in keys:
        tp[k] = params[k]
    tp["problem_name"] = params["matrix_params"]["problem_name"]
    return tp
