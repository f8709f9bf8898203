"""Drop-in for the reference's utils.py.  The per-sample kernel one_defl_Hutch_step
(utils.py:207-361) keeps its signature and additionally exists batched (defl_Hutch_batch):
k probes go through ONE fused device call (dmlmc_level_sample)."""
import os
import time

import numpy as np
from scipy.sparse.linalg import eigsh


# ---- utils.py:19-31 (kept for API parity; NOT a valid cost metric) --------
def flopsV_manual(bare_level, levels_info, level_id, mg_solver):
    if level_id == len(levels_info) - 2:
        if level_id == bare_level:
            return (2 * mg_solver.smooth_iters + 2) * levels_info[level_id].A.nnz + 0
        else:
            return (2 * mg_solver.smooth_iters + 1) * levels_info[level_id].A.nnz + 0
    else:
        if level_id == bare_level:
            return (2 * mg_solver.smooth_iters + 2) * levels_info[level_id].A.nnz + \
                   flopsV_manual(bare_level, levels_info, level_id + 1, mg_solver)
        else:
            return (2 * mg_solver.smooth_iters + 1) * levels_info[level_id].A.nnz + \
                   flopsV_manual(bare_level, levels_info, level_id + 1, mg_solver)


# ---- utils.py:36-69 ------------------------------------------------------------------------------
def print_post_results(A, params, result, example):
    if example == "mlmc":
        print(" -- matrix : " + params['matrix'])
        print(" -- matrix size : " + str(A.shape[0]) + "x" + str(A.shape[1]))
        print(" -- tr(A^{-1}) = " + str(result['trace']))
        cmplxity = result['total_complexity'] / (1.0e+6)
        print(" -- total MG complexity = " + str(cmplxity) + " MFLOPS")
        print(" -- std dev = ---")
        for i in range(result['nr_levels']):
            print(" -- level : " + str(i))
            print(" \t-- number of estimates = " + str(result['results'][i]['nr_ests']))
            print(" \t-- function iters = " + str(result['results'][i]['function_iters']))
            print(" \t-- trace = " + str(result['results'][i]['ests_avg']))
            print(" \t-- std dev = " + str(result['results'][i]['ests_dev']))
            print(" \t-- var = " + str(result['results'][i]['ests_dev'] * result['results'][i]['ests_dev']))
            cmplxity = result['results'][i]['level_complexity'] / (1.0e+6)
            print("\t-- level MG complexity = " + str(cmplxity) + " MFLOPS")
    elif example == "hutchinson":
        print(" -- matrix : " + params['matrix'])
        print(" -- matrix size : " + str(A.shape[0]) + "x" + str(A.shape[1]))
        print(" -- tr(A^{-1}) = " + str(result['trace']))
        cmplxity = result['total_complexity'] / (1.0e+6)
        print(" -- total MG complexity = " + str(cmplxity) + " MFLOPS")
        print(" -- std dev = " + str(result['std_dev']))
        print(" -- var = " + str(result['std_dev'] * result['std_dev']))
        print(" -- number of estimates = " + str(result['nr_ests']))
        print(" -- function iters = " + str(result['function_iters']))
    else:
        raise Exception("Value for parameter <example> not available.")


# ---- utils.py:73-125 ------------------------------------------------------------------------------
_B200_KEYS = ('probe_batch', 'smoother_degree', 'fgmres_restart', 'inner_precision', 'test_vectors',
              'deflation_eigpairs', 'mlmc_deflation_eigpairs', 'verbose', 'sequential_stop', 'device_probe_stream', 'pre_smooth',
              'mg_solver', 'geometric_precond', 'precond_degree')


def trace_params_from_params(params, example):
    if example not in ("mlmc", "hutchinson"):
        raise Exception("Value for parameter <example> not available.")
    trace_params = dict()
    function_params = dict()
    function_params['tol'] = params['function_tol']
    trace_params['function_params'] = function_params
    trace_params['tol'] = params['trace_tol']
    trace_params['max_nr_ests'] = 100000
    trace_params['max_nr_levels'] = params['max_nr_levels']
    trace_params['problem_name'] = params['matrix_params']['problem_name']
    trace_params['nr_deflat_vctrs'] = params['nr_deflat_vctrs']
    trace_params['defl_eigvs_tol_Hutch'] = params['defl_eigvs_tol_Hutch']
    if example == "mlmc":
        trace_params['mlmc_deflat_vctrs'] = params['mlmc_deflat_vctrs']
        trace_params['defl_eigvs_tol_MLMC'] = params['defl_eigvs_tol_MLMC']
        trace_params['diff_lev_op_tol'] = params['diff_lev_op_tol']
        trace_params['defl_type'] = params['defl_type']
        trace_params['coarsest_level_directly'] = params['coarsest_level_directly']
        trace_params['mlmc_levels_to_skip'] = params['mlmc_levels_to_skip']
    else:
        trace_params['defl-type'] = params['defl_type']            # sic, utils.py:113
        trace_params['defl_type'] = params['defl_type']
    trace_params['accuracy_mg_eigvs'] = params['accuracy_mg_eigvs']
    trace_params['aggrs'] = params['aggrs']
    trace_params['dof'] = params['dof']
    trace_params['use_permuted'] = params['use_permuted']
    trace_params['latt_dims'] = params['latt_dims']
    trace_params['x_displacement'] = params['x_displacement']
    trace_params['check_quality_MG'] = params['check_quality_MG']
    trace_params['test_vectors_type'] = params['test_vectors_type']
    for k in _B200_KEYS:
        if k in params:
            trace_params[k] = params[k]
    return trace_params


def device_deflation_eigenpairs(mg_solver, params, method, k, tolx, level_nr):
    """The eigensolves of utils.py:140,143 as block Lanczos on the device (eigensolve.largest_hermitian_eigenpairs): one
    operator application = one batched solve of p columns, where scipy's eigsh feeds the batched solver one vector at a time.
      "hutchinson": eigsh(Q, k, which='LM', sigma=0.0), Q = g3 A  ->  the k eigenvalues of Q closest to zero = largest of
                    Q^{-1} = A^{-1} g3, solved to 1e-2 * tolx;
      "mlmc":       eigsh(LinearOperator(diff_op_Q), k, which='LM')  ->  largest of (A_f^{-1} - P A_c^{-1} R) g3 with the
                    solves at diff_lev_op_tol, as the reference applies it.
    Deterministic (seeded start block).  Returns (Sy[k] ascending, Vx[n][k]) like eigsh."""
    import torch
    from . import eigensolve
    dev = mg_solver.dev
    lvl = 0 if method == "hutchinson" else level_nr
    n = mg_solver.level_shapes[lvl]
    p = int(params.get('deflation_block', min(max(k, 8), 32)))
    p = max(1, min(p, n // 4))
    gen = torch.Generator().manual_seed(777 + 13 * lvl + (0 if method == "hutchinson" else 1))
    X0 = torch.complex(torch.randn(n, p, dtype=torch.float64, generator=gen),
                       torch.randn(n, p, dtype=torch.float64, generator=gen)).to(dev.device)
    if method == "hutchinson":
        stol = min(1e-10, 1e-2 * tolx)
        h = n // 2

        def op(V):
            Vx = V.clone()
            Vx[h:] = -Vx[h:]
            X, _, _ = mg_solver.solve_batch(0, Vx.contiguous(), stol)
            return X
        lam, X, res, info = eigensolve.largest_hermitian_eigenpairs(op, X0, k, tol=tolx)
        Sy = 1.0 / lam
    else:
        mg_solver.level_for_diff_op = level_nr
        lam, X, res, info = eigensolve.largest_hermitian_eigenpairs(mg_solver.diff_op_Q_batch, X0, k, tol=tolx)
        Sy = lam
    if not info["converged"]:
        raise Exception("deflation eigensolver did not converge (%s, level %d): residuals %s" % (method, lvl, res))
    order = np.argsort(Sy, kind='stable')
    mg_solver.deflation_info = getattr(mg_solver, "deflation_info", {})
    mg_solver.deflation_info[(method, lvl)] = dict(info, residuals=res)
    from .multigrid import _same_on_all_ranks
    Sy = np.real(_same_on_all_ranks(np.asarray(Sy[order], dtype=np.complex128), mg_solver.device))
    return Sy, _same_on_all_ranks(X.cpu().numpy()[:, order], mg_solver.device)


# ---- utils.py:130-201 -----------------------------------------------------------------------------
def deflation_pre_computations(A, nr_deflat_vctrs, tolx, method, timer, params, mg_solver, lop=None, level_nr=0,
                               eigpairs=None):
    """Eigenpairs (host ARPACK as in the reference, utils.py:140,143 -- for "mlmc" every operator
    application is two device solves through mg_solver.diff_op_Q), sign fix, low-rank exact part.
    `eigpairs=(Sy, Vx)` injects the eigensolver output."""
    if nr_deflat_vctrs > 0:
        if eigpairs is not None:
            Sy, Vx = np.array(eigpairs[0]), np.array(eigpairs[1])
        elif method == "hutchinson":
            if params.get('host_eigensolver', False):
                Q = mg_solver.ml.levels[0].g3 * A
                Sy, Vx = eigsh(Q, k=nr_deflat_vctrs, which='LM', tol=tolx, sigma=0.0)
            else:
                Sy, Vx = device_deflation_eigenpairs(mg_solver, params, "hutchinson", nr_deflat_vctrs, tolx, 0)
        elif method == "mlmc":
            mg_solver.solve_tol = params['diff_lev_op_tol']
            if params.get('host_eigensolver', False):
                Sy, Vx = eigsh(lop, k=nr_deflat_vctrs, which='LM', tol=tolx)
            else:
                Sy, Vx = device_deflation_eigenpairs(mg_solver, params, "mlmc", nr_deflat_vctrs, tolx, level_nr)
        else:
            raise Exception("unknown method")
        sgnS = np.where(np.asarray(Sy) > 0, 1.0, -1.0)
        Sy = Sy * sgnS
        Ux = Vx * sgnS[None, :]
        if method == "hutchinson":
            Ux = mg_solver.ml.levels[0].g3 * Ux
            if params['use_permuted']:
                Ux = mg_solver.ml.levels[0].Pperm * Ux
        else:
            Vx = mg_solver.ml.levels[level_nr].g3 * Vx
        if os.getenv('OMP_NUM_THREADS') is None:
            raise Exception("Run : << export OMP_NUM_THREADS=N >>")          # utils.py:161-164
        mg_solver.solve_tol = params['function_params']['tol']
        d = np.einsum("ij,ij->j", np.conj(Ux), Vx)      # utils.py:173,176: `*` there is element-wise
        if method == "hutchinson":
            tr1 = np.sum(d / Sy)
        else:
            if params['defl_type'] == "exact":
                tr1 = np.sum(d * Sy)
            elif params['defl_type'] == "inexact_01":
                # utils.py:177-183: tr(Vx^H diff_op(Vx)) with the operator at the function tolerance, all columns in one batch
                mg_solver.level_for_diff_op = level_nr
                Vd = mg_solver._to_dev(np.ascontiguousarray(Vx))
                tr1 = complex((Vd.conj() * mg_solver.diff_op_batch(Vd)).sum().item())
            elif params['defl_type'] == "inexact_02":
                raise Exception("deflation type inexact_02 under construction")          # utils.py:184-185
            elif params['defl_type'] == "inexact_03":
                tr1 = 0.0                                                                # utils.py:186-187
            else:
                raise Exception("unknown deflation type")
    else:
        tr1 = 0.0
        Vx = None
        Ux = None
    if method == "hutchinson":
        return (Ux, tr1)
    return (Vx, Ux, tr1)


# ---- probes ------------------------------------------------------------------------------------------
from .sampling import draw_probe_bits  # noqa: E402  (one MT19937 word per element, LSB)


def pack_bits(bits01):
    return np.packbits(bits01, bitorder='little')


def _set_level_deflation(mg_solver, level, Vx, nr_deflat_vctrs):
    """Upload the deflation vectors of `level` once per distinct array."""
    cache = mg_solver.__dict__.setdefault("_defl_cache", {})
    key = None if (nr_deflat_vctrs == 0 or Vx is None) else id(Vx)
    if cache.get(level, "unset") != key:
        mg_solver.dev.set_deflation(level, None if key is None else np.asarray(Vx)[:, :nr_deflat_vctrs])
        cache[level] = key
        cache[("ref", level)] = Vx          # keep the array alive so id() stays unique


def defl_Hutch_batch(mg_solver, params, method, nr_deflat_vctrs, Vx, i, k, bits01=None, host_path=True, X0=None):
    """k samples of one_defl_Hutch_step in one device call.  bits01: k*n_i 0/1 values (probe-major);
    drawn from the global numpy stream if None.  X0: the probes as a complex128 CUDA tensor [n_i, k]
    (device-generated stream) instead of bits.  Returns (e[k] complex128, iters[2][k])."""
    n = mg_solver.level_shapes[i if method == "mlmc" else 0]
    if bits01 is None and X0 is None:
        bits01 = draw_probe_bits(k * n)
    lf = i if method == "mlmc" else 0
    if method == "mlmc":
        lc = lf + 2 if (mg_solver.skip_level and lf == 0) else lf + 1
    else:
        lc = lf
    _set_level_deflation(mg_solver, lf, Vx, nr_deflat_vctrs)
    tol = params['function_params']['tol']
    nlev = mg_solver.level_shapes[lf]
    maxiter = nlev if nlev < 1000 else 1000
    restart = min(mg_solver.restart, maxiter)
    if X0 is not None:
        e, iters = mg_solver.dev.level_sample(0 if method == "hutchinson" else 1, lf, lc, X0, tol, restart=restart,
                                              maxiter=maxiter)
        return e.cpu().numpy(), iters
    e, iters = mg_solver.dev.level_sample_host(0 if method == "hutchinson" else 1, lf, lc, pack_bits(bits01), k,
                                               tol, restart=restart, maxiter=maxiter)
    return e, iters


def exact_level_trace(mg_solver, params, method, i, k=256, nr_deflat_vctrs=0, Vx=None):
    """The EXACT value of what one_defl_Hutch_step estimates on level i: unit vectors instead of Rademacher
    probes through the same fused device call, e_j^H Op e_j summed over all j (the exact-trace validator of
    SURVEY.md 8f-4; gateway.py:100-104 quotes such a number for the shipped 128^2 set).
      method "hutchinson": tr(A_0^{-1} C (I - V V^H))
      method "mlmc"      : tr((A_i^{-1} - P A_c^{-1} R) C_i (I - V V^H)),  c = i+1 (or i+2 when level 1 is skipped)
    n_i solves in batches of k; returns a complex."""
    import torch
    dev = mg_solver.dev
    lf = i if method == "mlmc" else 0
    n = mg_solver.level_shapes[lf]
    if method == "mlmc":
        lc = lf + 2 if (mg_solver.skip_level and lf == 0) else lf + 1
    else:
        lc = lf
    _set_level_deflation(mg_solver, lf, Vx, nr_deflat_vctrs)
    tol = params['function_params']['tol']
    maxiter = n if n < 1000 else 1000
    restart = min(mg_solver.restart, maxiter)
    total = 0.0 + 0.0j
    for c0 in range(0, n, k):
        w = min(k, n - c0)
        X0 = torch.zeros((n, w), dtype=torch.complex128, device=dev.device)
        X0[c0:c0 + w, :] = torch.eye(w, dtype=torch.complex128, device=dev.device)
        e, _ = dev.level_sample(0 if method == "hutchinson" else 1, lf, lc, X0, tol, restart=restart, maxiter=maxiter)
        total += complex(e.sum().item())
        del X0
    return total


# ---- utils.py:207-361 ---------------------------------------------------------------------------------
def one_defl_Hutch_step(Af, Ac, mg_solver, params, method, nr_deflat_vctrs, Vx, Ux, i=0,
                        output_params=None, P=None, R=None, Pn=None, Rn=None):
    """One sample (k = 1 batch), same arguments / returns / side effects as the reference."""
    if method not in ("hutchinson", "mlmc"):
        raise Exception("unknown method")
    if method == "mlmc" and nr_deflat_vctrs > 0 and params['defl_type'] not in ("exact", "inexact_01"):
        if params['defl_type'] in ("inexact_02", "inexact_03"):
            raise Exception("deflation type " + params['defl_type'] + " under construction")
        raise Exception("unknown deflation type")
    e, iters = defl_Hutch_batch(mg_solver, params, method, nr_deflat_vctrs, Vx, i, 1)
    if method == "hutchinson":
        mg_solver.level_nr = 0
        mg_solver.num_iters = int(iters[0, 0])
        itrs = int(iters[0, 0])
    else:
        nl = len(mg_solver.ml.levels)
        lc = i + 2 if (mg_solver.skip_level and i == 0) else i + 1
        if output_params is not None:
            output_params['results'][i]['function_iters'] += int(iters[0, 0])
            output_params['results'][lc]['function_iters'] += int(iters[1, 0]) if lc < nl - 1 else 1
        itrs = 0
    if params.get('verbose', True):
        print('.', end='', flush=True)
    return (complex(e[0]), itrs)


# ---- utils.py:366-445 ------------------------------------------------------------------------------------
class CustomTimer:

    def __init__(self):
        self.reset()
        self.on = 0

    def reset(self):
        self.mvm = 0.0
        self.defl = 0.0
        self.P = 0.0
        self.R = 0.0
        self.mg_setup = 0.0
        self.defl_setup = 0.0
        self.axpy = 0.0
        self.tbuff = 0.0

    def start(self, part):
        if self.on == 1:
            raise Exception("Can't turn timer on, it's already timing")
        self.on = 1
        self.tbuff = time.time()

    def end(self, part):
        if self.on == 0:
            raise Exception("Can't turn timer off, it's already down")
        self.on = 0
        tot_t = time.time() - self.tbuff
        if part == "mvm":
            self.mvm += tot_t
        elif part == "defl":
            self.defl += tot_t
        elif part == "P":
            self.P += tot_t
        elif part == "R":
            self.R += tot_t
        elif part == "mg_setup":
            self.mg_setup += tot_t
        elif part == "defl_setup":
            self.defl_setup += tot_t
        elif part == "axpy":
            self.axpy += tot_t
        else:
            raise Exception("Uknown part to time")

    def __str__(self):
        str_out = ""
        str_out += "\nTimings specific to computations:\n"
        str_out += " -- matrix-vector multiplications : " + str(self.mvm) + "\n"
        str_out += " -- deflations : " + str(self.defl) + "\n"
        str_out += " -- applications of P : " + str(self.P) + "\n"
        str_out += " -- applications of R : " + str(self.R) + "\n"
        str_out += " -- applications of axpy : " + str(self.axpy) + "\n"
        str_out += " -- accumulated time : " + str(self.mvm + self.defl + self.P + self.R + self.mg_setup + self.defl_setup) + "\n"
        return str_out
