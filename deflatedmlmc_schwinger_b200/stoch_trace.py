"""Drop-in for the reference's stoch_trace.py: hutchinson(A, params) and mlmc(A, params) with
the same parameter dictionaries and result dictionaries (stoch_trace.py:165-179, 309-321,
462-471).  Probes are evaluated in batches of params['probe_batch'] (default 256) per GPU by
the fused device call of utils.defl_Hutch_batch, sharded over torch.distributed ranks when a
process group is initialised (sampling.py); the statistics, the stopping rule and the probe
stream are the reference's."""
import time
from math import sqrt, pow

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.linalg import LinearOperator

from .multigrid import MG
from .utils import flopsV_manual, deflation_pre_computations, defl_Hutch_batch
from . import sampling


def _say(params, *a, **kw):
    if params.get('verbose', True):
        print(*a, **kw)


def _make_solver(A, params):
    """params['mg_solver'] (not in the reference): a hierarchy that is already set up for A (bench.py times the sampling of
    the whole experiment on the solver it has just timed); otherwise MG(...).setup(...) as stoch_trace.py:55-75."""
    mg_solver = params.get('mg_solver')
    prebuilt = mg_solver is not None
    if not prebuilt:
        mg_solver = MG(A, smoother_degree=params.get('smoother_degree', 80), restart=params.get('fgmres_restart', 40),
                       inner_precision=params.get('inner_precision', 'c64'), pre_smooth=params.get('pre_smooth', False),
                       geometric_precond=params.get('geometric_precond', True), precond_degree=params.get('precond_degree', 36))
    mg_solver.coarsest_iters = 0
    mg_solver.coarsest_iters_tot = 0
    mg_solver.coarsest_iters_avg = 0
    mg_solver.nr_calls = 0
    if not prebuilt:
        _say(params, "MG setup phase ...", end='', flush=True)
        start = time.time()
        params.setdefault('skip_unused_inverses', True)
        mg_solver.setup(dof=params['dof'], aggrs=params['aggrs'], max_levels=params['max_nr_levels'], dim=2,
                        acc_eigvs=params['accuracy_mg_eigvs'], sys_type=params['problem_name'], params=params)
        end = time.time()
        mg_solver.setup_seconds = end - start
        _say(params, " done. Time : " + str(end - start) + " seconds")
        _say(params, mg_solver)
    nr_levels = len(mg_solver.ml.levels)
    mg_solver.total_levels = nr_levels
    if nr_levels < 3:
        raise Exception("Use three or more levels.")
    for i in range(nr_levels):
        mg_solver.coarsest_lev_iters[i] = 0
    for i in range(nr_levels - 1):
        mg_solver.ml.levels[i].P = csr_matrix(mg_solver.ml.levels[i].P)
        mg_solver.ml.levels[i].R = csr_matrix(mg_solver.ml.levels[i].R)
    return mg_solver


def _sampler(mg_solver, params, method, nr_deflat_vctrs, Vx, level, k):
    def fn(probes):
        if isinstance(probes, np.ndarray):      # k*n 0/1 values drawn on the host
            e, iters = defl_Hutch_batch(mg_solver, params, method, nr_deflat_vctrs, Vx, level, k, bits01=probes)
        else:                                   # the probes themselves, generated on the device
            e, iters = defl_Hutch_batch(mg_solver, params, method, nr_deflat_vctrs, Vx, level, k, X0=probes)
        fn.coarse_iters += int(iters[1].sum())
        return e, iters[0], iters[1]
    fn.coarse_iters = 0
    return fn


def _probe_source(mg_solver, params):
    """params['device_probe_stream'] (default True): advance the MT19937 probe stream on the GPU"""
    if params.get('device_probe_stream', True):
        return sampling.DeviceProbeSource(mg_solver.dev)
    return sampling.HostProbeSource()


def _rough_trace(A, mg_solver, params, nr_deflat_vctrs, Vx, tr1, comm, k):
    """5 deflated-Hutchinson samples from a freshly seeded stream (stoch_trace.py:103-114, 288-301)."""
    np.random.seed(123456)
    nr_rough_iters = 5
    fn = _sampler(mg_solver, params, "hutchinson", nr_deflat_vctrs, Vx, 0, min(k, 8))
    res = sampling.run_sampling(fn, A.shape[0], min(k, 8), 0.0, nr_rough_iters, comm, fixed_count=nr_rough_iters,
                                probe_source=_probe_source(mg_solver, params))
    return np.sum(res["ests"][0:nr_rough_iters]) / nr_rough_iters + tr1


# compute tr(A^{-1}) via Hutchinson                                  (stoch_trace.py:33-179)
def hutchinson(A, params):
    mg_solver = _make_solver(A, params)
    N = A.shape[0]
    comm = sampling.Comm(mg_solver.dev.device)
    k = int(params.get('probe_batch', 256))

    _say(params, "\nResetting timer to zero ...", end='')
    mg_solver.timer.reset()
    _say(params, " done\n")
    nr_deflat_vctrs = params['nr_deflat_vctrs']
    tolx = params['defl_eigvs_tol_Hutch']
    _say(params, "Computing deflation vectors ...", end='', flush=True)
    start = time.time()
    Vx, tr1 = deflation_pre_computations(A, nr_deflat_vctrs, tolx, "hutchinson", mg_solver.timer, params, mg_solver,
                                         eigpairs=params.get('deflation_eigpairs'))
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")

    _say(params, "\nComputing rough estimation of the trace ...", end='', flush=True)
    start = time.time()
    rough_trace = _rough_trace(A, mg_solver, params, nr_deflat_vctrs, Vx, tr1, comm, k)
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")
    rough_trace_tol = abs(params['tol'] * rough_trace)

    mg_solver.timer.reset()
    _say(params, "\nComputing the trace stochastically ...", end='', flush=True)
    start = time.time()
    mg_solver.coarsest_lev_iters[0] = 0
    fn = _sampler(mg_solver, params, "hutchinson", nr_deflat_vctrs, Vx, 0, k)
    if params.get('sequential_stop', True):
        res = sampling.run_sampling(fn, N, k, rough_trace_tol, params['max_nr_ests'], comm,
                                    probe_source=_probe_source(mg_solver, params))
    else:
        res = sampling.run_sampling_fixed(fn, N, k, rough_trace_tol, params['max_nr_ests'], comm,
                                          probe_source=_probe_source(mg_solver, params))
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")

    result = dict()
    result['trace'] = res["avg"] + tr1
    result['std_dev'] = res["dev"]
    result['nr_ests'] = res["j_stop"]
    result['function_iters'] = res["iters_sum"]
    result['total_complexity'] = flopsV_manual(len(mg_solver.ml.levels), mg_solver.ml.levels, 0, mg_solver) * result['function_iters']
    result['total_complexity'] += mg_solver.ml.levels[len(mg_solver.ml.levels) - 1].A.nnz * result['function_iters']
    result['total_complexity'] += result['nr_ests'] * (2 * N * nr_deflat_vctrs) / 3.0
    result['rough_trace'] = rough_trace
    result['sampling_seconds'] = end - start
    result['probes_evaluated'] = res["evaluated"]
    return result


# compute tr(A^{-1}) via MLMC                                        (stoch_trace.py:185-471)
def mlmc(A, params):
    if len(params['mlmc_levels_to_skip']) > 1:
        raise Exception("Only allowed to skip one level for now")
    skip_level = len(params['mlmc_levels_to_skip']) == 1
    if skip_level and not params['mlmc_levels_to_skip'][0] == 1:
        raise Exception("Only allowed to skip the second level for now")

    mg_solver = _make_solver(A, params)
    N = A.shape[0]
    nr_levels = len(mg_solver.ml.levels)
    mg_solver.skip_level = skip_level
    comm = sampling.Comm(mg_solver.dev.device)
    k = int(params.get('probe_batch', 256))

    mg_solver.timer.reset()
    _say(params, "Computing deflation vectors ...", end='', flush=True)
    start = time.time()
    nr_deflat_vctrs = params['mlmc_deflat_vctrs']
    tolx = params['defl_eigvs_tol_MLMC']
    Vxs, Uxs, tr1s = [], [], []
    inj = params.get('mlmc_deflation_eigpairs')
    for ix in range(nr_levels - 1):
        if skip_level and ix == 1:
            Vxs.append([]); Uxs.append([]); tr1s.append(0.0)
            continue
        mg_solver.level_for_diff_op = ix
        lop = LinearOperator(mg_solver.ml.levels[ix].A.shape, matvec=mg_solver.diff_op_Q, dtype=np.complex128)
        Vx, Ux, tr1 = deflation_pre_computations(A, nr_deflat_vctrs[ix], tolx, "mlmc", mg_solver.timer, params, mg_solver,
                                                 lop, level_nr=ix, eigpairs=None if inj is None else inj[ix])
        Vxs.append(Vx); Uxs.append(Ux); tr1s.append(tr1)
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")

    _say(params, "Computing deflation vectors (for rough trace estimation purposes only) ...", end='', flush=True)
    start = time.time()
    Vx, tr1 = deflation_pre_computations(A, params['nr_deflat_vctrs'], params['defl_eigvs_tol_Hutch'], "hutchinson",
                                         mg_solver.timer, params, mg_solver, eigpairs=params.get('deflation_eigpairs'))
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")

    _say(params, "\nComputing rough estimation of the trace ...", end='', flush=True)
    start = time.time()
    rough_trace = _rough_trace(A, mg_solver, params, params['nr_deflat_vctrs'], Vx, tr1, comm, k)
    end = time.time()
    _say(params, " done. Time : " + str(end - start) + " seconds")

    output_params = dict()
    output_params['nr_levels'] = nr_levels
    output_params['trace'] = 0.0
    output_params['total_complexity'] = 0.0
    output_params['std_dev'] = 0.0
    output_params['results'] = list()
    for i in range(nr_levels):
        output_params['results'].append(dict())
        output_params['results'][i]['function_iters'] = 0
        output_params['results'][i]['nr_ests'] = 0
        output_params['results'][i]['ests_avg'] = 0.0
        output_params['results'][i]['ests_dev'] = 0.0
        output_params['results'][i]['level_complexity'] = 0.0

    # delta factors for MLMC                                          (stoch_trace.py:327-336)
    if nr_levels == 3:
        tol_fraction0, tol_fraction1 = 0.8, 0.2
    else:
        tol_fraction0, tol_fraction1 = 0.45, 0.45
    if skip_level:
        tol_fraction0 = tol_fraction0 + tol_fraction1

    mg_solver.timer.reset()
    mg_solver.coarsest_lev_iters[0] = 0
    sampling_seconds = 0.0
    probes_evaluated = []
    for i in range(nr_levels - 1):
        if skip_level and i == 1:
            continue
        start = time.time()
        if i == 0:
            tol_fctr = sqrt(tol_fraction0)
        elif i == 1:
            tol_fctr = sqrt(tol_fraction1)
        elif skip_level:
            tol_fctr = sqrt(1.0 - tol_fraction0) / sqrt(nr_levels - 3)
        else:
            tol_fctr = sqrt(1.0 - tol_fraction0 - tol_fraction1) / sqrt(nr_levels - 3)
        level_trace_tol = abs(params['tol'] * rough_trace * tol_fctr)
        lc = i + 2 if (skip_level and i == 0) else i + 1
        _say(params, "Computing for level " + str(i) + " ...", end='', flush=True)
        n_i = mg_solver.ml.levels[i].A.shape[0]
        fn = _sampler(mg_solver, params, "mlmc", nr_deflat_vctrs[i], Vxs[i], i, k)
        if params.get('sequential_stop', True):
            res = sampling.run_sampling(fn, n_i, k, level_trace_tol, params['max_nr_ests'], comm,
                                        probe_source=_probe_source(mg_solver, params))
        else:
            res = sampling.run_sampling_fixed(fn, n_i, k, level_trace_tol, params['max_nr_ests'], comm,
                                              probe_source=_probe_source(mg_solver, params))
        output_params['results'][i]['nr_ests'] += res["j_stop"]
        output_params['results'][i]['ests_avg'] = res["avg"] + tr1s[i]
        output_params['results'][i]['ests_dev'] = res["dev"]
        output_params['results'][i]['function_iters'] += res["iters_sum"]
        output_params['results'][lc]['function_iters'] += (res["coarse_iters_sum"] if lc < nr_levels - 1 else res["j_stop"] + 1)
        mg_solver.coarsest_lev_iters[i] += res["iters_sum"]
        end = time.time()
        sampling_seconds += end - start
        probes_evaluated.append(res["evaluated"])
        _say(params, " done. Time : " + str(end - start) + " seconds")

    # coarsest level, directly                                        (stoch_trace.py:421-437)
    if mg_solver.ml.levels[nr_levels - 1].A.shape[0] == 1:
        raise Exception("your coarsest-level matrix is of size 1 ... is this what you want?")
    if params['coarsest_level_directly'] == True:
        output_params['results'][nr_levels - 1]['nr_ests'] += 1
        crst_mat = mg_solver.coarsest_inv
        if params["use_permuted"]:
            crst_mat = mg_solver.ml.levels[nr_levels - 1].Pperm.transpose().conjugate() * \
                       (crst_mat * mg_solver.ml.levels[nr_levels - 1].Bblock_perm)
        output_params['results'][nr_levels - 1]['ests_avg'] = np.trace(crst_mat)
        output_params['results'][nr_levels - 1]['ests_dev'] = 0
    else:
        raise Exception("Stochastic coarsest-level computation is disabled at the moment.")

    # complexity bookkeeping                                           (stoch_trace.py:443-467)
    for i in range(nr_levels - 1):
        output_params['results'][i]['level_complexity'] = \
            output_params['results'][i]['function_iters'] * flopsV_manual(i, mg_solver.ml.levels, i, mg_solver)
        output_params['results'][i]['level_complexity'] += \
            mg_solver.ml.levels[len(mg_solver.ml.levels) - 1].A.nnz * mg_solver.coarsest_lev_iters[i]
    output_params['results'][nr_levels - 1]['level_complexity'] = \
        pow(mg_solver.ml.levels[nr_levels - 1].A.shape[0], 3) + \
        output_params['results'][nr_levels - 1]['function_iters'] * pow(mg_solver.ml.levels[nr_levels - 1].A.shape[0], 2)
    for i in range(nr_levels):
        output_params['total_complexity'] += output_params['results'][i]['level_complexity']
    for i in range(nr_levels):
        output_params['trace'] += output_params['results'][i]['ests_avg']
    output_params['rough_trace'] = rough_trace
    output_params['sampling_seconds'] = sampling_seconds
    output_params['probes_evaluated'] = probes_evaluated
    output_params['level_shapes'] = list(mg_solver.level_shapes)
    output_params['setup_seconds'] = getattr(mg_solver, 'setup_seconds', None)
    return output_params


# exact-trace validator (not in the reference, which only quotes the number, gateway.py:100-104)
def exact_trace(A, params):
    """EXACT tr(A^{-1} C) level by level: the hierarchy of mlmc(A, params), but every level's estimator is summed
    over all unit vectors instead of sampled (n_l batched device solves per level).  Returns the same dictionary
    layout as mlmc() with 'ests_avg' = the exact level values and 'ests_dev' = 0."""
    from .utils import exact_level_trace
    skip_level = len(params['mlmc_levels_to_skip']) == 1
    mg_solver = _make_solver(A, params)
    nr_levels = len(mg_solver.ml.levels)
    mg_solver.skip_level = skip_level
    k = int(params.get('probe_batch', 256))
    out = {'nr_levels': nr_levels, 'trace': 0.0, 'std_dev': 0.0, 'total_complexity': 0.0, 'results': []}
    for i in range(nr_levels):
        out['results'].append({'function_iters': 0, 'nr_ests': 0, 'ests_avg': 0.0, 'ests_dev': 0.0, 'level_complexity': 0.0})
    for i in range(nr_levels - 1):
        if skip_level and i == 1:
            continue
        n_i = mg_solver.ml.levels[i].A.shape[0]
        out['results'][i]['ests_avg'] = exact_level_trace(mg_solver, params, "mlmc", i, k=k)
        out['results'][i]['nr_ests'] = n_i
    crst_mat = mg_solver.coarsest_inv
    if params["use_permuted"]:
        crst_mat = mg_solver.ml.levels[nr_levels - 1].Pperm.transpose().conjugate() * \
                   (crst_mat * mg_solver.ml.levels[nr_levels - 1].Bblock_perm)
    out['results'][nr_levels - 1]['ests_avg'] = np.trace(crst_mat)
    out['results'][nr_levels - 1]['nr_ests'] = 1
    out['trace'] = sum(r['ests_avg'] for r in out['results'])
    return out
