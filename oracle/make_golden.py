"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED
reference modules (oracle/ref_shim.py) in the authoring container, and cross-checks
oracle/refport.py against them on the same test vectors and probe stream.

    python -m oracle.make_golden            # ~3 min on one core

The reference has no tests / golden vectors of its own; the only
number in its tree is the exact trace of gateway.py:100-104, stored here too.
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np
from scipy.sparse import csr_matrix

from oracle import ref_shim, refport

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _ref_setup(ref, rec, p, method):
    with ref_shim.in_reference_dir():
        A = ref["matrix"].loadMatrix(p["matrix"], p["matrix_params"])
    tp = ref["utils"].trace_params_from_params(p, method)
    mg = ref["multigrid"].MG(A)
    rec.calls.clear()
    mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], dim=2,
             acc_eigvs=tp["accuracy_mg_eigvs"], sys_type="schwinger", params=tp)
    tvs = [c[2] for c in rec.calls]
    mg.total_levels = len(mg.ml.levels)
    for i in range(mg.total_levels - 1):
        mg.ml.levels[i].P = csr_matrix(mg.ml.levels[i].P)
        mg.ml.levels[i].R = csr_matrix(mg.ml.levels[i].R)
    return A, tp, mg, tvs


def _ref_probe(ref, mg, tp, method, i, nd, Vx, Ux, skip):
    lv = mg.ml.levels
    out = {"results": [{"function_iters": 0} for _ in range(len(lv))]}
    mg.skip_level = skip
    with _quiet():
        if method == "hutchinson":
            e, it = ref["utils"].one_defl_Hutch_step(lv[0].A, None, mg, tp, "hutchinson", nd, Vx, None)
        elif skip and i == 0:
            e, it = ref["utils"].one_defl_Hutch_step(lv[0].A, lv[2].A, mg, tp, "mlmc", nd, Vx, Ux, 0, out,
                                                     lv[0].P, lv[0].R, lv[1].P, lv[1].R)
        else:
            e, it = ref["utils"].one_defl_Hutch_step(lv[i].A, lv[i + 1].A, mg, tp, "mlmc", nd, Vx, Ux, i, out,
                                                     lv[i].P, lv[i].R)
    return e, np.array(mg.x)


def golden_16(ref, rec):
    g = {}
    for permuted in (False, True):
        tag = "perm" if permuted else "plain"
        p = ref_shim.params_16(permuted=permuted, nr_deflat_vctrs=16, mlmc_deflat_vctrs=(0, 0))
        A, tp, mg, tvs = _ref_setup(ref, rec, p, "mlmc")
        if not permuted:
            g["tv0"], g["tv1"] = tvs
            # Hutchinson deflation eigenpairs of Q = g3 A (utils.py:137-140)
            rec.calls.clear()
            with _quiet():
                Ux, tr1 = ref["utils"].deflation_pre_computations(A, 16, 1e-9, "hutchinson", mg.timer, tp, mg)
            g["defl_Sy"], g["defl_Vx"] = rec.calls[0][1], rec.calls[0][2]
            g["defl_Ux"], g["defl_tr1"] = Ux, tr1
        else:
            # same test vectors as the plain hierarchy
            rec2 = ref_shim.EigRecorder(replay=[("eigs", np.zeros(2), g["tv0"]), ("eigs", np.zeros(2), g["tv1"])])
            ref_shim.load_reference(rec2)
            A, tp, mg, tvs = _ref_setup(ref, rec2, p, "mlmc")
            ref_shim.load_reference(rec)
        mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
        mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp,
                 test_vectors=[g["tv0"], g["tv1"]])
        # per-probe: hutchinson (no deflation), 8 probes; mlmc level 0 and 1, 8 probes each
        np.random.seed(123456)
        rs = np.random.RandomState(123456)
        eh, zh = [], []
        for q in range(8):
            e, z = _ref_probe(ref, mg, tp, "hutchinson", 0, 0, None, None, False)
            e2, _ = refport.one_defl_hutch_step(mp.levels[0].A, None, mp, tp, "hutchinson", 0, None, None, rs)
            assert abs(e - e2) <= 1e-9 * abs(e), (e, e2)
            eh.append(e); zh.append(z)
        g[f"{tag}_hutch_e"], g[f"{tag}_hutch_z"] = np.array(eh), np.array(zh)
        for lvl in (0, 1):
            el = []
            for q in range(8):
                e, _ = _ref_probe(ref, mg, tp, "mlmc", lvl, 0, None, None, False)
                e2, _ = refport.one_defl_hutch_step(mp.levels[lvl].A, mp.levels[lvl + 1].A, mp, tp, "mlmc",
                                                    0, None, None, rs, lvl)
                assert abs(e - e2) <= 1e-9 * max(abs(e), 1.0), (e, e2)
                el.append(e)
            g[f"{tag}_mlmc_l{lvl}_e"] = np.array(el)
        if not permuted:
            # deflated hutchinson probes (16 vectors), stream restarts
            np.random.seed(123456)
            ed = []
            for q in range(8):
                e, _ = _ref_probe(ref, mg, tp, "hutchinson", 0, 16, g["defl_Ux"], None, False)
                ed.append(e)
            g["plain_hutch_defl16_e"] = np.array(ed)
        # full MLMC run through the reference driver (replay the same vectors)
        pr = ref_shim.params_16(permuted=permuted, nr_deflat_vctrs=0, mlmc_deflat_vctrs=(0, 0))
        rec2 = ref_shim.EigRecorder(replay=[("eigs", np.zeros(2), g["tv0"]), ("eigs", np.zeros(2), g["tv1"])])
        ref_shim.load_reference(rec2)
        with ref_shim.in_reference_dir():
            A = ref["matrix"].loadMatrix(pr["matrix"], pr["matrix_params"])
        tpr = ref["utils"].trace_params_from_params(pr, "mlmc")
        t = time.time()
        with _quiet():
            res = ref["stoch_trace"].mlmc(A, tpr)
        ref_shim.load_reference(rec)
        print(f"16^2 {tag} reference mlmc: {time.time()-t:.1f}s trace={res['trace']}",
              [r["nr_ests"] for r in res["results"]], file=sys.stderr)
        g[f"{tag}_mlmc_trace"] = res["trace"]
        g[f"{tag}_mlmc_nr_ests"] = np.array([r["nr_ests"] for r in res["results"]])
        g[f"{tag}_mlmc_ests_avg"] = np.array([r["ests_avg"] for r in res["results"]])
        g[f"{tag}_mlmc_ests_dev"] = np.array([r["ests_dev"] for r in res["results"]])
        # the port must reproduce the driver exactly (same stream, same stop rule)
        resp = refport.mlmc(refport.load_matrix(pr["matrix"], pr["matrix_params"]["mass"]), tpr,
                            test_vectors=[g["tv0"], g["tv1"]])
        assert [r["nr_ests"] for r in resp["results"]] == [r["nr_ests"] for r in res["results"]]
        assert abs(resp["trace"] - res["trace"]) <= 1e-9 * abs(res["trace"])
    # exact traces by dense algebra
    p = ref_shim.params_16()
    A = refport.load_matrix(p["matrix"], p["matrix_params"]["mass"])
    Ainv = np.linalg.inv(A.toarray())
    g["exact_trace"] = np.trace(Ainv)
    d0 = 16 * 2 * 2
    n = A.shape[0]
    g["exact_trace_perm"] = sum(Ainv[k, (k + d0) % n] for k in range(n))
    np.savez_compressed(os.path.join(OUT, "schwinger16.npz"), **g)


def golden_16_deflated_mlmc(ref, rec):
    """Deflated MLMC on 16^2 (mlmc_deflat_vctrs = [16, 16], the variant SURVEY.md 8d calls the valid deflated one):
    the reference's own deflation_pre_computations(method="mlmc") on its difference-level operator diff_op_Q
    (utils.py:141-158, multigrid.py:461-549) and 8 deflated level samples per level (utils.py:252-357 with
    nr_deflat_vctrs > 0).  Written to a separate file so that the other goldens stay byte-identical."""
    from scipy.sparse.linalg import LinearOperator
    g0 = np.load(os.path.join(OUT, "schwinger16.npz"))
    g = {}
    p = ref_shim.params_16(permuted=False, nr_deflat_vctrs=0, mlmc_deflat_vctrs=(16, 16))
    rec2 = ref_shim.EigRecorder(replay=[("eigs", np.zeros(2), g0["tv0"]), ("eigs", np.zeros(2), g0["tv1"])])
    ref_shim.load_reference(rec2)
    A, tp, mg, tvs = _ref_setup(ref, rec2, p, "mlmc")
    ref_shim.load_reference(rec)
    mg.skip_level = False
    mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
    mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp, test_vectors=[g0["tv0"], g0["tv1"]])
    for ix in range(2):
        mg.level_for_diff_op = ix
        lop = LinearOperator(mg.ml.levels[ix].A.shape, matvec=mg.diff_op_Q, dtype=np.complex128)
        rec.calls.clear()
        with _quiet():
            Vx, Ux, tr1 = ref["utils"].deflation_pre_computations(A, 16, tp["defl_eigvs_tol_MLMC"], "mlmc", mg.timer, tp,
                                                                  mg, lop, level_nr=ix)
        g[f"l{ix}_Sy"], g[f"l{ix}_eigvecs"] = rec.calls[0][1], rec.calls[0][2]     # raw eigsh output (injectable)
        g[f"l{ix}_Vx"], g[f"l{ix}_Ux"], g[f"l{ix}_tr1"] = Vx, Ux, tr1
        np.random.seed(123456 + ix)
        rs = np.random.RandomState(123456 + ix)
        el = []
        for q in range(8):
            e, _ = _ref_probe(ref, mg, tp, "mlmc", ix, 16, Vx, Ux, False)
            e2, _ = refport.one_defl_hutch_step(mp.levels[ix].A, mp.levels[ix + 1].A, mp, tp, "mlmc", 16, Vx, Ux, rs, ix)
            assert abs(e - e2) <= 1e-8 * max(abs(e), 1.0), (e, e2)
            el.append(e)
        g[f"l{ix}_e"] = np.array(el)
        print(f"16^2 deflated mlmc level {ix}: tr1 = {tr1}, first estimates {el[:2]}", file=sys.stderr)
    np.savez_compressed(os.path.join(OUT, "schwinger16_defl_mlmc.npz"), **g)


def golden_128(ref, rec):
    g = {}
    p = ref_shim.params_128()
    A, tp, mg, tvs = _ref_setup(ref, rec, p, "mlmc")
    g["tv0"], g["tv1"], g["tv2"] = tvs
    mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
    mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp, test_vectors=tvs)
    mp.skip_level = True
    g["level_sizes"] = np.array([l.A.shape[0] for l in mg.ml.levels])
    g["perm_shifts"] = np.array([l.perm_shift for l in mg.ml.levels])
    for i in range(3):
        Pi = mg.ml.levels[i].P.tocsr()
        Pi.sort_indices()
        g[f"P{i}_indices"] = Pi.indices.astype(np.int32)
        g[f"P{i}_data_head"] = Pi.data[:4096]
    np.random.seed(123456)
    rs = np.random.RandomState(123456)
    # stream order as in the reference driver: hutchinson probes first, then level 0, then level 2
    e, z = _ref_probe(ref, mg, tp, "hutchinson", 0, 0, None, None, True)
    e2, _ = refport.one_defl_hutch_step(mp.levels[0].A, None, mp, tp, "hutchinson", 0, None, None, rs)
    assert abs(e - e2) <= 1e-9 * abs(e)
    g["hutch_e"], g["hutch_z_sub"], g["hutch_z_norm"] = np.array([e]), z[::64].copy(), np.linalg.norm(z)
    el = []
    for q in range(3):
        e, z = _ref_probe(ref, mg, tp, "mlmc", 0, 0, None, None, True)
        e2, _ = refport.one_defl_hutch_step(mp.levels[0].A, mp.levels[2].A, mp, tp, "mlmc", 0, None, None, rs, 0)
        assert abs(e - e2) <= 1e-9 * abs(e), (e, e2)
        el.append(e)
        print("128^2 L0 probe", q, e, file=sys.stderr)
    g["mlmc_l0_e"] = np.array(el)
    el = []
    for q in range(16):
        e, z = _ref_probe(ref, mg, tp, "mlmc", 2, 0, None, None, True)
        e2, _ = refport.one_defl_hutch_step(mp.levels[2].A, mp.levels[3].A, mp, tp, "mlmc", 0, None, None, rs, 2)
        assert abs(e - e2) <= 1e-9 * max(abs(e), 1.0), (e, e2)
        el.append(e)
    g["mlmc_l2_e"] = np.array(el)
    lv = mg.ml.levels
    crst = lv[3].Pperm.transpose().conjugate() * (mg.coarsest_inv * lv[3].Bblock_perm)   # stoch_trace.py:433
    g["coarsest_term"] = np.trace(crst)
    g["exact_trace_perm"] = np.array(-8.748242701374695 + 50.215154098005584j)   # gateway.py:104
    g["exact_trace_plain"] = np.array(8326.432059538896)                          # SURVEY.md section 6 (sparse LU)
    np.savez_compressed(os.path.join(OUT, "schwinger128.npz"), **g)


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    rec = ref_shim.EigRecorder()
    ref = ref_shim.load_reference(rec)
    # probe stream pin
    np.random.seed(123456)
    first = np.random.randint(2, size=16) * 2 - 1
    assert list(first) == [1, -1, -1, 1, -1, 1, 1, 1, 1, -1, 1, -1, -1, -1, -1, -1]
    which = sys.argv[1:] or ["16", "128"]
    if "16" in which:
        golden_16(ref, rec)
    if "128" in which:
        golden_128(ref, rec)
    if "16defl" in which:
        golden_16_deflated_mlmc(ref, rec)
    print("golden written to", OUT, file=sys.stderr)


if __name__ == "__main__":
    main()
