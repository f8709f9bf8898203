"""GPU tier: the drop-in drivers hutchinson() / mlmc() against the reference driver's results on
identical test vectors, probe stream and stop rule."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mlmc_16_matches_reference_driver(g16):
    from conftest import params16
    from deflatedmlmc_schwinger_b200 import matrix, stoch_trace, utils
    p = params16()
    p["test_vectors"] = [g16["tv0"], g16["tv1"]]
    p["probe_batch"] = 16
    p["smoother_degree"] = 8
    tp = utils.trace_params_from_params(p, "mlmc")
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    res = stoch_trace.mlmc(A, tp)
    assert [r["nr_ests"] for r in res["results"]] == list(g16["plain_mlmc_nr_ests"])        # same stop index
    assert abs(res["trace"] - g16["plain_mlmc_trace"]) < 1e-8 * abs(g16["plain_mlmc_trace"])
    for i in range(3):
        assert abs(res["results"][i]["ests_avg"] - g16["plain_mlmc_ests_avg"][i]) < 1e-8 * max(1.0, abs(g16["plain_mlmc_ests_avg"][i]))
        assert abs(res["results"][i]["ests_dev"] - g16["plain_mlmc_ests_dev"][i]) < 1e-7 * max(1.0, g16["plain_mlmc_ests_dev"][i])
    assert set(res.keys()) >= {"nr_levels", "trace", "total_complexity", "std_dev", "results"}
    assert set(res["results"][0].keys()) >= {"function_iters", "nr_ests", "ests_avg", "ests_dev", "level_complexity"}


def test_mlmc_16_permuted_matches_reference_driver(g16):
    from conftest import params16
    from deflatedmlmc_schwinger_b200 import matrix, stoch_trace, utils
    p = params16(permuted=True)
    p["test_vectors"] = [g16["tv0"], g16["tv1"]]
    p["probe_batch"] = 256
    p["smoother_degree"] = 8
    tp = utils.trace_params_from_params(p, "mlmc")
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    res = stoch_trace.mlmc(A, tp)
    assert [r["nr_ests"] for r in res["results"]] == list(g16["perm_mlmc_nr_ests"])          # 2006 / 1648 samples
    assert abs(res["trace"] - g16["perm_mlmc_trace"]) < 1e-8 * max(1.0, abs(g16["perm_mlmc_trace"]))
    # and it is statistically consistent with the exact displaced trace (dense algebra)
    sig = sum((r["ests_dev"] ** 2) / (r["nr_ests"] + 1) for r in res["results"][:-1]) ** 0.5
    assert abs(res["trace"] - g16["exact_trace_perm"]) < 5 * sig


def test_hutchinson_16_deflated(g16):
    from conftest import params16
    from deflatedmlmc_schwinger_b200 import matrix, stoch_trace, utils
    p = params16(nd=16)
    p["test_vectors"] = [g16["tv0"], g16["tv1"]]
    p["deflation_eigpairs"] = (g16["defl_Sy"], g16["defl_Vx"])
    p["probe_batch"] = 64
    p["smoother_degree"] = 8
    tp = utils.trace_params_from_params(p, "hutchinson")
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    res = stoch_trace.hutchinson(A, tp)
    assert set(res.keys()) >= {"trace", "std_dev", "nr_ests", "function_iters", "total_complexity"}
    err = res["std_dev"] / np.sqrt(res["nr_ests"] + 1)
    assert abs(res["trace"] - g16["exact_trace"]) < 5 * err + 1e-6
    # the first five samples of the rough estimate are the golden deflated-Hutchinson probes
    assert abs(res["rough_trace"] - (g16["plain_hutch_defl16_e"][:5].mean() + g16["defl_tr1"])) < 1e-7 * abs(res["rough_trace"])


def test_mlmc_128_level2_and_coarsest(mg128, g128):
    """shipped 128^2 set: the exact coarsest term and the level-2 difference sampled to its target"""
    from deflatedmlmc_schwinger_b200 import sampling, utils
    mg, tp, A = mg128
    lv = mg.ml.levels
    crst = lv[3].Pperm.transpose().conjugate() * (mg.coarsest_inv * lv[3].Bblock_perm)
    assert abs(np.trace(crst) - g128["coarsest_term"]) < 1e-8 * abs(g128["coarsest_term"])
    # level 2 difference: exact value by dense algebra vs the sampled mean
    A2inv = np.linalg.inv(lv[2].A.toarray())
    C2 = (lv[2].Bblock_perm @ lv[2].Pperm.transpose()).toarray()
    exact_l2 = np.trace(A2inv @ C2) - np.trace(crst)
    np.random.seed(2024)
    k = 512

    def fn(bits):
        e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, 2, k, bits01=bits)
        return e, it[0]
    res = sampling.run_sampling(fn, 2048, k, 0.5, 100000)
    err = res["dev"] / np.sqrt(res["j_stop"] + 1)
    print("level-2: N =", res["j_stop"] + 1, "mean", res["avg"], "exact", exact_l2, "std", res["dev"])
    assert err < 0.5 and abs(res["avg"] - exact_l2) < 5 * err
