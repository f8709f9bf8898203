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


@pytest.mark.parametrize("consumed,skip,count,after", [(0, 0, 5000, 0), (7, 0, 1000, 0), (623, 5, 700, 11),
                                                      (624, 1300, 4096, 977), (100, 0, 0, 2000)])
def test_device_probe_stream_is_numpy_mt19937(mg16, consumed, skip, count, after):
    """dmlmc_mt19937_bits: element j of the device stream == np.random.randint(2) element j, and the state
    written back equals the host generator's state after the same number of draws."""
    import torch
    mg, tp, A = mg16
    dev = mg.dev
    np.random.seed(123456)
    if consumed:
        np.random.randint(2, size=consumed)
    st = np.random.get_state()
    words = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
    state = torch.from_numpy(words.view(np.int32).copy()).cuda()
    backup = torch.zeros(625, dtype=torch.int32, device="cuda")
    out = dev.mt19937_bits(state, skip, count, after, backup=backup)
    dev.rng_sync()
    ref_all = np.random.randint(2, size=skip + count + after)
    if count:
        assert np.array_equal(out.cpu().numpy(), ref_all[skip:skip + count].astype(np.uint8))
    st2 = np.random.get_state()
    got = state.cpu().numpy().view(np.uint32)
    # same stream position: compare what both generate next (the key arrays can differ by an un-applied twist)
    state2 = torch.from_numpy(got.view(np.int32).copy()).cuda()
    nxt = dev.mt19937_bits(state2, 0, 1500, 0)
    dev.rng_sync()
    assert np.array_equal(nxt.cpu().numpy(), np.random.randint(2, size=1500).astype(np.uint8))
    assert np.array_equal(backup.cpu().numpy().view(np.uint32), words)
    assert st2[0] == 'MT19937'


@pytest.mark.parametrize("consumed,skip,count,after", [(0, 0, 70000, 0), (7, 3 * 32768, 2 * 32768 + 5, 4 * 32768 - 5),
                                                      (624, (1 << 24) + 12345, 40000, (1 << 20) + 3),
                                                      (100, 5 * (1 << 21), 0, 0), (333, 0, 1 << 21, 7 * (1 << 21))])
def test_probe_stream_jump_ahead_equals_the_sequential_generator(mg16, consumed, skip, count, after):
    """dmlmc_mt19937_bits with the jump-ahead table (many CTAs, nobody generates the skipped words) == the sequential
    one-CTA kernel (option mt_jump = 0): same output bits, and EXACTLY the same state words (numpy's alignment of the
    key array and position), so the generator handed back to np.random is the one the reference would leave."""
    import torch
    mg, tp, A = mg16
    dev = mg.dev
    np.random.seed(123456)
    if consumed:
        np.random.randint(2, size=consumed)
    st = np.random.get_state()
    words = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
    res = []
    for jump in (1, 0):
        dev.set_option("mt_jump", jump)
        try:
            state = torch.from_numpy(words.view(np.int32).copy()).cuda()
            backup = torch.zeros(625, dtype=torch.int32, device="cuda")
            out = dev.mt19937_bits(state, skip, count, after, backup=backup)
            dev.rng_sync()
            res.append((None if out is None else out.cpu().numpy(), state.cpu().numpy().view(np.uint32).copy(),
                        backup.cpu().numpy().view(np.uint32).copy()))
        finally:
            dev.set_option("mt_jump", 1)
    (oj, sj, bj), (os_, ss, bs) = res
    if count:
        assert np.array_equal(oj, os_)
    assert np.array_equal(bj, words) and np.array_equal(bs, words)
    total = skip + count + after
    if total > 0 and int(words[624]) + total > 624:
        assert np.array_equal(sj, ss), "state after the jump differs from the sequential generator's"
    else:
        assert sj[624] == ss[624] and np.array_equal(sj[1:624], ss[1:624])
    # and against numpy itself on a sample of the range (the host walks the skipped words)
    if count and skip + count <= (1 << 25):
        sampling_skip = skip
        while sampling_skip > 0:
            c = min(sampling_skip, 1 << 22); np.random.bytes(4 * c); sampling_skip -= c
        assert np.array_equal(oj, np.random.randint(2, size=count).astype(np.uint8))


def test_device_and_host_probe_streams_give_the_same_run(g16):
    """the whole MLMC driver with the device-generated stream == with the host-generated stream, including the
    final state of the global numpy generator (the rewind of the sequential stopping rule)"""
    from conftest import params16
    from deflatedmlmc_schwinger_b200 import matrix, stoch_trace, utils
    out = []
    for dev_stream in (True, False):
        p = params16()
        p["test_vectors"] = [g16["tv0"], g16["tv1"]]
        p["probe_batch"] = 16
        p["smoother_degree"] = 8
        p["device_probe_stream"] = dev_stream
        tp = utils.trace_params_from_params(p, "mlmc")
        A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
        res = stoch_trace.mlmc(A, tp)
        out.append((res, np.random.randint(1 << 30, size=8)))
    (ra, na), (rb, nb) = out
    assert [r["nr_ests"] for r in ra["results"]] == [r["nr_ests"] for r in rb["results"]]
    assert abs(ra["trace"] - rb["trace"]) < 1e-9 * abs(rb["trace"])
    assert np.array_equal(na, nb)


def test_exact_trace_of_the_shipped_128_set(mg128, g128):
    """Unit vectors through the fused device sample = the exact value of every MLMC level; their telescoping
    sum must be the exact displaced trace the reference quotes (gateway.py:100-104).  32768 + 2048 solves."""
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg128
    assert mg.skip_level
    l0 = utils.exact_level_trace(mg, tp, "mlmc", 0, k=256)
    l2 = utils.exact_level_trace(mg, tp, "mlmc", 2, k=256)
    lv = mg.ml.levels
    crst = np.trace(lv[3].Pperm.transpose().conjugate() * (mg.coarsest_inv * lv[3].Bblock_perm))
    total = l0 + l2 + crst
    exact = -8.748242701374695 + 50.215154098005584j
    print("exact level traces:", l0, l2, crst, "sum", total)
    assert abs(total - exact) < 1e-8 * abs(exact)
    # level 2 against dense algebra
    A2inv = np.linalg.inv(lv[2].A.toarray())
    C2 = (lv[2].Bblock_perm @ lv[2].Pperm.transpose()).toarray()
    assert abs(l2 - (np.trace(A2inv @ C2) - crst)) < 1e-8 * abs(l2)
