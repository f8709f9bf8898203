"""CPU tier: host-side logic of the product (no GPU): probe stream handling, sampling loops and the
stop rule, aggregation maps, prolongator values, operator re-layout, parameter plumbing."""
import numpy as np
import pytest
import scipy.sparse as sp

from deflatedmlmc_schwinger_b200 import gateway, lattice, matrix, sampling, utils
from deflatedmlmc_schwinger_b200 import multigrid as mgm
from oracle import refport


def test_draw_probe_bits_consumes_the_reference_stream():
    np.random.seed(123456)
    a = np.random.randint(2, size=5000)
    b = np.random.randint(2, size=777)
    st = np.random.get_state()
    np.random.seed(123456)
    a2 = sampling.draw_probe_bits(5000)
    b2 = sampling.draw_probe_bits(777)
    assert np.array_equal(a, a2) and np.array_equal(b, b2)
    assert np.array_equal(np.random.get_state()[1], st[1]) and np.random.get_state()[2] == st[2]
    np.random.seed(1)
    sampling.skip_probe_words(1000)
    x = sampling.draw_probe_bits(10)
    np.random.seed(1)
    y = np.random.randint(2, size=1010)[1000:]
    assert np.array_equal(x, y)


def test_pack_bits_layout():
    bits = np.array([1, 0, 0, 1, 1, 1, 0, 0, 1, 0], dtype=np.uint8)
    p = utils.pack_bits(bits)
    for j in range(10):
        assert ((p[j >> 3] >> (j & 7)) & 1) == bits[j]


def _sequential_reference_loop(n, tol, max_nr, value_fn):
    """the reference's loop shape (stoch_trace.py:386-406) on a deterministic per-probe function"""
    ests = np.zeros(max_nr, dtype=complex)
    for j in range(max_nr):
        x = np.random.randint(2, size=n) * 2 - 1
        ests[j] = value_fn(x)
        avg, dev, err = sampling.reference_stats(ests, j)
        if j >= 5 and err < tol:
            break
    return j, avg, dev


def _value(x):
    n = x.shape[0]
    w = np.cos(np.arange(n)) + 1j * np.sin(0.3 * np.arange(n))
    return np.dot(w, x) * (1 + 0.01 * x[0]) + 3.0


def _sample_fn(n, k):
    def fn(bits01):
        xs = bits01.reshape(k, n).astype(np.int64) * 2 - 1
        return np.array([_value(x) for x in xs]), np.ones(k, dtype=np.int64)
    return fn


@pytest.mark.parametrize("k", [1, 4, 16, 64])
def test_run_sampling_equals_sequential_loop(k):
    n, tol = 64, 0.9
    np.random.seed(42)
    j_ref, avg_ref, dev_ref = _sequential_reference_loop(n, tol, 5000, _value)
    after_ref = np.random.randint(2, size=8)
    np.random.seed(42)
    res = sampling.run_sampling(_sample_fn(n, k), n, k, tol, 5000)
    after = np.random.randint(2, size=8)
    assert res["j_stop"] == j_ref
    assert abs(res["avg"] - avg_ref) < 1e-12 * abs(avg_ref) and abs(res["dev"] - dev_ref) < 1e-12 * dev_ref
    assert np.array_equal(after, after_ref)           # stream rewound to exactly after the last used probe
    assert res["iters_sum"] == j_ref + 1


def test_run_sampling_fixed_count_and_max():
    n, k = 32, 8
    np.random.seed(7)
    res = sampling.run_sampling(_sample_fn(n, k), n, k, 0.0, 5, fixed_count=5)
    assert res["j_stop"] == 4 and res["ests"].shape[0] == 5
    np.random.seed(7)
    ref = [_value(np.random.randint(2, size=n) * 2 - 1) for _ in range(5)]
    assert np.allclose(res["ests"], ref)
    np.random.seed(7)
    res = sampling.run_sampling(_sample_fn(n, k), n, k, 1e-30, 20)      # never converges -> max_nr_ests
    assert res["j_stop"] == 19


def test_reduce_level_sums_single_process():
    e = np.array([1 + 2j, 3 - 1j, -2 + 0.5j])
    m, s, N = sampling.reduce_level_sums(e)
    assert N == 3 and abs(m - e.mean()) < 1e-15 and abs(s - np.sqrt(np.mean(np.abs(e - e.mean()) ** 2))) < 1e-14


def test_aggregation_maps_closed_form_vs_oracle(port128):
    mp, tp = port128
    meta = [(32, 2, 4), (32, 4, 4), (32, 4, 4)]
    for i, (a, dofi, c) in enumerate(meta):
        P = mp.levels[i].P.tocsr(); P.sort_indices()
        n = P.shape[0]
        half, first = mgm.aggregation_maps(n, a, dofi, c)
        idx = (first[:, None] + np.arange(c)[None, :]).ravel()
        assert np.array_equal(idx, P.indices)                         # bit-exact index maps


def test_prolongator_values_vs_oracle(port128, g128):
    mp, tp = port128
    for i, (a, dofi, c, tv) in enumerate([(32, 2, 4, g128["tv0"]), (32, 4, 4, g128["tv1"]), (32, 4, 4, g128["tv2"])]):
        pv = mgm.build_prolongator_values(tv, a, dofi, c)
        P = mgm.prolongator_csr(pv, a, dofi, c)
        D = P - mp.levels[i].P
        assert (abs(D).max() if D.nnz else 0.0) == 0.0            # bit-identical to the oracle (and the reference)


def test_bsr_and_ell_padding_roundtrip(port128):
    mp, tp = port128
    for lvl, bs in ((1, 4), (2, 4)):
        A = mp.levels[lvl].A
        col, vals = mgm.bsr_padded(A, bs)
        nb = A.shape[0] // bs
        assert col.shape[0] == nb
        x = np.random.RandomState(0).standard_normal(A.shape[0]) + 0j
        y = np.zeros(A.shape[0], dtype=complex)
        for I in range(nb):
            for b in range(col.shape[1]):
                J = col[I, b]
                if J >= 0:
                    y[I * bs:(I + 1) * bs] += vals[I, b] @ x[J * bs:(J + 1) * bs]
        assert np.abs(y - A @ x).max() < 1e-12
    # structural claim of SURVEY.md 2.1: A1 has 9 4x4 blocks per block row
    col, _ = mgm.bsr_padded(mp.levels[1].A, 4)
    assert col.shape[1] == 9
    cols, vals = mgm.ell_padded(mp.levels[2].Bblock_perm)
    x = np.arange(2048) + 1j
    y = np.array([sum(vals[i, j] * x[cols[i, j]] for j in range(cols.shape[1]) if cols[i, j] >= 0) for i in range(2048)])
    assert np.abs(y - mp.levels[2].Bblock_perm @ x).max() < 1e-12


def test_perm_is_roll(port128):
    mp, tp = port128
    for l in range(4):
        n = mp.levels[l].A.shape[0]
        x = np.arange(n) + 0j
        assert np.array_equal(mp.levels[l].Pperm.transpose() @ x, np.roll(x, mp.levels[l].perm_shift))


def test_harmonic_ritz_polynomial_reduces_residual(port128):
    mp, tp = port128
    A = sp.csr_matrix(mp.levels[2].A)
    w = mgm.harmonic_ritz_inv_roots(A, 16)
    assert w.shape == (16,)
    r = np.random.RandomState(3).standard_normal(A.shape[0]) + 0j
    r0 = np.linalg.norm(r)
    for wi in w:
        r = r - wi * (A @ r)
    assert np.linalg.norm(r) < 0.7 * r0


def test_lattice_roundtrip_and_synthetic():
    links = lattice.random_u1_links(8, seed=5)
    A = lattice.wilson_matrix(links, -0.1)
    l2, d = lattice.links_from_matrix(A, 8, 8)
    assert np.array_equal(l2, links) and d == 4.0 - 0.1
    B = refport.wilson_from_links(links)
    D = A - (B + (-0.1) * sp.identity(128, format="csc"))
    assert (abs(D).max() if D.nnz else 0.0) < 1e-15
    with pytest.raises(Exception):
        lattice.links_from_matrix(sp.random(128, 128, 0.1, format="csr") + sp.identity(128), 8, 8)
    A2 = matrix.loadMatrix("synthetic:8:5", {"mass": -0.1})
    assert abs(A2 - A).max() == 0


def test_gateway_and_param_plumbing():
    p = gateway.set_params("schwinger128")
    assert p["aggrs"] == [16, 4, 4] and p["dof"] == [2, 8, 8, 8] and p["matrix_params"]["mass"] == -0.1320
    assert p["mlmc_levels_to_skip"] == [1] and p["nr_deflat_vctrs"] == 8 and p["use_permuted"] is True
    p["function_tol"] = 1e-12
    tp = utils.trace_params_from_params(p, "mlmc")
    assert tp["max_nr_ests"] == 100000 and tp["function_params"]["tol"] == 1e-12 and tp["tol"] == 1e-2
    th = utils.trace_params_from_params(p, "hutchinson")
    assert "mlmc_deflat_vctrs" not in th and th["defl-type"] == "exact"
    with pytest.raises(Exception):
        utils.trace_params_from_params(p, "nope")
    with pytest.raises(Exception):
        gateway.set_params("nope")
    t = utils.CustomTimer()
    t.start("mvm"); t.end("mvm")
    with pytest.raises(Exception):
        t.end("mvm")


def test_smoother_product_form_matches_richardson(port128):
    """multigrid.smoother_product_form: p0 prod_i (1 - nu_i A) b == sum_i omega_i prod_{j<i} (1 - omega_j A) b"""
    mp, tp = port128
    A = sp.csr_matrix(mp.levels[0].A)
    for d in (1, 2, 9, 32):
        w = mgm.harmonic_ritz_inv_roots(A, d)
        nu, p0 = mgm.smoother_product_form(w)
        assert nu.shape == (d - 1,)
        rs = np.random.RandomState(d)
        b = rs.standard_normal(A.shape[0]) + 1j * rs.standard_normal(A.shape[0])
        r = b.copy(); e = np.zeros_like(b)
        for i, wi in enumerate(w):
            e = e + wi * r
            if i < d - 1:
                r = r - wi * (A @ r)
        y = b.copy()
        for v in nu:
            y = y - v * (A @ y)
        assert np.linalg.norm(p0 * y - e) / np.linalg.norm(e) < 1e-12


def test_geometric_aggregates_of_the_preconditioner_hierarchy():
    """blocks of lattice sites split by spin: equal-sized aggregates, P^H P = I, chirality preserved, and the
    restricted test vectors are eigenvectors of the Galerkin operator (so the coarse levels need no eigensolve)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "schwinger128.npz"))
    L = 128
    n = 2 * L * L
    cb = mgm.geometric_blocks_level0(L, L, 4, 4)
    assert cb.dtype == np.int32 and cb.shape == (n,)
    cnt = np.bincount(cb)
    assert cnt.shape[0] == 2 * (L // 4) ** 2 and np.all(cnt == 16)
    # row i = s*V + x*L + t
    i = 1 * L * L + 37 * L + 102
    assert cb[i] == ((37 // 4) * (L // 4) + 102 // 4) * 2 + 1
    tv = g["tv0"]
    pv = mgm.block_orthonormal_values(tv, cb, 4)
    P = mgm.prolongator_csr_indexed(pv, cb)
    assert P.shape == (n, 8192)
    PhP = (P.conj().T @ P).toarray()
    assert np.abs(PhP - np.eye(8192)).max() < 1e-13
    # the test vectors lie in range(P)
    assert np.linalg.norm(P @ (P.conj().T @ tv) - tv) < 1e-12 * np.linalg.norm(tv)
    A = refport.load_matrix("schwinger128", -0.1320).tocsr()
    Ac = (P.conj().T @ A @ P).tocsr()
    tc = P.conj().T @ tv
    lam = np.einsum("ij,ij->j", tv.conj(), A @ tv) / np.einsum("ij,ij->j", tv.conj(), tv)
    assert np.linalg.norm(Ac @ tc - tc * lam[None, :]) < 1e-7 * np.linalg.norm(tc)
    # coarse level: rows ((X*LT + T)*2 + half)*nv + v, 2 x 2 blocks
    cb1 = mgm.geometric_blocks_coarse(32, 32, 4, 2, 2)
    assert cb1.shape == (8192,) and np.all(np.bincount(cb1) == 16)
    r = ((5 * 32 + 9) * 2 + 1) * 4 + 3
    assert cb1[r] == ((5 // 2) * 16 + 9 // 2) * 2 + 1
    # chirality: spin-0 rows only feed even coarse blocks
    assert np.all(cb[: L * L] % 2 == 0) and np.all(cb[L * L:] % 2 == 1)


def test_even_odd_schur_complement_reproduces_the_inverse():
    """A^{-1} r through the even-odd factorisation the device smoother uses: r^_e = r_e - H_eo r_o / c,
    x_e = S^{-1} r^_e, x_o = (r_o - H_oe x_e) / c  with  S = c - H_eo H_oe / c  (16^2, dense algebra)"""
    import scipy.sparse.linalg as spla
    L = 16
    A = sp.csr_matrix(refport.load_matrix("schwinger16", -1.00690114 * 0.99))
    S, c = mgm.even_odd_schur(A, L, L)
    s, x, t = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing="ij")
    par = ((x + t) & 1).ravel()
    ie, io = np.where(par == 0)[0], np.where(par == 1)[0]
    assert S.shape == (L * L, L * L)
    rs = np.random.RandomState(3)
    r = rs.standard_normal(2 * L * L) + 1j * rs.standard_normal(2 * L * L)
    Heo, Hoe = A[ie][:, io], A[io][:, ie]
    re_hat = r[ie] - Heo @ r[io] / c
    xe = spla.spsolve(S.tocsc(), re_hat)
    xo = (r[io] - Hoe @ xe) / c
    xfull = np.zeros_like(r); xfull[ie] = xe; xfull[io] = xo
    ref = spla.spsolve(A.tocsc(), r)
    assert np.linalg.norm(xfull - ref) < 1e-10 * np.linalg.norm(ref)


def test_geometric_blocks_of_the_estimators_level1():
    """the block map used to precondition the estimator's level-1 solves: every block must be the set of level-1 rows
    whose level-0 support (through the reference's own P_0) is one spin component of 2 neighbouring lattice columns x
    and one run of 32 sites in t -- checked against the oracle's P_0 on 128^2"""
    import os
    from deflatedmlmc_schwinger_b200 import gateway
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "schwinger128.npz"))
    L, nv = 128, 4
    cb, (gx, gt) = mgm.geometric_blocks_level1(L, L, 32, nv, 2)
    n1 = 8192
    assert cb.shape == (n1,) and (gx, gt) == (64, 4) and np.all(np.bincount(cb) == 16) and cb.max() + 1 == 2 * gx * gt
    # support of every level-1 row from the closed-form structure of P_0 (bit-exact with the reference, tested above)
    half, first = mgm.aggregation_maps(2 * L * L, 32, 2, nv)
    i = np.arange(2 * L * L)
    s, x, t = i // (L * L), (i % (L * L)) // L, i % L
    for v in range(nv):
        col = first + v                                   # level-1 row fed by fine row i (vector v)
        blk = cb[col]
        assert np.array_equal(blk, ((x // 2) * gt + t // 32) * 2 + s)
    # with the golden level-1 test vectors: orthonormal prolongator, test vectors in its range
    tv1 = g["tv1"]
    P = mgm.prolongator_csr_indexed(mgm.block_orthonormal_values(tv1, cb, nv), cb)
    assert abs((P.conj().T @ P) - sp.identity(P.shape[1])).max() < 1e-13
    assert np.linalg.norm(P @ (P.conj().T @ tv1) - tv1) < 1e-12 * np.linalg.norm(tv1)


def test_mt19937_jump_table_is_pinned_against_numpy():
    """mtjump: the characteristic polynomial recovered by Berlekamp-Massey has degree 19937 and 135 terms; the shipped table
    t^(2^b) mod phi equals a fresh computation; jumping with it lands on the words np.random itself produces."""
    from deflatedmlmc_schwinger_b200 import mtjump as mj
    phi = mj.char_poly()
    assert phi.bit_length() - 1 == mj.DEG and bin(phi).count("1") == 135
    tab = mj.table()
    p = 2
    for b in range(14):
        assert np.array_equal(tab[b], mj.poly_to_words(p))
        p = mj._reduce(mj._square(p), phi)
    rs = np.random.RandomState(123456)
    rs.bytes(4 * 77)
    st = rs.get_state()
    for skip in (0, 623, 1024, 3 * 32768 + 17, (1 << 22) + 999):
        arr, q = mj.jump_host(st[1], st[2], skip)
        r2 = np.random.RandomState()
        r2.set_state(st)
        if skip:
            r2.bytes(4 * skip)
        assert np.array_equal(np.frombuffer(r2.bytes(4 * 600), dtype=np.uint32), mj.temper(arr[q:q + 600]))


def test_geometric_blocks_of_the_coarse_levels_are_local(port128):
    """geometric_blocks_level1 with a_sites = the t-extent merged up to level l (32 at level 1, 128 at level 2 of the 128^2
    hierarchy) gives, for every row of the ESTIMATOR's A_l, the block (x pair, t range, spin) its lattice support lies in: the
    support of a row is read off the reference-aggregation prolongators themselves (|P_0|, |P_0| |P_1|)."""
    mp, tp = port128
    LX = LT = 128
    V = LX * LT
    P0 = abs(mp.levels[0].P).tocsc()
    P01 = (abs(mp.levels[0].P) @ abs(mp.levels[1].P)).tocsc()
    for lvl, Pc, a_sites in ((1, P0, 32), (2, P01, 128)):
        n_l = Pc.shape[1]
        cblk, (gx, gt) = mgm.geometric_blocks_level1(LX, LT, a_sites, 4, 2)
        assert cblk.shape[0] == n_l and (gx, gt) == (LX // 2, LT // a_sites)
        for r in list(range(0, n_l, 97)) + [n_l - 1]:
            rows = Pc.indices[Pc.indptr[r]:Pc.indptr[r + 1]]           # level-0 rows i = s V + x LT + t in the support
            s, x, t = rows // V, (rows % V) // LT, rows % LT
            b = int(cblk[r])
            bs, bq, bx = b % 2, (b // 2) % gt, (b // 2) // gt
            assert np.all(s == bs) and np.all(x // 2 == bx) and np.all(t // a_sites == bq), (lvl, r)
        assert np.bincount(cblk).min() == np.bincount(cblk).max() == 16      # 2 groups x 2 halves x 4 vectors per block


def test_mt19937_jump_polynomial_for_any_distance():
    """mtjump.poly_for_distance / chunk_polys: t^D mod phi by FFT products of the table entries -- one application of that
    polynomial lands on the words np.random produces D draws later (the multi-jump kernel applies one table entry per set bit
    of D instead; both are the same linear map)."""
    from deflatedmlmc_schwinger_b200 import mtjump as mj
    rs = np.random.RandomState(123456)
    rs.bytes(4 * 77)
    st = rs.get_state()
    key, pos = np.asarray(st[1], dtype=np.uint32), st[2]
    P = mj.chunk_polys(3 * (1 << 16) + 5, 1 << 12, 3, 12345)
    arr = mj.raw_words(key, mj.N + 1)[1:]                      # the window one word further on (see jump_host)
    for c, D in enumerate([3 * (1 << 16) + 5, 3 * (1 << 16) + 5 + (1 << 12), 3 * (1 << 16) + 5 + 2 * (1 << 12), 12345]):
        X = mj.raw_words(mj.apply_poly(arr, P[c]), (pos - 1) + 600)
        r2 = np.random.RandomState()
        r2.set_state(st)
        r2.bytes(4 * D)
        assert np.array_equal(np.frombuffer(r2.bytes(4 * 600 - 4 * (pos - 1) + 4 * (pos - 1)), dtype=np.uint32)[:600 - (pos - 1)],
                              mj.temper(X[pos - 1:600]))
    assert np.array_equal(mj.poly_for_distance(0)[:2], np.array([1, 0], dtype=np.uint32))


def test_padded_block_rows_round_trip():
    """bsr_padded (scipy -> the device layout) and csr_from_padded_bsr (the device layout, as dmlmc_galerkin writes it, ->
    scipy) are inverse to each other, rows with fewer blocks than the widest one included"""
    import scipy.sparse as sp
    from deflatedmlmc_schwinger_b200 import multigrid as mgm
    rs = np.random.RandomState(0)
    nb, bs = 12, 4
    mask = rs.rand(nb, nb) < 0.3
    mask[np.arange(nb), np.arange(nb)] = True
    mask[3, :] = False
    mask[3, 5] = True                           # one row with a single block
    D = np.zeros((nb * bs, nb * bs), dtype=np.complex128)
    for i, j in zip(*np.nonzero(mask)):
        D[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = rs.standard_normal((bs, bs)) + 1j * rs.standard_normal((bs, bs))
    A = sp.csr_matrix(D)
    col, vals = mgm.bsr_padded(A, bs)
    assert col.shape == (nb, mask.sum(axis=1).max()) and (col[3] >= 0).sum() == 1
    assert all(np.all(np.diff(r[r >= 0]) > 0) for r in col)
    B = mgm.csr_from_padded_bsr(col, vals)
    assert abs(B - A).max() == 0.0
    col2, vals2 = mgm.bsr_padded(B, bs)
    assert np.array_equal(col, col2) and np.array_equal(vals, vals2)


def test_strip_sites_of_the_coarse_levels():
    """a_l of MG._build_level1_preconditioner (t-sites per strip group of the rows of level l): a_1 = the level-0 aggregate,
    a_{l+1} = a_l times the strips a level-l aggregate merges; other aggregate structures give None"""
    import scipy.sparse as sp
    from deflatedmlmc_schwinger_b200 import multigrid as mgm
    mg = mgm.MG(sp.identity(8, format="csr", dtype=np.complex128))
    mg._transfer_meta = [(32, 2, 4, None), (32, 8, 4, None), (32, 8, 4, None)]       # (aggr_size, dofi, nvec, values)
    assert mg._strip_sites(1) == 32
    assert mg._strip_sites(2) == 32 * (32 // 8)
    assert mg._strip_sites(3) == 32 * (32 // 8) * (32 // 8)
    mg._transfer_meta = [(32, 4, 4, None), (32, 8, 4, None)]                         # dofi != 2 on level 0
    assert mg._strip_sites(1) is None
    mg._transfer_meta = [("indexed", None, 4, None)]
    assert mg._strip_sites(1) is None
    mg._transfer_meta = [(32, 2, 4, None), (36, 8, 4, None)]                         # aggregate not a whole number of strips
    assert mg._strip_sites(1) == 32 and mg._strip_sites(2) is None
