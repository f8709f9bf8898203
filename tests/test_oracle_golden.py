"""CPU tier: pins the oracle (oracle/refport.py) against the golden vectors produced by running the
UNMODIFIED reference modules (oracle/make_golden.py) and against the one number the reference's
tree holds (gateway.py:100-104)."""
import numpy as np
import pytest

from oracle import refport


def test_probe_stream_pin():
    # element j = 2*(MT19937(123456).next_u32() & 1) - 1
    rs = np.random.RandomState(123456)
    x = refport.rademacher(rs, 16)
    assert list(x.real.astype(int)) == [1, -1, -1, 1, -1, 1, 1, 1, 1, -1, 1, -1, -1, -1, -1, -1]


def test_stencil_rebuild_matches_structure():
    A = refport.load_matrix("schwinger128.mat", -0.1320)
    assert A.shape == (32768, 32768) and A.nnz == 294912
    assert np.all(np.diff(A.tocsr().indptr) == 9)
    A16 = refport.load_matrix("schwinger16.mat", 0.0)
    assert A16.nnz == 4608 and np.allclose(A16.diagonal(), 4.0)


def test_port16_hierarchy_invariants(port16):
    mp, tp = port16
    sizes = [l.A.shape[0] for l in mp.levels]
    assert sizes == [512, 256, 64]
    for l in mp.levels[:-1]:
        P = l.P.tocsr()
        assert np.all(np.diff(P.indptr) == 2)                       # nvec nnz per row
        RP = (l.R * l.P).toarray()
        assert np.abs(RP - np.eye(RP.shape[0])).max() < 1e-13       # orthonormal columns
        assert abs((l.R - l.P.conjugate().transpose())).max() == 0  # R = P^H exactly


def test_port16_probes_match_reference(port16, g16):
    mp, tp = port16
    rs = np.random.RandomState(123456)
    for q in range(8):
        tr = {}
        e, _ = refport.one_defl_hutch_step(mp.levels[0].A, None, mp, tp, "hutchinson", 0, None, None, rs, trace=tr)
        assert abs(e - g16["plain_hutch_e"][q]) <= 1e-9 * abs(g16["plain_hutch_e"][q])
        assert np.linalg.norm(tr["z"] - g16["plain_hutch_z"][q]) <= 1e-8 * np.linalg.norm(g16["plain_hutch_z"][q])
    for lvl in (0, 1):
        for q in range(8):
            e, _ = refport.one_defl_hutch_step(mp.levels[lvl].A, mp.levels[lvl + 1].A, mp, tp, "mlmc", 0, None, None, rs, lvl)
            ref = g16["plain_mlmc_l%d_e" % lvl][q]
            assert abs(e - ref) <= 1e-9 * max(abs(ref), 1.0)


def test_port16_deflated_hutchinson(port16, g16):
    mp, tp = port16
    A = mp.levels[0].A
    tpd = dict(tp); tpd["nr_deflat_vctrs"] = 16
    Ux, tr1 = refport.deflation_pre_computations(A, 16, 1e-9, "hutchinson", tpd, mp, eigpairs=(g16["defl_Sy"], g16["defl_Vx"]))
    assert np.abs(Ux - g16["defl_Ux"]).max() < 1e-14
    assert abs(tr1 - g16["defl_tr1"]) < 1e-9 * abs(g16["defl_tr1"])
    rs = np.random.RandomState(123456)
    for q in range(4):
        e, _ = refport.one_defl_hutch_step(A, None, mp, tpd, "hutchinson", 16, Ux, None, rs)
        assert abs(e - g16["plain_hutch_defl16_e"][q]) <= 1e-9 * max(abs(g16["plain_hutch_defl16_e"][q]), 1.0)


def test_port16_full_mlmc_matches_reference_driver(g16):
    from conftest import params16
    from deflatedmlmc_schwinger_b200 import utils
    p = params16()
    tp = utils.trace_params_from_params(p, "mlmc")
    A = refport.load_matrix(p["matrix"], p["matrix_params"]["mass"])
    res = refport.mlmc(A, tp, test_vectors=[g16["tv0"], g16["tv1"]])
    assert [r["nr_ests"] for r in res["results"]] == list(g16["plain_mlmc_nr_ests"])
    assert abs(res["trace"] - g16["plain_mlmc_trace"]) <= 1e-9 * abs(g16["plain_mlmc_trace"])
    # statistical consistency with the exact trace (dense inverse): |est - exact| within 5 sigma
    sig = sum((r["ests_dev"] ** 2) / (r["nr_ests"] + 1) for r in res["results"][:-1]) ** 0.5
    assert abs(res["trace"] - g16["exact_trace"]) < 5 * sig


def test_telescoping_identity_16(port16, g16):
    # sum over levels of the EXACT level terms equals tr(A^-1): C_{l+1} = R_l C_l P_l with C = I
    mp, tp = port16
    lv = mp.levels
    A0inv = np.linalg.inv(lv[0].A.toarray())
    A1inv = np.linalg.inv(lv[1].A.toarray())
    t0 = np.trace(A0inv - lv[0].P.toarray() @ A1inv @ lv[0].R.toarray())
    t1 = np.trace(A1inv - lv[1].P.toarray() @ mp.coarsest_inv @ lv[1].R.toarray())
    t2 = np.trace(mp.coarsest_inv)
    assert abs(t0 + t1 + t2 - g16["exact_trace"]) < 1e-9 * abs(g16["exact_trace"])


def test_port128_structure_bit_exact(port128, g128):
    mp, tp = port128
    assert [l.A.shape[0] for l in mp.levels] == list(g128["level_sizes"]) == [32768, 8192, 2048, 512]
    assert [l.perm_shift for l in mp.levels] == list(g128["perm_shifts"]) == [512, 128, 32, 8]
    for i in range(3):
        P = mp.levels[i].P.tocsr(); P.sort_indices()
        assert np.array_equal(P.indices.astype(np.int32), g128["P%d_indices" % i])        # index maps bit-exact
        assert np.array_equal(P.data[:4096], g128["P%d_data_head" % i])                   # and values (same vectors)
        assert np.all(np.diff(P.indptr) == 4)
        assert np.all(np.diff(P.tocsc().indptr) == 16)


def test_port128_level2_probes_and_coarsest(port128, g128):
    mp, tp = port128
    rs = np.random.RandomState(123456)
    rs.randint(2, size=32768 * 4)          # the golden stream: 1 hutchinson + 3 level-0 probes come first
    for q in range(6):
        e, _ = refport.one_defl_hutch_step(mp.levels[2].A, mp.levels[3].A, mp, tp, "mlmc", 0, None, None, rs, 2)
        assert abs(e - g128["mlmc_l2_e"][q]) <= 1e-9 * max(abs(g128["mlmc_l2_e"][q]), 1.0)
    lv = mp.levels
    crst = lv[3].Pperm.transpose().conjugate() * (mp.coarsest_inv * lv[3].Bblock_perm)
    assert abs(np.trace(crst) - g128["coarsest_term"]) <= 1e-8 * abs(g128["coarsest_term"])


def test_exact_level_split_sums_to_gateway_golden(port128, g128):
    """gateway.py:100-104: exact displaced trace.  L2-difference and L3 terms by dense algebra plus the
    level-0 difference obtained from the identity  total = L0 + L2 + L3  must be self-consistent with
    C_{l+1} = R_l C_l P_l; here: check the telescoping residual of the permutation operators."""
    mp, tp = port128
    lv = mp.levels
    for l in range(3):
        Cl = lv[l].Bblock_perm @ lv[l].Pperm.transpose()
        Cn = lv[l + 1].Bblock_perm @ lv[l + 1].Pperm.transpose()
        D = (lv[l].R @ Cl @ lv[l].P - Cn)
        assert abs(D).max() < 1e-12
    A2inv = np.linalg.inv(lv[2].A.toarray())
    C2 = (lv[2].Bblock_perm @ lv[2].Pperm.transpose()).toarray()
    C3 = (lv[3].Bblock_perm @ lv[3].Pperm.transpose()).toarray()
    t3 = np.trace(mp.coarsest_inv @ C3)
    assert abs(t3 - g128["coarsest_term"]) < 1e-8 * abs(t3)
    t2 = np.trace(A2inv @ C2) - t3
    # the level-2 probes of the golden stream scatter around t2 with std ~19.5
    assert abs(np.mean(g128["mlmc_l2_e"]) - t2) < 5 * 19.5 / np.sqrt(16) + 1.0


def test_port16_deflated_mlmc_probes_match_reference(port16, g16defl):
    """deflated MLMC level samples (utils.py:252-357 with nr_deflat_vctrs = 16) of the port against the unmodified
    reference, deflation vectors from the reference's own deflation_pre_computations on diff_op_Q"""
    mp, tp = port16
    mp.skip_level = False
    for ix in range(2):
        rs = np.random.RandomState(123456 + ix)
        ref = g16defl["l%d_e" % ix]
        for q in range(8):
            e, _ = refport.one_defl_hutch_step(mp.levels[ix].A, mp.levels[ix + 1].A, mp, tp, "mlmc", 16,
                                               g16defl["l%d_Vx" % ix], g16defl["l%d_Ux" % ix], rs, ix)
            assert abs(e - ref[q]) < 1e-8 * max(abs(ref[q]), 1.0), (ix, q)


# ---- round 2 fixtures ------------------------------------------------------------------------------------------------------
def test_port128_deflated_variant_level2_matches_reference(g128):
    """schwinger128_ext.npz, deflated variant (not permuted, mlmc_deflat_vctrs = [16, 0, 16]): tr1 and the 8 deflated level-2
    samples the UNMODIFIED reference produced from the stored (complex64) eigsh vectors of diff_op_Q -- reproduced by the
    port from the same stored vectors (utils.py:145-176, 252-357)"""
    import os
    from conftest import GOLDEN, params128
    from deflatedmlmc_schwinger_b200 import utils
    ge = np.load(os.path.join(GOLDEN, "schwinger128_ext.npz"))
    p = params128()
    p["use_permuted"] = False
    p["mlmc_deflat_vctrs"] = [16, 0, 16]
    tp = utils.trace_params_from_params(p, "mlmc")
    mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
    mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp, test_vectors=[g128["tv0"], g128["tv1"], g128["tv2"]])
    mp.skip_level = True
    V = ge["defl_l2_eigvecs_c64"].astype(np.complex128)
    Vx, Ux, tr1 = refport.deflation_pre_computations(mp.levels[0].A, 16, 1e-1, "mlmc", tp, mp, None, level_nr=2, eigpairs=(ge["defl_l2_Sy"], V))
    assert abs(tr1 - ge["defl_l2_tr1"]) < 1e-10 * abs(ge["defl_l2_tr1"])
    rs = np.random.RandomState(123456 + 2)
    for q in range(8):
        e, _ = refport.one_defl_hutch_step(mp.levels[2].A, mp.levels[3].A, mp, tp, "mlmc", 16, Vx, Ux, rs, 2)
        assert abs(e - ge["defl_l2_e"][q]) < 1e-8 * abs(ge["defl_l2_e"][q]), q


def test_synthetic_fixture_header_and_generators():
    """synthetic256.npz: the oracle's and the product's generators of the random-U(1) lattice agree bit for bit, both decode the
    BF16-stored test vectors identically, and the stored vectors are what the fixture says (4 per level, level sizes)"""
    import os
    from conftest import GOLDEN
    from oracle import make_golden
    from deflatedmlmc_schwinger_b200 import lattice
    g = np.load(os.path.join(GOLDEN, "synthetic256.npz"))
    L = int(g["L"])
    assert np.array_equal(refport.synthetic_links(L, int(g["seed"]), float(g["sigma"])), lattice.random_u1_links(L, int(g["seed"]), float(g["sigma"])))
    for i, n in enumerate(g["level_sizes"][:3]):
        a = make_golden.bf16_unpack(g["tv%d_bf16" % i])
        assert a.shape == (n, 4) and np.array_equal(a, lattice.unpack_bf16_vectors(g["tv%d_bf16" % i]))
        assert np.array_equal(make_golden.bf16_pack(a), g["tv%d_bf16" % i])           # decode/encode round trip
        assert np.all(np.abs(np.linalg.norm(a, axis=0) - 1.0) < 1e-2)                  # eigs returns unit vectors
    assert list(g["level_sizes"]) == [2 * L * L // 4 ** i for i in range(4)]


def test_packaged_links_equal_the_reference_mat_files():
    """deflatedmlmc_schwinger_b200/data/*_links.npy rebuild both reference matrices bit for bit (runs where /root/reference is
    present, i.e. in the authoring container; the GPU box has no reference tree)"""
    import os
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    import scipy.io as sio
    from deflatedmlmc_schwinger_b200 import lattice
    data = os.path.join(os.path.dirname(os.path.abspath(lattice.__file__)), "data")
    for name in ("schwinger16", "schwinger128"):
        S = sio.loadmat(os.path.join(ref_shim.REF_DIR, name + ".mat"))["S"].tocsc()
        if name == "schwinger16":                        # matrix.py:25-27: the 16^2 file stores gamma3 S
            h = S.shape[0] // 2
            S = S.tolil(); S[h:, :] = -S[h:, :]; S = S.tocsc()
        D = (lattice.wilson_matrix(np.load(os.path.join(data, name + "_links.npy")), 0.0) - S).tocsc()
        D.eliminate_zeros()
        assert D.nnz == 0, name
