"""GPU tier: parity of the batched solver and of the fused per-batch sample against the oracle and
the golden vectors generated from the unmodified reference.  Tolerance (north_star): per-probe
solves to a relative residual of 1e-8 in complex128; estimates within that tolerance propagated."""
import numpy as np
import pytest
import torch

from oracle import refport

pytestmark = pytest.mark.gpu

# every column leaves the solve on its TRUE complex128 residual: <= tol on the Arnoldi estimate in the first cycle, and
# <= 1.25 tol + 1e-14 when a later cycle re-checks it against b - A x (gmres_init_kernel) -- so with tol = 1e-12:
RES_MAX = 1.3e-12


def host(t):
    return t.cpu().numpy().astype(np.complex128)


def probes(n, k, seed=123456):
    rs = np.random.RandomState(seed)
    return (rs.randint(2, size=(k, n)) * 2 - 1).T.astype(np.complex128)


@pytest.mark.parametrize("prec", ["c64", "c128"])
def test_fgmres_16_solutions_match_reference(mg16, g16, prec):
    from deflatedmlmc_schwinger_b200 import _lib
    mg, tp, A = mg16
    mg.dev.set_inner_precision(_lib.C64 if prec == "c64" else _lib.C128)
    B = probes(512, 8)
    X, iters, relres = mg.dev.fgmres(0, torch.from_numpy(np.ascontiguousarray(B)).cuda(), 1e-12, restart=40, maxiter=512)
    x = host(X)
    assert np.all(relres < 1e-12) and np.all(iters > 0)
    res = np.linalg.norm(B - A @ x, axis=0) / np.linalg.norm(B, axis=0)
    assert res.max() <= RES_MAX
    for q in range(8):                       # the reference's own solutions for the same probes
        z = g16["plain_hutch_z"][q]
        assert np.linalg.norm(x[:, q] - z) / np.linalg.norm(z) < 1e-8
    mg.dev.set_inner_precision(_lib.C64)


def test_fgmres_restart_and_maxiter(mg16):
    mg, tp, A = mg16
    B = probes(512, 4)
    Bd = torch.from_numpy(np.ascontiguousarray(B)).cuda()
    X, iters, relres = mg.dev.fgmres(0, Bd, 1e-12, restart=3, maxiter=512)     # forces restarts
    res = np.linalg.norm(B - A @ host(X), axis=0) / np.linalg.norm(B, axis=0)
    assert res.max() <= RES_MAX
    X, iters, relres = mg.dev.fgmres(0, Bd, 1e-12, restart=40, maxiter=2)      # hits maxiter
    assert np.all(iters == 2) and np.all(relres > 1e-12)
    Z = torch.zeros_like(Bd)                                                    # zero rhs -> zero solution
    X, iters, relres = mg.dev.fgmres(0, Z, 1e-12, restart=10, maxiter=20)
    assert np.all(iters == 0) and float(X.abs().max()) == 0.0


def test_fgmres_batch_independence(mg128):
    """a probe's solution must not depend on which batch it is solved in (deterministic reductions,
    per-column convergence): the 1-GPU / N-GPU comparability of SURVEY.md section 4"""
    mg, tp, A = mg128
    B = torch.from_numpy(np.ascontiguousarray(probes(2048, 32))).cuda()
    X, it, _ = mg.dev.fgmres(2, B, 1e-12)
    X2, it2, _ = mg.dev.fgmres(2, B[:, 5:13].contiguous(), 1e-12)
    assert np.array_equal(it[5:13], it2)
    assert torch.equal(X[:, 5:13], X2)


def test_fgmres_128_all_levels(mg128, g128):
    mg, tp, A = mg128
    for lvl, k in ((0, 16), (1, 16), (2, 32)):
        Al = mg.ml.levels[lvl].A
        B = probes(Al.shape[0], k, seed=7 + lvl)
        X, iters, relres = mg.solve_batch(lvl, torch.from_numpy(np.ascontiguousarray(B)).cuda(), 1e-12)
        res = np.linalg.norm(B - Al @ host(X), axis=0) / np.linalg.norm(B, axis=0)
        print("level", lvl, "iters", iters.min(), iters.max(), "true relres", res.max())
        assert res.max() <= RES_MAX and relres.max() <= RES_MAX


def test_level_samples_16_match_reference(mg16, g16):
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg16
    np.random.seed(123456)
    e, it = utils.defl_Hutch_batch(mg, tp, "hutchinson", 0, None, 0, 8)
    assert np.abs(e - g16["plain_hutch_e"]).max() < 1e-8 * np.abs(g16["plain_hutch_e"]).max()
    mg.skip_level = False
    for lvl in (0, 1):
        e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, lvl, 8)
        ref = g16["plain_mlmc_l%d_e" % lvl]
        assert np.abs(e - ref).max() < 1e-8 * max(np.abs(ref).max(), 1.0), lvl
    # deflated Hutchinson (16 vectors), stream restarts
    np.random.seed(123456)
    e, it = utils.defl_Hutch_batch(mg, tp, "hutchinson", 16, g16["defl_Ux"], 0, 8)
    assert np.abs(e - g16["plain_hutch_defl16_e"]).max() < 1e-8 * np.abs(g16["plain_hutch_defl16_e"]).max()
    # the k = 1 reference-shaped entry point walks the same stream
    np.random.seed(123456)
    out = {"results": [{"function_iters": 0} for _ in range(3)]}
    e1, _ = utils.one_defl_Hutch_step(A, None, mg, tp, "hutchinson", 0, None, None)
    assert abs(e1 - g16["plain_hutch_e"][0]) < 1e-8 * abs(g16["plain_hutch_e"][0])


def test_level_samples_16_permuted(mg16perm, g16):
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg16perm
    np.random.seed(123456)
    e, it = utils.defl_Hutch_batch(mg, tp, "hutchinson", 0, None, 0, 8)
    assert np.abs(e - g16["perm_hutch_e"]).max() < 1e-8 * np.abs(g16["perm_hutch_e"]).max()
    for lvl in (0, 1):
        e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, lvl, 8)
        ref = g16["perm_mlmc_l%d_e" % lvl]
        assert np.abs(e - ref).max() < 1e-8 * max(np.abs(ref).max(), 1.0), lvl


def test_level_samples_128_match_reference(mg128, g128):
    """the shipped configuration (permuted, skip level 1): stream order of the golden file is
    1 Hutchinson probe, 3 level-0 probes, 16 level-2 probes"""
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg128
    np.random.seed(123456)
    e, it = utils.defl_Hutch_batch(mg, tp, "hutchinson", 0, None, 0, 1)
    assert abs(e[0] - g128["hutch_e"][0]) < 1e-8 * abs(g128["hutch_e"][0])
    e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, 0, 3)
    print("level-0 iters", it)
    assert np.abs(e - g128["mlmc_l0_e"]).max() < 1e-8 * np.abs(g128["mlmc_l0_e"]).max()
    e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, 2, 16)
    assert np.abs(e - g128["mlmc_l2_e"]).max() < 1e-8 * np.abs(g128["mlmc_l2_e"]).max()


def test_hutch_solution_128_subsample(mg128, g128):
    mg, tp, A = mg128
    np.random.seed(123456)
    x0 = (np.random.randint(2, size=32768) * 2 - 1).astype(np.complex128)
    rhs = np.roll(x0, 512)
    mg.level_nr = 0
    mg.solve(A, rhs, 1e-12)
    z = mg.x
    assert abs(np.linalg.norm(z) - g128["hutch_z_norm"]) < 1e-8 * g128["hutch_z_norm"]
    assert np.linalg.norm(z[::64] - g128["hutch_z_sub"]) < 1e-8 * np.linalg.norm(g128["hutch_z_sub"])
    assert mg.num_iters > 0


def test_full_size_batch_properties_128(mg128):
    """BASELINE config 2/3 size (k = 256 probes): residuals of every column, and the estimate of the
    level-0 difference equals x0^H z - x0^H P P A2^{-1} R R rhs recomputed from device pieces."""
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg128
    k = 256
    np.random.seed(99)
    st = np.random.get_state()
    e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 0, None, 0, k)
    assert np.all(np.isfinite(e.real)) and it[0].min() > 0 and it[0].max() < 60
    np.random.set_state(st)
    X0 = (np.random.randint(2, size=(k, 32768)) * 2 - 1).T.astype(np.complex128)
    rhs = np.roll(X0, 512, axis=0)
    Z, iters, relres = mg.solve_batch(0, torch.from_numpy(np.ascontiguousarray(rhs)).cuda(), 1e-12)
    res = np.linalg.norm(rhs - A @ host(Z), axis=0) / np.linalg.norm(rhs, axis=0)
    assert res.max() <= RES_MAX
    lv = mg.ml.levels
    xc = lv[1].R @ (lv[0].R @ rhs)
    Y, _, _ = mg.solve_batch(2, torch.from_numpy(np.ascontiguousarray(xc)).cuda(), 1e-12)
    w = lv[0].P @ (lv[1].P @ host(Y))
    e_ref = np.einsum("ij,ij->j", X0.conj(), host(Z)) - np.einsum("ij,ij->j", X0.conj(), w)
    assert np.abs(e - e_ref).max() < 1e-8 * np.abs(e_ref).max()


def test_mg_object_api(mg16, port16):
    """MG.solve / one_mg_step / matvec / diff_op keep the reference's shapes and semantics"""
    mg, tp, A = mg16
    mp, _ = port16
    b = probes(512, 1)[:, 0]
    mg.level_nr = 0
    y = mg.matvec(b)
    assert y.shape == (512,) and np.abs(y - A @ b).max() < 1e-12
    x = mg.one_mg_step(b)
    assert x.shape == (512,) and np.linalg.norm(b - A @ x) < np.linalg.norm(b)
    mg.solve(A, b, 1e-10)
    assert mg.x.shape == (512,) and np.linalg.norm(b - A @ mg.x) / np.linalg.norm(b) < 1e-9
    mg.level_for_diff_op = 0
    mg.skip_level = False
    mg.solve_tol = 1e-12
    mp.level_for_diff_op = 0; mp.skip_level = False; mp.solve_tol = 1e-12
    d_gpu = mg.diff_op(b)
    d_cpu = mp.diff_op(b)
    assert np.linalg.norm(d_gpu - d_cpu) < 1e-8 * np.linalg.norm(d_cpu)
    v = b.copy()
    mg.diff_op_Q(v)
    assert np.array_equal(v, b)              # input not mutated (the reference's quirk is not replicated)
    assert "size(A) = (512, 512)" in str(mg)


def test_deflated_mlmc_level_samples_16_match_reference(mg16, g16defl):
    """the deflated-MLMC branch: x_def = x - V V^H x on the device, then the same fused sample; estimates against
    the unmodified reference's for identical deflation vectors and probes (tolerance 1e-8), and the host part of
    deflation_pre_computations (sign fix, gamma3, tr1; utils.py:145-176) from the same raw eigensolver output"""
    from deflatedmlmc_schwinger_b200 import utils
    mg, tp, A = mg16
    mg.skip_level = False
    tp = dict(tp); tp["defl_type"] = "exact"; tp["diff_lev_op_tol"] = 1e-3
    for ix in range(2):
        Vx, Ux, tr1 = utils.deflation_pre_computations(A, 16, 1e-1, "mlmc", mg.timer, tp, mg, None, level_nr=ix,
                                                       eigpairs=(g16defl["l%d_Sy" % ix], g16defl["l%d_eigvecs" % ix]))
        assert np.abs(Vx - g16defl["l%d_Vx" % ix]).max() < 1e-12 and np.abs(Ux - g16defl["l%d_Ux" % ix]).max() < 1e-12
        assert abs(tr1 - g16defl["l%d_tr1" % ix]) < 1e-10 * abs(tr1)
        np.random.seed(123456 + ix)
        e, it = utils.defl_Hutch_batch(mg, tp, "mlmc", 16, Vx, ix, 8)
        ref = g16defl["l%d_e" % ix]
        assert np.abs(e - ref).max() < 1e-8 * max(np.abs(ref).max(), 1.0), ix
    mg.dev.set_deflation(0, None); mg.dev.set_deflation(1, None)
    mg.__dict__.pop("_defl_cache", None)


@pytest.mark.parametrize("k", [1, 8, 48, 128])
def test_cuda_graph_replay_of_the_vcycle_is_bit_identical(mg128, k):
    """the V-cycle's launch sequence is replayed as a CUDA graph (eager, capture, replay ...; k <= 64 through staging
    buffers, k = 128 with one graph per Krylov vector pair): solutions and iteration counts must equal the eager
    path's bit for bit, on levels 0 and 2"""
    mg, tp, A = mg128
    for lvl, n in ((0, 32768), (2, 2048)):
        B = torch.from_numpy(np.ascontiguousarray(probes(n, k, seed=77 + k))).cuda()
        mg.dev.set_option("use_graphs", 0)
        X0, it0, rr0 = mg.dev.fgmres(lvl, B, 1e-12)
        mg.dev.set_option("use_graphs", 1)
        X1, it1, rr1 = mg.dev.fgmres(lvl, B, 1e-12)
        X2, it2, rr2 = mg.dev.fgmres(lvl, B, 1e-12)          # graphs already instantiated: replay from the first iteration
        assert torch.equal(X0, X1) and torch.equal(X0, X2)
        assert np.array_equal(it0, it1) and np.array_equal(it0, it2)
        assert np.all(rr1 < 1.3e-12)


def test_geometric_preconditioner_128(mg128):
    """the level-0 solve preconditioned by the geometric hierarchy: same solutions as with the estimator's own
    hierarchy as preconditioner (both to the true residual 1e-12), in far fewer outer iterations"""
    mg, tp, A = mg128
    assert mg.precond_mg is not None
    A0 = mg.ml.levels[0].A
    B = probes(A0.shape[0], 16, seed=3)
    Bd = torch.from_numpy(np.ascontiguousarray(B)).cuda()
    X, it, relres = mg.dev.fgmres(0, Bd, 1e-12)
    res = np.linalg.norm(B - A0 @ host(X), axis=0) / np.linalg.norm(B, axis=0)
    assert res.max() <= RES_MAX and relres.max() <= RES_MAX
    mg.dev.set_preconditioner(0, None)
    try:
        X2, it2, relres2 = mg.dev.fgmres(0, Bd, 1e-12)
    finally:
        mg.dev.set_preconditioner(0, mg.precond_mg.dev, 0)
    print("outer iterations: geometric", it.min(), it.max(), " reference aggregation", it2.min(), it2.max())
    assert relres2.max() < 1e-12
    assert np.abs(host(X) - host(X2)).max() < 1e-8 * np.abs(host(X2)).max()
    assert it.max() < it2.min()
    # a probe's solution does not depend on its batch with the second hierarchy either
    X3, it3, _ = mg.dev.fgmres(0, Bd[:, 3:9].contiguous(), 1e-12)
    assert np.array_equal(it[3:9], it3) and torch.equal(X[:, 3:9], X3)


@pytest.mark.parametrize("k", [2, 16])
def test_fused_vcycle_io_is_bit_identical(mg128, k):
    """option fuse_io (complex64 copy of the basis vector written by the normalisation kernel, complex128 output
    written by the last smoother factor, prolongation without the zero read) changes no bit of the solve"""
    mg, tp, A = mg128
    B = torch.from_numpy(np.ascontiguousarray(probes(mg.level_shapes[0], k, seed=11))).cuda()
    X2, it2, _ = mg.dev.fgmres(0, B, 1e-12)          # default: also fuse_res (the residual stored as BF16: not bit-identical)
    mg.set_option("fuse_res", 0)
    mg.set_option("smoother_eo", 0)                  # (the even-odd smoother needs fuse_io)
    try:
        X1, it1, _ = mg.dev.fgmres(0, B, 1e-12)
        mg.set_option("fuse_io", 0)
        X0, it0, _ = mg.dev.fgmres(0, B, 1e-12)
    finally:
        mg.set_option("fuse_io", 1)
        mg.set_option("fuse_res", 1)
        mg.set_option("smoother_eo", 1)
    assert np.array_equal(it0, it1) and torch.equal(X0, X1)
    assert np.abs(it2.astype(int) - it1.astype(int)).max() <= 1
    assert np.abs(host(X2) - host(X1)).max() < 1e-9 * np.abs(host(X1)).max()


def test_geometric_preconditioner_of_the_level1_solve(g128):
    """lattices whose level 1 is too large for a dense inverse get a geometric hierarchy for the level-1 solve as well
    (2 strips in x, both halves, split by spin); forced here on 128^2 with dense_coarse_threshold = 2048"""
    from conftest import make_mg, params128
    mg, tp, A = make_mg(params128(), "mlmc", [g128["tv0"], g128["tv1"], g128["tv2"]], smoother_degree=32,
                        dense_coarse_threshold=2048)
    assert mg.precond_mg1 is not None and mg.precond_mg1.level_shapes == [8192, 2048, 512]
    assert mg.precond_mg1.level0_format == "bsr4" and 1 not in mg.dense_levels
    A1 = mg.ml.levels[1].A
    B = probes(A1.shape[0], 12, seed=21)
    Bd = torch.from_numpy(np.ascontiguousarray(B)).cuda()
    X, it, relres = mg.dev.fgmres(1, Bd, 1e-12)
    res = np.linalg.norm(B - A1 @ host(X), axis=0) / np.linalg.norm(B, axis=0)
    assert res.max() <= RES_MAX and relres.max() <= RES_MAX
    mg.dev.set_preconditioner(1, None)
    X2, it2, relres2 = mg.dev.fgmres(1, Bd, 1e-12)
    print("level-1 outer iterations: geometric", it.min(), it.max(), " estimator's hierarchy", it2.min(), it2.max())
    assert relres2.max() < 1e-12
    assert np.abs(host(X) - host(X2)).max() < 1e-8 * np.abs(host(X2)).max()
    assert it.max() < it2.min()
    # the level-0 solve of this hierarchy (its geometric level 1 is not dense either: BSR smoother level inside the cycle)
    A0 = mg.ml.levels[0].A
    B0 = probes(A0.shape[0], 4, seed=22)
    X0, it0, rr0 = mg.dev.fgmres(0, torch.from_numpy(np.ascontiguousarray(B0)).cuda(), 1e-12)
    res0 = np.linalg.norm(B0 - A0 @ host(X0), axis=0) / np.linalg.norm(B0, axis=0)
    assert res0.max() <= RES_MAX and it0.max() <= 14


@pytest.mark.parametrize("k", [2, 16])
def test_outer_solve_on_the_even_odd_schur_complement(mg128, k):
    """option outer_eo (the default): FGMRES on S x_e = b^_e with half-lattice Krylov vectors gives the solution of A x = b to the
    same tolerance in (about) the same number of iterations as the solve on the full system (outer_eo = 0)"""
    mg, tp, A = mg128
    A0 = mg.ml.levels[0].A
    B = probes(A0.shape[0], k, seed=31)
    Bd = torch.from_numpy(np.ascontiguousarray(B)).cuda()
    mg.set_option("outer_eo", 0)
    try:
        X0, it0, rr0 = mg.dev.fgmres(0, Bd, 1e-12)
    finally:
        mg.set_option("outer_eo", 1)
    X1, it1, rr1 = mg.dev.fgmres(0, Bd, 1e-12)
    res = np.linalg.norm(B - A0 @ host(X1), axis=0) / np.linalg.norm(B, axis=0)
    print("outer_eo iterations", it1.min(), it1.max(), "full", it0.min(), it0.max(), "true relres", res.max())
    assert res.max() <= RES_MAX
    assert np.abs(host(X1) - host(X0)).max() < 1e-8 * np.abs(host(X0)).max()
    assert np.abs(it1.astype(int) - it0.astype(int)).max() <= 2
