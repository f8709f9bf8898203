"""Set-up on the device (reference multigrid.py:190-277, 342-344): the Galerkin product R A P, the
batched orthonormalisation of the geometric aggregates and the dense coarsest inverse, each against its host counterpart
(scipy / numpy -- what the reference itself calls)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import params128, params16, make_mg

pytestmark = pytest.mark.gpu


def _rel(a, b):
    d = abs(sp.csr_matrix(a) - sp.csr_matrix(b))
    return (d.max() if d.nnz else 0.0) / abs(sp.csr_matrix(b)).max()


def test_device_galerkin_matches_scipy_on_every_level_128(g128, mg128):
    """the hierarchy the drivers build (device R A P, device coarsest inverse) against the same set-up with scipy's R*A*P and
    np.linalg.inv: the estimator's levels AND the geometric preconditioner hierarchy"""
    mg, _, _ = mg128
    p = params128()
    p["host_galerkin"] = True
    mh, _, _ = make_mg(p, "mlmc", [g128["tv0"], g128["tv1"], g128["tv2"]], smoother_degree=32)
    assert len(mg.ml.levels) == len(mh.ml.levels) == 4
    for l in range(1, 4):
        # the transfer values of the two builds agree to rounding, so the products do (relative to the largest entry)
        assert _rel(mg.ml.levels[l].A, mh.ml.levels[l].A) < 1e-13, l
        # and the device product IS R A P of the device build's own (host-visible) factors
        lv = mg.ml.levels[l - 1]
        assert _rel(mg.ml.levels[l].A, lv.R @ sp.csr_matrix(lv.A) @ lv.P) < 1e-14, l
    assert np.abs(mg.coarsest_inv - mh.coarsest_inv).max() < 1e-11 * np.abs(mh.coarsest_inv).max()
    assert np.abs(mg.coarsest_inv @ mg.ml.levels[-1].A.toarray() - np.eye(512)).max() < 1e-11
    pg, ph = mg.precond_mg, mh.precond_mg
    assert pg is not None and ph is not None and len(pg.ml.levels) == len(ph.ml.levels)
    for l in range(1, len(pg.ml.levels)):
        lv = pg.ml.levels[l - 1]
        assert _rel(pg.ml.levels[l].A, lv.R @ sp.csr_matrix(lv.A) @ lv.P) < 1e-14, l
        # same coarse SPACE as the host QR (the bases may differ by a unitary factor per block): P P^H agrees
        Pg, Ph = pg.ml.levels[l - 1].P, ph.ml.levels[l - 1].P
        assert _rel(Pg @ Pg.conj().T, Ph @ Ph.conj().T) < 1e-10, l
        assert _rel(Pg.conj().T @ Pg, sp.identity(Pg.shape[1])) < 1e-13


def test_device_galerkin_of_a_small_stencil_with_wrapping_neighbours():
    """LT = 2: forward and backward t-neighbour are the same site (two entries on one column); indexed aggregates in an
    arbitrary (non-contiguous) arrangement; the BSR -> BSR product one level further down"""
    import torch
    from deflatedmlmc_schwinger_b200 import _lib, lattice, multigrid
    LX, LT, nv = 8, 2, 2
    links = lattice.random_u1_links(LX, 5, LT=LT)
    A0 = sp.csr_matrix(lattice.wilson_matrix(links, -0.05))
    n0 = A0.shape[0]
    rs = np.random.RandomState(3)
    cblk0 = rs.permutation(np.repeat(np.arange(n0 // 4), 4)).astype(np.int32)
    V0 = rs.standard_normal((n0, nv)) + 1j * rs.standard_normal((n0, nv))
    dev = _lib.Hierarchy(3)
    dev.set_stencil(0, links, 4.0 - 0.05)
    pv0, rows0 = dev.block_orthonormal_values(V0, cblk0, nv)
    pv0 = pv0.cpu().numpy()
    assert np.array_equal(np.sort(rows0.cpu().numpy().ravel()), np.arange(n0))
    P0 = multigrid.prolongator_csr_indexed(pv0, cblk0)
    assert _rel(P0.conj().T @ P0, sp.identity(P0.shape[1])) < 1e-14
    dev.set_transfer_indexed(0, n0, nv, pv0, cblk0)
    col1, vals1 = dev.galerkin(0, nv, P0.shape[1], cap=2)          # cap too small on purpose: the binding retries
    A1 = multigrid.csr_from_padded_bsr(col1.cpu().numpy(), vals1.cpu().numpy())
    ref1 = (P0.conj().T @ A0 @ P0).tocsr()
    assert _rel(A1, ref1) < 1e-14
    c1 = col1.cpu().numpy()
    assert all(np.all(np.diff(r[r >= 0]) > 0) for r in c1), "block columns of a row must be sorted and distinct"
    # level 1 -> 2 with the reference's closed-form aggregates (aggr_size 8, dofi 2 * nv)
    n1 = A1.shape[0]
    V1 = rs.standard_normal((n1, nv)) + 1j * rs.standard_normal((n1, nv))
    aggr, dofi = 8, 2 * nv
    pv1 = dev.prolongator_values(V1, aggr, dofi, nv).cpu().numpy()
    P1 = multigrid.prolongator_csr(pv1, aggr, dofi, nv)
    dev.set_transfer(1, n1, aggr, dofi, nv, pv1)
    col2, vals2 = dev.galerkin(1, nv, P1.shape[1])
    A2 = multigrid.csr_from_padded_bsr(col2.cpu().numpy(), vals2.cpu().numpy())
    assert _rel(A2, (P1.conj().T @ ref1 @ P1).tocsr()) < 1e-14
    # the installed coarse operators act like the matrices
    X = torch.from_numpy(rs.standard_normal((n1, 3)) + 1j * rs.standard_normal((n1, 3))).cuda()
    assert np.abs(dev.spmm(1, X).cpu().numpy() - ref1 @ X.cpu().numpy()).max() < 1e-13
    dev.close()


@pytest.mark.parametrize("n", [1, 7, 64, 512, 1000])
def test_dense_inverse_on_the_device(n):
    from deflatedmlmc_schwinger_b200 import _lib
    rs = np.random.RandomState(n)
    M = rs.standard_normal((n, n)) + 1j * rs.standard_normal((n, n))
    if n >= 7:
        M[np.arange(n), np.arange(n)] = 0.0        # zero diagonal: every step needs its row exchange
    dev = _lib.Hierarchy(1)
    X = dev.dense_inverse(M).cpu().numpy()
    ref = np.linalg.inv(M)
    assert np.abs(X - ref).max() < 1e-10 * np.abs(ref).max() * max(1.0, np.linalg.cond(M) / 1e4)
    assert np.abs(X @ M - np.eye(n)).max() < 1e-9
    with pytest.raises(_lib.DmlmcError):
        dev.dense_inverse(np.zeros((4, 4), dtype=np.complex128))
    dev.close()


def test_setup_16_device_and_host_builds_give_the_same_samples(g16):
    """end to end: per-probe level samples of the hierarchy built on the device equal those of the host-built one"""
    import torch
    tvs = [g16["tv0"], g16["tv1"]]
    md, tp, _ = make_mg(params16(), "mlmc", tvs, smoother_degree=8)
    p = params16()
    p["host_galerkin"] = True
    p["host_prolongator"] = True
    mh, _, _ = make_mg(p, "mlmc", tvs, smoother_degree=8)
    n0 = md.level_shapes[0]
    rs = np.random.RandomState(1)
    X0 = torch.from_numpy((2.0 * rs.randint(2, size=(n0, 8)) - 1.0).astype(np.complex128)).cuda()
    for lf in range(len(md.level_shapes) - 1):
        Xl = X0 if lf == 0 else torch.from_numpy(
            (2.0 * rs.randint(2, size=(md.level_shapes[lf], 8)) - 1.0).astype(np.complex128)).cuda()
        ed, _ = md.dev.level_sample(1, lf, lf + 1, Xl, 1e-12, 40, 1000)
        eh, _ = mh.dev.level_sample(1, lf, lf + 1, Xl, 1e-12, 40, 1000)
        assert np.abs(ed.cpu().numpy() - eh.cpu().numpy()).max() < 1e-9 * np.abs(eh.cpu().numpy()).max()


def test_even_odd_smoother_setup_on_the_device_matches_the_host(mg128):
    """the Schur complement as a device callable against scipy's S; the GMRES polynomial of S from the device Arnoldi run against
    the host's (same start vector): the same polynomial as a function on the spectrum; the BF16-storage check agrees"""
    import torch
    from deflatedmlmc_schwinger_b200 import multigrid as mgm
    mg, _, A = mg128
    pm = mg.precond_mg                      # the hierarchy whose level 0 carries the even-odd smoother
    L = 128
    c = 4.0 + params128()["matrix_params"]["mass"]
    S, c_host = mgm.even_odd_schur(A, L, L)
    assert abs(c - c_host) < 1e-14
    apply_S, even_rows = mgm.even_odd_schur_device(pm.dev, L, L, c)
    n0 = A.shape[0]
    rs = np.random.RandomState(4)
    xe = rs.standard_normal((n0 // 2, 2)) + 1j * rs.standard_normal((n0 // 2, 2))
    X = torch.zeros((n0, 2), dtype=torch.complex128, device=pm.dev.device)
    X[even_rows] = torch.from_numpy(xe).to(X.device)
    Y = apply_S(X).cpu().numpy()
    ie = even_rows.cpu().numpy()
    assert np.abs(Y[ie] - S @ xe).max() < 1e-12
    assert np.abs(np.delete(Y, ie, axis=0)).max() == 0.0
    de = 16
    om_d = mgm.harmonic_ritz_inv_roots_device(apply_S, n0, de, pm.dev.device, support=even_rows)
    om_h = mgm.harmonic_ritz_inv_roots(S, de)
    z = np.linspace(0.05, 7.5, 200)
    res = lambda om: np.prod(1.0 - np.outer(z, om), axis=1)          # the GMRES residual polynomial prod (1 - omega_i z)
    assert np.abs(res(om_d) - res(om_h)).max() < 1e-6
    nu, p0 = mgm.smoother_product_form(om_d)
    e_d = mgm.smoother_storage_error_op(apply_S, n0, even_rows, om_d, nu, p0, pm.dev.device)
    e_h = mgm.smoother_storage_error(S, om_h, *mgm.smoother_product_form(om_h), 'bf16')
    assert e_d < 0.15 and e_h < 0.15 and abs(e_d - e_h) < 0.5 * max(e_d, e_h) + 1e-3
    # and the polynomial the set-up installed is this one
    assert pm.eo_poly is not None and len(pm.eo_poly[0]) == de - 1
    assert np.abs(np.sort_complex(pm.eo_poly[0]) - np.sort_complex(nu)).max() < 1e-9 * np.abs(nu).max()
