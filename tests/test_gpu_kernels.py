"""GPU tier: every kernel behind the C ABI against scipy/numpy (the oracle's arithmetic) on the
same inputs.  complex128 kernels must agree to rounding (1e-13 relative), complex64 ones to 2e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.complex128: 1e-13, torch.complex64: 2e-5}


def rnd(n, k, dtype, seed=0):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal((n, k)) + 1j * rs.standard_normal((n, k))
    return torch.from_numpy(a).to("cuda").to(dtype).contiguous()


def relerr(a, b):
    a = np.asarray(a, dtype=np.complex128); b = np.asarray(b, dtype=np.complex128)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def host(t):
    return t.cpu().numpy().astype(np.complex128)


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("k", [1, 3, 8, 64])
def test_spmm_all_levels_128(mg128, dtype, k):
    mg, tp, A = mg128
    for lvl in range(3):
        Al = mg.ml.levels[lvl].A
        X = rnd(Al.shape[0], k, dtype, seed=lvl)
        Y = mg.dev.spmm(lvl, X)
        assert relerr(host(Y), Al @ host(X)) < TOL[dtype], (lvl, k)


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
def test_spmm_16_bs2(mg16, dtype):
    mg, tp, A = mg16
    for lvl in range(2):
        Al = mg.ml.levels[lvl].A
        X = rnd(Al.shape[0], 5, dtype)
        assert relerr(host(mg.dev.spmm(lvl, X)), Al @ host(X)) < TOL[dtype]


def test_spmm_probe_input_is_exact_linear(mg128):
    # +-1 inputs (the actual probes), k = 256 (BASELINE config 3), linearity A(x+y) = Ax + Ay
    mg, tp, A = mg128
    rs = np.random.RandomState(1)
    X = torch.from_numpy((rs.randint(2, size=(32768, 256)) * 2 - 1).astype(np.complex128)).cuda()
    Y = mg.dev.spmm(0, X)
    ref = A @ host(X[:, :4])
    assert relerr(host(Y[:, :4]), ref) < 1e-14
    Z = rnd(32768, 256, torch.complex128, 5)
    lhs = mg.dev.spmm(0, X + Z)
    rhs = Y + mg.dev.spmm(0, Z)
    assert relerr(host(lhs), host(rhs)) < 1e-13


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("k", [1, 6, 32])
def test_restrict_prolong(mg128, dtype, k):
    mg, tp, A = mg128
    for lvl in range(3):
        P, R = mg.ml.levels[lvl].P, mg.ml.levels[lvl].R
        Xf = rnd(P.shape[0], k, dtype, 10 + lvl)
        Xc = mg.dev.restrict(lvl, Xf)
        assert relerr(host(Xc), R @ host(Xf)) < TOL[dtype]
        Yc = rnd(P.shape[1], k, dtype, 20 + lvl)
        Yf = rnd(P.shape[0], k, dtype, 30 + lvl)
        ref = host(Yf) + P @ host(Yc)
        mg.dev.prolong_add(lvl, Yc, Yf)
        assert relerr(host(Yf), ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("k", [1, 7, 64])
def test_coarsest_apply(mg128, dtype, k):
    mg, tp, A = mg128
    B = rnd(512, k, dtype, 3)
    X = mg.dev.coarsest_apply(B)
    ref = mg.coarsest_inv @ host(B)
    tol = 1e-12 if dtype == torch.complex128 else 5e-5
    assert relerr(host(X), ref) < tol


def test_perm_and_probe_expand(mg128):
    mg, tp, A = mg128
    for lvl in range(4):
        lv = mg.ml.levels[lvl]
        n = lv.A.shape[0]
        X = rnd(n, 5, torch.complex128, 40 + lvl)
        Y = mg.dev.apply_perm(lvl, X)
        ref = lv.Bblock_perm @ (lv.Pperm.transpose() @ host(X))
        assert relerr(host(Y), ref) < 1e-13, lvl
    np.random.seed(123456)
    from deflatedmlmc_schwinger_b200 import utils, sampling
    bits = sampling.draw_probe_bits(2048 * 6)
    dev_bits = torch.from_numpy(utils.pack_bits(bits)).cuda()
    X0 = mg.dev.probe_expand(dev_bits, 2048, 6)
    ref = (bits.reshape(6, 2048).T.astype(np.float64) * 2 - 1)
    assert np.array_equal(host(X0).real, ref) and np.all(host(X0).imag == 0)


def test_dotc_and_deflate(mg16, g16):
    mg, tp, A = mg16
    X = rnd(512, 9, torch.complex128, 1); Y = rnd(512, 9, torch.complex128, 2)
    d = mg.dev.dotc(X, Y)
    ref = np.einsum("ij,ij->j", np.conj(host(X)), host(Y))
    assert relerr(host(d), ref) < 1e-13
    V = g16["defl_Ux"]
    mg.dev.set_deflation(0, V)
    Xd = X.clone()
    mg.dev.deflate(0, Xd)
    ref = host(X) - V @ (V.conj().T @ host(X))
    assert relerr(host(Xd), ref) < 1e-13
    mg.dev.set_deflation(0, None)
    mg.__dict__.pop("_defl_cache", None)


@pytest.mark.parametrize("d", [4, 8, 16, 36, 64, 6])
@pytest.mark.parametrize("k", [1, 30, 256])
def test_deflation_projection_tensor_cores(mg128, d, k):
    """x - V (V^H x), complex128: FP64 tensor-core kernels (d % 4 == 0) and the SIMT kernels (d = 6, or
    option defl_tensor = 0) against numpy, ragged column counts included"""
    mg, tp, A = mg128
    n = mg.level_shapes[0]
    rs = np.random.RandomState(d)
    V, _ = np.linalg.qr(rs.standard_normal((n, d)) + 1j * rs.standard_normal((n, d)))
    X = rnd(n, k, torch.complex128, 90 + k)
    ref = host(X) - V @ (V.conj().T @ host(X))
    mg.dev.set_deflation(0, V)
    Xt = X.clone(); mg.dev.deflate(0, Xt)
    mg.dev.set_option("defl_tensor", 0)
    Xs = X.clone(); mg.dev.deflate(0, Xs)
    mg.dev.set_option("defl_tensor", 1)
    mg.dev.set_deflation(0, None)
    mg.__dict__.pop("_defl_cache", None)
    assert relerr(host(Xt), ref) < 1e-13 and relerr(host(Xs), ref) < 1e-13


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
def test_smoother_is_the_polynomial(mg128, dtype):
    from deflatedmlmc_schwinger_b200.multigrid import harmonic_ritz_inv_roots
    from scipy.sparse import csr_matrix
    mg, tp, A = mg128
    for lvl in range(3):
        Al = csr_matrix(mg.ml.levels[lvl].A)
        w = harmonic_ritz_inv_roots(Al, mg.level_degree(lvl))
        R = rnd(Al.shape[0], 4, dtype, 50 + lvl)
        E = mg.dev.smooth(lvl, R)
        r = host(R); e = np.zeros_like(r)
        for wi in w:
            e = e + wi * r
            r = r - wi * (Al @ r)
        tol = 1e-11 if dtype == torch.complex128 else 2e-3
        if dtype == torch.complex64:
            # the complex64 smoother keeps the intermediate vectors of the product in BF16 (FP32 arithmetic)
            assert relerr(host(E), e) < 1e-1, lvl
            mg.dev.set_option("stencil_fast", 0)            # generic kernel on the same BF16-stored data
            Eg = mg.dev.smooth(lvl, R)
            mg.dev.set_option("stencil_fast", 1)
            assert relerr(host(Eg), e) < 1e-1 and relerr(host(Eg), host(E)) < 5e-2, lvl
            if lvl == 0:                                    # shared-memory-tiled variant: same arithmetic
                mg.dev.set_option("stencil_smem", 1)
                Es = mg.dev.smooth(lvl, R)
                mg.dev.set_option("stencil_smem", 0)
                assert relerr(host(Es), host(E)) < 1e-6, "smem variant"
                for t2 in (0, 1):                               # 1 / 2 sites per thread: same arithmetic
                    mg.dev.set_option("stencil_t2", t2)
                    Et = mg.dev.smooth(lvl, R)
                    assert relerr(host(Et), host(E)) < 1e-6, "sites-per-thread variant %d" % t2
                mg.dev.set_option("stencil_t2", 1)
            mg.dev.set_option("smoother_half", 0)
            E = mg.dev.smooth(lvl, R)
            mg.dev.set_option("smoother_half", 1)
        assert relerr(host(E), e) < tol, lvl


_DENSE_LU = {}


def _vcycle_numpy(mg, b, l0, dense_from=None):
    """numpy restatement of the device V-cycle; levels >= dense_from are solved exactly"""
    from deflatedmlmc_schwinger_b200.multigrid import harmonic_ritz_inv_roots
    from scipy.sparse import csr_matrix, csc_matrix
    from scipy.sparse.linalg import splu
    lv = mg.ml.levels
    nl = len(lv)
    if dense_from is None:
        dense_from = min(l for l, kind in mg.dense_levels.items() if kind == "host")
    if l0 == nl - 1:
        return mg.coarsest_inv @ b
    if l0 >= dense_from:
        if (id(mg), l0) not in _DENSE_LU:
            _DENSE_LU[(id(mg), l0)] = splu(csc_matrix(lv[l0].A))
        return _DENSE_LU[(id(mg), l0)].solve(b)
    Al = csr_matrix(lv[l0].A)
    w = harmonic_ritz_inv_roots(Al, mg.level_degree(l0))
    r = b.copy(); x = np.zeros_like(b)
    if mg.pre_smooth:
        for wi in w:
            x = x + wi * r; r = r - wi * (Al @ r)
    x = x + lv[l0].P @ _vcycle_numpy(mg, lv[l0].R @ r, l0 + 1, dense_from)
    r = b - Al @ x
    for wi in w:
        x = x + wi * r; r = r - wi * (Al @ r)
    return x


@pytest.mark.parametrize("pre", [False, True])
@pytest.mark.parametrize("l0", [0, 1, 2, 3])
def test_vcycle_matches_numpy_restatement(mg128, l0, pre):
    mg, tp, A = mg128
    mg.pre_smooth = pre
    mg.dev.set_option("pre_smooth", 1 if pre else 0)
    try:
        _check_vcycle(mg, l0)
    finally:
        mg.pre_smooth = False
        mg.dev.set_option("pre_smooth", 0)


def _check_vcycle(mg, l0):
    n = mg.level_shapes[l0]
    B = rnd(n, 3, torch.complex128, 60 + l0)
    X = mg.dev.vcycle(l0, B)
    ref = _vcycle_numpy(mg, host(B), l0)
    assert relerr(host(X), ref) < 1e-8
    # the complex64 cycle bottoms out at the first dense level of any kind; a tensor-core (BF16) inverse
    # is a preconditioner-grade solve
    first = min(mg.dense_levels)
    bf16 = (l0 < first and (mg.dense_levels[first] == "tensor" or mg.level_shapes[first] >= 1024)) or \
           (l0 == first and mg.dense_levels[first] == "tensor")
    Xf = mg.dev.vcycle(l0, B.to(torch.complex64))
    assert relerr(host(Xf), _vcycle_numpy(mg, host(B), l0, first)) < (1e-1 if (bf16 or l0 == 0) else 5e-3)


def _bf16_round(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.real.float().bfloat16().double() + 1j * t.imag.float().bfloat16().double()).numpy()


@pytest.mark.parametrize("k", [256, 64, 19, 2])
def test_tensor_core_dense_apply(mg128, k):
    """tcgen05 kernel: X = Minv B with BF16 operands and FP32 accumulation == the same product formed in
    float64 from BF16-rounded operands (level 2, n = 2048: 32 CTAs x 64 k-blocks, ragged N for k = 19, 2)."""
    mg, tp, A = mg128
    lvl = 2
    n = mg.level_shapes[lvl]
    Minv = np.linalg.inv(mg.ml.levels[lvl].A.toarray())
    B = rnd(n, k, torch.complex64, 80 + k)
    mg.dev.set_option("dense_tensor_min_n", 1024)
    mg.dev.set_option("dense_direct_exact", 0)
    X = mg.dev.vcycle(lvl, B)
    mg.dev.set_option("dense_direct_exact", 1)
    ref = _bf16_round(Minv) @ _bf16_round(host(B))
    assert relerr(host(X), ref) < 2e-5
    assert relerr(host(X), Minv @ host(B)) < 2e-2
    exact = Minv @ host(B)
    Xs = mg.dev.vcycle(lvl, B)                # direct solve of the level: split-BF16 tensor-core GEMM
    e_split = relerr(host(Xs), exact)
    mg.dev.set_option("dense_split_bf16", 0)
    Xf = mg.dev.vcycle(lvl, B)                # FP32 SIMT kernel
    mg.dev.set_option("dense_split_bf16", 1)
    e_f32 = relerr(host(Xf), exact)
    print("k", k, "split-BF16 error", e_split, "FP32 error", e_f32)
    assert e_f32 < 2e-5 and e_split < 5e-5
    assert not torch.equal(Xs, Xf)            # the tensor-core path was taken


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("k", [70, 129])
def test_vcycle_column_chunks(mg128, dtype, k):
    """k = 70 / 129: level-0 chunks of 64 columns plus a ragged (even / odd) remainder; every column
    must equal the numpy restatement and must not depend on how the batch was chunked."""
    mg, tp, A = mg128
    B = rnd(mg.level_shapes[0], k, dtype, 70)
    mg.dev.set_option("chunk_cols", 64)
    X = mg.dev.vcycle(0, B)
    cols = [0, 1, 63, 64, k - 1]
    first = min(mg.dense_levels)
    ref = _vcycle_numpy(mg, host(B)[:, cols], 0, None if dtype == torch.complex128 else first)
    assert relerr(host(X)[:, cols], ref) < (1e-8 if dtype == torch.complex128 else 1e-1)
    if k % 2 == 0:      # same kernel variants whatever the chunking: bitwise identical
        mg.dev.set_option("chunk_cols", 100000)
        X1 = mg.dev.vcycle(0, B)
        assert torch.equal(X, X1)
    mg.dev.set_option("chunk_cols", 0)
    mg.dev.set_option("l2_budget_mb", 72)
    X2 = mg.dev.vcycle(0, B)
    mg.dev.set_option("l2_budget_mb", 0)
    assert relerr(host(X2), host(X)) < (1e-12 if dtype == torch.complex128 else 1e-4)


def test_smoother_product_form_equals_richardson(mg128):
    """host: p0 prod (1 - nu_i z) is the same polynomial as the Richardson accumulation over omega"""
    from deflatedmlmc_schwinger_b200.multigrid import harmonic_ritz_inv_roots, smoother_product_form
    from scipy.sparse import csr_matrix
    mg, tp, A = mg128
    Al = csr_matrix(mg.ml.levels[1].A)
    w = harmonic_ritz_inv_roots(Al, 24)
    nu, p0 = smoother_product_form(w)
    assert nu.shape == (23,)
    b = host(rnd(Al.shape[0], 1, torch.complex128, 3))[:, 0]
    r = b.copy(); e = np.zeros_like(b)
    for wi in w:
        e = e + wi * r; r = r - wi * (Al @ r)
    y = b.copy()
    for v in nu:
        y = y - v * (Al @ y)
    assert relerr(p0 * y, e) < 1e-11


def test_argument_errors_are_reported(mg16):
    from deflatedmlmc_schwinger_b200 import _lib
    mg, tp, A = mg16
    X = rnd(512, 2, torch.complex128)
    with pytest.raises(_lib.DmlmcError):
        mg.dev.lib.dmlmc_spmm.restype  # noqa: B018
        _lib._check(mg.dev.lib.dmlmc_spmm(mg.dev.h, 7, 0, X.data_ptr(), X.data_ptr(), 2))
    with pytest.raises(_lib.DmlmcError):
        _lib._check(mg.dev.lib.dmlmc_fgmres(mg.dev.h, 2, X.data_ptr(), X.data_ptr(), 2, 1e-8, 10, 10, None, None))


@pytest.mark.parametrize("dtype", [torch.complex128, torch.complex64])
@pytest.mark.parametrize("k", [1, 6, 32])
def test_indexed_transfers_of_the_geometric_hierarchy(mg128, dtype, k):
    """restrict / prolong on the geometric (4 x 4 sites, spin-split; then 2 x 2) aggregates of the preconditioner
    hierarchy against its scipy P, R"""
    mg, tp, A = mg128
    pm = mg.precond_mg
    assert pm is not None and pm.level_shapes == [32768, 8192, 2048, 512]
    for lvl in range(3):
        P, R = pm.ml.levels[lvl].P, pm.ml.levels[lvl].R
        Xf = rnd(P.shape[0], k, dtype, 40 + lvl)
        Xc = pm.dev.restrict(lvl, Xf)
        assert relerr(host(Xc), R @ host(Xf)) < TOL[dtype]
        Yc = rnd(P.shape[1], k, dtype, 50 + lvl)
        Yf = rnd(P.shape[0], k, dtype, 60 + lvl)
        ref = host(Yf) + P @ host(Yc)
        pm.dev.prolong_add(lvl, Yc, Yf)
        assert relerr(host(Yf), ref) < TOL[dtype]
    # Galerkin operators of the geometric levels (BSR on the device) against scipy
    for lvl in (1, 2):
        Al = pm.ml.levels[lvl].A
        X = rnd(Al.shape[0], k, dtype, 70 + lvl)
        assert relerr(host(pm.dev.spmm(lvl, X)), Al @ host(X)) < TOL[dtype]


def test_even_odd_smoother_cycle_matches_numpy(mg128):
    """the preconditioner of the level-0 solve as the solver applies it (geometric two-grid cycle, tcgen05 coarse solve,
    even-odd Schur-complement post-smoother in BF16 storage) against its complex128 restatement"""
    from scipy.sparse import csc_matrix
    from scipy.sparse.linalg import splu
    from deflatedmlmc_schwinger_b200.multigrid import even_odd_schur
    mg, tp, A = mg128
    pm = mg.precond_mg
    assert pm is not None and pm.eo_poly is not None
    A0 = pm.ml.levels[0].A.tocsr()
    P, R = pm.ml.levels[0].P, pm.ml.levels[0].R
    lu = splu(csc_matrix(pm.ml.levels[1].A))
    L = 128
    S, c = even_odd_schur(A0, L, L)
    s, x, t = np.meshgrid(np.arange(2), np.arange(L), np.arange(L), indexing="ij")
    par = ((x + t) & 1).ravel()
    ie, io = np.where(par == 0)[0], np.where(par == 1)[0]
    Heo, Hoe = A0[ie][:, io], A0[io][:, ie]
    nu, p0 = pm.eo_poly
    B = rnd(A0.shape[0], 4, torch.complex128, 91)
    b = host(B)
    xc = P @ lu.solve(R @ b)
    r = b - A0 @ xc
    y = r[ie] - Heo @ r[io] / c
    for v in nu:
        y = y - v * (S @ y)
    xe = p0 * y
    xo = (r[io] - Hoe @ xe) / c
    ref = xc.copy(); ref[ie] += xe; ref[io] += xo
    Z = host(mg.dev.precondition(0, B))
    err = relerr(Z, ref)
    mg.set_option("smoother_eo", 0)
    try:
        Z2 = host(mg.dev.precondition(0, B))
    finally:
        mg.set_option("smoother_eo", 1)
    print("even-odd cycle vs restatement", err, " vs the degree-36 cycle in A", relerr(Z2, ref))
    assert err < 3e-2                      # BF16-grade (storage of the intermediates, BF16 coarse operand)
    assert not np.array_equal(Z, Z2)       # the even-odd path was taken


# ---- bit-exact test of the even-odd hop kernel --------------------------------------------------
def _hop_case(LX, LT, seed):
    """a stencil level whose links are fourth roots of unity (products with small integers are exact in BF16 / FP32) and
    the matching scipy operator H = A - diag"""
    from deflatedmlmc_schwinger_b200 import _lib, lattice
    rs = np.random.RandomState(seed)
    links = (1j ** rs.randint(4, size=(2, LX, LT))).astype(np.complex128)
    diag = 4.0 - 0.25
    dev = _lib.Hierarchy(2)
    dev.set_stencil(0, links, diag)
    H = (lattice.wilson_matrix(links, -0.25) - diag * __import__("scipy.sparse").sparse.identity(2 * LX * LT)).tocsr()
    return dev, H


def _eo_rows(LX, LT, p):
    """full-lattice row of entry [s][x][th] of the parity-p half-lattice layout (t = 2 th + ((x + p) & 1))"""
    s, x, th = np.meshgrid(np.arange(2), np.arange(LX), np.arange(LT // 2), indexing="ij")
    t = 2 * th + ((x + p) & 1)
    return (s * LX * LT + x * LT + t).ravel()


def _int_cplx(rs, shape, lim=3):
    return (rs.randint(-lim, lim + 1, size=shape) + 1j * rs.randint(-lim, lim + 1, size=shape)).astype(np.complex128)


def _to_bf16_half(v, LX, LT, k):
    """complex [nh][k] -> torch.bfloat16 [2, LX, LT/2, k, 2] on the device"""
    a = np.stack([v.real, v.imag], axis=-1).astype(np.float32).reshape(2, LX, LT // 2, k, 2)
    return torch.from_numpy(a).cuda().to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("LX,LT", [(4, 4), (6, 8), (16, 12)])
@pytest.mark.parametrize("k", [2, 4, 6, 64])
@pytest.mark.parametrize("variant", ["y,y,n", "n,y,n", "y,y,y", "y,n,y"])
@pytest.mark.parametrize("parity", [0, 1])
def test_wilson_hop_eo_kernel_is_bit_exact_on_integer_data(LX, LT, k, variant, parity):
    """wilson_hop_eo_kernel, every template variant the solver launches (k % 4 == 0: four columns per thread, else two), both
    checkerboard colours, lattices small enough that every row is a wrap row in x or t for some thread: small Gaussian
    integers, links in {1, i, -1, -i}, a, b in {0, +-1, +-i} -- every product and sum is exact in BF16 storage / FP32
    arithmetic, so the output must EQUAL the scipy operator's, bit for bit, and an indexing error anywhere shows."""
    dev, H = _hop_case(LX, LT, 100 * LX + LT)
    n, nh = 2 * LX * LT, LX * LT
    has2, hout, zout = [c == "y" for c in variant.split(",")]
    rows_p, rows_q = _eo_rows(LX, LT, parity), _eo_rows(LX, LT, 1 - parity)
    rs = np.random.RandomState(7 * k + parity)
    units = [1, -1, 1j, -1j]
    for trial, (a, b) in enumerate([(units[rs.randint(4)], units[rs.randint(4)]), (0, units[rs.randint(4)]), (1, -1), (-1j, 1j)]):
        if not has2:
            a = 0
        vq, v2 = _int_cplx(rs, (nh, k)), _int_cplx(rs, (nh, k))
        full = np.zeros((n, k), dtype=np.complex128)
        full[rows_q] = vq
        ref = b * (H @ full)[rows_p] + (a * v2 if has2 else 0)
        inq, in2 = _to_bf16_half(vq, LX, LT, k), (_to_bf16_half(v2, LX, LT, k) if has2 else None)
        out = torch.full((2, LX, LT // 2, k, 2), 777.0, dtype=torch.bfloat16, device="cuda") if hout else None
        xc = z = None
        if zout:
            xcn = _int_cplx(rs, (n, k), 5)
            xc = torch.from_numpy(xcn.astype(np.complex64)).cuda().contiguous()
            z = torch.full((n, k), 1234.5 + 0j, dtype=torch.complex128, device="cuda")
        dev.hop_eo(0, parity, inq, in2, out, a, b, k, xc, z)
        torch.cuda.synchronize()
        if hout:
            o = out.float().cpu().numpy().reshape(nh, k, 2)
            got = o[..., 0] + 1j * o[..., 1]
            assert np.array_equal(got, ref), (variant, parity, trial, np.argwhere(got != ref)[:4])
        if zout:
            zz = z.cpu().numpy()
            want = np.full((n, k), 1234.5 + 0j)
            want[rows_p] = xcn[rows_p] + ref
            assert np.array_equal(zz, want), (variant, parity, trial)
    dev.close()


def test_device_prolongator_values_match_the_host_builder(g128):
    """dmlmc_prolongator_values (batched per-aggregate classical Gram-Schmidt, multigrid.py:232-259) against the host builder,
    which is bit-identical to the reference: all three levels of the 128^2 hierarchy.  Same order of operations; only the
    order of the additions inside an inner product differs, and classical Gram-Schmidt amplifies that 1e-16 by the
    conditioning of the 16 x 4 block it works on (measured 1.1e-12 on level 0), hence 1e-10 here; the columns of every
    (aggregate, half) block are orthonormal to the same level in both builders."""
    from deflatedmlmc_schwinger_b200 import _lib, multigrid as mgm
    dev = _lib.Hierarchy(1)
    for a, dofi, c, tv in [(32, 2, 4, g128["tv0"]), (32, 4, 4, g128["tv1"]), (32, 4, 4, g128["tv2"])]:
        ref = mgm.build_prolongator_values(tv, a, dofi, c)
        got = dev.prolongator_values(tv, a, dofi, c).cpu().numpy()
        P, Pr = mgm.prolongator_csr(got, a, dofi, c), mgm.prolongator_csr(ref, a, dofi, c)
        G, Gr = (P.conj().T @ P).toarray(), (Pr.conj().T @ Pr).toarray()
        print("level with dofi", dofi, ": max |device - host|", np.abs(got - ref).max(), " orthonormality device / host",
              np.abs(G - np.eye(G.shape[0])).max(), np.abs(Gr - np.eye(G.shape[0])).max())
        assert np.abs(got - ref).max() < 1e-10
        assert np.abs(G - np.eye(G.shape[0])).max() < 10 * max(np.abs(Gr - np.eye(G.shape[0])).max(), 1e-14)
    dev.close()


@pytest.mark.parametrize("LX,LT", [(2, 8), (4, 16), (6, 24)])
@pytest.mark.parametrize("k", [256, 512])
@pytest.mark.parametrize("has2", [True, False])
@pytest.mark.parametrize("parity", [0, 1])
def test_tma_staged_hop_kernel_is_bit_exact(LX, LT, k, has2, parity):
    """wilson_hop_eo_tma_kernel (the sweep with its neighbour halo staged in shared memory by cp.async.bulk; option hop_tma,
    engaged when LX % 2 == 0, (LT / 2) % 4 == 0 and k % 256 == 0): exact integer data as above, lattices of one tile (every
    halo row is a periodic wrap) up to 3 x 3 tiles -- equal, bit for bit, to the scipy operator AND to the direct kernel (option hop_tma = 0)."""
    dev, H = _hop_case(LX, LT, 31 * LX + LT)
    n, nh = 2 * LX * LT, LX * LT
    rows_p, rows_q = _eo_rows(LX, LT, parity), _eo_rows(LX, LT, 1 - parity)
    rs = np.random.RandomState(11 * k + parity)
    units = [1, -1, 1j, -1j]
    for trial in range(2):
        a, b = (units[rs.randint(4)] if has2 else 0), units[rs.randint(4)]
        vq, v2 = _int_cplx(rs, (nh, k)), _int_cplx(rs, (nh, k))
        full = np.zeros((n, k), dtype=np.complex128)
        full[rows_q] = vq
        ref = b * (H @ full)[rows_p] + (a * v2 if has2 else 0)
        inq, in2 = _to_bf16_half(vq, LX, LT, k), (_to_bf16_half(v2, LX, LT, k) if has2 else None)
        outs = []
        for tma in (1, 0):
            dev.set_option("hop_tma", tma)
            out = torch.full((2, LX, LT // 2, k, 2), 777.0, dtype=torch.bfloat16, device="cuda")
            dev.hop_eo(0, parity, inq, in2, out, a, b, k)
            torch.cuda.synchronize()
            o = out.float().cpu().numpy().reshape(nh, k, 2)
            outs.append(o[..., 0] + 1j * o[..., 1])
        dev.set_option("hop_tma", 0)
        assert np.array_equal(outs[0], ref), (LX, LT, k, has2, parity, trial, np.argwhere(outs[0] != ref)[:4])
        assert np.array_equal(outs[1], ref)
    # random (non-integer) data: the two kernels do the same FP32 operations in the same order
    vq = rs.standard_normal((nh, k)) + 1j * rs.standard_normal((nh, k))
    v2 = rs.standard_normal((nh, k)) + 1j * rs.standard_normal((nh, k))
    inq, in2 = _to_bf16_half(vq, LX, LT, k), (_to_bf16_half(v2, LX, LT, k) if has2 else None)
    outs = []
    for tma in (1, 0):
        dev.set_option("hop_tma", tma)
        out = torch.zeros((2, LX, LT // 2, k, 2), dtype=torch.bfloat16, device="cuda")
        dev.hop_eo(0, parity, inq, in2, out, 0.3 - 0.2j if has2 else 0, -0.7 + 0.1j, k)
        torch.cuda.synchronize()
        outs.append(out.view(torch.int16).cpu().numpy().copy())
    dev.set_option("hop_tma", 0)
    assert np.array_equal(outs[0], outs[1])
    dev.close()
