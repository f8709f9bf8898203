"""ctypes binding of libdmlmc_sm100.so (C ABI: include/dmlmc.h).

PyTorch is used for device memory, streams and (in stoch_trace) torch.distributed only;
all arithmetic of the hot path happens in the hand-written sm_100a kernels behind the C ABI.
There is NO fallback: if the shared library is missing or no CUDA device is present the
calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmlmc_sm100.so")

C128, C64 = 0, 1

_lib = None

_SIGS = {
    "dmlmc_abi_version": (ctypes.c_int, []),
    "dmlmc_last_error": (ctypes.c_char_p, []),
    "dmlmc_hier_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "dmlmc_hier_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dmlmc_set_stencil": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_double, ctypes.c_double]),
    "dmlmc_set_bsr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "dmlmc_set_transfer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_set_transfer_indexed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "dmlmc_set_preconditioner": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_set_coarsest_inverse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_set_dense_inverse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_set_dense_inverse_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_prolongator_values": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_block_orthonormal_values": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_galerkin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_int)]),
    "dmlmc_set_bsr_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "dmlmc_dense_inverse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_set_dense_inverse_device_full": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_set_smoother": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_double, ctypes.c_double]),
    "dmlmc_set_smoother_eo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_double, ctypes.c_double]),
    "dmlmc_set_smoother_storage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "dmlmc_set_perm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "dmlmc_set_deflation": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_spmm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_restrict": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_prolong_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_coarsest_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_smooth": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_vcycle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_precondition": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_dotc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_deflate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_probe_expand": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_mt19937_bits": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                          ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p]),
    "dmlmc_set_mt_jump_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_probe_expand_bytes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "dmlmc_rng_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "dmlmc_apply_perm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dmlmc_set_workspace": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "dmlmc_set_inner_precision": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dmlmc_fgmres": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dmlmc_level_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "dmlmc_level_sample_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "dmlmc_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]),
    "dmlmc_hop_eo": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "dmlmc_unconverged_columns": (ctypes.c_int, [ctypes.c_void_p]),
    "dmlmc_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "dmlmc_vcycle_chunk_cols": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def load():
    """Load libdmlmc_sm100.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libdmlmc_sm100.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "this package has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.dmlmc_abi_version() != 6:
        raise RuntimeError("libdmlmc_sm100.so ABI version mismatch")
    _lib = lib
    return lib


class DmlmcError(Exception):
    pass


def _check(rc):
    if rc != 0:
        msg = load().dmlmc_last_error()
        raise DmlmcError("libdmlmc_sm100 error %d: %s" % (rc, msg.decode() if msg else "?"))


def _host_c128(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _host_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


class Hierarchy:
    """Device-resident multigrid hierarchy + batched solver (one per GPU / stream)."""

    def __init__(self, n_levels, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("deflatedmlmc_schwinger_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        self.n_levels = n_levels
        h = ctypes.c_void_p()
        _check(self.lib.dmlmc_hier_create(self.device.index, ctypes.c_void_p(self.stream.cuda_stream), n_levels,
                                          ctypes.byref(h)))
        self.h = h
        self.sizes = [0] * n_levels
        self._ws = None
        self._ws_key = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.dmlmc_hier_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ---------------------------------------------------------------------
    def set_stencil(self, level, links, diag):
        links, p = _host_c128(links)
        _, LX, LT = links.shape
        _check(self.lib.dmlmc_set_stencil(self.h, level, LX, LT, p, float(np.real(diag)), float(np.imag(diag))))
        self.sizes[level] = 2 * LX * LT

    def set_bsr(self, level, n, bs, colidx, vals):
        colidx, pc = _host_i32(colidx)
        vals, pv = _host_c128(vals)
        _check(self.lib.dmlmc_set_bsr(self.h, level, n, bs, colidx.shape[1], pc, pv))
        self.sizes[level] = n

    def set_transfer(self, level, n_f, aggr_size, dofi, nvec, pvals):
        pvals, p = _host_c128(pvals)
        assert pvals.shape == (n_f, nvec)
        _check(self.lib.dmlmc_set_transfer(self.h, level, n_f, aggr_size, dofi, nvec, p))

    def set_transfer_indexed(self, level, n_f, nvec, pvals, cblk):
        """aggregation prolongator on arbitrary equal-sized aggregates: fine row r -> coarse block cblk[r]"""
        pvals, p = _host_c128(pvals)
        cblk, pc = _host_i32(cblk)
        assert pvals.shape == (n_f, nvec) and cblk.shape == (n_f,)
        _check(self.lib.dmlmc_set_transfer_indexed(self.h, level, n_f, int(cblk.max()) + 1, nvec, p, pc))

    def set_preconditioner(self, level, other, other_level=0):
        """FGMRES solves of `level` are preconditioned by the V-cycle of hierarchy `other` from its `other_level`"""
        _check(self.lib.dmlmc_set_preconditioner(self.h, level, None if other is None else other.h, other_level))
        self._prec_keepalive = getattr(self, "_prec_keepalive", {})
        self._prec_keepalive[level] = other
        self._ws_key = None          # the work space may have to grow

    def set_coarsest_inverse(self, minv):
        minv, p = _host_c128(minv)
        _check(self.lib.dmlmc_set_coarsest_inverse(self.h, minv.shape[0], p))
        self.sizes[self.n_levels - 1] = minv.shape[0]

    def set_dense_inverse(self, level, minv):
        minv, p = _host_c128(minv)
        _check(self.lib.dmlmc_set_dense_inverse(self.h, level, minv.shape[0], p))

    def prolongator_values(self, eig_vecs, aggr_size, dofi, nvec):
        """multigrid.py:232-259 on the device: eig_vecs numpy or torch complex128 [n, >= nvec] -> pvals torch [n, nvec]"""
        torch = self.torch
        ev = eig_vecs if torch.is_tensor(eig_vecs) else torch.from_numpy(np.ascontiguousarray(eig_vecs, dtype=np.complex128))
        ev = ev.to(self.device).contiguous()
        pv = torch.empty((ev.shape[0], nvec), dtype=torch.complex128, device=self.device)
        _check(self.lib.dmlmc_prolongator_values(self.h, ctypes.c_void_p(ev.data_ptr()), int(ev.shape[1]), int(ev.shape[0]),
                                                 int(aggr_size), int(dofi), int(nvec), ctypes.c_void_p(pv.data_ptr())))
        return pv

    def block_orthonormal_values(self, vecs, cblk, nvec, passes=2):
        """the first nvec columns of `vecs` orthonormalised within every coarse block cblk[r] (equal-sized blocks) on the
        device -> (pvals torch [n, nvec], rows torch int32 [n_blocks, m])"""
        torch = self.torch
        ev = vecs if torch.is_tensor(vecs) else torch.from_numpy(np.ascontiguousarray(vecs, dtype=np.complex128))
        ev = ev.to(self.device).contiguous()
        cb = cblk if torch.is_tensor(cblk) else torch.from_numpy(np.ascontiguousarray(cblk, dtype=np.int32))
        cb = cb.to(self.device)
        n = int(ev.shape[0])
        nb = int(cb.max().item()) + 1
        m = n // nb
        rows = torch.argsort(cb.to(torch.int64), stable=True).to(torch.int32).contiguous()
        if m * nb != n or not bool((cb[rows.long()].reshape(nb, m) == torch.arange(nb, device=self.device, dtype=cb.dtype)[:, None]).all()):
            raise DmlmcError("block_orthonormal_values: the coarse blocks must all have the same number of rows")
        pv = torch.empty((n, nvec), dtype=torch.complex128, device=self.device)
        _check(self.lib.dmlmc_block_orthonormal_values(self.h, ctypes.c_void_p(ev.data_ptr()), int(ev.shape[1]), n, m, int(nvec),
                                                       ctypes.c_void_p(rows.data_ptr()), int(passes), ctypes.c_void_p(pv.data_ptr())))
        return pv, rows.reshape(nb, m)

    def galerkin(self, level, nvec, n_coarse, cap=24):
        """A_{level+1} = R A P (multigrid.py:276) on the device from this level's operator and transfer; the coarse operator is
        installed on level + 1 (padded block-sparse rows) and returned as (col torch int32 [nb, bpr], vals torch complex128
        [nb, bpr, nvec, nvec])."""
        torch = self.torch
        nb = n_coarse // nvec
        while True:
            col = torch.empty((nb, cap), dtype=torch.int32, device=self.device)
            vals = torch.empty((nb, cap, nvec, nvec), dtype=torch.complex128, device=self.device)
            slots = ctypes.c_int(0)
            _check(self.lib.dmlmc_galerkin(self.h, level, cap, ctypes.c_void_p(col.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                                           ctypes.byref(slots)))
            if slots.value <= cap:
                break
            del col, vals
            cap *= 2
        bpr = max(int(slots.value), 1)
        col = col[:, :bpr].contiguous()
        vals = vals[:, :bpr].contiguous()
        self.set_bsr_device(level + 1, n_coarse, nvec, col, vals)
        return col, vals

    def set_bsr_device(self, level, n, bs, col, vals):
        assert col.is_cuda and col.is_contiguous() and col.dtype == self.torch.int32
        assert vals.is_cuda and vals.is_contiguous() and vals.dtype == self.torch.complex128
        _check(self.lib.dmlmc_set_bsr_device(self.h, level, int(n), int(bs), int(col.shape[1]), ctypes.c_void_p(col.data_ptr()),
                                             ctypes.c_void_p(vals.data_ptr())))
        self.sizes[level] = int(n)

    def dense_inverse(self, M):
        """multigrid.py:342-344 on the device: the inverse of a dense complex128 matrix (numpy or torch, n <= 8192) -> torch [n, n]"""
        torch = self.torch
        A = M if torch.is_tensor(M) else torch.from_numpy(np.ascontiguousarray(M, dtype=np.complex128))
        A = A.to(self.device).clone().contiguous()
        _check(self.lib.dmlmc_dense_inverse(self.h, int(A.shape[0]), ctypes.c_void_p(A.data_ptr())))
        return A

    def set_dense_inverse_device(self, level, minv_dev, full=False):
        """minv_dev: torch complex128 CUDA tensor [n, n] (row-major inverse); full: keep it in all precisions (else only as
        the BF16 tensor-core operand of the complex64 V-cycle)"""
        assert minv_dev.is_cuda and minv_dev.is_contiguous() and minv_dev.dtype == self.torch.complex128
        fn = self.lib.dmlmc_set_dense_inverse_device_full if full else self.lib.dmlmc_set_dense_inverse_device
        _check(fn(self.h, level, minv_dev.shape[0], ctypes.c_void_p(minv_dev.data_ptr())))
        self.sizes[level] = minv_dev.shape[0]
        self.torch.cuda.current_stream(self.device).synchronize()

    def set_smoother(self, level, nu, p0, storage16=True):
        """p(A) = p0 * prod_i (I - nu[i] A); storage16: the complex64 cycle may keep the intermediates in BF16"""
        nu, p = _host_c128(np.asarray(nu).reshape(-1))
        _check(self.lib.dmlmc_set_smoother(self.h, level, nu.shape[0], p if nu.shape[0] else None,
                                           float(np.real(p0)), float(np.imag(p0))))
        _check(self.lib.dmlmc_set_smoother_storage(self.h, level, 1 if storage16 else 0))

    def set_smoother_eo(self, level, nu, p0):
        """polynomial p(S) = p0 * prod_i (I - nu[i] S) in the even-odd Schur complement S of a stencil level"""
        nu, p = _host_c128(np.asarray(nu).reshape(-1))
        _check(self.lib.dmlmc_set_smoother_eo(self.h, level, nu.shape[0], p if nu.shape[0] else None,
                                              float(np.real(p0)), float(np.imag(p0))))

    def set_perm(self, level, shift, cols=None, vals=None):
        if cols is None:
            _check(self.lib.dmlmc_set_perm(self.h, level, int(shift), 0, None, None))
        else:
            cols, pc = _host_i32(cols)
            vals, pv = _host_c128(vals)
            _check(self.lib.dmlmc_set_perm(self.h, level, int(shift), cols.shape[1], pc, pv))

    def set_deflation(self, level, V):
        if V is None:
            _check(self.lib.dmlmc_set_deflation(self.h, level, 0, None))
        else:
            V, p = _host_c128(V)
            _check(self.lib.dmlmc_set_deflation(self.h, level, V.shape[1], p))

    def set_option(self, name, value):
        _check(self.lib.dmlmc_set_option(self.h, name.encode(), float(value)))
        self._ws_key = None          # the work-space size depends on options

    def set_inner_precision(self, prec):
        _check(self.lib.dmlmc_set_inner_precision(self.h, prec))
        self._ws_key = None

    # ---- helpers -------------------------------------------------------------------
    def _dtype(self, prec):
        return self.torch.complex128 if prec == C128 else self.torch.complex64

    def _prec(self, t):
        if t.dtype == self.torch.complex128:
            return C128
        if t.dtype == self.torch.complex64:
            return C64
        raise TypeError("expected a complex64/complex128 CUDA tensor")

    def _chk(self, t, n=None):
        assert t.is_cuda and t.is_contiguous() and t.dim() == 2, "expected a contiguous CUDA tensor [n, k]"
        if n is not None:
            assert t.shape[0] == n, (t.shape, n)
        return ctypes.c_void_p(t.data_ptr())

    def empty(self, n, k, prec=C128):
        return self.torch.empty((n, k), dtype=self._dtype(prec), device=self.device)

    def ensure_workspace(self, level, k, restart):
        """The library keeps the pointer it is given: it is re-registered whenever the tensor behind it changes (the
        size query depends on options as well as on (level, k, restart)), never only when the key changes."""
        key = (level, k, restart)
        need = self.lib.dmlmc_workspace_bytes(self.h, level, k, restart)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws_reg = None
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        reg = (self._ws.data_ptr(), self._ws.numel())
        if self._ws_key != key or getattr(self, "_ws_reg", None) != reg:
            _check(self.lib.dmlmc_set_workspace(self.h, ctypes.c_void_p(reg[0]), reg[1]))
            self._ws_key = key
            self._ws_reg = reg

    def release_workspace(self):
        """drop the work space (the library forgets the pointer before the tensor is freed)"""
        _check(self.lib.dmlmc_set_workspace(self.h, None, 0))
        self._ws = None
        self._ws_key = None
        self._ws_reg = None

    # ---- operators -------------------------------------------------------------------
    def spmm(self, level, X, Y=None):
        Y = self.torch.empty_like(X) if Y is None else Y
        _check(self.lib.dmlmc_spmm(self.h, level, self._prec(X), self._chk(X, self.sizes[level]), self._chk(Y), X.shape[1]))
        return Y

    def restrict(self, level, Xf):
        Xc = self.empty(self.sizes[level + 1], Xf.shape[1], self._prec(Xf))
        _check(self.lib.dmlmc_restrict(self.h, level, self._prec(Xf), self._chk(Xf, self.sizes[level]), self._chk(Xc), Xf.shape[1]))
        return Xc

    def prolong_add(self, level, Xc, Xf):
        _check(self.lib.dmlmc_prolong_add(self.h, level, self._prec(Xc), self._chk(Xc, self.sizes[level + 1]),
                                          self._chk(Xf, self.sizes[level]), Xc.shape[1]))
        return Xf

    def coarsest_apply(self, B):
        X = self.torch.empty_like(B)
        _check(self.lib.dmlmc_coarsest_apply(self.h, self._prec(B), self._chk(B, self.sizes[-1]), self._chk(X), B.shape[1]))
        return X

    def smooth(self, level, R):
        self.ensure_workspace(level, R.shape[1], 1)
        E = self.torch.empty_like(R)
        _check(self.lib.dmlmc_smooth(self.h, level, self._prec(R), self._chk(R, self.sizes[level]), self._chk(E), R.shape[1]))
        return E

    def vcycle(self, level, B):
        self.ensure_workspace(level, B.shape[1], 1)
        X = self.torch.empty_like(B)
        _check(self.lib.dmlmc_vcycle(self.h, level, self._prec(B), self._chk(B, self.sizes[level]), self._chk(X), B.shape[1]))
        return X

    def precondition(self, level, V):
        """Z = M^{-1} V (complex128): the preconditioner of the level's FGMRES, as the solver applies it"""
        assert V.dtype == self.torch.complex128
        self.ensure_workspace(level, V.shape[1], 1)
        Z = self.torch.empty_like(V)
        _check(self.lib.dmlmc_precondition(self.h, level, self._chk(V, self.sizes[level]), self._chk(Z), V.shape[1]))
        return Z

    def dotc(self, X, Y):
        self.ensure_workspace(0, X.shape[1], 1)
        out = self.torch.empty(X.shape[1], dtype=self.torch.complex128, device=self.device)
        _check(self.lib.dmlmc_dotc(self.h, self._chk(X), self._chk(Y, X.shape[0]), X.shape[0], X.shape[1],
                                   ctypes.c_void_p(out.data_ptr())))
        return out

    def deflate(self, level, X):
        self.ensure_workspace(level, X.shape[1], 1)
        _check(self.lib.dmlmc_deflate(self.h, level, self._chk(X, self.sizes[level]), X.shape[1]))
        return X

    def probe_expand(self, bits_dev, n, k):
        X0 = self.empty(n, k, C128)
        _check(self.lib.dmlmc_probe_expand(self.h, ctypes.c_void_p(bits_dev.data_ptr()), n, k, self._chk(X0)))
        return X0

    def mt19937_bits(self, state, skip_before, count, skip_after, out=None, backup=None):
        """state: int32 CUDA tensor [625] (key + position, bit pattern of uint32); out: uint8 CUDA tensor [count]"""
        if out is None and count > 0:
            out = self.torch.empty(int(count), dtype=self.torch.uint8, device=self.device)
        if not getattr(self, "_mt_table_set", False):
            from . import mtjump
            tab = np.ascontiguousarray(mtjump.table())
            _check(self.lib.dmlmc_set_mt_jump_table(self.h, tab.ctypes.data_as(ctypes.c_void_p), int(tab.shape[0])))
            self._mt_table_set = True
        _check(self.lib.dmlmc_mt19937_bits(self.h, ctypes.c_void_p(state.data_ptr()),
                                           ctypes.c_void_p(backup.data_ptr()) if backup is not None else None,
                                           int(skip_before), int(count), int(skip_after),
                                           ctypes.c_void_p(out.data_ptr()) if out is not None else None))
        return out

    def probe_expand_bytes(self, lsb, n, k):
        X0 = self.empty(n, k, C128)
        _check(self.lib.dmlmc_probe_expand_bytes(self.h, ctypes.c_void_p(lsb.data_ptr()), n, k, self._chk(X0)))
        return X0

    def rng_sync(self):
        _check(self.lib.dmlmc_rng_sync(self.h))

    def hop_eo(self, level, parity, in_q, in2, out_p, a, b, k, xc=None, z=None):
        """one half-lattice sweep of the even-odd smoother (dmlmc_hop_eo); BF16 arrays as torch.bfloat16 tensors
        [2, LX, LT/2, k, 2]"""
        ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        a, b = complex(a), complex(b)
        _check(self.lib.dmlmc_hop_eo(self.h, level, int(parity), ptr(in_q), ptr(in2), ptr(out_p), a.real, a.imag, b.real, b.imag,
                                     int(k), ptr(xc), ptr(z)))

    def unconverged_columns(self):
        return int(self.lib.dmlmc_unconverged_columns(self.h))

    def apply_perm(self, level, X):
        Y = self.torch.empty_like(X)
        _check(self.lib.dmlmc_apply_perm(self.h, level, self._chk(X, self.sizes[level]), self._chk(Y), X.shape[1]))
        return Y

    # ---- solver ----------------------------------------------------------------------
    def fgmres(self, level, B, tol, restart=40, maxiter=1000):
        k = B.shape[1]
        self.ensure_workspace(level, k, restart)
        X = self.torch.empty_like(B)
        iters = np.zeros(k, dtype=np.int32)
        relres = np.zeros(k, dtype=np.float64)
        _check(self.lib.dmlmc_fgmres(self.h, level, self._chk(B, self.sizes[level]), self._chk(X), k, float(tol),
                                     int(restart), int(maxiter), iters.ctypes.data_as(ctypes.c_void_p),
                                     relres.ctypes.data_as(ctypes.c_void_p)))
        return X, iters, relres

    def level_sample(self, method, level_f, level_c, X0, tol, restart=40, maxiter=1000):
        k = X0.shape[1]
        self.ensure_workspace(level_f, k, restart)
        e = self.torch.empty(k, dtype=self.torch.complex128, device=self.device)
        iters = np.zeros(2 * k, dtype=np.int32)
        _check(self.lib.dmlmc_level_sample(self.h, method, level_f, level_c, self._chk(X0, self.sizes[level_f]), k,
                                           float(tol), int(restart), int(maxiter), ctypes.c_void_p(e.data_ptr()),
                                           iters.ctypes.data_as(ctypes.c_void_p)))
        self._check_converged(tol, maxiter)
        return e, iters.reshape(2, k)

    def level_sample_host(self, method, level_f, level_c, bits_host, k, tol, restart=40, maxiter=1000):
        """bits_host: packed uint8 (numpy, bitorder='little') of k probes of length n_f; returns (e[k] numpy, iters[2,k])."""
        self.ensure_workspace(level_f, k, restart)
        bits_host = np.ascontiguousarray(bits_host, dtype=np.uint8)
        assert bits_host.size * 8 >= self.sizes[level_f] * k
        e = np.zeros(k, dtype=np.complex128)
        iters = np.zeros(2 * k, dtype=np.int32)
        _check(self.lib.dmlmc_level_sample_host(self.h, method, level_f, level_c, bits_host.ctypes.data_as(ctypes.c_void_p),
                                                k, float(tol), int(restart), int(maxiter),
                                                e.ctypes.data_as(ctypes.c_void_p), iters.ctypes.data_as(ctypes.c_void_p)))
        self._check_converged(tol, maxiter)
        return e, iters.reshape(2, k)

    def _check_converged(self, tol, maxiter):
        bad = self.unconverged_columns()
        if bad:
            raise DmlmcError("%d probe column(s) did not reach the relative residual %g within %d iterations" % (bad, tol, maxiter))

    def launch_count(self):
        return int(self.lib.dmlmc_launch_count(self.h))

    def vcycle_chunk_cols(self, level, prec, k):
        return int(self.lib.dmlmc_vcycle_chunk_cols(self.h, level, prec, k))
