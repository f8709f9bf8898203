"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference modules from /root/reference (authoring container) or
from its byte-for-byte copy oracle/_ref/ (oracle/vendor_reference.py; that copy travels to the GPU box).

Used to (a) validate oracle/refport.py, (b) generate tests/golden/* through oracle/make_golden.py and (c) time the
reference itself in `bench.py --impl reference`.  Nothing in the product imports this file.

What has to be patched for the reference to run here:
  1. `pyamg` and `matplotlib` are absent            -> oracle/shims on sys.path
  2. scipy >= 1.14 dropped lgmres(tol=)             -> translate to rtol
     (multigrid.py:393,438 call lgmres(lop, r, tol=1e-20, maxiter=smooth_iters))
  3. matrix.py:21,25 use the bare file name          -> chdir(/root/reference)
  4. utils.py:161-164 raises without OMP_NUM_THREADS
  5. eigs/eigsh start vectors are unseeded in scipy  -> optional deterministic v0 and
     optional INJECTION of previously recorded eigenvectors (hierarchies are only
     comparable with identical test vectors)
"""
import contextlib
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
# the read-only reference tree of the authoring container, else the byte-for-byte copy made by oracle/vendor_reference.py
# (oracle/_ref/, git-ignored, travels to the GPU box)
REF_DIR = "/root/reference" if os.path.isfile("/root/reference/multigrid.py") else os.path.join(_HERE, "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "multigrid.py"))


class EigRecorder:
    """Wraps scipy eigs/eigsh: seeds v0, records every call's output, or replays
    recorded outputs (injection)."""

    def __init__(self, seed=20261018, replay=None):
        self.rng = np.random.RandomState(seed)
        self.calls = []            # list of (kind, eigvals, eigvecs)
        self.replay = list(replay) if replay is not None else None

    def wrap(self, fn, kind):
        def wrapped(A, *args, **kwargs):
            if self.replay is not None:
                k_, w, v = self.replay.pop(0)
                assert k_ == kind, (k_, kind)
                self.calls.append((kind, w, v))
                return w.copy(), v.copy()
            n = A.shape[0]
            if "v0" not in kwargs or kwargs["v0"] is None:
                kwargs["v0"] = self.rng.standard_normal(n) + 0.0
            w, v = fn(A, *args, **kwargs)
            self.calls.append((kind, np.array(w), np.array(v)))
            return w, v
        return wrapped


_loaded = {}


def load_reference(eig_recorder=None):
    """Import (once) and return the reference modules as a dict; (re)install the
    eigen-solver wrappers of `eig_recorder` when given."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_DIR)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    sys.dont_write_bytecode = True
    if not _loaded:
        import scipy.sparse.linalg as spla
        _orig_lgmres = spla.lgmres

        def lgmres_compat(A, b, x0=None, tol=None, **kw):
            if tol is not None and "rtol" not in kw:
                kw["rtol"] = tol
            return _orig_lgmres(A, b, x0=x0, **kw)

        saved_path = list(sys.path)
        saved_mods = {k: sys.modules.get(k) for k in
                      ("utils", "matrix", "multigrid", "stoch_trace", "examples", "gateway")}
        sys.path.insert(0, REF_DIR)
        sys.path.insert(0, _SHIMS)
        spla.lgmres = lgmres_compat       # bound by `from ... import lgmres` at multigrid.py:2
        try:
            for name in ("utils", "matrix", "multigrid", "stoch_trace", "examples", "gateway"):
                sys.modules.pop(name, None)
                _loaded[name] = importlib.import_module(name)
        finally:
            spla.lgmres = _orig_lgmres
            sys.path[:] = saved_path
            for k, v in saved_mods.items():
                if v is not None:
                    sys.modules[k] = v
                else:
                    sys.modules.pop(k, None)
        import scipy.sparse.linalg as _spla
        _loaded["_orig_eigs"] = _spla.eigs
        _loaded["_orig_eigsh"] = _spla.eigsh
    if eig_recorder is not None:
        _loaded["multigrid"].eigs = eig_recorder.wrap(_loaded["_orig_eigs"], "eigs")
        _loaded["multigrid"].eigsh = eig_recorder.wrap(_loaded["_orig_eigsh"], "eigsh")
        _loaded["utils"].eigsh = eig_recorder.wrap(_loaded["_orig_eigsh"], "eigsh")
    return _loaded


@contextlib.contextmanager
def in_reference_dir():
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        yield
    finally:
        os.chdir(cwd)


def params_128():
    """gateway.set_params('schwinger128') + G202's function_tol (gateway.py:52-59)."""
    ref = load_reference()
    p = ref["gateway"].set_params("schwinger128")
    p["function_tol"] = 1e-12
    return p


def params_16(permuted=False, nr_deflat_vctrs=0, mlmc_deflat_vctrs=(0, 0)):
    """The shipped 16^2 set is broken: missing keys and
    dof=[2,2,2] gives P=0.  Overlay the missing keys and dof=[2,4,4]."""
    ref = load_reference()
    p = ref["gateway"].set_params("schwinger16")
    p["function_tol"] = 1e-12
    p["dof"] = [2, 4, 4]
    p["aggrs"] = [4, 4]
    p["nr_deflat_vctrs"] = nr_deflat_vctrs
    p["mlmc_deflat_vctrs"] = list(mlmc_deflat_vctrs)
    p["mlmc_levels_to_skip"] = []
    p["defl_eigvs_tol_Hutch"] = 1e-9
    p["defl_eigvs_tol_MLMC"] = 1e-1
    p["diff_lev_op_tol"] = 1e-3
    p["defl_type"] = "exact"
    p["use_permuted"] = permuted
    p["latt_dims"] = [16, 16]
    p["x_displacement"] = 2
    p["check_quality_MG"] = False
    p["test_vectors_type"] = "EVs"
    p["accuracy_mg_eigvs"] = "high"
    return p
