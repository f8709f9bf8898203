import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def g16():
    return np.load(os.path.join(GOLDEN, "schwinger16.npz"))


@pytest.fixture(scope="session")
def g16defl():
    """deflated MLMC on 16^2 from the unmodified reference (oracle.make_golden 16defl)"""
    return np.load(os.path.join(GOLDEN, "schwinger16_defl_mlmc.npz"))


@pytest.fixture(scope="session")
def g128():
    return np.load(os.path.join(GOLDEN, "schwinger128.npz"))


def params16(permuted=False, nd=0, mlmc_nd=(0, 0)):
    from deflatedmlmc_schwinger_b200 import gateway, utils
    p = gateway.set_params("schwinger16")
    p["function_tol"] = 1e-12
    p["accuracy_mg_eigvs"] = "high"
    p["use_permuted"] = permuted
    p["nr_deflat_vctrs"] = nd
    p["mlmc_deflat_vctrs"] = list(mlmc_nd)
    p["verbose"] = False
    return p


def params128():
    from deflatedmlmc_schwinger_b200 import gateway
    p = gateway.set_params("schwinger128")
    p["function_tol"] = 1e-12
    p["verbose"] = False
    return p


@pytest.fixture(scope="session")
def port16(g16):
    """oracle hierarchy, 16^2 plain, golden test vectors"""
    from oracle import refport
    from deflatedmlmc_schwinger_b200 import utils
    p = params16()
    tp = utils.trace_params_from_params(p, "mlmc")
    mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
    mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp, test_vectors=[g16["tv0"], g16["tv1"]])
    return mp, tp


@pytest.fixture(scope="session")
def port128(g128):
    from oracle import refport
    from deflatedmlmc_schwinger_b200 import utils
    p = params128()
    tp = utils.trace_params_from_params(p, "mlmc")
    mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
    mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp,
             test_vectors=[g128["tv0"], g128["tv1"], g128["tv2"]])
    mp.skip_level = True
    return mp, tp


def make_mg(p, method, tvs, **kw):
    from deflatedmlmc_schwinger_b200 import matrix, multigrid, utils
    tp = utils.trace_params_from_params(p, method)
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    mg = multigrid.MG(A, **kw)
    mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"],
             params=tp, test_vectors=tvs)
    mg.total_levels = len(mg.ml.levels)
    return mg, tp, A


@pytest.fixture(scope="session")
def mg16(g16):
    mg, tp, A = make_mg(params16(), "mlmc", [g16["tv0"], g16["tv1"]], smoother_degree=8)
    return mg, tp, A


@pytest.fixture(scope="session")
def mg16perm(g16):
    mg, tp, A = make_mg(params16(permuted=True), "mlmc", [g16["tv0"], g16["tv1"]], smoother_degree=8)
    return mg, tp, A


@pytest.fixture(scope="session")
def mg128(g128):
    mg, tp, A = make_mg(params128(), "mlmc", [g128["tv0"], g128["tv1"], g128["tv2"]], smoother_degree=32)
    mg.skip_level = True
    return mg, tp, A
