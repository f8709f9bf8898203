"""Dev tool (GPU): CUDA-event timings of the individual kernels behind the C ABI on the 128^2
hierarchy, k probes (default 256).  Prints one JSON object per kernel: us, algorithmic GB/s."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np, torch
import bench
from deflatedmlmc_schwinger_b200 import matrix, multigrid

k = int(os.environ.get("K", "256"))
deg = int(os.environ.get("DEG", "32"))
p, tp = bench.params128()
A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
mg = multigrid.MG(A, smoother_degree=deg)
mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp, test_vectors=bench.golden_tvs())
dev = mg.dev
for kv in os.environ.get("OPTS", "").split(","):
    if kv:
        dev.set_option(kv.split("=")[0], float(kv.split("=")[1]))
stream = torch.cuda.current_stream()

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps   # us

def rnd(n, dt):
    return torch.randn(n, k, device="cuda", dtype=torch.float64).to(dt).contiguous()

out = []
for lvl in range(3):
    n = mg.level_shapes[lvl]
    for name, dt, s in (("c64", torch.complex64, 8), ("c128", torch.complex128, 16)):
        X = rnd(n, dt); Y = torch.empty_like(X)
        us = timeit(lambda: dev.spmm(lvl, X, Y))
        out.append({"kernel": "spmm", "level": lvl, "prec": name, "us": us, "GBps": 2 * n * k * s / us / 1e3})
        us = timeit(lambda: dev.smooth(lvl, X), reps=5, warm=2)
        d = mg.level_degree(lvl)
        by = (4 + 2 * (d - 1)) * n * k * s        # copy in, d-1 factor kernels (read x, write x'), copy out
        out.append({"kernel": "smooth(deg %d)" % d, "level": lvl, "prec": name, "us": us, "us_per_step": us / max(d - 1, 1), "GBps": by / us / 1e3})
        Xc = dev.restrict(lvl, X)
        us = timeit(lambda: dev.restrict(lvl, X))
        out.append({"kernel": "restrict", "level": lvl, "prec": name, "us": us, "GBps": 1.25 * n * k * s / us / 1e3})
        us = timeit(lambda: dev.prolong_add(lvl, Xc, X))
        out.append({"kernel": "prolong_add", "level": lvl, "prec": name, "us": us, "GBps": 2.25 * n * k * s / us / 1e3})
        us = timeit(lambda: dev.vcycle(lvl, X), reps=3, warm=1)
        out.append({"kernel": "vcycle", "level": lvl, "prec": name, "us": us})
for name, dt, s in (("c64", torch.complex64, 8), ("c128", torch.complex128, 16)):
    B = rnd(512, dt)
    us = timeit(lambda: dev.coarsest_apply(B))
    out.append({"kernel": "coarsest_apply", "prec": name, "us": us, "TFLOPs": 8 * 512 * 512 * k / us / 1e6})
X = rnd(32768, torch.complex128); Y = rnd(32768, torch.complex128)
us = timeit(lambda: dev.dotc(X, Y))
out.append({"kernel": "dotc", "us": us, "GBps": 2 * 32768 * k * 16 / us / 1e3})
# deflation projections x - V (V^H x): FP64 tensor cores (DMMA) vs SIMT, bytes 2 n s (d + k) + n s k, flops 16 n d k
for d in (8, 16, 64):
    rs = np.random.RandomState(d)
    V, _ = np.linalg.qr(rs.standard_normal((32768, d)) + 1j * rs.standard_normal((32768, d)))
    dev.set_deflation(0, V)
    Xd = rnd(32768, torch.complex128)
    for tens in (1, 0):
        dev.set_option("defl_tensor", tens)
        us = timeit(lambda: dev.deflate(0, Xd), reps=10)
        out.append({"kernel": "deflate (dot + axpy)", "d": d, "tensor_cores": bool(tens), "us": us,
                    "GBps": (2 * 32768 * 16 * (d + k) + 32768 * 16 * k) / us / 1e3, "TFLOPs": 16 * 32768 * d * k / us / 1e6})
    dev.set_option("defl_tensor", 1)
    dev.set_deflation(0, None)
for o in out:
    print(json.dumps(o))
