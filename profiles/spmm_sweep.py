"""BASELINE config 3: schwinger128, X in {+-1}^{n x k} from the probe stream, k = 1..256;
CUDA-event timings of Y = A_l X (l = 0, 1, 2), R_0, P_0 and achieved algorithmic GB/s against the measured HBM
peak (MEASURED_PEAKS.json).  Bytes per call: level 0 n0*s*(1+2k) (link form), level 1
n1*s*(36+2k)+9*n1, level 2 n2*s*(48+2k)+3*n2, restrict n0*s*(4+1.25k), prolong n0*s*(4+2.25k).
Between timed calls a 256 MB buffer is overwritten (L2 flush), because below k ~ 64 the operands fit in L2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np, torch
import bench
from deflatedmlmc_schwinger_b200 import matrix, multigrid, sampling, utils

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
p, tp = bench.params128()
A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
mg = multigrid.MG(A, smoother_degree=8, dense_coarse_threshold=0)
mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp, test_vectors=bench.golden_tvs())
dev = mg.dev
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=10):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

np.random.seed(123456)
n = mg.level_shapes
for k in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    bits = sampling.draw_probe_bits(k * n[0]).reshape(k, n[0]).T.astype(np.float64) * 2 - 1
    for name, dt, s in (("c128", torch.complex128, 16), ("c64", torch.complex64, 8)):
        X0 = torch.from_numpy(bits).cuda().to(dt).contiguous()
        row = {"k": k, "prec": name}
        for lvl, by in ((0, n[0] * s * (1 + 2 * k)), (1, n[1] * s * (36 + 2 * k) + 9 * n[1]), (2, n[2] * s * (48 + 2 * k) + 3 * n[2])):
            X = X0[:n[lvl]].contiguous(); Y = torch.empty_like(X)
            us = timeit(lambda: dev.spmm(lvl, X, Y))
            row["A%d_us" % lvl] = round(us, 2); row["A%d_GBps" % lvl] = round(by / us / 1e3, 1); row["A%d_frac" % lvl] = round(by / us / 1e3 / peak, 3)
        Xc = dev.restrict(0, X0)
        us = timeit(lambda: dev.restrict(0, X0)); by = n[0] * s * (4 + 1.25 * k)
        row["R0_us"] = round(us, 2); row["R0_GBps"] = round(by / us / 1e3, 1)
        us = timeit(lambda: dev.prolong_add(0, Xc, X0)); by = n[0] * s * (4 + 2.25 * k)
        row["P0_us"] = round(us, 2); row["P0_GBps"] = round(by / us / 1e3, 1)
        print(json.dumps(row), flush=True)
