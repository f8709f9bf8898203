"""The config-3 operators (Y = A_l X for l = 0, 1, 2, R_0, P_0; complex128 and complex64, k = 256 and 512 columns) launched a
few times each with an L2 flush in between, inside a cudaProfilerStart/Stop bracket -- the target of
    ncu --profile-from-start off --set full --clock-control none -k regex:"stencil_kernel|bsr_kernel|restrict_kernel|prolong_add_kernel" ...
so that dram__bytes of each kernel can be compared with its algorithmic bytes (profiles/spmm_sweep.py has the formulas)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np, torch
import bench
from deflatedmlmc_schwinger_b200 import matrix, multigrid, sampling

p, tp = bench.params128()
A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
mg = multigrid.MG(A, smoother_degree=8, dense_coarse_threshold=0, geometric_precond=False)
mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp,
         test_vectors=bench.golden_tvs())
dev = mg.dev
n = mg.level_shapes
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
np.random.seed(123456)
cases = []
for k in (256, 512):
    bits = sampling.draw_probe_bits(k * n[0]).reshape(k, n[0]).T.astype(np.float64) * 2 - 1
    for dt in (torch.complex128, torch.complex64):
        X0 = torch.from_numpy(bits).cuda().to(dt).contiguous()
        Xs = [X0[:n[l]].contiguous() for l in range(3)]
        Ys = [torch.empty_like(x) for x in Xs]
        Xc = dev.restrict(0, X0)
        cases.append((X0, Xs, Ys, Xc))
torch.cuda.synchronize()
torch.cuda.profiler.start()
for X0, Xs, Ys, Xc in cases:
    for rep in range(2):
        for l in range(3):
            flush.fill_(1)
            dev.spmm(l, Xs[l], Ys[l])
        flush.fill_(1)
        dev.restrict(0, X0)
        flush.fill_(1)
        dev.prolong_add(0, Xc, X0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
