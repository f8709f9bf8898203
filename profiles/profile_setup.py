"""Where the hierarchy set-up time goes (cProfile, cumulative).
    python profiles/profile_setup.py            # the bench's solver (128^2, golden test vectors injected)
    python profiles/profile_setup.py --L 512    # synthetic random-U(1) lattice, NOTHING injected: device eigensolver for the
                                                # test vectors of every level, then the hierarchy (profiles/run_synthetic.py's)"""
import argparse, cProfile, io, json, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "4")
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=0)
ap.add_argument("--mass", type=float, default=-0.062)
ap.add_argument("--lines", type=int, default=45)
ap.add_argument("--uninjected", action="store_true", help="128^2 without the golden test vectors")
ap.add_argument("--no-two-stage", action="store_true", help="level-0 eigensolve preconditioned by the polynomial only (round-2 run 5)")
ap.add_argument("--host-galerkin", action="store_true", help="scipy R*A*P, host QR and np.linalg.inv instead of the device set-up")
args = ap.parse_args()
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import bench
from deflatedmlmc_schwinger_b200 import lattice, multigrid

bench.build_solver()           # warm: CUDA context, library load, torch kernels
torch.cuda.synchronize()


def build():
    if args.L == 0:
        if not args.uninjected:
            return bench.build_solver()[0]
        from deflatedmlmc_schwinger_b200 import matrix
        p, tp = bench.params128()
        tp["skip_unused_inverses"] = True
        A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
        mg = multigrid.MG(A, smoother_degree=80, precond_degree=36, geometric_precond=True)
        mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp)
        return mg
    L = args.L
    A = lattice.wilson_matrix(lattice.random_u1_links(L, seed=L), args.mass)
    sizes = [A.shape[0] // 4 ** i for i in range(16) if A.shape[0] // 4 ** i >= 2048]
    nlev = len(sizes)
    params = {"use_permuted": False, "latt_dims": [L, L], "x_displacement": 2, "test_vectors_type": "EVs",
              "function_params": {"tol": 1e-12}}
    if args.no_two_stage:
        params["two_stage_min_n"] = 1 << 62
    if args.host_galerkin:
        params["host_galerkin"] = True
    mg = multigrid.MG(A, smoother_degree=80, geometric_precond=True, precond_degree=36)
    mg.setup(dof=[2] + [8] * (nlev - 1), aggrs=[16] + [4] * (nlev - 2), max_levels=nlev, acc_eigvs="low", params=params)
    return mg


pr = cProfile.Profile()
t0 = time.time()
pr.enable()
mg = build()
torch.cuda.synchronize()
pr.disable()
print("setup wall %.3f s, levels %s" % (time.time() - t0, mg.level_shapes))
for lvl, info in sorted(getattr(mg, "test_vector_info", {}).items()):
    print("test vectors level %d: |theta| %s residual max %.2e, %d block solves of %d columns, %d FGMRES iterations (bootstrap degree %d)"
          % (lvl, np.round(np.abs(info["theta"]), 6), info["residuals"].max(), info["block_solves"], info["block"], info["fgmres_iters"],
             info["bootstrap_degree"]))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(args.lines)
print(s.getvalue())
