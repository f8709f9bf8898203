"""BASELINE config 5: synthetic random-U(1) Schwinger lattice, L = 256 / 512 (/1024), deeper
hierarchy (aggrs=[16,4,4,...], dof=[2,8,8,...]).  Sets the hierarchy up (test vectors from
gpurun_cache/synthetic_L<L>_tvs.npz if present -- see make_synthetic_tvs.py -- else the host eigensolver),
then times batches of k level-0 MLMC difference samples (fine level 0, coarse level 1, tol 1e-12) and checks
the true residuals.  One JSON line.   python profiles/run_synthetic.py --L 512 --probes 64 [--mass -0.062]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "4")
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=256)
ap.add_argument("--probes", type=int, default=64)
ap.add_argument("--mass", type=float, default=-0.062)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--degree", type=int, default=80, help="smoother degree on the estimator's own hierarchy")
ap.add_argument("--precond-degree", type=int, default=36)
ap.add_argument("--coarse-degree", type=int, default=0, help="smoother degree on the coarse levels of the geometric hierarchies (0: the default, 16)")
ap.add_argument("--no-geometric", action="store_true")
args = ap.parse_args()

import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:                      # one process per GPU, probes sharded: rank g takes the g-th block of k probes
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as ge
if rank == 0:
    ge.build()
if world > 1:
    dist.barrier()
from deflatedmlmc_schwinger_b200 import lattice, multigrid, sampling, utils

L = args.L
links = lattice.random_u1_links(L, seed=L)
A = lattice.wilson_matrix(links, args.mass)
n0 = A.shape[0]
nlev = 2
n = n0 // 4
while n > 2048:
    n //= 4; nlev += 1
nlev += 1 if n0 // 4 ** (nlev - 1) > 2048 else 0
sizes = [n0 // 4 ** i for i in range(16) if n0 // 4 ** i >= 2048]
nlev = len(sizes)
dof = [2] + [8] * (nlev - 1)
aggrs = [16] + [4] * (nlev - 2)
tvs = None
cache = os.path.join(ROOT, "gpurun_cache", "synthetic_L%d_tvs.npz" % L)
if os.path.isfile(cache):
    g = np.load(cache)
    if abs(float(g["mass"]) - args.mass) < 1e-12:
        tvs = [g["tv%d" % i] for i in range(nlev - 1)]
params = {"use_permuted": False, "latt_dims": [L, L], "x_displacement": 2, "test_vectors_type": "EVs",
          "function_params": {"tol": 1e-12}}
t0 = time.time()
mg = multigrid.MG(A, smoother_degree=args.degree, geometric_precond=not args.no_geometric, precond_degree=args.precond_degree,
                  precond_coarse_degree=args.coarse_degree or None)
mg.setup(dof=dof, aggrs=aggrs, max_levels=nlev, acc_eigvs="low", params=params, test_vectors=tvs)
torch.cuda.synchronize()
setup_s = time.time() - t0
dev = mg.dev
k = args.probes
restart, maxiter = 40, 1000
np.random.seed(123456)
sampling.skip_probe_words(rank * k * n0)
bits = torch.from_numpy(utils.pack_bits(sampling.draw_probe_bits(k * n0))).cuda()
X0 = dev.probe_expand(bits, n0, k)
# one solve with residual check
Xs, iters, relres = dev.fgmres(0, X0, 1e-12, restart=restart, maxiter=maxiter)
R = X0 - dev.spmm(0, Xs)
true_rel = float((torch.linalg.vector_norm(R, dim=0) / torch.linalg.vector_norm(X0, dim=0)).max())
del Xs, R
# time of the two solves of a sample separately (level 0: geometric preconditioner; level 1: the estimator's own hierarchy)
def _time_solve(level, B, reps=2):
    dev.fgmres(level, B, 1e-12, restart=restart, maxiter=maxiter)
    torch.cuda.synchronize(); t_ = time.time()
    for _ in range(reps):
        _, it_, _ = dev.fgmres(level, B, 1e-12, restart=restart, maxiter=maxiter)
    torch.cuda.synchronize()
    return 1e3 * (time.time() - t_) / reps, [int(it_.min()), int(it_.max())]
ms_solve0, it_solve0 = _time_solve(0, X0)
X1 = dev.restrict(0, X0)
ms_solve1, it_solve1 = _time_solve(1, X1)
del X1
e, it = dev.level_sample(1, 0, 1, X0, 1e-12, restart, maxiter)      # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = dev.launch_count(); t = time.time()
for _ in range(args.steps):
    e, it = dev.level_sample(1, 0, 1, X0, 1e-12, restart, maxiter)
torch.cuda.synchronize()
dt = (time.time() - t) / args.steps
if world > 1:                      # the level's single collective + max time over ranks
    red = torch.tensor([e.real.sum().item(), e.imag.sum().item(), float(e.numel())], device="cuda", dtype=torch.float64)
    dist.all_reduce(red)
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    mean_est = [float(red[0] / red[2]), float(red[1] / red[2])]
else:
    mean_est = [float(e.real.mean()), float(e.imag.mean())]
if rank == 0:
  print(json.dumps({"n_gpus": world, "workload": "synthetic random-U(1) Schwinger %dx%d, m=%g, level-0 MLMC difference samples (coarse level 1)" % (L, L, args.mass),
                  "levels": mg.level_shapes, "dense_levels": {str(a): b for a, b in mg.dense_levels.items()},
                  "smoother_degrees": mg.smoother_degrees_used, "probes_per_gpu": k, "ms_per_batch": 1e3 * dt,
                  "probes_per_s": world * k / dt, "fgmres_iters_level0": [int(it[0].min()), int(it[0].max())],
                  "fgmres_iters_level1": [int(it[1].min()), int(it[1].max())],
                  "solve_iters": [int(iters.min()), int(iters.max())], "ms_solve_level0": ms_solve0, "ms_solve_level1": ms_solve1,
                  "precond_levels": mg.precond_mg.level_shapes if mg.precond_mg is not None else None,
                  "precond_dense_levels": {str(a): b for a, b in mg.precond_mg.dense_levels.items()} if mg.precond_mg is not None else None,
                  "precond_degrees": mg.precond_mg.smoother_degrees_used if mg.precond_mg is not None else None,
                  "precond1_levels": mg.precond_mg1.level_shapes if mg.precond_mg1 is not None else None,
                  "precond1_degrees": mg.precond_mg1.smoother_degrees_used if mg.precond_mg1 is not None else None, "max_true_relres": true_rel,
                  "launches_per_batch": (dev.launch_count() - l0) // args.steps, "setup_s": setup_s,
                  "test_vectors": "cache" if tvs is not None else "host eigs",
                  "mean_estimate": mean_est}), flush=True)
if world > 1:
    dist.destroy_process_group()
