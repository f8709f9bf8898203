"""Dev tool (CPU): test vectors of a synthetic random-U(1) lattice, computed exactly as
MG.setup would (scipy eigs, multigrid.py:174) and saved to gpurun_cache/synthetic_L<L>_tvs.npz so that the
GPU box does not spend minutes in a host eigensolver.  python profiles/make_synthetic_tvs.py L mass"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "4")
import numpy as np
from scipy.sparse.linalg import eigs
from deflatedmlmc_schwinger_b200 import lattice
from deflatedmlmc_schwinger_b200.multigrid import build_prolongator_values, prolongator_csr

L = int(sys.argv[1]); mass = float(sys.argv[2])
aggrs = [16] + [4] * 8
dof = [2] + [8] * 9
links = lattice.random_u1_links(L, seed=L)
A = lattice.wilson_matrix(links, mass).tocsr()
tvs = []
Al = A
lvl = 0
while Al.shape[0] > 2048:
    dofi = dof[lvl] if lvl == 0 else dof[lvl] // 2
    nvec = dof[lvl + 1] // 2
    t = time.time()
    w, V = eigs(Al, k=nvec, which='LM', tol=1e-3, maxiter=1000000, sigma=0.0, ncv=nvec + 2)
    print("level", lvl, "n", Al.shape[0], "eigs", np.round(w, 5), "%.1f s" % (time.time() - t), flush=True)
    tvs.append(V)
    aggr = aggrs[lvl] * dofi if lvl == 0 else aggrs[lvl] * dofi * 2
    P = prolongator_csr(build_prolongator_values(V, aggr, dofi, nvec), aggr, dofi, nvec)
    Al = (P.conj().T.tocsr() @ Al @ P).tocsr()
    lvl += 1
os.makedirs(os.path.join(ROOT, "gpurun_cache"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_cache", "synthetic_L%d_tvs.npz" % L), mass=mass, **{"tv%d" % i: v for i, v in enumerate(tvs)})
print("levels:", lvl + 1, "coarsest", Al.shape[0])
