"""Dev tool (GPU): time one batch of k level-0 MLMC samples of the 128^2 set for a grid of smoother
degrees / solver options.  python profiles/tune.py "32,32,32;1" "64,32,32;0;40;chunk_cols=128,stencil_by=8" ...
(degrees;reorth;restart;option=value,...)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np, torch
import bench
from deflatedmlmc_schwinger_b200 import matrix, multigrid, sampling, utils

k = int(os.environ.get("K", "256"))
p, tp = bench.params128()
A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
tvs = bench.golden_tvs()
np.random.seed(123456)
bits = torch.from_numpy(utils.pack_bits(sampling.draw_probe_bits(k * 32768))).cuda()
e_ref = None
for cfg in sys.argv[1:]:
    parts = cfg.split(";")
    deg = [int(x) for x in parts[0].split(",")]
    reorth = int(parts[1]) if len(parts) > 1 else 1
    restart = int(parts[2]) if len(parts) > 2 else 40
    mg = multigrid.MG(A, smoother_degree=deg, restart=restart)
    mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp, test_vectors=tvs)
    mg.skip_level = True
    dev = mg.dev
    dev.set_option("reorth", reorth)
    opts = {}
    if len(parts) > 3 and parts[3]:
        for kv in parts[3].split(","):
            name, val = kv.split("=")
            opts[name] = float(val)
            dev.set_option(name, float(val))
    X0 = dev.probe_expand(bits, 32768, k)
    for _ in range(2):
        e, it = dev.level_sample(1, 0, 2, X0, 1e-12, restart, 1000)
    torch.cuda.synchronize()
    l0 = dev.launch_count(); t = time.time()
    reps = 3
    for _ in range(reps):
        e, it = dev.level_sample(1, 0, 2, X0, 1e-12, restart, 1000)
    torch.cuda.synchronize()
    dt = (time.time() - t) / reps
    e = e.cpu().numpy()
    if e_ref is None:
        e_ref = e
    print(json.dumps({"deg": deg, "reorth": reorth, "restart": restart, "opts": opts, "ms": 1e3 * dt, "probes_per_s": k / dt,
                      "it0": [int(it[0].min()), int(it[0].max())], "it2": [int(it[1].min()), int(it[1].max())],
                      "launches": (dev.launch_count() - l0) // reps,
                      "max_rel_diff_vs_first": float(np.abs(e - e_ref).max() / np.abs(e_ref).max())}), flush=True)
    del mg, dev
    torch.cuda.empty_cache()
