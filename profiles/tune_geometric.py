"""Dev tool (GPU): outer iterations and time per 256-probe step of the 128^2 level-0 MLMC sample for several
degrees of the geometric preconditioner's smoother and storage options.  One JSON line per configuration."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from deflatedmlmc_schwinger_b200 import matrix, multigrid, sampling, utils

p, tp = bench.params128()
A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
tvs = bench.golden_tvs()
k = 256
configs = [dict(degree=d) for d in (28, 32, 36, 40)] + [dict(degree=32, smoother_half=0), dict(degree=32, dense_tensor_min_n=1 << 30),
                                                         dict(degree=32, blocks=(4, 8)), dict(degree=32, blocks=(2, 4)), dict(degree=16, blocks=(2, 4))]
if len(sys.argv) > 1:
    configs = json.loads(sys.argv[1])
for cfg in configs:
    mg = multigrid.MG(A, smoother_degree=80, precond_degree=cfg["degree"], precond_blocks=tuple(cfg.get("blocks", (4, 4))),
                      precond_eo_degree=cfg.get("eo_degree", 16))
    t0 = time.time()
    mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"], params=tp, test_vectors=tvs)
    setup_s = time.time() - t0
    mg.skip_level = True
    for name in ("smoother_half", "dense_tensor_min_n", "fuse_io", "dense_split_bf16", "fuse_res", "adaptive_poll", "dot32", "smoother_eo", "eo_by", "eo_bz", "eo_packs"):
        if name in cfg:
            mg.set_option(name, cfg[name])
    dev = mg.dev
    n0 = mg.level_shapes[0]
    np.random.seed(123456)
    bits = torch.from_numpy(utils.pack_bits(sampling.draw_probe_bits(k * n0))).cuda()
    X0 = dev.probe_expand(bits, n0, k)
    for _ in range(3):
        e, it = dev.level_sample(1, 0, 2, X0, 1e-12, 40, 1000)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(4):
        e, it = dev.level_sample(1, 0, 2, X0, 1e-12, 40, 1000)
    torch.cuda.synchronize()
    dt = (time.time() - t) / 4
    print(json.dumps({"config": cfg, "levels_precond": mg.precond_mg.level_shapes if mg.precond_mg else None, "ms_per_step": 1e3 * dt, "probes_per_s": k / dt,
                      "iters_level0": [int(it[0].min()), int(it[0].max())], "iters_level2": [int(it[1].min()), int(it[1].max())],
                      "setup_s": setup_s}), flush=True)
    del mg, dev
    torch.cuda.empty_cache()
