"""End-to-end run of the reference's shipped experiment G202 (gateway.py:52-59: Schwinger 128^2, deflated MLMC,
permuted, level 1 skipped, variance target trace_tol = 1e-2) and of G102 (deflated Hutchinson) on the GPU path,
through the drop-in modules.  One JSON line per experiment on rank 0.

    python profiles/run_e2e.py [--golden-tvs] [--skip-hutchinson] [--batch 256]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/run_e2e.py ...

--golden-tvs injects the test vectors of tests/golden/schwinger128.npz (skips the host eigensolve of
multigrid.py:174, so that the hierarchy equals the oracle's)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np

EXACT_DISPLACED = -8.748242701374695 + 50.215154098005584j      # gateway.py:104
EXACT_PLAIN = 8326.432059538896 + 0j                             # tr(A^-1), not permuted (SURVEY.md 8c pin 2)

ap = argparse.ArgumentParser()
ap.add_argument("--golden-tvs", action="store_true")
ap.add_argument("--skip-hutchinson", action="store_true")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--fixed", action="store_true", help="sequential_stop=False: sample count from a pilot round, one all_reduce per level")
ap.add_argument("--exact", action="store_true", help="also compute the EXACT level values with unit vectors (stoch_trace.exact_trace)")
ap.add_argument("--set", default="schwinger128", help="gateway.set_params name: schwinger128 (G202/G102) or synthetic<L> (BASELINE configs[4])")
ap.add_argument("--coarse-geo", type=int, default=-1, help="params['geometric_coarse_levels']: deepest estimator level that gets a geometric preconditioner hierarchy (default: the package's, 2)")
ap.add_argument("--trace-tol", type=float, default=0.0, help="params['trace_tol'] (default: the parameter set's 1e-2)")
ap.add_argument("--deflated", action="store_true",
                help="the valid deflated-MLMC variant of SURVEY.md 8d cfg-2: not permuted, mlmc_deflat_vctrs=[16,0,16]")
args = ap.parse_args()

import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as ge
if rank == 0:
    ge.build()
if world > 1:
    dist.barrier()
from deflatedmlmc_schwinger_b200 import gateway, matrix, stoch_trace, utils


def run(method):
    p = gateway.set_params(args.set)
    p["function_tol"] = 1e-12
    p["verbose"] = False
    p["probe_batch"] = args.batch
    p["sequential_stop"] = not args.fixed
    if args.trace_tol > 0:
        p["trace_tol"] = args.trace_tol
    if args.deflated:
        p["use_permuted"] = False
        p["mlmc_deflat_vctrs"] = [16, 0, 16]
    tp = utils.trace_params_from_params(p, method)
    if args.coarse_geo >= 0:
        tp["geometric_coarse_levels"] = args.coarse_geo
    if args.golden_tvs:
        g = np.load(os.path.join(ROOT, "tests", "golden", "schwinger128.npz"))
        tp["test_vectors"] = [g["tv0"], g["tv1"], g["tv2"]]
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    t0 = time.time()
    res = (stoch_trace.mlmc if method == "mlmc" else stoch_trace.hutchinson)(A, tp)
    torch.cuda.synchronize()
    wall = time.time() - t0
    exact = EXACT_PLAIN if args.deflated else EXACT_DISPLACED
    if args.set != "schwinger128":
        exact = complex("nan")
    out = {"experiment": (("G202 (mlmc)" if method == "mlmc" else "G102 (hutchinson)") if args.set == "schwinger128" else
                          args.set + " (" + method + ")") +
                         (", not permuted, mlmc_deflat_vctrs=[16,0,16]" if args.deflated else ""), "n_gpus": world,
           "level_shapes": res.get("level_shapes"), "setup_s": res.get("setup_seconds"),
           "trace": [float(np.real(res["trace"])), float(np.imag(res["trace"]))],
           "exact": [exact.real, exact.imag],
           "abs_err": float(abs(res["trace"] - exact)),
           "target_err": float(abs(p["trace_tol"] * res["rough_trace"])),
           "rough_trace": [float(np.real(res["rough_trace"])), float(np.imag(res["rough_trace"]))],
           "wall_s": wall, "sampling_s": float(res["sampling_seconds"]), "probes_evaluated": res["probes_evaluated"],
           "sequential_stop": not args.fixed, "probe_batch_per_gpu": args.batch}
    if method == "mlmc":
        out["levels"] = [{"nr_ests": int(r["nr_ests"]), "avg": [float(np.real(r["ests_avg"])), float(np.imag(r["ests_avg"]))],
                          "dev": float(r["ests_dev"]), "function_iters": int(r["function_iters"])} for r in res["results"]]
        out["probes_per_s"] = float(sum(res["probes_evaluated"]) / res["sampling_seconds"])
    else:
        out["nr_ests"] = int(res["nr_ests"]); out["std_dev"] = float(res["std_dev"])
        out["probes_per_s"] = float(res["probes_evaluated"] / res["sampling_seconds"])
    if rank == 0:
        print(json.dumps(out), flush=True)


run("mlmc")
if args.exact and rank == 0:
    p = gateway.set_params(args.set); p["function_tol"] = 1e-12; p["verbose"] = False
    p["probe_batch"] = args.batch
    if args.deflated:
        p["use_permuted"] = False
    tp = utils.trace_params_from_params(p, "mlmc")
    if args.golden_tvs:
        g = np.load(os.path.join(ROOT, "tests", "golden", "schwinger128.npz"))
        tp["test_vectors"] = [g["tv0"], g["tv1"], g["tv2"]]
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    t0 = time.time()
    ex = stoch_trace.exact_trace(A, tp)
    exact = (EXACT_PLAIN if args.deflated else EXACT_DISPLACED) if args.set == "schwinger128" else complex("nan")
    print(json.dumps({"experiment": "exact level traces (unit vectors through the batched solver)",
                      "levels": [[float(np.real(r["ests_avg"])), float(np.imag(r["ests_avg"]))] for r in ex["results"]],
                      "trace": [float(np.real(ex["trace"])), float(np.imag(ex["trace"]))], "reference_exact": [exact.real, exact.imag],
                      "abs_diff": float(abs(ex["trace"] - exact)), "wall_s": time.time() - t0}), flush=True)
if not args.skip_hutchinson:
    run("hutchinson")
if world > 1:
    dist.destroy_process_group()
