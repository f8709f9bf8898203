#!/usr/bin/env python
"""bench.py -- throughput of the hot path on B200(s):  tr(D^-1) probes/sec, Schwinger 128^2.

A "step" is one batch of K_PROBES deflated-MLMC level-0 samples of the shipped 128^2
configuration (BASELINE.json configs[1]; gateway.set_params('schwinger128'): permuted, level 1
skipped): per probe one FGMRES solve on level 0 (n = 32768) to 1e-12, one on level 2 (n = 2048),
the transfers R1 R0 / P0 P1, the permutation and the two inner products -- i.e. one call of the
fused dmlmc_level_sample on k probes (reference: utils.one_defl_Hutch_step, utils.py:252-357).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--probes k] [--impl reference]

One JSON line on stdout (rank 0).  `value` times the device-resident path (probe bits already in
HBM); `e2e` times the host-buffer C-ABI call dmlmc_level_sample_host (packed probe bits H2D and the
estimates D2H inside the timed region).  N > 1: one process per GPU (torchrun), probes sharded by
rank, no data-path collective except the single all_reduce of the level's partial sums.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
# stdout carries ONE JSON line (rank 0) and nothing else: whatever libraries write to file descriptor 1 (NCCL prints its
# version banner there, whatever NCCL_DEBUG_FILE says) is sent to stderr, the line itself goes to the saved descriptor
_JSON_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)
# NCCL's log (whatever level the caller asks for through NCCL_DEBUG) goes to stderr: rank 0 prints ONE JSON line on stdout
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

METRIC = "tr(D^-1) probes/sec, Schwinger 128^2 (deflated-MLMC level-0 difference samples)"
WORKLOAD = "schwinger128.mat deflated MLMC level-0 samples (configs[1]; permuted, skip level 1, tol 1e-12)"


def golden_tvs():
    g = np.load(os.path.join(ROOT, "tests", "golden", "schwinger128.npz"))
    return [g["tv0"], g["tv1"], g["tv2"]]


def params128():
    from deflatedmlmc_schwinger_b200 import gateway, utils
    p = gateway.set_params("schwinger128")
    p["function_tol"] = 1e-12
    p["verbose"] = False
    return p, utils.trace_params_from_params(p, "mlmc")


def build_solver(precond="geometric", degree=0, options=()):
    """The hierarchy of the bench workload, exactly as the timed run builds it (tests/test_gpu_parity_round2.py checks
    this very object against the reference's golden samples).  Returns (mg, tp, A, degree)."""
    from deflatedmlmc_schwinger_b200 import matrix, multigrid
    p, tp = params128()
    A = matrix.loadMatrix(p["matrix"], p["matrix_params"])
    geo = precond == "geometric"
    tp["skip_unused_inverses"] = True
    if degree <= 0:
        degree = 36 if geo else 80
    mg = multigrid.MG(A, smoother_degree=80, precond_degree=degree, geometric_precond=True) if geo else \
        multigrid.MG(A, smoother_degree=degree, geometric_precond=False)
    mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], acc_eigvs=tp["accuracy_mg_eigvs"],
             params=tp, test_vectors=golden_tvs())
    mg.skip_level = True
    for name, value in options:
        mg.set_option(name, value)
    return mg, tp, A, degree


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference algorithm (the reference is Python and cannot travel)

_CPU = {}


def _cpu_setup():
    if "mp" not in _CPU:
        from oracle import refport
        p, tp = params128()
        mp = refport.MGPort(refport.load_matrix(p["matrix"], p["matrix_params"]["mass"]))
        mp.setup(tp["dof"], tp["aggrs"], tp["max_nr_levels"], tp["accuracy_mg_eigvs"], tp, test_vectors=golden_tvs())
        mp.skip_level = True
        _CPU["mp"], _CPU["tp"] = mp, tp
    return _CPU["mp"], _CPU["tp"]


def _cpu_probe(seed):
    """one level-0 MLMC sample with the reference algorithm (FGMRES + V-cycle with lgmres smoother)"""
    from oracle import refport
    mp, tp = _cpu_setup()
    rs = np.random.RandomState(seed)
    t = time.time()
    e, _ = refport.one_defl_hutch_step(mp.levels[0].A, mp.levels[2].A, mp, tp, "mlmc", 0, None, None, rs, 0)
    return time.time() - t, complex(e)


def cpu_baseline_pool(min_probes=16):
    """the in-line cpu_baseline of the GPU arm: >= 16 level-0 MLMC samples of the same workload with the oracle port of the
    reference algorithm on all host cores (one probe per process at a time), SURVEY.md 8d"""
    import multiprocessing as mpx
    cores = max(1, (os.cpu_count() or 1))
    cores = min(cores, int(os.environ.get("DMLMC_REF_CORES", cores)))
    n_probes = cores * max(1, -(-min_probes // cores))
    real = _ref_available() and not os.environ.get("DMLMC_REF_PORT")
    if real:
        try:
            _ref_setup()
        except Exception as why:
            print("[bench] reference set-up failed (%s); timing the oracle port instead" % why, file=sys.stderr, flush=True)
            real = False
    if not real:
        _cpu_setup()
    probe = _ref_probe if real else _cpu_probe
    with mpx.get_context("fork").Pool(cores) as pool:
        pool.map(probe, list(range(900, 900 + cores)))            # warm-up (page-in, scipy imports)
        t = time.time()
        pool.map(probe, list(range(1000, 1000 + n_probes)), chunksize=1)
        dt = time.time() - t
    what = ("the unmodified reference from oracle/_ref (multigrid.MG.setup + utils.one_defl_Hutch_step: pyamg-style FGMRES + "
            "V-cycle with scipy lgmres smoother, tol 1e-12; golden test vectors replayed into its eigs calls)" if real else
            "oracle port of the reference algorithm (oracle/_ref did not travel)")
    return {"value": n_probes / dt, "unit": "probes/s", "cores": cores, "kind": "reference" if real else "port",
            "sample": "%d level-0 MLMC samples of the same workload on %d processes: %s" % (n_probes, cores, what)}


_REF = {}


def _ref_available():
    try:
        from oracle import ref_shim
        return ref_shim.reference_available()
    except Exception:
        return False


def _ref_setup():
    """The UNMODIFIED reference (oracle/_ref, a byte-for-byte copy made by oracle/vendor_reference.py; loaded through
    oracle/ref_shim.py): MG(A).setup(...) of gateway.set_params('schwinger128') with the golden test vectors replayed into its
    eigs calls (same hierarchy as the GPU arm and the golden fixtures; skips minutes of ARPACK)."""
    if "mg" not in _REF:
        import contextlib, io
        from scipy.sparse import csr_matrix
        from oracle import ref_shim
        tv = golden_tvs()
        rec = ref_shim.EigRecorder(replay=[("eigs", np.zeros(t.shape[1]), t) for t in tv])
        ref = ref_shim.load_reference(rec)
        p = ref_shim.params_128()
        with ref_shim.in_reference_dir():
            A = ref["matrix"].loadMatrix(p["matrix"], p["matrix_params"])
        tp = ref["utils"].trace_params_from_params(p, "mlmc")
        mg = ref["multigrid"].MG(A)
        with contextlib.redirect_stdout(io.StringIO()):
            mg.setup(dof=tp["dof"], aggrs=tp["aggrs"], max_levels=tp["max_nr_levels"], dim=2, acc_eigvs=tp["accuracy_mg_eigvs"],
                     sys_type="schwinger", params=tp)
        mg.total_levels = len(mg.ml.levels)
        for i in range(mg.total_levels - 1):                       # stoch_trace.py:262-264
            mg.ml.levels[i].P = csr_matrix(mg.ml.levels[i].P)
            mg.ml.levels[i].R = csr_matrix(mg.ml.levels[i].R)
        mg.skip_level = True
        _REF.update({"ref": ref, "mg": mg, "tp": tp})
    return _REF["ref"], _REF["mg"], _REF["tp"]


def _ref_probe(seed):
    """one level-0 MLMC sample by the reference's own utils.one_defl_Hutch_step (utils.py:207-361), probe drawn inside"""
    import contextlib, io
    ref, mg, tp = _ref_setup()
    lv = mg.ml.levels
    out = {"results": [{"function_iters": 0} for _ in range(len(lv))]}
    np.random.seed(seed)
    t = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        e, _ = ref["utils"].one_defl_Hutch_step(lv[0].A, lv[2].A, mg, tp, "mlmc", 0, None, None, 0, out,
                                                lv[0].P, lv[0].R, lv[1].P, lv[1].R)
    return time.time() - t, complex(e)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, one probe per core per step:
    the unmodified reference from oracle/_ref when it travelled with the snapshot (kind "reference"), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mpx
    cores = max(1, (os.cpu_count() or 1))
    cores = min(cores, int(os.environ.get("DMLMC_REF_CORES", cores)))
    ctx = mpx.get_context("fork")
    real = _ref_available() and not os.environ.get("DMLMC_REF_PORT")
    if real:
        try:
            _ref_setup()
        except Exception as why:                 # e.g. not enough host memory for the reference's dense prolongator
            print("[bench] reference set-up failed (%s); timing the oracle port instead" % why, file=sys.stderr, flush=True)
            real = False
    if not real:
        _cpu_setup()
    probe = _ref_probe if real else _cpu_probe
    with ctx.Pool(cores) as pool:
        step = 0
        for _ in range(args.warmup):
            pool.map(probe, [step * cores + c for c in range(cores)]); step += 1
        t = time.time()
        for _ in range(args.steps):
            pool.map(probe, [step * cores + c for c in range(cores)]); step += 1
        dt = time.time() - t
    value = args.steps * cores / dt
    what = ("the unmodified reference (oracle/_ref: multigrid.MG.setup + utils.one_defl_Hutch_step, golden test vectors replayed "
            "into its eigs calls)" if real else "oracle port of the reference algorithm (oracle/_ref did not travel)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "schwinger128 gauge field (reference input) + MT19937 Rademacher probes",
            "config": {"workload": WORKLOAD, "probes_per_step": cores},
            "cpu_baseline": {"value": value, "unit": "probes/s", "cores": cores, "kind": "reference" if real else "port",
                             "sample": "%d level-0 MLMC samples per step, one per core: %s" % (cores, what)},
            "e2e": {"value": value, "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
EXACT_DISPLACED = -8.748242701374695 + 50.215154098005584j      # /root/reference gateway.py:104


def run_experiment(mg, A, k, world, runs=2):
    """The reference's shipped experiment G202 (gateway.py:52-59 -> stoch_trace.mlmc: rough trace, level-0 and level-2
    difference samples with the sequential stopping rule of stoch_trace.py:386-406 up to the variance target trace_tol = 1e-2,
    coarsest term directly) on the solver just timed, probes from the device MT19937 stream INSIDE the timed region, the probes
    of every round sharded over the ranks: total work is fixed, so this is the STRONG-scaling figure of the run.  The hierarchy
    (injected test vectors) is not rebuilt.  Returns the record of the last of `runs` runs (+ the first run's time)."""
    import torch
    import torch.distributed as dist
    from deflatedmlmc_schwinger_b200 import stoch_trace
    recs = []
    for _ in range(runs):
        p, tp = params128()
        tp.update({"mg_solver": mg, "probe_batch": k, "device_probe_stream": True, "sequential_stop": True})
        torch.cuda.synchronize()
        t0 = time.time()
        res = stoch_trace.mlmc(A, tp)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ts = torch.tensor([res["sampling_seconds"], wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        recs.append((res, float(ts[0].item()), float(ts[1].item())))
    res, samp, wall = recs[-1]
    used = [int(r["nr_ests"]) + 1 for i, r in enumerate(res["results"][:-1]) if r["nr_ests"] > 0]
    return {"name": "G202: schwinger128.mat deflated MLMC to the reference variance target (trace_tol 1e-2, sequential stop rule, "
                    "device probe stream, levels 0 and 2 sampled, level 1 skipped, coarsest term direct)",
            "scaling": "strong", "n_gpus": world, "probe_batch_per_gpu": k,
            "sampling_s": samp, "sampling_s_first_run": recs[0][1], "wall_s_without_hierarchy_setup": wall,
            "stop_indices": [int(r["nr_ests"]) for r in res["results"]],
            "probes_used": used, "probes_evaluated": [int(x) for x in res["probes_evaluated"]],
            "probes_per_s": float(sum(used) / samp), "evaluated_probes_per_s": float(sum(res["probes_evaluated"]) / samp),
            "trace": [float(np.real(res["trace"])), float(np.imag(res["trace"]))],
            "abs_err_vs_exact": float(abs(res["trace"] - EXACT_DISPLACED)),
            "target_err": float(abs(1e-2 * res["rough_trace"])),
            "exact": [EXACT_DISPLACED.real, EXACT_DISPLACED.imag]}


def spmm_sweep(mg, peak, ks=(1, 8, 32, 128, 256, 512)):
    """BASELINE.json configs[2] ("batched multi-RHS probe sweep ... measuring SpMV/SpMM HBM roofline fraction"): Y = A_l X for
    the estimator's levels 0 (link-form stencil), 1, 2 (padded BSR), the restriction R_0 and the prolongation P_0 on k
    Rademacher columns, complex128, CUDA events, a 256 MB buffer overwritten between timed calls (L2 flush: below k ~ 64 the
    operands fit in L2).  Algorithmic bytes per call as SURVEY.md 8d: level 0 n0 s (1 + 2k), level 1 n1 s (36 + 2k) + 9 n1,
    level 2 n2 s (48 + 2k) + 3 n2, restrict n0 s (4 + 1.25 k), prolong n0 s (4 + 2.25 k), s = 16 B."""
    import torch
    dev = mg.dev
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n = mg.level_shapes
    s = 16

    def timeit(fn, reps=7):
        fn(); fn()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts))
    rows = []
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    for k in ks:
        X0 = (torch.randint(0, 2, (n[0], k), device="cuda", generator=g).to(torch.float64) * 2 - 1).to(torch.complex128).contiguous()
        row = {"k": k}
        for lvl, by in ((0, n[0] * s * (1 + 2 * k)), (1, n[1] * s * (36 + 2 * k) + 9 * n[1]), (2, n[2] * s * (48 + 2 * k) + 3 * n[2])):
            X = X0[:n[lvl]].contiguous(); Y = torch.empty_like(X)
            us = timeit(lambda: dev.spmm(lvl, X, Y))
            row["A%d" % lvl] = {"us": round(us, 2), "GBps": round(by / us / 1e3, 1), "frac": round(by / us / 1e3 / peak, 3)}
        Xc = dev.restrict(0, X0)
        us = timeit(lambda: dev.restrict(0, X0)); by = n[0] * s * (4 + 1.25 * k)
        row["R0"] = {"us": round(us, 2), "GBps": round(by / us / 1e3, 1), "frac": round(by / us / 1e3 / peak, 3)}
        us = timeit(lambda: dev.prolong_add(0, Xc, X0)); by = n[0] * s * (4 + 2.25 * k)
        row["P0"] = {"us": round(us, 2), "GBps": round(by / us / 1e3, 1), "frac": round(by / us / 1e3 / peak, 3)}
        rows.append(row)
        del X0, Xc
    del flush
    return rows


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, through NVML in-process (an `nvidia-smi`
    subprocess per sample stalls the CUDA launches of every rank for ~100 ms on an 8-GPU box)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self._halt = threading.Event()
        self.max_mhz = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            # the first query of each kind is slow (7 ms on a 1-GPU box, 32 ms with 8 ranks asking at once; measured, run
            # r2_11) and stalls this process's launches while it lasts: make it here, outside the timed region, and drop it
            self._sample_nvml()
            self.samples, self.reasons = [], set()
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for nm, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0])); self.max_mhz = float(out[1])
        for nm, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        self.call_ms = []
        while not self._halt.is_set():
            t = time.time()
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.call_ms.append(1e3 * (time.time() - t))
            self._halt.wait(1.0)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "query_ms_max": max(self.call_ms) if getattr(self, "call_ms", None) else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--probes", type=int, default=512, help="probes per step per GPU (batch of the fused level sample)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--degree", type=int, default=0, help="smoother degree of the level-0 V-cycle (0: 36 geometric / 80 reference)")
    ap.add_argument("--precond", default="geometric", choices=["geometric", "reference"],
                    help="hierarchy of the V-cycle that preconditions the level-0 solve: geometric 4x4-site aggregates "
                         "(default) or the estimator's own (reference aggregation)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-experiment", action="store_true", help="skip the strong-scaling G202 run")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="solver option (dmlmc_set_option), repeatable")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from deflatedmlmc_schwinger_b200 import sampling, utils, _lib

    t0 = time.time()
    options = [(o.split("=")[0], float(o.split("=")[1])) for o in args.opt]
    mg, tp, A, args.degree = build_solver(args.precond, args.degree, options)
    setup_s = time.time() - t0
    dev = mg.dev
    pmg = mg.precond_mg if mg.precond_mg is not None else mg     # the hierarchy whose V-cycle preconditions level 0
    pdev = pmg.dev
    n0, k = mg.level_shapes[0], args.probes
    tol, restart, maxiter = 1e-12, 40, 1000
    total_steps = args.warmup + args.steps

    # probes: rank g owns the g-th block of k probes of every round of the MT19937(123456) stream
    np.random.seed(123456)
    nbytes = (n0 * k + 7) // 8
    host_bits = torch.empty((total_steps, nbytes), dtype=torch.uint8).pin_memory()
    for s in range(total_steps):
        sampling.skip_probe_words(rank * k * n0)
        host_bits[s].copy_(torch.from_numpy(utils.pack_bits(sampling.draw_probe_bits(k * n0))))
        sampling.skip_probe_words((world - 1 - rank) * k * n0)
    dev_bits = host_bits.cuda()
    dev.ensure_workspace(0, k, restart)
    stream = torch.cuda.current_stream()

    def step_device(s):
        X0 = dev.probe_expand(dev_bits[s], n0, k)
        return dev.level_sample(1, 0, 2, X0, tol, restart, maxiter)

    def step_host(s):
        return dev.level_sample_host(1, 0, 2, host_bits[s].numpy(), k, tol, restart, maxiter)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    def level_sums(e_list):
        """the single collective of the level: all_reduce of [sum Re e, sum Im e, sum |e|^2, N]"""
        e_all = torch.cat(e_list)
        sums = torch.stack([e_all.real.sum(), e_all.imag.sum(), (e_all.abs() ** 2).sum(),
                            torch.tensor(float(e_all.numel()), device=e_all.device, dtype=torch.float64)])
        if world > 1:
            dist.all_reduce(sums)
        return sums

    warm = [step_device(s)[0] for s in range(args.warmup)]
    if warm:             # also outside the timed region once: torch loads its kernels lazily (first call of each op
        level_sums(warm)  # costs ~10-50 ms) and NCCL sets up its channels on first use
    del warm
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.profiler.start()      # `ncu --profile-from-start off` captures the timed region only
    l0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    dbg = [] if os.environ.get("BENCH_DEBUG") else None
    es, its = [], []
    for s in range(args.warmup, total_steps):
        if s == args.warmup + 1 or total_steps - args.warmup == 1:
            sampler.start()          # first NVML query (and then one per second) while the GPU is under load
        e, it = step_device(s)
        es.append(e); its.append(it)
        if dbg is not None:
            evs = torch.cuda.Event(enable_timing=True); evs.record(stream); dbg.append(evs)
    sums = level_sums(es)
    ev1.record(stream)
    barrier()
    torch.cuda.profiler.stop()
    launches = dev.launch_count() - l0
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if dbg:
        print("[rank %d] timed region: steps end at %s ms, total %.1f ms" %
              (rank, ["%.1f" % ev0.elapsed_time(x) for x in dbg], ms), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * k * args.steps / (ms * 1e-3)
    iters = np.concatenate(its, axis=1)

    # ---- end-to-end timing through the host-buffer C-ABI call ----------------------------------------
    for s in range(min(args.warmup, 2)):
        step_host(s)
    barrier()
    ev0.record(stream)
    for s in range(args.warmup, total_steps):
        e_h, _ = step_host(s)
    ev1.record(stream)
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {"value": world * k * args.steps / (ms_e2e * 1e-3), "unit": "probes/s",
           "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(16 * k + 8 * k),
           "api": "dmlmc_level_sample_host (utils.defl_Hutch_batch)"}
    same = float(np.abs(e_h - es[-1].cpu().numpy()).max())

    # ---- the product's default path: probes drawn by the device MT19937 stream (utils.py:255-258 is part of the reference's
    # one_defl_Hutch_step), rank g jumping over the other ranks' blocks, estimates read back to the host every step
    src = sampling.DeviceProbeSource(dev)
    comm = sampling.Comm(dev.device)
    np.random.seed(123456)
    src.begin()
    for s in range(min(args.warmup, 2)):
        dev.level_sample(1, 0, 2, src.next_round(comm, n0, k), tol, restart, maxiter)[0].cpu()
    barrier()
    ev0.record(stream)
    for s in range(args.steps):
        e_s = dev.level_sample(1, 0, 2, src.next_round(comm, n0, k), tol, restart, maxiter)[0].cpu()
    ev1.record(stream)
    barrier()
    src.end()
    ms_ds = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_ds], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_ds = float(t.item())
    e2e["device_stream"] = {"value": world * k * args.steps / (ms_ds * 1e-3), "unit": "probes/s", "h2d_bytes_per_step": 0,
                            "d2h_bytes_per_step": int(16 * k + 8 * k),
                            "api": "sampling.DeviceProbeSource + dmlmc_level_sample (the drivers' default: probe generation "
                                   "by dmlmc_mt19937_bits with jump-ahead inside the timed region, next round's probes "
                                   "generated beside the solve)"}

    # ---- the whole experiment (strong scaling): G202 to the reference variance target ----------------------------
    experiment = None if args.no_experiment else run_experiment(mg, A, k, world)
    dev.ensure_workspace(0, k, restart)

    # ---- roofline of the dominant kernel: the level-0 fused stencil + Richardson-update step (c64) ------
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        R = torch.randn(n0, k, device="cuda", dtype=torch.float32).to(torch.complex64).contiguous()
        nu, p0 = pmg.smoother_polys[0]
        m = len(nu)
        kc = pdev.vcycle_chunk_cols(0, _lib.C64, k)
        nchunks = (k + kc - 1) // kc

        def time_smooth(reps=5):
            for _ in range(2):
                pdev.smooth(0, R)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(reps):
                pdev.smooth(0, R)
            ev1.record(stream)
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) * 1e-3 / reps

        # dmlmc_smooth = per chunk [copy in, m factor kernels, copy out]; t(m) - t(1) isolates the factor kernel
        t_full = time_smooth()
        pdev.set_smoother(0, nu[:1], p0)
        t_one = time_smooth()
        pdev.set_smoother(0, nu, p0)
        dev.set_option("use_graphs", 1)      # (drops the CUDA graphs that embed the preconditioner's kernels)
        t_step = (t_full - t_one) / max(m - 1, 1) / nchunks
        # one factor kernel on kc columns, vectors stored as BF16 (4 B per complex): read x, write x'
        # (n0 * kc * 4 B each) + 4 pre-splatted links per site (16 B each)
        sb = 4
        alg_bytes = 2 * n0 * kc * sb + 4 * (n0 // 2) * 16
        achieved = alg_bytes / t_step / 1e9
        roof = {"bound": "hbm",
                "kernel": "stencil_step_bf16_t2_kernel (level-0 operator + polynomial-factor update x' = x - nu A x of the "
                          "complex64 V-cycle, BF16-stored vectors, FP32 packed arithmetic, two sites per thread, %d columns)" % kc,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "avg_launch_us": 1e6 * t_step, "alg_bytes_per_launch": alg_bytes, "traffic": 36.8e6,
                "traffic_source": "ncu --set full, profiles/r1_run22_step_t2_ncu.md: dram__bytes_read.sum (34.6 MB) + "
                                  "dram__bytes_write.sum (1.8-3.1 MB) per launch, k = 256 (ncu replays each launch cold: the "
                                  "input comes from DRAM there, the output stays in L2)",
                "limiter": "instruction issue / integer ALU pipe, not HBM: the two %.1f MB ping-pong vectors of the polynomial "
                           "product stay in the 126 MB L2 across the %d consecutive launches (DRAM traffic per launch is about "
                           "half the algorithmic bytes) and ncu shows issue slots 62 %%, integer ALU pipe 51 %% (BF16 <-> FP32 "
                           "conversions), FMA pipe 26 %%, L2 26 %% busy, 78 registers, 31 %% of the warp slots; "
                           "spmm_level0 below is the same operator streaming complex128 / complex64 from HBM"
                           % (n0 * kc * sb / 1e6, m - 1),
                "smooth_call_us": 1e6 * t_full, "chunk_cols": kc}
        if getattr(pmg, "eo_poly", None) is not None and mg.precond_mg is not None:
            # The dominant kernel of the step is now wilson_hop_eo_kernel (the even-odd post-smoother: every sweep of the
            # polynomial in the Schur complement S).  One application of S = two half-lattice sweeps; isolated as
            # [t(m factors) - t(1 factor)] / (m - 1) of the preconditioner call the solver itself makes.
            t2_entry = roof
            nue, p0e = pmg.eo_poly
            me = len(nue)
            Vz = torch.randn(n0, k, device="cuda", dtype=torch.float64).to(torch.complex128).contiguous()

            def time_precond(reps=5):
                for _ in range(2):
                    dev.precondition(0, Vz)
                torch.cuda.synchronize()
                ev0.record(stream)
                for _ in range(reps):
                    dev.precondition(0, Vz)
                ev1.record(stream)
                torch.cuda.synchronize()
                return ev0.elapsed_time(ev1) * 1e-3 / reps

            tp_full = time_precond()
            # the sweeps themselves, launched through dmlmc_hop_eo exactly as smooth_eo launches them (w_o = H_oe y_e, then
            # y_e' = a y_e + b H_eo w_o with the factor's own a, b; ping-pong buffers), CUDA events around the launches
            LXl = LTl = int(round((n0 // 2) ** 0.5))
            cdiag = float(np.real(pmg.ml.levels[0].A.diagonal()[0]))
            Yb = [torch.randn(2, LXl, LTl // 2, k, 2, device="cuda", dtype=torch.float32).to(torch.bfloat16).contiguous() for _ in range(2)]
            Wb = torch.empty_like(Yb[0])

            def sweeps():
                cur = 0
                for nu_i in nue:
                    pdev.hop_eo(0, 1, Yb[cur], None, Wb, 1.0, 1.0, k)
                    pdev.hop_eo(0, 0, Wb, Yb[cur], Yb[cur ^ 1], 1.0 - nu_i * cdiag, nu_i / cdiag, k)
                    cur ^= 1
            for _ in range(2):
                sweeps()
            torch.cuda.synchronize()
            ev0.record(stream)
            reps_h = 5
            for _ in range(reps_h):
                sweeps()
            ev1.record(stream)
            torch.cuda.synchronize()
            t_pair = ev0.elapsed_time(ev1) * 1e-3 / (reps_h * me)
            del Yb, Wb
            half = (n0 // 2) * k * 4                      # one BF16 half-lattice vector of k columns (4 B per complex)
            links = 4 * (n0 // 4) * 16                    # 4 pre-splatted links for each of the V/2 sites of a sweep
            # sweep 1: w_o = H_oe y_e (read 1, write 1); sweep 2: y_e' = a y_e + b H_eo w_o (read 2, write 1)
            pair_bytes = 5 * half + 2 * links
            ach = pair_bytes / t_pair / 1e9
            roof = {"bound": "hbm",
                    "kernel": "wilson_hop_eo_kernel (even-odd post-smoother of the complex64 V-cycle: Out_p = a In2_p + b H In_q on "
                              "BF16 checkerboard half-lattice vectors, packed FP32 arithmetic, four columns per thread; figures "
                              "per sweep, averaged over the two sweeps of one Schur-complement application, %d columns)" % k,
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                    "frac_of_nominal_8TBs": ach / 8000.0,
                    "avg_launch_us": 1e6 * t_pair / 2, "alg_bytes_per_launch": pair_bytes / 2,
                    "timing": "CUDA events around %d consecutive launches through dmlmc_hop_eo (the V-cycle's own launch "
                              "configuration and coefficients)" % (2 * me * 5),
                    "traffic": {512: 57.1e6, 256: 26.3e6}.get(k),
                    "traffic_source": ("ncu --set full of this command at k = 512 (profiles/r2_run29_ncu_full_k512.md, 30 launches "
                                       "each; the same figures in r2_run1 / r2_run8): dram__bytes_read.sum + dram__bytes_write.sum = "
                                       "68.6 + 8.6 MB (sweep with In2) / 34.7 + 2.3 MB (sweep without), averaged over the two sweeps; "
                                       "k = 256: profiles/r1_run36_hop_eo_ncu.md; "
                                       "no capture at other batch sizes (null).  ncu replays each launch cold: the inputs come from "
                                       "DRAM there, most of the output stays in L2"),
                    "limiter": "the five %.1f MB half-lattice vectors of a Schur-complement application %s the 126 MB L2 across the "
                               "%d consecutive sweeps; ncu (k = 512): issue slots 58-61 %%, integer ALU pipe 51-54 %% (BF16 <-> FP32 "
                               "conversions), FMA pipe 46-48 %%, 70-72 registers, 36-37 %% of the warp slots, L2 hit 22-30 %% -- "
                               "instruction issue / L2 latency, not HBM; gram_schmidt_dot and spmm_level0 below are the HBM-streaming "
                               "kernels of the step"
                               % (half / 1e6, "stay in" if 5 * half < 120e6 else "no longer all fit in", 2 * me + 2),
                    "precondition_call_us": 1e6 * tp_full, "sweeps_per_vcycle": 2 * me + 2,
                    "stencil_step_bf16_t2_kernel": {kk: t2_entry[kk] for kk in ("achieved", "frac", "avg_launch_us",
                                                                                "alg_bytes_per_launch", "traffic", "traffic_source")}}
            roof["stencil_step_bf16_t2_kernel"]["note"] = ("the factor kernel of the polynomial in A itself (option smoother_eo = 0, "
                                                           "and the estimator's own hierarchy)")
            # Gram-Schmidt kernels (29 % of the step, HBM-streaming complex128): per-column dot of two vectors = multi_dot_kernel, nv = 1
            Wz = torch.randn(n0, k, device="cuda", dtype=torch.float64).to(torch.complex128).contiguous()
            for _ in range(3):
                dev.dotc(Vz, Wz)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(20):
                dev.dotc(Vz, Wz)
            ev1.record(stream)
            torch.cuda.synchronize()
            tdot = ev0.elapsed_time(ev1) * 1e-3 / 20
            roof["gram_schmidt_dot"] = {"kernel": "multi_dot_kernel (+ sum_partials_kernel), nv = 1: conj(V)^T W per column, complex128",
                                        "us": 1e6 * tdot, "GBps": 2 * n0 * k * 16 / tdot / 1e9, "frac": 2 * n0 * k * 16 / tdot / 1e9 / peak,
                                        "ncu": "profiles/r1_run33_gs_umma_ncu.md"}
            del Vz, Wz
        # plain SpMM Y = A X (config 3), c128 and c64, bytes n0*s*(2k) + links
        spmm = {}
        for name, dt, sb in (("c128", torch.complex128, 16), ("c64", torch.complex64, 8)):
            X = R.to(dt).contiguous(); Y = torch.empty_like(X)
            for _ in range(3):
                dev.spmm(0, X, Y)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(20):
                dev.spmm(0, X, Y)
            ev1.record(stream)
            torch.cuda.synchronize()
            tt = ev0.elapsed_time(ev1) * 1e-3 / 20
            by = n0 * sb * (1 + 2 * k)
            spmm[name] = {"us": 1e6 * tt, "GBps": by / tt / 1e9, "frac": by / tt / 1e9 / peak}
        roof["spmm_level0"] = spmm
        roof["spmm_sweep_c128"] = spmm_sweep(mg, peak)
        # the tensor-core kernel of the path: dense coarse solve of the V-cycle (tcgen05, BF16 x BF16 -> FP32),
        # real GEMM [2n x 2n] x [2n x k]; timed through dmlmc_vcycle on the dense level (includes the RHS pack kernel)
        dl = pmg.dense_level
        if pmg.dense_levels.get(dl) == "tensor":
            nd = pmg.level_shapes[dl]
            Xd = torch.randn(nd, k, device="cuda", dtype=torch.float32).to(torch.complex64).contiguous()
            for _ in range(3):
                pdev.vcycle(dl, Xd)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(10):
                pdev.vcycle(dl, Xd)
            ev1.record(stream)
            torch.cuda.synchronize()
            td = ev0.elapsed_time(ev1) * 1e-3 / 10
            fl = 2.0 * (2 * nd) * (2 * nd) * k
            tpeak = float(peaks.get("bf16_tflops", 1662.5))
            roof["dense_umma_kernel"] = {"bound": "tensor", "n": nd, "us": 1e6 * td, "achieved": fl / td / 1e12, "peak": tpeak,
                                         "unit": "TFLOP/s", "frac": fl / td / 1e12 / tpeak,
                                         "ncu": "sm__pipe_tensor_cycles_active 60.6 % of active cycles, DRAM read = the BF16 matrix once "
                                                "(profiles/r1_run10_umma_ncu_full.csv)"}

    line = {"metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (FGMRES, transfers, dots) + c64 (V-cycle)",
            "data": "schwinger128 gauge field (reference input) + MT19937(123456) Rademacher probes",
            "config": {"workload": WORKLOAD, "probes_per_step_per_gpu": k, "solver_tol": tol,
                       "preconditioner": ("geometric hierarchy (4x4-site spin-split aggregates, the estimator's level-0 test "
                                          "vectors)" if mg.precond_mg is not None else "the estimator's hierarchy (reference aggregation)"),
                       "smoother": ("V-cycle = dense tcgen05 coarse solve at level %d + even-odd post-smoother: fixed GMRES polynomial of "
                                    "degree %d in the Schur complement, product form" % (pmg.dense_level, len(pmg.eo_poly[0]) + 1))
                                   if getattr(pmg, "eo_poly", None) is not None else
                                   ("V-cycle = dense tcgen05 coarse solve at level %d + fixed GMRES polynomial of degree %d in "
                                    "product form as post-smoother" % (pmg.dense_level, args.degree)),
                       "fgmres_restart": restart, "l2_flush": "inputs larger than L2 (Krylov basis %.1f GB per step)"
                       % (2 * (int(iters[0].max()) + 1) * n0 * k * 16 / 1e9), "parallelism": "probes sharded x%d" % world},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "fgmres_iters": {"level0": [int(iters[0].min()), int(iters[0].max())],
                             "level2": [int(iters[1].min()), int(iters[1].max())]},
            "setup_s": setup_s, "e2e_vs_device_max_abs_diff": same}
    if experiment is not None:
        line["experiment"] = experiment
    if roof is not None:
        line["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_pool(16)
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
