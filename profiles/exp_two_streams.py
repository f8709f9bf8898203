"""Experiment: do two INDEPENDENT half-batches on two CUDA streams overlap usefully?  A 512-probe step is a serial chain of
kernels with different limiters (the even-odd sweeps: issue / L2 latency at 36 % of the warp slots; the tcgen05 coarse solve:
tensor pipe, HBM idle; Gram-Schmidt: HBM).  Two hierarchies (own stream, own work space, own copy of the operators), each
sampling k probes from its own host thread (the C ABI blocks in its convergence polls; ctypes releases the GIL), against one
hierarchy sampling 2k.

    python profiles/exp_two_streams.py [--k 256] [--steps 12]
"""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=256)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
import bench

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)


def make(stream):
    with torch.cuda.stream(stream):
        mg, tp, A, _ = bench.build_solver()
    return mg, tp


def probes(n, k, seed):
    rs = np.random.RandomState(seed)
    return torch.from_numpy((2.0 * rs.randint(2, size=(n, k)) - 1.0).astype(np.complex128)).to(dev)


def run(mgs, streams, k, steps, warmup):
    """every hierarchy samples `steps` batches of k probes on its stream from its own thread; returns probes/s overall"""
    n0 = mgs[0][0].level_shapes[0]
    X = [[probes(n0, k, 17 * i + s) for s in range(2)] for i in range(len(mgs))]
    out = [None] * len(mgs)

    def work(i, count):
        mg, tp = mgs[i]
        with torch.cuda.stream(streams[i]):
            for s in range(count):
                e, _ = mg.dev.level_sample(1, 0, 2, X[i][s & 1], 1e-12, 40, 1000)
            out[i] = e
            streams[i].synchronize()

    def all_threads(count):
        th = [threading.Thread(target=work, args=(i, count)) for i in range(len(mgs))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    all_threads(warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    all_threads(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return len(mgs) * steps * k / dt, dt / steps * 1e3, [o.cpu().numpy() for o in out]


s0 = torch.cuda.Stream()
s1 = torch.cuda.Stream()
a = make(s0)
b = make(s1)
res = {}
v1, ms1, e_single = run([a], [s0], 2 * args.k, args.steps, args.warmup)
res["one_stream_%d" % (2 * args.k)] = {"probes_per_s": v1, "ms_per_step": ms1}
vh, msh, _ = run([a], [s0], args.k, args.steps, args.warmup)
res["one_stream_%d" % args.k] = {"probes_per_s": vh, "ms_per_step": msh}
v2, ms2, e_two = run([a, b], [s0, s1], args.k, args.steps, args.warmup)
res["two_streams_%d_each" % args.k] = {"probes_per_s": v2, "ms_per_step": ms2}
# same probes -> same estimates, whichever hierarchy / stream computed them (batch-independent reductions)
ea = run([a], [s0], args.k, 1, 0)[2][0]
eb = run([b], [s1], args.k, 1, 0)[2][0]
res["estimates_equal_across_hierarchies"] = bool(np.array_equal(ea, eb))
res["max_abs_diff"] = float(np.abs(ea - eb).max())
print(json.dumps(res))
