"""Where the device-stream e2e loses against the host-bits e2e (34.0k vs 37.9k probes/s): per-step time of
 (a) level_sample on fixed probes + D2H of the estimates, (b) + probe_expand_bytes from a fixed byte buffer,
 (c) + one generator call per step whose output is not used, (d) the drivers' DeviceProbeSource loop."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
import bench
from deflatedmlmc_schwinger_b200 import sampling
mg, tp, A, _ = bench.build_solver()
dev = mg.dev
n0, k = mg.level_shapes[0], 512
tol, restart, maxiter = 1e-12, 40, 1000
dev.ensure_workspace(0, k, restart)
stream = torch.cuda.current_stream()
np.random.seed(123456)
st = np.random.get_state()
words = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
state = torch.from_numpy(words.view(np.int32).copy()).cuda()
lsb = dev.mt19937_bits(state, 0, k * n0, 0)
dev.rng_sync()
X0 = dev.probe_expand_bytes(lsb, n0, k)
out2 = torch.empty(k * n0, dtype=torch.uint8, device="cuda")


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def a():
    dev.level_sample(1, 0, 2, X0, tol, restart, maxiter)[0].cpu()


def b():
    X = dev.probe_expand_bytes(lsb, n0, k)
    dev.level_sample(1, 0, 2, X, tol, restart, maxiter)[0].cpu()


def c():
    X = dev.probe_expand_bytes(lsb, n0, k)
    dev.mt19937_bits(state, 0, k * n0, 0, out=out2)
    dev.level_sample(1, 0, 2, X, tol, restart, maxiter)[0].cpu()


src = sampling.DeviceProbeSource(dev)
comm = sampling.Comm(dev.device)
src.begin()


def d():
    dev.level_sample(1, 0, 2, src.next_round(comm, n0, k), tol, restart, maxiter)[0].cpu()


def e():          # generator call with nothing to generate (only the end-state CTA)
    X = dev.probe_expand_bytes(lsb, n0, k)
    dev.mt19937_bits(state, 0, 0, k * n0)
    dev.level_sample(1, 0, 2, X, tol, restart, maxiter)[0].cpu()


def g():          # generator forced to finish before the solve is launched: its full duration exposed
    X = dev.probe_expand_bytes(lsb, n0, k)
    dev.mt19937_bits(state, 0, k * n0, 0, out=out2)
    dev.rng_sync()
    dev.level_sample(1, 0, 2, X, tol, restart, maxiter)[0].cpu()


side = torch.cuda.Stream()
tiny = torch.zeros(1024, device="cuda")


def h():          # any small kernel on a side stream per step
    X = dev.probe_expand_bytes(lsb, n0, k)
    with torch.cuda.stream(side):
        tiny.add_(1.0)
    dev.level_sample(1, 0, 2, X, tol, restart, maxiter)[0].cpu()


def host_call_ms():
    torch.cuda.synchronize()
    t = time.time()
    dev.mt19937_bits(state, 0, k * n0, 0, out=out2)
    dt = 1e3 * (time.time() - t)
    dev.rng_sync()
    return dt


res = {"host_side_call_ms": [host_call_ms() for _ in range(3)], "h_tiny_side_stream_kernel_ms": timed(h), "e_generator_call_without_outputs_ms": timed(e),
       "g_generator_then_sync_ms": timed(g), "a_fixed_probes_ms": timed(a), "b_plus_expand_ms": timed(b), "c_plus_generator_ms": timed(c), "d_device_probe_source_ms": timed(d)}
src.end()
dev.set_option("mt_prio", 1)
res["c_high_priority_ms"] = timed(c)
src.begin()
res["d_high_priority_ms"] = timed(d)
src.end()
dev.set_option("mt_prio", 0)
mg.set_option("use_graphs", 0)
res["a_no_graphs_ms"] = timed(a)
res["c_no_graphs_ms"] = timed(c)
mg.set_option("use_graphs", 1)
for prio in (1,):
    dev.set_option("mt_jump", 0)
    res["c_sequential_generator_ms"] = timed(c)
    dev.set_option("mt_jump", 1)
print(json.dumps(res))
