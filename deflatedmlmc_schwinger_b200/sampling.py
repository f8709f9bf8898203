"""Host logic of the sampling loops (stoch_trace.py:137-154 and 386-406 of the reference),
re-organised for batches of k probes per GPU and G ranks, without changing the result:

  * probes are numbered along ONE MT19937 stream (np.random global state, seeded by the caller
    exactly where the reference seeds it); round r covers probe indices [r*G*k, (r+1)*G*k) and
    rank g owns the g-th block of k of them;
  * the reference's sequential stopping rule (`j >= 5 and err < tol`, population std) is applied
    to the ORDERED prefix of estimates, the overshoot of the last round is discarded and the
    stream is rewound to just after the last used probe, so the next level sees the same words
    as in the reference;
  * communication: sequential_stop=True needs the ordered estimates -> one all_gather of k
    complex numbers per round; sequential_stop=False fixes the sample count from the first
    round and uses a single all_reduce of [sum Re e, sum Im e, sum |e|^2, N] per level.

Nothing here touches the GPU: `sample_fn(bits01[k*n]) -> (e[k], iters[k])` is the device call.
"""
from math import sqrt

import numpy as np


class Comm:
    """Thin wrapper over torch.distributed (NCCL on GPUs, gloo in CPU tests); single process
    when torch.distributed is not initialised."""

    def __init__(self, device=None):
        self.rank, self.world = 0, 1
        self.dist = None
        self.device = device
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass

    def _dev(self):
        import torch
        if self.dist is not None and self.dist.get_backend() == "nccl":
            return self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def all_gather(self, arr):
        """arr: float64 numpy [m] -> [world*m] in rank order."""
        if self.world == 1:
            return np.asarray(arr, dtype=np.float64).copy()
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self._dev())
        out = torch.empty(self.world * t.numel(), dtype=torch.float64, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()

    def all_reduce_sum(self, arr):
        if self.world == 1:
            return np.asarray(arr, dtype=np.float64).copy()
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


def draw_probe_bits(count):
    """count 0/1 values, consuming exactly the MT19937 words np.random.randint(2,size=count) consumes."""
    raw = np.frombuffer(np.random.bytes(4 * count), dtype=np.uint8)
    return raw[0::4] & 1


def skip_probe_words(count):
    """advance the global stream by `count` words without keeping them"""
    step = 1 << 24
    while count > 0:
        c = min(step, count)
        np.random.bytes(4 * c)
        count -= c


def reference_stats(ests, j):
    """stoch_trace.py:143-147 / 394-398 verbatim"""
    ests_avg = np.sum(ests[0:(j + 1)]) / (j + 1)
    ests_dev = sqrt(np.sum(np.square(np.abs(ests[0:(j + 1)] - ests_avg))) / (j + 1))
    return ests_avg, ests_dev, ests_dev / sqrt(j + 1)


def first_stop_index(ests, start, tol, min_index=5):
    """smallest j >= start with j >= min_index and err_j < tol (reference formulas), or None.
    Candidates come from running sums; each candidate is confirmed with the exact two-pass formula."""
    n = ests.shape[0]
    if n == 0:
        return None
    idx = np.arange(1, n + 1, dtype=np.float64)
    s1 = np.cumsum(ests)
    s2 = np.cumsum(np.abs(ests) ** 2)
    var = np.maximum(s2 / idx - np.abs(s1 / idx) ** 2, 0.0)
    err = np.sqrt(var / idx)
    lo = max(start, min_index)
    # a generous band around the threshold, then exact confirmation in order
    for j in range(lo, n):
        if err[j] < tol * (1.0 + 1e-9) + 1e-300:
            if reference_stats(ests, j)[2] < tol:
                return j
    return None


def _round_bits(comm, n, k):
    """Draw this rank's k probes of the next round; the stream advances by the whole round."""
    G, g = comm.world, comm.rank
    skip_probe_words(g * k * n)
    bits = draw_probe_bits(k * n)
    skip_probe_words((G - 1 - g) * k * n)
    return bits


def run_sampling(sample_fn, n, k, tol, max_nr_ests, comm=None, fixed_count=None):
    """Sampling loop of one level with the reference's sequential stopping rule.
    Returns dict(ests, j_stop, avg, dev, iters_sum, rounds, evaluated).
    fixed_count: take exactly that many samples (no stop rule; used for the rough estimate)."""
    comm = comm or Comm()
    G = comm.world
    ests = np.zeros(0, dtype=np.complex128)
    iters_all = np.zeros(0, dtype=np.int64)
    rounds = 0
    while True:
        start_state = np.random.get_state()
        base = ests.shape[0]
        bits = _round_bits(comm, n, k)
        e_loc, it_loc = sample_fn(bits)
        rounds += 1
        payload = np.concatenate([np.real(e_loc), np.imag(e_loc), np.asarray(it_loc, dtype=np.float64)])
        allp = comm.all_gather(payload).reshape(G, 3, k)
        ests = np.concatenate([ests, (allp[:, 0, :] + 1j * allp[:, 1, :]).reshape(-1)])
        iters_all = np.concatenate([iters_all, allp[:, 2, :].reshape(-1).astype(np.int64)])
        if fixed_count is not None:
            if ests.shape[0] >= fixed_count:
                j = fixed_count - 1
                break
            continue
        j = first_stop_index(ests[:max_nr_ests], base, tol)
        if j is not None:
            break
        if ests.shape[0] >= max_nr_ests:
            j = max_nr_ests - 1
            break
    # rewind the stream to just after probe j (the reference never draws the overshoot)
    used_in_round = (j + 1) - base
    np.random.set_state(start_state)
    skip_probe_words(used_in_round * n)
    avg, dev, _ = reference_stats(ests, j)
    return {"ests": ests[:j + 1], "j_stop": j, "avg": avg, "dev": dev,
            "iters_sum": int(iters_all[:j + 1].sum()), "rounds": rounds, "evaluated": int(ests.shape[0])}


def run_sampling_fixed(sample_fn, n, k, tol, max_nr_ests, comm=None):
    """Throughput mode (sequential_stop=False): the sample count is fixed from a pilot round,
    N = max(6, ceil((sigma_pilot / tol)^2)) rounded up to whole rounds; after the pilot there is no
    communication until the single all_reduce of [sum Re e, sum Im e, sum |e|^2, N] at the end."""
    comm = comm or Comm()
    G = comm.world
    bits = _round_bits(comm, n, k)
    e_loc, it_loc = sample_fn(bits)
    mean, dev, N = reduce_level_sums(e_loc, comm)
    target = int(min(max_nr_ests, max(6, np.ceil((dev / tol) ** 2))))
    n_rounds = max(1, -(-target // (G * k)))
    es = [np.asarray(e_loc, dtype=np.complex128)]
    it_sum = int(np.sum(it_loc))
    for _ in range(n_rounds - 1):
        bits = _round_bits(comm, n, k)
        e_loc, it_loc = sample_fn(bits)
        es.append(np.asarray(e_loc, dtype=np.complex128))
        it_sum += int(np.sum(it_loc))
    mean, dev, N = reduce_level_sums(np.concatenate(es), comm)
    it_sum = int(comm.all_reduce_sum(np.array([float(it_sum)]))[0]) if n_rounds > 0 else it_sum
    return {"ests": np.concatenate(es), "j_stop": N - 1, "avg": mean, "dev": dev, "iters_sum": it_sum,
            "rounds": n_rounds, "evaluated": N}


def reduce_level_sums(e_local, comm=None):
    """The single collective of the fixed-count mode: all_reduce of [sum Re e, sum Im e, sum |e|^2, N]
    -> (mean, population std, N).  Exposed for the multi-GPU bench and tests."""
    comm = comm or Comm()
    e_local = np.asarray(e_local, dtype=np.complex128)
    s = np.array([np.real(e_local).sum(), np.imag(e_local).sum(), (np.abs(e_local) ** 2).sum(), e_local.shape[0]])
    s = comm.all_reduce_sum(s)
    N = s[3]
    mean = (s[0] + 1j * s[1]) / N
    var = max(s[2] / N - abs(mean) ** 2, 0.0)
    return mean, sqrt(var), int(N)
