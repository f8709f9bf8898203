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
    for j in lo + np.nonzero(err[lo:] < tol * (1.0 + 1e-9) + 1e-300)[0]:
        if reference_stats(ests, int(j))[2] < tol:
            return int(j)
    return None


def _round_bits(comm, n, k):
    """Draw this rank's k probes of the next round; the stream advances by the whole round."""
    G, g = comm.world, comm.rank
    skip_probe_words(g * k * n)
    bits = draw_probe_bits(k * n)
    skip_probe_words((G - 1 - g) * k * n)
    return bits


class HostProbeSource:
    """Probes drawn on the host from the global numpy stream (one MT19937 word per element)."""

    def begin(self):
        pass

    def next_round(self, comm, n, k):
        """this rank's k probes of the next round as k*n 0/1 values; the stream advances by the whole round"""
        self._start_state = np.random.get_state()
        return _round_bits(comm, n, k)

    def rewind(self, used_in_round, n):
        """leave the stream just after the `used_in_round`-th probe of the last round"""
        np.random.set_state(self._start_state)
        skip_probe_words(used_in_round * n)

    def end(self):
        pass


class DeviceProbeSource:
    """The same stream advanced ON THE GPU (dmlmc_mt19937_bits): bit-identical to np.random, no host RNG work and
    no host->device copy of probes.  begin() uploads the numpy generator's state, end() writes the final state
    back, so the global numpy RNG -- part of the reference's interface -- ends where the reference leaves it.
    The generator of round r+1 runs on a side stream while round r is being solved."""

    def __init__(self, dev):
        self.dev = dev
        self.torch = dev.torch
        self.state = None
        self.prefetched = None
        self.last_backup = None

    def begin(self):
        st = np.random.get_state()
        if st[0] != 'MT19937':
            raise Exception("the probe stream is numpy's legacy MT19937 generator")
        words = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
        self._gauss = (st[3], st[4])
        self.state = self.torch.from_numpy(words.view(np.int32).copy()).to(self.dev.device)
        self.prefetched = None
        self.last_backup = None

    def _issue(self, comm, n, k):
        G, g = comm.world, comm.rank
        backup = self.torch.empty(625, dtype=self.torch.int32, device=self.dev.device)
        lsb = self.dev.mt19937_bits(self.state, g * k * n, k * n, (G - 1 - g) * k * n, backup=backup)
        self.prefetched = (lsb, backup, n, k)

    def next_round(self, comm, n, k):
        if self.prefetched is None or self.prefetched[2:] != (n, k):
            if self.prefetched is not None:        # shape changed: discard the prefetch, go back to its start
                self._restore(self.prefetched[1])
            self._issue(comm, n, k)
        lsb, backup, _, _ = self.prefetched
        self.prefetched = None
        X0 = self.dev.probe_expand_bytes(lsb, n, k)
        self.last_backup = backup
        self._keep = lsb                            # alive until the expand kernel has run
        self._issue(comm, n, k)                     # next round, beside the solve of this one
        return X0

    def _restore(self, backup):
        self.dev.rng_sync()
        self.torch.cuda.current_stream(self.dev.device).synchronize()
        self.state.copy_(backup)

    def rewind(self, used_in_round, n):
        self._restore(self.last_backup)
        self.prefetched = None
        if used_in_round > 0:
            self.dev.mt19937_bits(self.state, used_in_round * n, 0, 0)

    def end(self):
        if self.prefetched is not None:
            self._restore(self.prefetched[1])
            self.prefetched = None
        self.dev.rng_sync()
        self.torch.cuda.current_stream(self.dev.device).synchronize()
        words = self.state.cpu().numpy().view(np.uint32)
        np.random.set_state(('MT19937', words[:624].copy(), int(words[624]), self._gauss[0], self._gauss[1]))


def run_sampling(sample_fn, n, k, tol, max_nr_ests, comm=None, fixed_count=None, probe_source=None):
    """Sampling loop of one level with the reference's sequential stopping rule.
    Returns dict(ests, j_stop, avg, dev, iters_sum, rounds, evaluated).
    fixed_count: take exactly that many samples (no stop rule; used for the rough estimate).
    probe_source: HostProbeSource (default; sample_fn gets k*n 0/1 values) or DeviceProbeSource (sample_fn gets
    the probes as a complex128 CUDA tensor [n, k])."""
    comm = comm or Comm()
    src = probe_source or HostProbeSource()
    G = comm.world
    ests = np.zeros(0, dtype=np.complex128)
    iters_all = np.zeros(0, dtype=np.int64)
    coarse_all = np.zeros(0, dtype=np.int64)
    rounds = 0
    src.begin()
    while True:
        base = ests.shape[0]
        probes = src.next_round(comm, n, k)
        out = sample_fn(probes)                    # (e[k], fine-level iterations[k] [, coarse-level iterations[k]])
        e_loc, it_loc = out[0], out[1]
        itc_loc = out[2] if len(out) > 2 else np.zeros(k)
        rounds += 1
        payload = np.concatenate([np.real(e_loc), np.imag(e_loc), np.asarray(it_loc, dtype=np.float64),
                                  np.asarray(itc_loc, dtype=np.float64)])
        allp = comm.all_gather(payload).reshape(G, 4, k)
        ests = np.concatenate([ests, (allp[:, 0, :] + 1j * allp[:, 1, :]).reshape(-1)])
        iters_all = np.concatenate([iters_all, allp[:, 2, :].reshape(-1).astype(np.int64)])
        coarse_all = np.concatenate([coarse_all, allp[:, 3, :].reshape(-1).astype(np.int64)])
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
    src.rewind((j + 1) - base, n)
    src.end()
    avg, dev, _ = reference_stats(ests, j)
    # iteration counts of the USED samples only, the same on every rank (the overshoot of the last round is not counted)
    return {"ests": ests[:j + 1], "j_stop": j, "avg": avg, "dev": dev,
            "iters_sum": int(iters_all[:j + 1].sum()), "coarse_iters_sum": int(coarse_all[:j + 1].sum()),
            "rounds": rounds, "evaluated": int(ests.shape[0])}


def run_sampling_fixed(sample_fn, n, k, tol, max_nr_ests, comm=None, probe_source=None):
    """Throughput mode (sequential_stop=False): the sample count is fixed from a pilot round,
    N = max(6, ceil((sigma_pilot / tol)^2)) rounded up to whole rounds; after the pilot there is no
    communication until the single all_reduce of [sum Re e, sum Im e, sum |e|^2, N] at the end."""
    comm = comm or Comm()
    src = probe_source or HostProbeSource()
    G = comm.world
    src.begin()
    out = sample_fn(src.next_round(comm, n, k))
    e_loc = out[0]
    mean, dev, N = reduce_level_sums(e_loc, comm)
    target = int(min(max_nr_ests, max(6, np.ceil((dev / tol) ** 2))))
    n_rounds = max(1, -(-target // (G * k)))
    es = [np.asarray(e_loc, dtype=np.complex128)]
    it_sum = int(np.sum(out[1]))
    itc_sum = int(np.sum(out[2])) if len(out) > 2 else 0
    for _ in range(n_rounds - 1):
        out = sample_fn(src.next_round(comm, n, k))
        es.append(np.asarray(out[0], dtype=np.complex128))
        it_sum += int(np.sum(out[1]))
        itc_sum += int(np.sum(out[2])) if len(out) > 2 else 0
    src.end()
    mean, dev, N = reduce_level_sums(np.concatenate(es), comm)
    sums = comm.all_reduce_sum(np.array([float(it_sum), float(itc_sum)]))
    return {"ests": np.concatenate(es), "j_stop": N - 1, "avg": mean, "dev": dev, "iters_sum": int(sums[0]),
            "coarse_iters_sum": int(sums[1]), "rounds": n_rounds, "evaluated": N}


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
