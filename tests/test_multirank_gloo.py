"""CPU tier: the N > 1 path of the sampling driver with torch.distributed (gloo, world_size 2):
sharded rounds + all_gather reproduce the single-process result bit for bit, and the fixed-count
mode's single all_reduce gives the global mean / std."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _value(x):
    n = x.shape[0]
    w = np.cos(np.arange(n)) + 1j * np.sin(0.3 * np.arange(n))
    return np.dot(w, x) * (1 + 0.01 * x[0]) + 3.0


def _worker(rank, world, port, n, k, tol, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from deflatedmlmc_schwinger_b200 import sampling
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fn(bits01):
        xs = bits01.reshape(k, n).astype(np.int64) * 2 - 1
        return np.array([_value(x) for x in xs]), np.ones(k, dtype=np.int64)
    comm = sampling.Comm()
    np.random.seed(42)
    res = sampling.run_sampling(fn, n, k, tol, 5000, comm)
    after = np.random.randint(2, size=8)
    np.random.seed(43)
    resf = sampling.run_sampling_fixed(fn, n, k, tol, 5000, comm)
    m, s, N = sampling.reduce_level_sums(np.array([1.0 + rank, 2.0j]), comm)
    # set-up arrays (test vectors, deflation vectors) are rank 0's on every rank
    from deflatedmlmc_schwinger_b200 import multigrid
    mine = (np.arange(12).reshape(6, 2) * (1 + 1j) + 100 * rank).astype(np.complex128)
    shared = multigrid._same_on_all_ranks(mine)
    q.put((rank, res["j_stop"], complex(res["avg"]), res["dev"], after.tolist(), res["ests"].tolist(),
           complex(resf["avg"]), resf["dev"], resf["evaluated"], resf["ests"].tolist(), complex(m), s, N, shared.tolist()))
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process():
    import torch.multiprocessing as mp
    from deflatedmlmc_schwinger_b200 import sampling
    n, k, tol, world = 64, 8, 0.9, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, tol, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    def fn1(bits01):
        xs = bits01.reshape(k, n).astype(np.int64) * 2 - 1
        return np.array([_value(x) for x in xs]), np.ones(k, dtype=np.int64)
    np.random.seed(42)
    ref = sampling.run_sampling(fn1, n, k, tol, 5000)
    after_ref = np.random.randint(2, size=8).tolist()
    for o in out:
        assert np.array_equal(np.array(o[13]), np.arange(12).reshape(6, 2) * (1 + 1j))     # rank 0's array everywhere
        assert o[1] == ref["j_stop"]                                   # same stop index on every rank
        assert np.array_equal(np.array(o[5]), ref["ests"])             # bit-identical ordered estimates
        assert o[2] == complex(ref["avg"]) and o[3] == ref["dev"]
        assert o[4] == after_ref                                       # stream rewound identically
    # fixed-count mode: the union of the ranks' local samples is the first N probes of the stream
    N = out[0][8]
    assert N == out[1][8] and N % (world * k) == 0
    np.random.seed(43)
    allv = np.array([_value(np.random.randint(2, size=n) * 2 - 1) for _ in range(N)])
    union = np.concatenate([np.array(out[0][9]).reshape(-1, k), np.array(out[1][9]).reshape(-1, k)], axis=1)
    assert np.allclose(np.sort_complex(union.ravel()), np.sort_complex(allv))
    assert abs(out[0][6] - allv.mean()) < 1e-12 * abs(allv.mean())
    assert abs(out[0][7] - np.sqrt(np.mean(np.abs(allv - allv.mean()) ** 2))) < 1e-10
    # the single collective: [sum Re, sum Im, sum |e|^2, N]
    vals = np.array([1.0, 2.0j, 2.0, 2.0j])
    assert abs(out[0][10] - vals.mean()) < 1e-15 and out[0][12] == 4
