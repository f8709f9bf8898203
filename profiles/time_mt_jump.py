"""Timing of the probe-stream generator (dmlmc_mt19937_bits) for the block of rank g of G in a round of G * k * n words,
k = 512 probes of n = 32768 elements: jump-ahead kernel (default) against the sequential one-CTA kernel (option mt_jump = 0).
One JSON line per (G, g, mode).  Timed with host clocks around rng_sync (the kernel runs on the library's side stream)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from deflatedmlmc_schwinger_b200 import _lib
dev = _lib.Hierarchy(1)
np.random.seed(123456)
st = np.random.get_state()
words = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
n, k = 32768, 512
out = torch.empty(n * k, dtype=torch.uint8, device="cuda")
for G, g in ((1, 0), (2, 1), (8, 0), (8, 7)):
    for mode in (1, 0):
        dev.set_option("mt_jump", mode)
        ts = []
        for rep in range(3):
            state = torch.from_numpy(words.view(np.int32).copy()).cuda()
            torch.cuda.synchronize()
            t = time.time()
            dev.mt19937_bits(state, g * k * n, k * n, (G - 1 - g) * k * n, out=out)
            dev.rng_sync()
            ts.append(time.time() - t)
        print(json.dumps({"G": G, "g": g, "mode": "jump-ahead" if mode else "sequential, 1 CTA", "words_kept": n * k,
                          "words_skipped": (G - 1) * n * k, "ms": round(1e3 * min(ts), 3), "checksum": int(out.sum().item())}), flush=True)
dev.set_option("mt_jump", 1)
