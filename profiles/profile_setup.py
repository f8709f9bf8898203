"""Where the hierarchy set-up time goes (cProfile, cumulative) for the bench's solver: python profiles/profile_setup.py [L]
L = 128: the shipped configuration with the golden test vectors injected."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
import torch
import __graft_entry__ as ge
ge.build()
import bench
bench.build_solver()           # warm: CUDA context, library load, torch kernels
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
mg, tp, A, _ = bench.build_solver()
torch.cuda.synchronize()
pr.disable()
print("setup wall %.3f s" % (time.time() - t0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
