"""Time the three Metropolis samplers (register / warp / thread) at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fermiflow_b200 import HO2D, FreeFermion
dev = torch.device("cuda:0")
ho = HO2D()
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nup, B in ((10, 65536), (10, 16384), (10, 8192), (6, 65536), (3, 65536), (3, 8192)):
    row = []
    for name, env in (("reg", "FF_METRO_REG"), ("warp", "FF_METRO_WARP"), ("thread", "FF_METRO_THREAD")):
        for k in ("FF_METRO_REG", "FF_METRO_WARP", "FF_METRO_THREAD"):
            os.environ.pop(k, None)
        os.environ[env] = "1"
        ff = FreeFermion(dev)
        row.append("%s %.2f ms" % (name, timeit(lambda: ff.sample(ho.orbitals[:nup], ho.orbitals[:nup], (B,)))))
    print("n = %d + %d, %d walkers, 100 moves: " % (nup, nup, B) + " | ".join(row), flush=True)
