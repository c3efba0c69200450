"""E_loc sweep time (CUDA events, preallocated stash) at N = 20."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
_, x = model.sample((B,))
def t(stash):
    r = model.local_energy(x, stash=stash); del r; torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); r = model.local_energy(x, stash=stash); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del r
    return min(ts)
print("eloc %.1f ms   eloc+stash %.1f ms" % (t(False), t(True)))
