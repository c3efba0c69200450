"""Backward pass (adjoint sweep + parameter gradient) time at N = 20 with and without the bulk L2 prefetch of the stash."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fermiflow_b200 import _lib
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
def t():
    ts = []
    for _ in range(4):
        g = model(B); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
for rep in range(2):
    for np_ in (0, 1):
        with _lib.options(adjoint_no_prefetch=np_):
            print("walkers %d, adjoint_no_prefetch=%d: backward %.2f ms" % (B, np_, t()), flush=True)
