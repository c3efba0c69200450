"""Backward time (adjoint sweep + parameter gradient) at N = 20 for pgrad_tile values (and FF_DEV_LIB builds)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
if os.environ.get("FF_DEV_LIB"):
    L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ["FF_DEV_LIB"])
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
stages = [int(v) for v in sys.argv[2:]] or [0]
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
def t():
    ts = []
    for _ in range(4):
        for p in model.parameters(): p.grad = None
        g = model(B); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
for rs in stages:
    with L.options(pgrad_tile=rs):
        tt = t()
        print("%s walkers %d pgrad_tile %d: backward %.2f ms" % (os.environ.get("FF_DEV_LIB", "default"), B, rs, tt), flush=True)
