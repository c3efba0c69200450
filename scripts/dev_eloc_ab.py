"""E_loc sweep time (CUDA events) at N = 20: default (eloc4) against option eloc_v2, and agreement of the two."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fermiflow_b200 import _lib
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
res = {}
variants = [("eloc5", {}), ("eloc4", dict(eloc_v4=1)), ("eloc2", dict(eloc_v2=1))]
for rep in range(2):
    for v2, (name, kw) in enumerate(variants):
        with _lib.options(**kw):
            print("%s: eloc %.2f ms   eloc+stash %.2f ms   eloc %.2f ms" % (name, t(False), t(True), t(False)), flush=True)
            res[v2] = model.local_energy(x[:4096], stash=True)
for k in ("z", "logp", "grad", "lap", "kinetic", "potential", "eloc"):
    a, b = getattr(res[0], k), getattr(res[2], k)
    print(k, float((a - b).abs().max() / b.abs().max()))
print("stash y", float((res[0].stash.y - res[2].stash.y).abs().max()), "c", float((res[0].stash.c - res[2].stash.c).abs().max()))
