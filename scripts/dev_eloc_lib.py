"""E_loc sweep time (CUDA events) at N = 20 for the library named by FF_DEV_LIB (dev A/B builds of scripts/build_variant.sh)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
if os.environ.get("FF_DEV_LIB"):
    L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ["FF_DEV_LIB"])
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
    for _ in range(4):
        e0.record(); r = model.local_energy(x, stash=stash); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del r
    return min(ts)
r = model.local_energy(x[:4096], stash=True)
print("%s: eloc+stash %.2f ms  eloc %.2f ms  checksum eloc %.15e lap %.15e" % (os.environ.get("FF_DEV_LIB", "default"), t(True), t(False), float(r.eloc.sum()), float(r.lap.sum())), flush=True)
