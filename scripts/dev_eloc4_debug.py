"""eloc4 against eloc2 at small sizes (debugging aid)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fermiflow_b200 import _lib
if os.environ.get("FF_DEV_LIB"): _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ["FF_DEV_LIB"])
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
nup = int(sys.argv[1]); B = int(sys.argv[2]); S = int(sys.argv[3])
args = argparse.Namespace(hidden=50, ode_steps=S, nup=nup, ndown=nup, Z=2.0)
model = bench.build_model(args, dev)
torch.manual_seed(1)
x = 1.5 * torch.randn(B, 2 * nup, 2, device=dev)
res = {}
for v2 in (1, 0):
    with _lib.options(eloc_v2=v2):
        res[v2] = model.local_energy(x, stash=True)
        torch.cuda.synchronize()
        print("ran eloc_v2 =", v2, flush=True)
for k in ("z", "logp", "grad", "lap", "kinetic", "potential", "eloc"):
    a, b = getattr(res[0], k), getattr(res[1], k)
    print(k, float((a - b).abs().max() / b.abs().max()))
print("stash y", float((res[0].stash.y - res[1].stash.y).abs().max()), "c", float((res[0].stash.c - res[1].stash.c).abs().max()))
