"""A/B: second-generation E_loc sweep (ff_eloc2.cuh) against the generic flow_kernel<MODE_ELOC>
(oracle-validated) on the same walkers; prints max relative differences and timings."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
_, x = model.sample((B,))
def run(v1):
    var = os.environ.get("FF_AB_VARIANT", "FF_ELOC_V3")
    if v1: os.environ.pop(var, None); os.environ["FF_NO_STATIC"] = "1"
    else: os.environ[var] = "1"; os.environ.pop("FF_NO_STATIC", None)
    r = model.local_energy(x, stash=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = model.local_energy(x, stash=True); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
ra, ta = run(True)
rb, tb = run(False)
print("walkers %d: generic %.2f ms, variant %.2f ms" % (B, ta, tb))
for k in ("z", "logp", "grad", "lap", "kinetic", "potential", "eloc"):
    a, b = getattr(ra, k), getattr(rb, k)
    print("  %-10s max rel diff %.3e" % (k, float((a - b).abs().max() / a.abs().max())))
for k, (a, b) in enumerate(((ra.stash.y, rb.stash.y), (ra.stash.c, rb.stash.c))):
    print("  stash[%d]   max rel diff %.3e" % (k, float((a - b).abs().max() / a.abs().max())))
