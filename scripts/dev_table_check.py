"""Radial Taylor tables: A/B of every sweep with tables (default) and without (FF_NO_TABLE=1), N = 20."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
z = model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,))
def run():
    x = model.cnf.generate(z)
    zz, dl = model.cnf.delta_logp(x)
    r = model.local_energy(x, stash=True)
    return dict(x=x, z=zz, dl=dl, logp=r.logp, grad=r.grad, lap=r.lap, eloc=r.eloc, stash_c=r.stash.c)
os.environ["FF_NO_TABLE"] = "1"; ref = run()
os.environ.pop("FF_NO_TABLE"); got = run()
for k in ref:
    print("  %-8s max rel diff %.3e" % (k, float((ref[k] - got[k]).abs().max() / ref[k].abs().max())))
def t(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
x = got["x"]
print("walkers %d: generate %.2f ms, delta_logp %.2f ms, eloc %.2f ms" % (B, t(lambda: model.cnf.generate(z)), t(lambda: model.cnf.delta_logp(x)), t(lambda: model.local_energy(x))))
