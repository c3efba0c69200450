"""Run-to-run reproducibility of every stage of a small VMC iteration (fresh models, same seeds)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
def run(nup, ndn, B, hidden=10, steps=2):
    args = argparse.Namespace(hidden=hidden, ode_steps=steps, nup=nup, ndown=ndn, Z=2.0)
    model = bench.build_model(args, dev)
    z = model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,))
    x = model.cnf.generate(z)
    r = model.local_energy(x, stash=True)
    return dict(z=z, x=x, logp=r.logp, grad=r.grad, lap=r.lap, eloc=r.eloc)
for nup, ndn, B in ((10, 10, 5), (7, 6, 3), (3, 3, 4), (10, 10, 300), (3, 3, 64)):
    a = run(nup, ndn, B)
    junk = torch.randn(1 << 24, device=dev)
    b = run(nup, ndn, B)
    bad = [k for k in a if not torch.equal(a[k], b[k])]
    print("N = %d B = %d: differing outputs %s; checksums z %.12e x %.12e eloc %.12e" % (
        nup + ndn, B, bad, float(a["z"].sum()), float(a["x"].sum()), float(a["eloc"].sum())), flush=True)
