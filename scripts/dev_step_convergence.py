"""Discretisation error of the fixed-step RK4 sweep vs number of steps (N=20 bench model)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
for scale_tag, mult in (("bench weights (w2 ~ 1e-2)", 1.0), ("10x stronger flow (w2 ~ 1e-1)", 10.0)):
    args = argparse.Namespace(hidden=50, ode_steps=128, nup=10, ndown=10, Z=2.0)
    model = bench.build_model(args, dev)
    with torch.no_grad():
        model.cnf.v.eta.fc2.weight.mul_(mult); model.cnf.v.mu.fc2.weight.mul_(mult)
    model.basedist.manual_seed(3)
    _, x = model.sample((2048,))
    ref = model.local_energy(x)
    print(scale_tag, " mean |x - z| =", float((x - ref.z).abs().mean()))
    for S in (4, 8, 12, 16, 32, 64):
        model.cnf.nsteps = S
        r = model.local_energy(x)
        e = lambda a, b: (float(((a - b).abs() / b.abs().clamp_min(1e-300)).median()), float(((a - b).abs().max() / b.abs().max())))
        print("  S=%3d  rel err median/max:  logp %.1e/%.1e  grad %.1e  lap %.1e/%.1e  eloc %.1e/%.1e" % (
            S, *e(r.logp, ref.logp), e(r.grad, ref.grad)[1], *e(r.lap, ref.lap), *e(r.eloc, ref.eloc)))
    model.cnf.nsteps = 128
