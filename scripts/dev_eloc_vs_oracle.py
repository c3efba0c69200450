"""Which E_loc sweep is closer to the oracle at ill-conditioned random coordinates (N = 16, x = 1.2 randn)?"""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fermiflow_b200 import _lib
from oracle import fermiflow_oracle as O
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
for nup, ndn, S in ((8, 8, 4), (7, 7, 4), (10, 10, 4)):
    args = argparse.Namespace(hidden=12, ode_steps=S, nup=nup, ndown=ndn, Z=2.0)
    model = bench.build_model(args, dev)
    torch.manual_seed(0)
    x = 1.2 * torch.randn(6, nup + ndn, 2, device=dev)
    eta, mu = model.cnf.v.eta, model.cnf.v.mu
    pe = tuple(t.cpu() for t in eta.kernel_params()); pm = tuple(t.cpu() for t in mu.kernel_params())
    ref = O.local_energy(x.cpu(), list(range(nup)), list(range(ndn)), pe, pm, (0.0, 1.0), S, 2.0)
    for name, kw in (("eloc5 + warp finale", {}), ("eloc5 + CTA finale", dict(finale_cta=1)), ("eloc2", dict(eloc_v2=1)), ("generic", dict(eloc_generic=1))):
        with _lib.options(**kw):
            r = model.local_energy(x)
        errs = {k: float((getattr(r, k).cpu() - ref[k]).abs().max() / ref[k].abs().max()) for k in ("logp", "grad", "lap", "eloc")}
        print("N = %d  %-22s" % (nup + ndn, name), "  ".join("%s %.1e" % kv for kv in errs.items()), flush=True)
