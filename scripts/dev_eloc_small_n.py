"""E_loc sweep at the small BASELINE configs (N = 6, 12): fused static kernel (default) vs the generic kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fermiflow_b200 import MLP, Backflow, CNF, HO2D, FreeFermion, GSVMC, HO, CoulombPairPotential
dev = torch.device("cuda:0")
for nup, ndn in ((3, 3), (6, 6)):
    g = torch.Generator().manual_seed(1)
    eta, mu = MLP(1, 50), MLP(1, 50)
    with torch.no_grad():
        for m in (eta, mu):
            m.fc1.weight.copy_(torch.randn(50, 1, generator=g)); m.fc1.bias.copy_(torch.randn(50, generator=g)); m.fc2.weight.copy_(1e-2 * torch.randn(1, 50, generator=g))
    cnf = CNF(Backflow(eta, mu=mu), (0.0, 1.0), nsteps=16)
    model = GSVMC(nup, ndn, HO2D(), FreeFermion(dev), cnf, CoulombPairPotential(2.0), sp_potential=HO()).to(dev)
    B = 262144
    _, x = model.sample((B,))
    out = {}
    for tag, env in (("fused static", {}), ("generic", {"FF_ELOC_V1": "1", "FF_NO_STATIC": "1"})):
        for k in ("FF_ELOC_V1", "FF_NO_STATIC"): os.environ.pop(k, None)
        os.environ.update(env)
        r = model.local_energy(x); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = model.local_energy(x); e1.record(); torch.cuda.synchronize()
        out[tag] = (e0.elapsed_time(e1), r.eloc.clone())
    d = float((out["fused static"][1] - out["generic"][1]).abs().max() / out["generic"][1].abs().max())
    print("N=%d, %d walkers: fused %.2f ms, generic %.2f ms, max rel diff of E_loc %.2e" % (nup + ndn, B, out["fused static"][0], out["generic"][0], d))
