"""E_loc sweep time by particle number: register-resident sweep (default from N = 10 on) against the generic kernel."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fermiflow_b200 import _lib
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = 32768
for nup, ndn in ((3, 3), (4, 3), (4, 4), (5, 4), (5, 5), (6, 6), (7, 7), (8, 8), (10, 10)):
    args = argparse.Namespace(hidden=50, ode_steps=16, nup=nup, ndown=ndn, Z=2.0)
    model = bench.build_model(args, dev)
    torch.manual_seed(0)
    x = 1.2 * torch.randn(B, nup + ndn, 2, device=dev)
    out = []
    res = []
    for gen in (0, 1):
        with _lib.options(eloc_generic=gen):
            r = model.local_energy(x, stash=True); res.append(r.eloc.clone()); del r; torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = model.local_energy(x, stash=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del r
            out.append(min(ts))
    print("N = %2d (%d + %d), %d walkers: default %.2f ms, generic %.2f ms, max rel diff of E_loc %.1e" % (
        nup + ndn, nup, ndn, B, out[0], out[1], float((res[0] - res[1]).abs().max() / res[1].abs().max())), flush=True)
