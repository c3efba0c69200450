"""Small runs of the hot path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
for nup, ndn, B in ((10, 10, 5), (7, 6, 3), (3, 3, 4)):
    args = argparse.Namespace(hidden=10, ode_steps=2, nup=nup, ndown=ndn, Z=2.0)
    model = bench.build_model(args, dev)
    g = model(B); g.backward(); torch.cuda.synchronize()
    print("N =", nup + ndn, "E =", model.E, flush=True)
