"""Profiling driver: a few VMC iterations at a given walker count (used under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench
p = argparse.ArgumentParser()
p.add_argument("--walkers", type=int, default=8192)
p.add_argument("--iters", type=int, default=2)
p.add_argument("--ode-steps", type=int, default=16)
p.add_argument("--nup", type=int, default=10)
p.add_argument("--ndown", type=int, default=10)
a = p.parse_args()
args = argparse.Namespace(hidden=int(os.environ.get("HID", "50")), ode_steps=a.ode_steps, nup=a.nup, ndown=a.ndown, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
for i in range(a.iters):
    g = model(a.walkers)
    opt.zero_grad(); g.backward(); opt.step()
    torch.cuda.synchronize()
    print("iter", i, "E", model.E, "+-", model.E_std)
