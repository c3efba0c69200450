"""Times of the stages of one VMC iteration at N = 20 (CUDA events)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L0
if os.environ.get("FF_DEV_LIB"): L0.LIB_PATH = os.path.join(os.path.dirname(L0.LIB_PATH), os.environ["FF_DEV_LIB"])
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
def ev(f, n=4):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), r
t_m, z = ev(lambda: model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,)))
t_g, x = ev(lambda: model.cnf.generate(z))
t_d, _ = ev(lambda: model.cnf.delta_logp(x))
t_e, _ = ev(lambda: model.local_energy(x, stash=True))
ts = []
for _ in range(4):
    g = model(B); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("walkers %d: metropolis %.2f  generate %.2f  delta_logp %.2f  eloc+stash %.2f  backward %.2f ms" % (B, t_m, t_g, t_d, t_e, min(ts[1:])))
