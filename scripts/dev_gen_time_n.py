"""generate / delta_logp kernel times at a given particle number (CUDA events): dev_gen_time_n.py nup ndown B nsteps"""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
nup, ndn, B, ns = (int(v) for v in sys.argv[1:5])
args = argparse.Namespace(hidden=50, ode_steps=ns, nup=nup, ndown=ndn, Z=2.0)
model = bench.build_model(args, dev)
z = model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,))
def t(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
x = model.cnf.generate(z)
print("n = %d + %d, %d walkers, %d steps: generate %.2f ms   delta_logp %.2f ms   eloc %.2f ms" % (
    nup, ndn, B, ns, t(lambda: model.cnf.generate(z)), t(lambda: model.cnf.delta_logp(x)), t(lambda: model.local_energy(x))))
