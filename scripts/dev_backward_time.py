"""CUDA-event timing of the backward pieces (adjoint sweep alone / with parameter gradients) at the bench shape."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
from fermiflow_b200.flow import _backward_through_flow
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
zz, xx = model.sample((B,))
res = model.local_energy(xx, stash=True)
g0 = torch.randn(B, 20, 2, device=dev)
gd = -torch.ones(B, device=dev)
print("logp_backward (adjoint + parameter gradient) %.2f ms" % timeit(lambda: _backward_through_flow(model.cnf, res.model, res.stash, B, g0, gd, False)))
print("full step fwd+bwd %.2f ms" % timeit(lambda: model(B).backward()))
