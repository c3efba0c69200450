"""Time individual stages (CUDA events) at a given hidden size."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
hidden = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
args = argparse.Namespace(hidden=hidden, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
z = model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,))
print("hidden", hidden, "walkers", B)
print("  metropolis  %.1f ms" % timeit(lambda: model.basedist.sample(model.orbitals_up, model.orbitals_down, (B,))))
print("  generate    %.1f ms" % timeit(lambda: model.cnf.generate(z)))
x = model.cnf.generate(z)
print("  delta_logp  %.1f ms" % timeit(lambda: model.cnf.delta_logp(x)))
print("  eloc        %.1f ms" % timeit(lambda: model.local_energy(x)))
print("  eloc+stash  %.1f ms" % timeit(lambda: model.local_energy(x, stash=True)))
def full():
    g = model(B); g.backward()
print("  model+bwd   %.1f ms" % timeit(full))
import time
def fwd():
    global g
    g = model(B)
def bwd():
    g.backward()
fwd(); bwd(); torch.cuda.synchronize()
for name, f in (("forward model(B)", fwd), ("backward", bwd)):
    torch.cuda.synchronize(); t = time.time(); f(); torch.cuda.synchronize(); print("  %-18s %.1f ms (wall)" % (name, 1e3 * (time.time() - t)))
# pieces of forward
torch.cuda.synchronize(); t = time.time(); zz, xx = model.sample((B,)); torch.cuda.synchronize(); print("  sample             %.1f ms" % (1e3 * (time.time() - t)))
t = time.time(); res = model.local_energy(xx, stash=True); torch.cuda.synchronize(); print("  local_energy+stash %.1f ms" % (1e3 * (time.time() - t)))
from fermiflow_b200.VMC import _global_mean_std
t = time.time(); m = _global_mean_std(res.eloc); torch.cuda.synchronize(); print("  moments            %.1f ms" % (1e3 * (time.time() - t)))
# pieces of backward
from fermiflow_b200.base_dist import _FreeFermionLogp
from fermiflow_b200.flow import _backward_through_flow
zr = res.z.detach().requires_grad_(True)
t = time.time(); lp0 = _FreeFermionLogp.apply(zr, model._orb(dev), None, 10, 10); g0, = torch.autograd.grad(lp0, zr, grad_outputs=torch.ones(B, device=dev)); torch.cuda.synchronize(); print("  slater grad        %.1f ms" % (1e3 * (time.time() - t)))
t = time.time(); _backward_through_flow(model.cnf, res.model, res.stash, B, g0, -torch.ones(B, device=dev), False); torch.cuda.synchronize(); print("  logp_backward      %.1f ms" % (1e3 * (time.time() - t)))
