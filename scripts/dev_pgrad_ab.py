"""Parameter gradient: binned Taylor moments (default) against the direct kernel (FF_NO_BINNED_PGRAD=1), N = 20."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
model.basedist.manual_seed(5)
def grads():
    model.basedist.manual_seed(5)
    for p in model.parameters(): p.grad = None
    g = model(B); g.backward(); torch.cuda.synchronize()
    return [p.grad.clone() for p in model.parameters()]
os.environ["FF_NO_BINNED_PGRAD"] = "1"; ref = grads()
os.environ.pop("FF_NO_BINNED_PGRAD"); got = grads()
for (nm, _), a, b in zip(model.named_parameters(), ref, got):
    print("  %-22s max rel diff %.3e   (|g| max %.3e)" % (nm, float((a - b).abs().max() / a.abs().max()), float(a.abs().max())))
def t(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        g = model(B); torch.cuda.synchronize()
        e0.record(); g.backward(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
print("walkers %d: backward %.2f ms (binned)" % (B, t(lambda: None)))
os.environ["FF_NO_BINNED_PGRAD"] = "1"
print("walkers %d: backward %.2f ms (direct)" % (B, t(lambda: None)))
