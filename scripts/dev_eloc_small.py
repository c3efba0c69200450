"""Kernel time of the E_loc sweep for 1 / 2 / 4 walkers per SM: cycles per RK stage of a lone
CTA versus two co-resident CTAs (production build, CUDA events)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
model = bench.build_model(args, dev)
_, xall = model.sample((148 * 8,))
for per_sm in (1, 2, 4, 8):
    x = xall[: 148 * per_sm].contiguous()
    model.local_energy(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); model.local_energy(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print("walkers/SM %d: %.3f ms -> %.0f cycles per stage per CTA-slot pass" % (per_sm, ms, ms * 1e-3 * 1.965e9 / 64 / max(1, per_sm // 2)))
