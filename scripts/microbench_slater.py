"""BASELINE.json config 5: batched log|det| + gradient + exact Laplacian of the free-fermion state,
N = 6..30 (N/2 up, N/2 down), 1e6 walkers, against the reference's autograd path on the host CPU
(oracle: log|det| through torch.linalg.slogdet + 1 + 2N autograd passes, utils.py:44-65) on a
bounded walker sample.  Prints one JSON line per N."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fermiflow_b200 import HO2D, FreeFermion

B = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
cpu_walkers = int(sys.argv[2]) if len(sys.argv) > 2 else 256
Ns = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(6, 31, 2))
dev = torch.device("cuda:0")
ho, ff = HO2D(), FreeFermion(dev)
from oracle import fermiflow_oracle as O
for N in Ns:
    nup = N // 2
    up, dn = ho.orbitals[:nup], ho.orbitals[:N - nup]
    x = 1.2 * torch.randn(B, N, 2, device=dev)
    ff.log_prob_grad_laplacian(up, dn, x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); logp, grad, lap = ff.log_prob_grad_laplacian(up, dn, x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    # CPU: the reference's way (nested autograd), bounded sample, all host threads
    xc = x[:cpu_walkers].cpu()
    t = time.time()
    lp_ref, g_ref, l_ref = O.free_fermion_grad_laplacian(list(range(nup)), list(range(N - nup)), xc)
    tc = time.time() - t
    err = max(float((logp[:cpu_walkers].cpu() - lp_ref).abs().max() / lp_ref.abs().max()),
              float((grad[:cpu_walkers].cpu() - g_ref).abs().max() / g_ref.abs().max()),
              float((lap[:cpu_walkers].cpu() - l_ref).abs().max() / l_ref.abs().max()))
    print(json.dumps({"N": N, "walkers": B, "gpu_ms": round(ms, 3), "gpu_walkers_per_s": round(B / ms * 1e3),
                      "cpu_walkers": cpu_walkers, "cpu_s": round(tc, 3), "cpu_walkers_per_s": round(cpu_walkers / tc, 1),
                      "cpu_threads": torch.get_num_threads(), "speedup": round(B / ms * 1e3 / (cpu_walkers / tc)),
                      "max_rel_err_vs_oracle": err}))
