import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctypes as C
from fermiflow_b200 import _lib as L
from oracle import fermiflow_oracle as O
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
def mk(H, seed, sc):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(H, generator=g), torch.randn(H, generator=g), sc*torch.randn(H, generator=g))
def rel(a, b):
    a = a.cpu(); return float((a-b).abs().max() / b.abs().max())
for (nup, ndn, H, Hm, S, B) in [(3,2,8,6,16,5), (3,0,8,0,8,7), (6,6,16,16,8,9), (10,10,50,50,4,3)]:
    n = nup+ndn
    eta = mk(H, 1, 0.05); mu = mk(Hm, 2, 0.05) if Hm else None
    g = torch.Generator().manual_seed(5)
    z0 = 0.9*torch.randn(B, n, 2, generator=g)
    ts = (0.0, 1.0)
    eta_d = tuple(t.to(dev) for t in eta); mu_d = tuple(t.to(dev) for t in mu) if mu else None
    m = L.make_model(nup, ndn, eta_d, mu_d, ts, S)
    zd = z0.to(dev); xd = torch.empty_like(zd)
    L.check(L.lib().ff_cnf_generate(C.byref(m), L.ptr(zd), B, 0, L.ptr(xd), L.stream()))
    torch.cuda.synchronize()
    x_ref = O.cnf_generate(z0, eta, mu, ts, S)
    print("n=%d generate rel err %.2e" % (n, rel(xd, x_ref)))
    zb = torch.empty_like(xd); dl = torch.empty(B, device=dev)
    L.check(L.lib().ff_cnf_delta_logp(C.byref(m), L.ptr(xd), B, L.ptr(zb), L.ptr(dl), None, None, L.stream()))
    z_ref, dl_ref = O.cnf_delta_logp(x_ref, eta, mu, ts, S)
    print("   delta_logp z err %.2e dl err %.2e" % (rel(zb, z_ref), rel(dl, dl_ref)))
    orb = torch.tensor(list(range(nup)) + list(range(ndn)), dtype=torch.int32, device=dev)
    outs = {k: torch.empty(B, device=dev) for k in ("dl","logp","lap","kin","pot","eloc")}
    grad = torch.empty(B, n, 2, device=dev); z2 = torch.empty(B, n, 2, device=dev)
    t0=time.time()
    L.check(L.lib().ff_eloc(C.byref(m), L.ptr(xd), B, L.ptr(orb, torch.int32), None, 2.0, 1,
            L.ptr(z2), L.ptr(outs["dl"]), L.ptr(outs["logp"]), L.ptr(grad), L.ptr(outs["lap"]),
            L.ptr(outs["kin"]), L.ptr(outs["pot"]), L.ptr(outs["eloc"]), None, None, L.stream()))
    torch.cuda.synchronize()
    if n <= 12:
        r = O.local_energy(x_ref, list(range(nup)), list(range(ndn)), eta, mu, ts, S, 2.0)
        print("   eloc: logp %.2e grad %.2e lap %.2e kin %.2e pot %.2e eloc %.2e" % (
            rel(outs["logp"], r["logp"]), rel(grad, r["grad"]), rel(outs["lap"], r["lap"]),
            rel(outs["kin"], r["kinetic"]), rel(outs["pot"], r["potential"]), rel(outs["eloc"], r["eloc"])))
    else:
        print("   eloc values", outs["eloc"].cpu().numpy(), "time", time.time()-t0)
fl = C.c_double()
L.check(L.lib().ff_fp64_peak(20000, C.byref(fl), None))
print("fp64 peak TFLOP/s", fl.value/1e12)
