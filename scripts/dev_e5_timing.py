"""Per-role cycle counts of eloc5_kernel (build with scripts/build_variant.sh timing5 -DFF_E5_TIMING; FF_DEV_LIB=libff_timing5.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ.get("FF_DEV_LIB", "libff_timing5.so"))
import argparse, torch, bench
walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 296 * 8
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
_, x = model.sample((walkers,))
lib = L.lib()
lib.ff_debug_e5_cycles.argtypes = [C.POINTER(C.c_ulonglong * 64), C.c_int]
out = (C.c_ulonglong * 64)()
model.local_energy(x, stash=True); lib.ff_debug_e5_cycles(C.byref(out), 1)
model.local_energy(x, stash=True); lib.ff_debug_e5_cycles(C.byref(out), 1)
names = {0: ["wait M free + Gram", "wait items", "row sums + y", "bar 1", "init + K.A (DMMA)", "K.u + Ks copy", "-", "owner barrier"],
         2: ["A blocks (after the wait)", "wait M", "contraction", "bar 3 + sums + L", "wait row sums", "scalars", "items + wait A free", "next radial functions"]}
nb = walkers * 64
for obs, role in enumerate(["owner warp 0 (no Gram)", "owner warp 1", "first worker warp", "last worker warp"]):
    v = out[16 * obs:16 * obs + 16]
    tot = sum(v)
    print(role, "total %.0f cycles/stage" % (tot / nb))
    nm = names[0 if obs < 2 else 2]
    for k in range(8):
        print("   %-22s %7.0f  %5.1f%%" % (nm[k], v[k] / nb, 100.0 * v[k] / tot))
    print("   %-22s %7.0f" % ("loop top", v[15] / nb))
