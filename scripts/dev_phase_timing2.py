"""Per-phase cycle counts of the second-generation E_loc kernel (build with -DFF_PHASE_TIMING);
thread 0 (an item thread of warp 0) is the observer."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ.get("FF_TIMING_LIB", "libfermiflow_b200_timing0.so"))
import argparse, torch, bench
hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 50
walkers = int(sys.argv[2]) if len(sys.argv) > 2 else 296 * 8
args = argparse.Namespace(hidden=hidden, ode_steps=16, nup=10, ndown=10, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
_, x = model.sample((walkers,))
lib = L.lib()
lib.ff_debug_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong * 16), C.c_int]
out = (C.c_ulonglong * 16)()
model.local_energy(x); lib.ff_debug_phase_cycles(C.byref(out), 1)
model.local_energy(x); lib.ff_debug_phase_cycles(C.byref(out), 1)
names = ["loop", "-", "A MLP|gram", "sync", "B contract", "sync", "C gather", "sync", "D gemm+rk", "D matvec", "sync"]
nb = walkers * 64
tot = sum(out)
for k, nm in enumerate(names):
    print("%-16s %8.0f cycles/stage  %5.1f%%" % (nm, out[k] / nb, 100.0 * out[k] / tot))
print("total per stage", tot / nb)
