"""Per-phase cycle counts of the barrier-synchronous fused E_loc kernel (eloc2_kernel; build with
-DFF_PHASE_TIMING=<observer warp>; run with FF_TIMING_LIB=<lib>)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ["FF_TIMING_LIB"])
import argparse, torch, bench
walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 296 * 8
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
_, x = model.sample((walkers,))
lib = L.lib()
lib.ff_debug_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong * 16), C.c_int]
out = (C.c_ulonglong * 16)()
model.local_energy(x, stash=True); lib.ff_debug_phase_cycles(C.byref(out), 1)
model.local_energy(x, stash=True); lib.ff_debug_phase_cycles(C.byref(out), 1)
names = ["loop top", "A: r, rsqrt", "A: table look-up", "A: geometry, G | helper: Gram", "barrier 1", "B: M contraction", "barrier 2",
         "C: A blocks + gather", "barrier 3", "D: A.J + RK | vectors", "-", "barrier 4"]
nb = walkers * 64
tot = sum(out)
print("observer warp of", os.environ["FF_TIMING_LIB"], "walkers", walkers)
for k, nm in enumerate(names):
    print("%-32s %8.0f cycles/stage  %5.1f%%" % (nm, out[k] / nb, 100.0 * out[k] / tot))
print("total per stage", tot / nb)
