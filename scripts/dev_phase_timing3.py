"""Per-phase cycle counts of the pipelined E_loc kernel (build with -DFF_PHASE_TIMING=<observer warp>)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), os.environ["FF_TIMING_LIB"])
os.environ["FF_ELOC_V3"] = "1"
import argparse, torch, bench
walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 296 * 8
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
_, x = model.sample((walkers,))
lib = L.lib()
lib.ff_debug_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong * 16), C.c_int]
out = (C.c_ulonglong * 16)()
model.local_energy(x); lib.ff_debug_phase_cycles(C.byref(out), 1)
model.local_energy(x); lib.ff_debug_phase_cycles(C.byref(out), 1)
item = ["(turn end->wait)", "wait EMPTY", "load+MLP", "shuffle+geom+M contraction", "wait ITEM bar", "A writes+arrive"]
mat = ["(turn end->wait)", "wait FULL", "-", "-", "-", "gather", "sync", "A.J + RK", "matvec", "sync", "gram | finale"]
names = item if os.environ.get("FF_OBS", "item") == "item" else mat
nb = walkers * 64
tot = sum(out)
for k, nm in enumerate(names):
    print("%-28s %8.0f cycles/stage  %5.1f%%" % (nm, out[k] / nb, 100.0 * out[k] / tot))
print("total per walker-stage", tot / nb)
