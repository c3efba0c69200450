"""Per-warp phase cycles of pgrad_binned_kernel (capi.cu built with -DFF_PG_TIMING into libff_pgt.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermiflow_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "libff_pgt.so")
import argparse, torch, bench
walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
args = argparse.Namespace(hidden=50, ode_steps=16, nup=10, ndown=10, Z=2.0)
dev = torch.device("cuda:0")
model = bench.build_model(args, dev)
lib = L.lib()
lib.ff_debug_pg_cycles.argtypes = [C.POINTER(C.c_ulonglong * 164), C.c_int]
out = (C.c_ulonglong * 164)()
for it in range(2):
    g = model(walkers); g.backward(); lib.ff_debug_pg_cycles(C.byref(out), 1)
names = ["walk->top (direct tail)", "top barrier wait", "records", "barrier", "next-tile fetch", "walk", "-", "-"]
ntile = out[162]
print("tiles", ntile, "max bin count %d, mean over tiles and warps of the largest bin of the warp %.1f (mean bin ~8)" % (out[160], out[163] / max(ntile, 1) / 20))
for w in (0, 2, 4, 6, 8, 10, 12, 14, 16, 19):
    v = out[8 * w:8 * w + 8]; tot = sum(v)
    print("warp %2d: total %8.0f cycles/tile | " % (w, tot / max(ntile, 1) * 148) + "  ".join("%s %4.1f%%" % (names[k][:14], 100.0 * v[k] / max(tot, 1)) for k in range(6)))
