import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fermiflow_b200 import _lib as L
import torch
torch.cuda.init()
for name in ("ff_fp64_peak", "ff_fp64_mma_peak"):
    v = C.c_double()
    L.check(getattr(L.lib(), name)(20000, C.byref(v), None))
    print(name, v.value / 1e12, "TFLOP/s")
