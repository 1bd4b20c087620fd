"""One config-2 time step at mesh size n (default 192) -- the smallest program that launches every kernel of
the hot path once; used as the ncu target.  usage: python tools/one_step.py [n] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.pressure_method = "cg"
if not os.environ.get("FX_NO_COARSE"):
    drv.enable_pressure_coarse(256)
for _ in range(steps):
    out = drv.step()
torch.cuda.synchronize()
print(out, [(lbl, i.iterations) for lbl, i in zip(drv.solve_labels, drv.last_info)])
