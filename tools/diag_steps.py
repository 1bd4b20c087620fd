"""GPU diagnostic: run config-2 steps and print per-step solver iterations / energies / wall time.
usage: python tools/diag_steps.py [n] [steps] [rtol] [pressure_method] [t0]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
pm = sys.argv[4] if len(sys.argv) > 4 else "cg"
t0 = float(sys.argv[5]) if len(sys.argv) > 5 else None
kind = int(sys.argv[6]) if len(sys.argv) > 6 else 2
m = pmesh.Skew_Mesh(n, 1, 1)
drv = (ldc.CompLDC if kind == 2 else ldc.IncompLDC)(m)
drv.solve_rtol = rtol
drv.pressure_method = pm
drv.check = False
if t0 is not None:
    drv.t = t0
print("n=%d rtol=%g pressure=%s t0=%s kind=%d" % (n, rtol, pm, drv.t, kind))
for k in range(steps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    try:
        r = drv.step()
    except RuntimeError as e:
        print("step", k, "FAILED:", e)
        print([(l, i.iterations, i.converged, "%.2e" % i.residual, "%.2e" % i.residual0)
               for l, i in zip(drv.solve_labels, drv.last_info)])
        break
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("step %d %.2f ms E_k %.6e E_e %.6e its %s" % (k, dt * 1e3, r["E_k"], r["E_e"],
                                                         [i.iterations for i in drv.last_info]))
print(drv.solve_labels)
