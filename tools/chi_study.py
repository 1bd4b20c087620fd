"""Soft physics pin: the journal-bearing stability measure chi = F_x / F_y at the end of the
FENE-P-MP loop for a few Weissenberg numbers, on a script-sized O-grid.  The only numbers the reference itself holds
for this path are the chi curves of plot_data.py:10-13 (0.019 ... 0.55, increasing in We).
usage: python tools/chi_study.py [n_theta] [n_r] [t_end]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200.drivers import jbp  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
t_end = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
out = []
for We in (0.1, 0.25, 0.5, 1.0):
    d = jbp.FenepmpJBP(n_theta=nt, n_r=nr, We=We)
    d.pressure_method = "cg"
    d.enable_pressure_coarse()
    d.check = False
    t0 = time.time()
    steps = 0
    hist = []
    while d.t < t_end:
        r = d.step()
        steps += 1
        if steps % 200 == 0:
            hist.append((round(r["t"], 3), r["F_x"] / r["F_y"], r["torque"]))
    torch.cuda.synchronize()
    row = dict(We=We, steps=steps, wall_s=round(time.time() - t0, 1), chi=r["F_x"] / r["F_y"], F_x=r["F_x"], F_y=r["F_y"],
               torque=r["torque"], E_k=r["E_k"], history=hist[-6:])
    print(json.dumps(row), flush=True)
    out.append(row)
