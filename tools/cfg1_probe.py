"""config 1 on the GPU: per-step iteration counts and energies for a few (n, dt) -- is the 330-iteration stress solve of
the bench configuration (n = 192, dt = 0.005) a property of the system or of a transient?
usage: python tools/cfg1_probe.py [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
out = []
for n, dt, t0 in ((192, 0.005, 0.45), (192, 0.0025, 0.45), (128, 0.005, 0.45), (192, 0.005, 0.0)):
    d = ldc.IncompLDC(pmesh.Skew_Mesh(n, 1, 1), dt=dt, device=0)
    d.t = t0 if t0 > 0 else dt
    d.pressure_method = "cg"
    d.enable_pressure_coarse(256)
    rows = []
    for k in range(steps):
        r = d.step()
        rows.append(dict(t=round(r["t"], 4), E_k=r["E_k"], E_e=r["E_e"], err=r["err"],
                         its=[i.iterations for i in d.last_info]))
    out.append(dict(n=n, dt=dt, t0=t0, labels=list(d.solve_labels), rows=rows))
    del d
    torch.cuda.empty_cache()
print(json.dumps(out))
