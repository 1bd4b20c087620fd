"""ILU(0) vs Jacobi on the GPU, measured: iterations and milliseconds per solve of the three systems of config 2
(velocity-DEVSS W, stress Zc, pressure Q) at rtol 1e-5 from a zero initial guess, on growing meshes.  Appends one JSON
line per mesh to gpurun_out/ilu_study.jsonl (the level-by-level dependency chain of natural-ordering ILU(0) grows with
the mesh; DESIGN.md section 4 quotes these numbers).

    python tools/ilu_study.py 32 64 96
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

la.parameters["preconditioner_substitution"] = "silent"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "ilu_study.jsonl")
os.makedirs(os.path.dirname(out), exist_ok=True)
t_end = time.time() + float(os.environ.get("ILU_STUDY_BUDGET_S", "25"))
for n in [int(a) for a in sys.argv[1:]] or [32, 64]:
    drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
    drv.t = 0.45
    drv.step()
    drv.step()
    rec = {"n": n, "systems": {}}
    rng = np.random.default_rng(0)
    for name, fid, bc in (("Q", abi.FX_FORM_C2_A5, False), ("Zc", abi.FX_FORM_C2_A4, False), ("W", abi.FX_FORM_C2_A1, True)):
        A = dapi.assemble(drv.pr.form(fid))
        b = la.Vector(drv.plan, name)
        b.set_local(rng.standard_normal(b.size()))
        if bc:
            for c in drv.bcu:
                c.apply(A, b)
        row = {"dofs": b.size()}
        for pc in ("jacobi", "ilu"):
            x = la.Vector(drv.plan, name)
            for rep in range(2):                           # second solve timed (workspaces exist)
                x.zero()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                info = la.krylov_solve(A, x, b, "bicgstab", pc, rtol=1e-5, maxit=20000)
                e1.record()
                torch.cuda.synchronize()
            row[pc] = {"iterations": info.iterations, "ms": round(e0.elapsed_time(e1), 3)}
        rec["systems"][name] = row
    with open(out, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)
    if time.time() > t_end:
        break
