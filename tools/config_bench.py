"""GPU timing of all five BASELINE configs at ~1M DOFs (single GPU, synthetic structured meshes, the scripts'
own parameters unless noted): steps/s, iteration counts, diagnostics.  Not the headline bench (bench.py =
config 2); evidence that every loop runs at scale.  usage: python tools/config_bench.py [steps] [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import dgp, ewm, jbp, ldc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
out_path = sys.argv[2] if len(sys.argv) > 2 else None


def timed(name, drv, warm=3, extra=None):
    drv.check = False
    for _ in range(warm):
        drv.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        r = drv.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    dofs = sum(drv.S[k].dim() for k in ("W", "Zc", "Q"))
    row = dict(config=name, dofs=dofs, cells=drv.mesh.num_cells(), ms_per_step=round(ms, 3), steps_per_s=round(1e3 / ms, 2),
               wall_ms_per_step=round((time.perf_counter() - t0) * 1e3 / steps, 3),
               iterations=dict(zip(drv.solve_labels, [i.iterations for i in drv.last_info])),
               result={k: float(v) for k, v in dict(r).items()})
    if extra:
        row.update(extra(drv))
    print(json.dumps(row), flush=True)
    return row


rows = []
# cfg1: incompressible cavity (singular pressure matrix: BiCGStab + Jacobi as in the script)
d = ldc.IncompLDC(pmesh.Skew_Mesh(192, 1, 1))
d.t = 0.45
rows.append(timed("cfg1 incomp_LDC n=192", d, extra=lambda q: dict(eye=q.min_location(q.stream_function())[0].tolist())))
del d
# cfg2: compressible cavity (the bench workload; CG + two-level pressure preconditioner)
d = ldc.CompLDC(pmesh.Skew_Mesh(192, 1, 1))
d.t = 0.45
d.pressure_method = "cg"
d.enable_pressure_coarse(256)
rows.append(timed("cfg2 comp_LDC n=192", d, extra=lambda q: dict(eye=q.min_location(q.stream_function(True))[0].tolist())))
del d
# cfg3: EWM journal bearing, main pass (ramped spin), annulus 768 x 48
d = ewm.EwmJBP(n_theta=768, n_r=48, Ma=0.01, We=0.1, betav=0.5)  # script values We = Ma = 1e-6 switch the polymer off
d.t = 0.45
d.pressure_method = "cg"
print("cfg3 two-level pressure preconditioner:", d.enable_pressure_coarse(256), flush=True)
rows.append(timed("cfg3 EWM journal bearing 768x48 (We=0.1, Ma=0.01, betav=0.5)", d))
del d
# cfg4: FENE-P-MP journal bearing, CSV member (We=0.1, Ma=0.01)
d = jbp.FenepmpJBP(n_theta=768, n_r=48)
d.t = 0.45
d.pressure_method = "cg"
print("cfg4 two-level pressure preconditioner:", d.enable_pressure_coarse(256), flush=True)
rows.append(timed("cfg4 FENE-P-MP journal bearing 768x48", d))
del d
# cfg5: buoyancy-driven compressible flow
d = dgp.CompDGP(mesh_resolution=192)
d.pressure_method = "cg"
print("cfg5 two-level pressure preconditioner:", d.enable_pressure_coarse(256), flush=True)
rows.append(timed("cfg5 compressible DGP n=192", d))
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
