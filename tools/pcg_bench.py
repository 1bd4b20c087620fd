"""GPU micro-benchmark: the pressure solve (config 2, A5 on P1) under the pipelined-CG kernel variants
(FX_PCG_VARIANT bit0 = LL barrier-reductions, bit1 = 1024-thread CTAs).  usage: FX_PCG_VARIANT=v python tools/pcg_bench.py [n]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.pressure_method = "cg"
drv.step()
A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A5))
b = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_L5))
x = la.Vector(drv.plan, "Q")
pc = "jacobi"
if os.environ.get("FX_COARSE"):
    assert drv.enable_pressure_coarse(int(os.environ["FX_COARSE"]))
    pc = "jacobi_coarse"
res = []
for rep in range(6):
    x.t.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info = la.krylov_solve(A, x, b, "cg", pc, rtol=1e-5)
    e1.record()
    torch.cuda.synchronize()
    res.append((e0.elapsed_time(e1), info.iterations))
ms, its = min(res)
print(json.dumps(dict(pc=pc, variant=os.environ.get("FX_PCG_VARIANT", "default"), n=n, dofs=A.n, ms=round(ms, 3), iterations=its,
                      us_per_iteration=round(1e3 * ms / its, 3), checksum=float(x.t.double().sum()))))
