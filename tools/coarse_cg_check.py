"""GPU check of the two-level CG outside the resident kernel (k_cg with coarse correction):
   FX_PCG_RESIDENT=0 python tools/coarse_cg_check.py [n]      (without the variable: the resident kernel)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import scipy.sparse.linalg as spla  # noqa: E402

from fenics_b200 import _abi as abi, dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.pressure_method = "cg"
for _ in range(2):
    drv.step()                                   # plain Jacobi here
print("coarse accepted:", drv.enable_pressure_coarse(k), "aggregates", drv._coarse_size(drv.S["Q"].dim(), k))
A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A5))
b = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_L5))
ref = spla.splu(A.to_scipy().tocsc()).solve(b.get_local())
la.parameters["krylov_solver"]["error_on_nonconvergence"] = False
for pc in ("jacobi", "jacobi_coarse"):
    for rtol in (1e-5, 1e-9):
        x = la.Vector(drv.plan, "Q")
        info = la.krylov_solve(A, x, b, "cg", pc, rtol=rtol, maxit=50000)
        err = np.abs(x.get_local() - ref).max() / np.abs(ref).max()
        print(pc, "rtol", rtol, "iterations", info.iterations, "converged", info.converged, "residual %.3e / %.3e" %
              (info.residual, info.residual0), "err vs splu %.2e" % err)
