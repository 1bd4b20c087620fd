"""probe: CG iteration counts of the cfg3 pressure system with / without the two-level preconditioner.
usage: python tools/coarse_probe.py n [Ma]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la  # noqa: E402
from fenics_b200.drivers import ewm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
Ma = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
la.parameters["krylov_solver"]["error_on_nonconvergence"] = False
d = ewm.EwmJBP(n_theta=4 * n, n_r=n // 4, Ma=Ma)
d.t = 0.45
d.pressure_method = "cg"
d.step()
for k in (256, 1024):
    ok = d.enable_pressure_coarse(k)
    A = dapi.assemble(d.pr.form(abi.FX_FORM_C3_A5))
    b = dapi.assemble(d.pr.form(abi.FX_FORM_C3_L5))
    for bc in getattr(d, "bcp", []):
        bc.apply(A, b)
    As = A.to_scipy()
    rs = np.abs(As @ np.ones(As.shape[0])).max() / np.abs(As.diagonal()).max()
    for pc in ("jacobi", "jacobi_coarse"):
        for rtol in (1e-5, 1e-10):
            x = la.Vector(d.plan, "Q")
            info = la.krylov_solve(A, x, b, "cg", pc, rtol=rtol, maxit=20000)
            v = x.get_local()
            res = np.linalg.norm(As @ v - b.get_local()) / np.linalg.norm(b.get_local())
            print("n", n, "Ma", Ma, "k", k, ok, pc, "rtol", rtol, "its", info.iterations, "conv", info.converged,
                  "true rel res %.2e" % res, "rowsum/diag %.1e" % rs, flush=True)
