"""BiCGStab solves of the cfg2 W / Zc systems on an n x n mesh against splu; prints one JSON line.  Run with FX_NB=0/1
to compare the scalar compacted-CSR kernel with the node-block kernel (tests/test_gpu_nb.py).
usage: python tools/nb_check.py [n]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import scipy.sparse.linalg as spla  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.step()
rng = np.random.default_rng(5)
out = {}
for space, fid, bcs in (("W", abi.FX_FORM_C2_A1, drv.bcu), ("W", abi.FX_FORM_C2_A7, drv.bcu), ("Zc", abi.FX_FORM_C2_A4, [])):
    A = dapi.assemble(drv.pr.form(fid))
    b = la.Vector(drv.plan, space)
    b.set_local(rng.standard_normal(b.size()))
    for bc in bcs:
        bc.apply(A, b)
    ref = spla.splu(A.to_scipy().tocsc()).solve(b.get_local())
    for pc in ("jacobi", "none"):
        for rtol in (1e-5, 1e-12):
            x = la.Vector(drv.plan, space)
            info = la.krylov_solve(A, x, b, "bicgstab", pc, rtol=rtol, maxit=20000)
            err = float(np.abs(x.get_local() - ref).max() / np.abs(ref).max())
            out["%s/%d/%s/%g" % (space, fid, pc, rtol)] = dict(iterations=info.iterations, converged=info.converged,
                                                              err_vs_splu=err if rtol < 1e-9 else 0.0)
print(json.dumps(out))
