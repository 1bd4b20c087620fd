"""Small ncu target: fixed-length BiCGStab solves (K/4 and K iterations) of the cfg2 W and Zc systems at n (default 192);
the time step before them launches 5 BiCGStab kernels (skip them with ncu -s 5).
usage: python tools/nb_profile_target.py [n] [iterations]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1), dt=0.001 * min(1.0, 192.0 / n))  # dt ~ h beyond the bench mesh (stability)
drv.t = 0.45
drv.step()
torch.manual_seed(7)
for space, fid, bcs in (("W", abi.FX_FORM_C2_A1, drv.bcu), ("Zc", abi.FX_FORM_C2_A4, [])):
    A = dapi.assemble(drv.pr.form(fid))
    b, x = la.Vector(drv.plan, space), la.Vector(drv.plan, space)
    b.t.normal_()
    for bc in bcs:
        bc.apply(A, b)
    for k in (K // 4, K):  # two lengths: DRAM bytes per iteration = difference / (K - K/4), the rest is set-up
        x.t.zero_()
        la.krylov_solve(A, x, b, "bicgstab", "jacobi", rtol=0.0, atol=0.0, maxit=k, ticket=0)
    torch.cuda.synchronize()
print("ok")
