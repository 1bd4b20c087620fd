"""GPU micro-benchmark of the node-block SpMV (fx_spmv_variant 8) against the stand-alone CSR-stream SpMV, and of the
BiCGStab kernels (FX_NB is read once per process: run twice to A/B).  usage: python tools/nb_bench.py [n] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh  # noqa: E402
from fenics_b200._lib import check, lib  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1), dt=0.001 * min(1.0, 192.0 / n))  # dt ~ h beyond the bench mesh (stability)
drv.t = 0.45
for _ in range(3):
    drv.step()
out = {"n": n, "FX_NB": os.environ.get("FX_NB", "1")}


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


for space, fid, bcs in (("W", abi.FX_FORM_C2_A1, drv.bcu), ("Zc", abi.FX_FORM_C2_A4, [])):
    A = dapi.assemble(drv.pr.form(fid))
    x, y, b = la.Vector(drv.plan, space), la.Vector(drv.plan, space), la.Vector(drv.plan, space)
    x.t.normal_()
    torch.manual_seed(7)
    b.t.normal_()
    for bc in bcs:
        bc.apply(A, b)
    sid = abi.SPACE_ID[space]

    def run(variant):
        check(lib.fx_spmv_variant(drv.plan.h, sid, la._ptr(A.val), la._ptr(x.t), la._ptr(y.t), variant, la._stream()))
    t_csr = timed(lambda: run(0), 50)
    t1 = timed(lambda: run(8 + (1 << 8)))
    tR = timed(lambda: run(8 + (REPS << 8)))
    per = (tR - t1) / (REPS - 1)  # one product + one grid barrier, values already gathered
    bs = 2 if space == "W" else 3
    # the layout's own bytes: block entries x (bs^2 x 8 + 4) + in/out vectors
    row = {"csr_stream_us": round(t_csr, 2), "nb_setup_plus_one_us": round(t1, 2), "nb_product_us": round(per, 2)}
    xs = la.Vector(drv.plan, space)

    def solve_fixed(k):  # exactly k iterations (rtol 0 is never met): time per iteration without the set-up
        def f():
            xs.t.zero_()
            la.krylov_solve(A, xs, b, "bicgstab", "jacobi", rtol=0.0, atol=0.0, maxit=k, ticket=0)
        return f
    t10, t60 = timed(solve_fixed(10), 10), timed(solve_fixed(60), 10)
    row["bicgstab_us_per_iteration"] = round((t60 - t10) / 50, 2)
    row["bicgstab_setup_us"] = round(t10 - 10 * (t60 - t10) / 50, 1)
    xs.t.zero_()
    info = la.krylov_solve(A, xs, b, "bicgstab", "jacobi", rtol=1e-5, maxit=20000)
    row["iterations_rtol1e-5"] = info.iterations
    out[space] = row
    print(space, row, flush=True)
print(json.dumps(out))
