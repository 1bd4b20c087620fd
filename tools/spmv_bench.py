"""GPU micro-benchmark: stand-alone SpMV variants on the n x n cavity matrices (algorithmic GB/s).
usage: python tools/spmv_bench.py [n]"""
import ctypes as C
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
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.step()
out = {}
for space, fid in (("W", abi.FX_FORM_C2_A1), ("Zc", abi.FX_FORM_C2_A4), ("Q", abi.FX_FORM_C2_A5)):
    A = dapi.assemble(drv.pr.form(fid))
    x, y = la.Vector(drv.plan, space), la.Vector(drv.plan, space)
    x.t.normal_()
    nbytes = 12 * A.nnz + 4 * (A.n + 1) + 16 * A.n
    ref = None
    for variant in range(8):
        def run():
            check(lib.fx_spmv_variant(drv.plan.h, abi.SPACE_ID[space], la._ptr(A.val), la._ptr(x.t), la._ptr(y.t), variant,
                                      la._stream()))
        for _ in range(5):
            run()
        if variant == 0:
            ref = y.t.clone()
        else:
            err = float((y.t - ref).abs().max() / ref.abs().max())
            assert err < 1e-12, (space, variant, err)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out["%s/v%d" % (space, variant)] = dict(us=round(us, 2), gbs=round(nbytes / us / 1e3, 1))
        print(space, "variant", variant, "%.1f us  %.0f GB/s (algorithmic, %d MB)" % (us, nbytes / us / 1e3, nbytes / 1e6))
print(json.dumps(out))
