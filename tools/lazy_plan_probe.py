"""GPU check: per-step time with synchronous results vs pipelined results (pipeline=True)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from fenics_b200 import mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

def fresh(pipeline):
    d = ldc.CompLDC(pmesh.Skew_Mesh(192, 1, 1))
    d.pressure_method = "cg"
    d.check = False
    d.pipeline = pipeline
    d.t = 0.45
    for _ in range(3):
        d.step()
    d.sync()
    return d


def timed(drv, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    cpu = []
    for _ in range(n):
        c0 = time.perf_counter()
        drv.step()
        cpu.append((time.perf_counter() - c0) * 1e3)
    e1.record()
    torch.cuda.synchronize()
    drv.sync()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, min(cpu), max(cpu)


for mode in (False, True, False, True):
    print("pipeline=%s: gpu %.2f ms wall %.2f ms; step() call %.2f..%.2f ms" % ((mode,) + timed(fresh(mode))))
