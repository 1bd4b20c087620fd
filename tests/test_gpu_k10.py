"""Long K-step parity (K = 10) on meshes of the reference's own first refinement level: fields and physics
diagnostics of the GPU path against the oracle after 10 time steps, 1e-8 relative (north_star).  The oracle solves
exactly (LU); the GPU Krylov solves are tightened to 1e-13, so what is compared is the discretisation + the loop, and
errors would have 10 steps to compound.

  cfg1 / cfg2: n = 40 is incomp_LDC.py's / comp_LDC_mesh_convergence.py's first mesh (mm = 40, :96 / :73)
  cfg3 / cfg4: 80 x 20 O-grid annulus (3 200 cells, the size of the scripts' first JBP_mesh); cfg5: n = 40
"""
import numpy as np
import pytest
import torch

import fxt_common as lc

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

K = 10


def _run(ocl, drv, t0, scalars, e_e_scale=None, fields=None):
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    ocl.t = drv.t = t0
    for step in range(K):
        ro, rd = ocl.step(), drv.step()
        # the two load components are compared on the scale of the load VECTOR: near symmetric flow states one of
        # them is a cancellation-limited small number (F_x = 8.7e-3 against F_y = -50 in the EWM main pass)
        load = max(abs(ro.get("F_x", 0.0)), abs(ro.get("F_y", 0.0)))
        for k in scalars:
            scale = load if k in ("F_x", "F_y") else abs(ro[k])
            assert abs(ro[k] - rd[k]) < 1e-8 * max(scale, 1e-10), (step, k, ro[k], rd[k])
        if e_e_scale is not None:
            assert abs(ro["E_e"] - rd["E_e"]) < 1e-8 * max(abs(ro["E_e"]), e_e_scale), (step, ro["E_e"], rd["E_e"])
    fo, fd = ocl.fields(), drv.fields()
    for k in (fields or fo):
        # kapp = sqrt(|(I - Pi) residual|^2) is an indicator derived from the fields through a cancelling projection and
        # a square root (relative error amplified ~1/kapp): 1e-7; the solution fields: 1e-8 (north_star)
        assert lc.rel_err(fd[k], fo[k]) < (1e-7 if k == "kapp" else 1e-8), k
    assert all(i.converged == 1 for i in drv.last_info)
    return ro, rd


@pytest.mark.parametrize("kind", [1, 2])
def test_cavity_k10_n40(kind):
    from test_gpu_parity import _pair
    ocl, drv = _pair(kind, n=40)
    ro, rd = _run(ocl, drv, 0.45, ("E_k",), e_e_scale=1e-12)
    # the diagnostic the north_star names for the cavity: location of the stream-function minimum (primary vortex eye)
    # (compared with the oracle's in test_stream_function_and_min_location; here: it sits where a lid-driven eye does)
    psi = drv.stream_function(kind == 2)
    xy = drv.min_location(psi)[0]
    assert 0.2 < xy[0] < 0.8 and 0.5 < xy[1] < 1.0 and psi.vector().get_local().min() < 0.0


@pytest.mark.parametrize("prepass", [False, True])
def test_ewm_journal_bearing_k10(prepass):
    import fxt_ewm as ew
    from test_gpu_ewm import _pair
    ocl, drv = _pair(nt=80, nr=20, prepass=prepass, **ew.GENERAL)
    _run(ocl, drv, 0.45, ("E_k", "torque", "F_x", "F_y"), e_e_scale=10.0, fields=("w1", "p1", "tau1", "rho1", "T1"))


@pytest.mark.parametrize("lam", [0.0, 0.1])
def test_fenepmp_journal_bearing_k10(lam):
    from test_gpu_jbp import _pair
    ocl, drv = _pair(nt=80, nr=20, We=0.5, lambda_d=lam)
    _run(ocl, drv, 0.45, ("E_k", "torque", "F_x", "F_y"))


def test_buoyancy_k10_n40():
    from test_gpu_dgp import _pair
    ocl, drv = _pair(n=40)
    _run(ocl, drv, 1.4, ("E_k", "Nus"), e_e_scale=2.0)
