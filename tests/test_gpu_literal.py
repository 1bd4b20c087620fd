"""GPU: the velocity half step of lid-driven-cavity/incomp_LDC.py:317-331 written exactly as the script writes it --
symbolic forms built from the drop-in module's names, `assemble(a1)`, `bc.apply`, `solve(A1, w12.vector(), b1,
"bicgstab", "default")` -- against the same step of the driver (fx_form ids).  Same functors underneath: the matrices
agree to the atomics' rounding, the solutions to the solver tolerance."""
import warnings

import numpy as np
import pytest
import torch

import fxt_common as lc
from fenics_b200 import _abi as abi

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from fenics_b200 import dolfin_api as dapi  # noqa: E402
from fenics_b200 import la, mesh as pmesh, symbolic  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402


def test_literal_velocity_half_step_equals_the_driver():
    import fenics_b200.shims.lid_driven_cavity.fenics_fem as ff
    from fenics_b200.shims.lid_driven_cavity.fenics_fem import (Dcomp, Dincomp, Parameter, TestFunctions, assemble, div, dx,
                                                                grad, inner, reshape_elements, solve, trial_functions)
    drv = ldc.IncompLDC(pmesh.Skew_Mesh(12, 1, 1))
    drv.t = 0.45
    drv.step()                                            # a state with every field non-trivial
    S = drv.S
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    pv = drv.p
    Re, dt, betav, th, conv = (Parameter(k, pv[k]) for k in ("Re", "dt", "betav", "th", "conv"))
    # set-up section of the script (:142-170, :259-262) on the driver's Functions
    _, p, T, tau_vec, u, D_vec, D, tau = trial_functions(Q, Zc, W)
    (v, R_vec) = TestFunctions(W)
    R = symbolic.as_matrix([[R_vec[0], R_vec[1]], [R_vec[1], R_vec[2]]])
    w0, w12, p0, tau0_vec = drv.w0, dapi.Function(W, drv.plan), drv.p0, drv.tau0_vec
    (u0, D0_vec) = w0.split()
    D0, _, _, _, tau0, _, _ = reshape_elements(D0_vec, D0_vec, D0_vec, D0_vec, tau0_vec, tau0_vec, tau0_vec)
    bcu = drv.bcu
    DEVSSl_u12 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
    DEVSSr_u12 = 2*(1-betav)*inner(D0,Dincomp(v))*dx
    # ---- incomp_LDC.py:317-331 -------------------------------------------------------------------------
    visc_u12 = betav*grad(u)
    lhs_u12 = (Re/(dt/2.0))*u
    rhs_u12 = (Re/(dt/2.0))*u0 - Re*conv*grad(u0)*u0

    a1=inner(lhs_u12,v)*dx + inner(visc_u12,grad(v))*dx + (inner(D-Dincomp(u),R)*dx)
    L1=inner(rhs_u12,v)*dx + inner(p0,div(v))*dx - inner(tau0,grad(v))*dx

    a1+= th*DEVSSl_u12
    L1+= th*DEVSSr_u12

    A1 = assemble(a1)
    b1= assemble(L1)
    [bc.apply(A1, b1) for bc in bcu]
    w12.vector().t.copy_(drv.w12.vector().t)              # same initial guess as the driver will use
    la.parameters["krylov_solver"]["relative_tolerance"], old = 1e-12, la.parameters["krylov_solver"]["relative_tolerance"]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", la.PreconditionerSubstitution)
            solve(A1, w12.vector(), b1, "bicgstab", "default")
    finally:
        la.parameters["krylov_solver"]["relative_tolerance"] = old
    # ---- the same sub-step through the driver's fx_form ids ----------------------------------------------
    Ad, bd = dapi.assemble(drv.pr.form(abi.FX_FORM_C1_A1)), dapi.assemble(drv.pr.form(abi.FX_FORM_C1_L1))
    [bc.apply(Ad, bd) for bc in bcu]
    assert lc.rel_err(A1.val.cpu().numpy(), Ad.val.cpu().numpy()) < 1e-13
    assert lc.rel_err(b1.get_local(), bd.get_local()) < 1e-13
    xd = la.Vector(drv.plan, "W")
    xd.t.copy_(drv.w12.vector().t)
    la.krylov_solve(Ad, xd, bd, "bicgstab", "jacobi", rtol=1e-12)
    assert lc.rel_err(w12.vector().get_local(), xd.get_local()) < 1e-9
    # ... and a form nobody wrote a functor for is refused by name, not mis-assembled
    with pytest.raises(symbolic.UnknownForm):
        assemble(inner(u, v) * dx)
    # functionals: E_k = assemble(0.5*dot(u1,u1)*dx) (:420)
    (u1, D1_vec) = drv.w1.split()
    E_k = assemble(0.5*ff.dot(u1,u1)*dx)
    assert abs(E_k - dapi.assemble(drv.pr.form(abi.FX_FORM_C1_EK))) <= 1e-14 * abs(E_k)
