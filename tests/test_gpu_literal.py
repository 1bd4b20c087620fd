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


def test_literal_config2_forms_through_lhs_rhs_equal_the_driver():
    """comp_LDC_mesh_convergence.py:414-549: every form of the compressible loop written as the script writes it -- one
    residual per momentum / stress step, split with lhs() / rhs() -- assembles to what the driver's fx_form ids give."""
    from fenics_b200.shims.lid_driven_cavity.fenics_fem import (CellDiameter, Dcomp, Dincomp, FacetNormal, Fdef, Parameter,
                                                                TestFunction, TestFunctions, assemble, div, dot, ds, dx,
                                                                grad, inner, lhs, nabla_grad, rhs, sigma, trial_functions)
    drv = ldc.CompLDC(pmesh.Skew_Mesh(12, 1, 1))
    drv.t = 0.45
    drv.step()
    drv.step()                                            # every field non-trivial, w0 != w1
    S = drv.S
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    pv = drv.p
    Re, We, Ma, dt, betav, th, conv, c1, c2 = (Parameter(k, pv[k]) for k in
                                               ("Re", "We", "Ma", "dt", "betav", "th", "conv", "c1", "c2"))
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = trial_functions(Q, Zc, W)
    rho, q = p, TestFunction(Q)
    Rt = m3(TestFunction(Zc))
    (v, R_vec) = TestFunctions(W)
    R = m3(R_vec)
    h, n = CellDiameter(drv.mesh), FacetNormal(drv.mesh)
    w0, w12, ws, w1, p0, p1, rho0, rho1 = drv.w0, drv.w12, drv.ws, drv.w1, drv.p0, drv.p1, drv.rho0, drv.rho1
    tau0_vec, tau1_vec, kapp = drv.tau0_vec, drv.tau1_vec, drv.kapp
    (u0, D0_vec), (u12, D12_vec), (us, Ds_vec), (u1, D1_vec) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D12, tau0 = m3(D0_vec), m3(D12_vec), m3(symbolic.as_expr(tau0_vec))
    DEVSSl_u12 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
    DEVSSl_u1 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
    LPSl_stress = inner(kapp*h*c1*grad(tau),grad(Rt))*dx + inner(kapp*h*c2*div(tau),div(Rt))*dx
    DEVSSr_u12 = 2*(1-betav)*inner(D0,Dincomp(v))*dx
    DEVSSr_u1 = 2*(1-betav)*inner(D12,Dincomp(v))*dx
    U = 0.5*(u + u0)
    forms = {}
    # ---- :417-421
    rho_eq = 2.0*(rho - rho0)/dt + dot(u0, grad(rho0)) - rho0*div(u0)
    rho_weak = inner(rho_eq,q)*dx
    forms["A0"], forms["L0"] = lhs(rho_weak), rhs(rho_weak)
    # ---- :430-443
    lhsFu12 = Re*rho0*(2.0*(u - u0) / dt + conv*dot(u0, nabla_grad(u0)))
    Fu12 = dot(lhsFu12, v)*dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \
        + dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \
        - dot(tau0*n, v)*ds \
        +  inner(D-Dcomp(u),R)*dx
    a1 = lhs(Fu12)
    L1 = rhs(Fu12)
    a1+= th*DEVSSl_u12
    L1+= th*DEVSSr_u12
    forms["A1"], forms["L1"] = a1, L1
    # ---- :458-473
    lhsFus = Re*rho0*((u - u0)/dt + conv*dot(u12, nabla_grad(u12)))
    Fus = dot(lhsFus, v)*dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \
        + 0.5*dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \
        - dot(tau0*n, v)*ds\
        +  inner(D-Dcomp(u),R)*dx
    a2= lhs(Fus)
    L2= rhs(Fus)
    a2+= th*DEVSSl_u1
    L2+= th*DEVSSr_u1
    forms["A2"], forms["L2"] = a2, L2
    # ---- :483-490
    lhs_p_1 = (Ma*Ma/(dt*Re))*p
    rhs_p_1 = (Ma*Ma/(dt*Re))*p0 - rho0*div(us) + dot(grad(rho0),us)
    lhs_p_2 = 0.5*dt*grad(p)
    rhs_p_2 = 0.5*dt*grad(p0)
    forms["A5"] = inner(lhs_p_1,q)*dx + inner(lhs_p_2,grad(q))*dx
    forms["L5"] = inner(rhs_p_1,q)*dx + inner(rhs_p_2,grad(q))*dx
    # ---- :506-512
    lhs_u1 = (Re/dt)*rho1*u
    rhs_u1 = (Re/dt)*rho1*us
    forms["A7"] = inner(lhs_u1,v)*dx  + inner(D-Dcomp(u),R)*dx
    forms["L7"] = inner(rhs_u1,v)*dx + 0.5*inner(p1-p0,div(v))*dx - 0.5*dot(p1*n, v)*ds
    # ---- :528-537
    lhs_tau1 = (We/dt+1.0)*tau  +  We*Fdef(u1,tau)
    rhs_tau1= (We/dt)*tau0 + 2.0*(1.0-betav)*Dcomp(u1)
    A = inner(lhs_tau1,Rt)*dx - inner(rhs_tau1,Rt)*dx
    a4 = lhs(A)
    L4 = rhs(A)
    a4 += LPSl_stress
    L4 += 0
    forms["A4"], forms["L4"] = a4, L4
    for name, form in forms.items():
        fid = getattr(abi, "FX_FORM_C2_" + name)
        got, want = assemble(form), dapi.assemble(drv.pr.form(fid))
        a, b = (got.val, want.val) if name.startswith("A") else (got.t, want.t)
        assert lc.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-13, name
        assert float(b.abs().max()) > 0.0, name
    E_k = assemble(0.5*rho1*dot(u1,u1)*dx)
    E_e = assemble((tau1_vec[0]+tau1_vec[2])*dx)
    assert abs(E_k - dapi.assemble(drv.pr.form(abi.FX_FORM_C2_EK))) <= 1e-14 * abs(E_k)
    assert abs(E_e - dapi.assemble(drv.pr.form(abi.FX_FORM_C2_EE))) <= 1e-13 * max(abs(E_e), 1e-12)


# the loop body of comp_LDC_mesh_convergence.py:414-549 (statements as the script writes them; the LPS block :392-402 and
# the field roll :558-562 are the driver's -- they are projections / assignments, not forms)
CFG2_LOOP_BODY = '''
U = 0.5*(u + u0)
rho_eq = 2.0*(rho - rho0)/dt + dot(u0, grad(rho0)) - rho0*div(u0)
rho_weak = inner(rho_eq,q)*dx
a0 = lhs(rho_weak)
L0 = rhs(rho_weak)
A0 = assemble(a0)
b0 = assemble(L0)
[bc.apply(A0, b0) for bc in bcp]
solve(A0, rho12.vector(), b0, "bicgstab", "default")
end()
lhsFu12 = Re*rho0*(2.0*(u - u0) / dt + conv*dot(u0, nabla_grad(u0)))
Fu12 = dot(lhsFu12, v)*dx + \\
    + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \\
    + dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \\
    - dot(tau0*n, v)*ds \\
    +  inner(D-Dcomp(u),R)*dx
a1 = lhs(Fu12)
L1 = rhs(Fu12)
a1+= th*DEVSSl_u12
L1+= th*DEVSSr_u12
A1 = assemble(a1)
b1= assemble(L1)
[bc.apply(A1, b1) for bc in bcu]
solve(A1, w12.vector(), b1, "bicgstab", "default")
end()
(u12, D12_vec) = w12.split()
D12 = as_matrix([[D12_vec[0], D12_vec[1]],
                [D12_vec[1], D12_vec[2]]])
DEVSSr_u1 = 2*(1-betav)*inner(D12,Dincomp(v))*dx
lhsFus = Re*rho0*((u - u0)/dt + conv*dot(u12, nabla_grad(u12)))
Fus = dot(lhsFus, v)*dx + \\
    + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \\
    + 0.5*dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \\
    - dot(tau0*n, v)*ds\\
    +  inner(D-Dcomp(u),R)*dx
a2= lhs(Fus)
L2= rhs(Fus)
a2+= th*DEVSSl_u1
L2+= th*DEVSSr_u1
A2 = assemble(a2)
b2 = assemble(L2)
[bc.apply(A2, b2) for bc in bcu]
solve(A2, ws.vector(), b2, "bicgstab", "default")
end()
(us, Ds_vec) = ws.split()
lhs_p_1 = (Ma*Ma/(dt*Re))*p
rhs_p_1 = (Ma*Ma/(dt*Re))*p0 - rho0*div(us) + dot(grad(rho0),us)
lhs_p_2 = 0.5*dt*grad(p)
rhs_p_2 = 0.5*dt*grad(p0)
a5=inner(lhs_p_1,q)*dx + inner(lhs_p_2,grad(q))*dx
L5=inner(rhs_p_1,q)*dx + inner(rhs_p_2,grad(q))*dx
A5 = assemble(a5)
b5 = assemble(L5)
[bc.apply(A5, b5) for bc in bcp]
solve(A5, p1.vector(), b5, "bicgstab", "default")
end()
rho1 = rho0 + (Ma*Ma/Re)*(p1-p0)
rho1 = project(rho1,Q)
lhs_u1 = (Re/dt)*rho1*u
rhs_u1 = (Re/dt)*rho1*us
a7=inner(lhs_u1,v)*dx  + inner(D-Dcomp(u),R)*dx
L7=inner(rhs_u1,v)*dx + 0.5*inner(p1-p0,div(v))*dx - 0.5*dot(p1*n, v)*ds
a7+= 0
L7+= 0
A7 = assemble(a7)
b7 = assemble(L7)
[bc.apply(A7, b7) for bc in bcu]
solve(A7, w1.vector(), b7, "bicgstab", "default")
end()
(u1, D1_vec) = w1.split()
lhs_tau1 = (We/dt+1.0)*tau  +  We*Fdef(u1,tau)
rhs_tau1= (We/dt)*tau0 + 2.0*(1.0-betav)*Dcomp(u1)
A = inner(lhs_tau1,Rt)*dx - inner(rhs_tau1,Rt)*dx
a4 = lhs(A)
L4 = rhs(A)
a4 += LPSl_stress
L4 += 0
A4=assemble(a4)
b4=assemble(L4)
[bc.apply(A4, b4) for bc in bctau]
solve(A4, tau1_vec.vector(), b4, "bicgstab", "default")
end()
E_k=assemble(0.5*rho1*dot(u1,u1)*dx)
E_e=assemble((tau1_vec[0]+tau1_vec[2])*dx)
'''


def test_literal_config2_time_step_equals_the_driver_step():
    """One whole time step of the bench configuration executed from the script's statements (assemble / bc.apply /
    solve / project, 6 linear systems + the density projection) against the driver's step from the same state."""
    import fenics_b200.shims.lid_driven_cavity.fenics_fem as ff
    mk = lambda: ldc.CompLDC(pmesh.Skew_Mesh(12, 1, 1))  # noqa: E731
    drv, twin = mk(), mk()
    drv.t = 0.45
    drv.step()
    drv.step()
    for slot, fn in drv.pr.fields.items():                 # the twin starts from the same state (all bound Functions)
        twin.pr.fields[slot].vector().t.copy_(fn.vector().t)
    twin.t = drv.t
    twin.solve_rtol = 1e-12
    want = twin.step()
    # ---- the script's set-up section on the driver's Functions (:201-217, :343-349) and the top of the loop (:392-412)
    ns = {k: getattr(ff, k) for k in dir(ff) if not k.startswith("_")}
    S, pv = drv.S, drv.p
    ns.update({k: ff.Parameter(k, pv[k]) for k in ("Re", "We", "Ma", "dt", "betav", "th", "conv", "c1", "c2")})
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = ff.trial_functions(Q, Zc, W)
    (v, R_vec) = ff.TestFunctions(W)
    ns.update(Q=Q, p=p, rho=p, q=ff.TestFunction(Q), u=u, D=D, tau=tau, v=v, R=m3(R_vec), Rt=m3(ff.TestFunction(Zc)),
              h=ff.CellDiameter(drv.mesh), n=ff.FacetNormal(drv.mesh), bcu=drv.bcu, bcp=[], bctau=[])
    for name in ("w0", "w12", "ws", "w1", "p0", "p1", "rho0", "rho12", "tau0_vec", "tau1_vec", "kapp"):
        ns[name] = getattr(drv, name)
    drv._bct["t"] = drv.t                                  # ulidreg.t = t (:379)
    drv.lps_indicator()                                    # :392-401 -> kapp
    (ns["u0"], D0_vec) = drv.w0.split()                    # :408-412
    ns["D0"], ns["tau0"] = m3(D0_vec), m3(symbolic.as_expr(drv.tau0_vec))
    exec("DEVSSl_u12 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx\n"
         "DEVSSl_u1 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx\n"
         "LPSl_stress = inner(kapp*h*c1*grad(tau),grad(Rt))*dx + inner(kapp*h*c2*div(tau),div(Rt))*dx\n"
         "DEVSSr_u12 = 2*(1-betav)*inner(D0,Dincomp(v))*dx\n", ns)
    ks = la.parameters["krylov_solver"]
    old = ks["relative_tolerance"], la.parameters.get("preconditioner_substitution")
    ks["relative_tolerance"], la.parameters["preconditioner_substitution"] = 1e-12, "silent"
    try:
        exec(CFG2_LOOP_BODY, ns)
    finally:
        ks["relative_tolerance"], la.parameters["preconditioner_substitution"] = old
    for name, ref in (("w1", twin.w1), ("p1", twin.p1), ("tau1_vec", twin.tau1_vec), ("rho1", twin.rho1),
                      ("w12", twin.w12), ("ws", twin.ws)):
        assert lc.rel_err(ns[name].vector().get_local(), ref.vector().get_local()) < 1e-8, name
    assert abs(ns["E_k"] - want["E_k"]) < 1e-9 * abs(want["E_k"])
    assert abs(ns["E_e"] - want["E_e"]) < 1e-9 * max(abs(want["E_e"]), 1e-12)


def test_literal_sweep_script_right_hand_sides_bind_re_times_conv():
    """comp_LDC.py:418-431 on the GPU: the literal L1 with `Re*conv*dot(u0, nabla_grad(u0))` equals the sweep driver's
    (CompLDCSweep: the functor's conv slot holds Re*conv), and differs from the mesh-convergence form with the same conv."""
    from fenics_b200.shims.lid_driven_cavity.fenics_fem import (Dcomp, Dincomp, FacetNormal, Parameter, TestFunctions,
                                                                assemble, div, dot, ds, dx, inner, nabla_grad, rhs, sigma,
                                                                trial_functions)
    drv = ldc.CompLDCSweep(pmesh.Skew_Mesh(12, 1, 1), Re=25.0, conv=0.5)
    drv.t = 0.45
    drv.step()
    drv.step()
    S, pv = drv.S, drv.p
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    Re, We, dt, betav, th = (Parameter(k, pv[k]) for k in ("Re", "We", "dt", "betav", "th"))
    conv = Parameter("conv", 0.5)                          # the script's conv (the driver's table holds Re*conv = 12.5)
    assert pv["conv"] == 12.5
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = trial_functions(Q, Zc, W)
    (v, R_vec) = TestFunctions(W)
    R, n = m3(R_vec), FacetNormal(drv.mesh)
    rho0, p0 = drv.rho0, drv.p0
    (u0, D0_vec) = drv.w0.split()
    D0, tau0 = m3(D0_vec), m3(symbolic.as_expr(drv.tau0_vec))
    DEVSSr_u12 = 2*(1-betav)*inner(D0,Dincomp(v))*dx
    U = 0.5*(u + u0)
    lhsFu12 = Re*rho0*(2.0*(u - u0) / dt + Re*conv*dot(u0, nabla_grad(u0)))
    Fu12 = dot(lhsFu12, v)*dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \
        + dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \
        - dot(tau0*n, v)*ds \
        +  inner(D-Dcomp(u),R)*dx
    L1 = rhs(Fu12)
    L1+= th*DEVSSr_u12
    got = assemble(L1).t.cpu().numpy()
    want = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_L1)).t.cpu().numpy()
    assert lc.rel_err(got, want) < 1e-13
    lhs_mc = Re*rho0*(2.0*(u - u0) / dt + conv*dot(u0, nabla_grad(u0)))   # the mesh-convergence form with conv = 0.5
    F_mc = dot(lhs_mc, v)*dx + inner(sigma(U, p0, tau0, We, betav), Dincomp(v))*dx \
        + dot(p0*n, v)*ds - dot(betav*nabla_grad(U)*n, v)*ds - (1.0/3)*betav*dot(div(U)*n,v)*ds \
        - dot(tau0*n, v)*ds + inner(D-Dcomp(u),R)*dx
    other = assemble(rhs(F_mc) + th*DEVSSr_u12).t.cpu().numpy()
    assert lc.rel_err(other, want) > 1e-6


def test_literal_config5_forms_equal_the_driver():
    """buoyancy-driven-flow/compressible_dgp.py:405-583: the forms of the buoyancy-driven loop written as the script
    writes them (sigmacom, the Boussinesq source Ra*Pr*rho0*therm_inv*k, `lhs` / `rhs`, ds(3) / ds(4) / ds(2)) assemble
    to what the driver's fx_form ids give; so do the four projections the script hands to project()."""
    import fenics_b200.shims.buoyancy_driven_flow.fenics_fem as ff
    from fenics_b200.drivers import dgp
    drv = dgp.CompDGP(mesh_resolution=8)
    drv.t = 1.4
    for _ in range(3):
        drv.step()
    ns = {k: getattr(ff, k) for k in dir(ff) if not k.startswith("_")}
    S, pv = drv.S, drv.p
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    P = ff.Parameter
    ns.update({k: P(k, pv[k]) for k in ("Ra", "Pr", "We", "Ma", "dt", "betav", "th", "Vh", "Bi", "c1", "c2")})
    ns.update(T_0=P("T0", pv["T0"]), T_h=P("th_hot", pv["th_hot"]), al_1=P("al1", pv["al1"]), al_2=P("al2", pv["al2"]))
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = ff.trial_functions(Q, Zc, W)
    (v, R_vec) = ff.TestFunctions(W)
    q = ff.TestFunction(Q)
    ns.update(Q=Q, p=p, rho=p, T=p, q=q, r=q, u=u, D=D, tau=tau, v=v, R=m3(R_vec), Rt=m3(ff.TestFunction(Zc)),
              h=ff.CellDiameter(drv.mesh), n=ff.FacetNormal(drv.mesh), k=ff.as_vector([0.0, 1.0]))
    for name in ("w0", "w12", "ws", "w1", "p0", "p1", "rho0", "rho12", "rho1", "T0", "T1", "theta0", "therm_inv", "theta1",
                 "tau0_vec", "tau1_vec", "kapp"):
        ns[name] = getattr(drv, name)
    thetar = dapi.Function(Q, drv.plan)
    thetar.vector().t.fill_(pv["T0"] / (pv["th_hot"] - pv["T0"]))        # project(T_0/(T_h-T_0), Q): exact for a constant
    ns["thetar"] = thetar
    (ns["u0"], D0_vec), (ns["u12"], D12_vec), (ns["us"], _), (ns["u1"], _) = (drv.w0.split(), drv.w12.split(), drv.ws.split(),
                                                                            drv.w1.split())
    ns.update(D0=m3(D0_vec), D12=m3(D12_vec), tau0=m3(symbolic.as_expr(drv.tau0_vec)), tau1=m3(symbolic.as_expr(drv.tau1_vec)))
    exec('''
thetal = (T)/(T_h-T_0)
DEVSSl_u12 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
DEVSSl_u1 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
LPSl_stress = inner(kapp*h*c1*grad(tau),grad(Rt))*dx + inner(kapp*h*c2*div(tau),div(Rt))*dx
therm = (al_1+al_2*theta0)
DEVSSr_u12 = 2*(1-betav)*inner(D0,Dincomp(v))*dx
U = 0.5*(u + u0)
rho_eq = 2.0*(rho - rho0)/dt + dot(u0, grad(rho0)) - rho0*div(u0)
rho_weak = inner(rho_eq,q)*dx
a0 = lhs(rho_weak)
L0 = rhs(rho_weak)
Du12Dt = rho0*(2.0*(u - u0) / dt + dot(u0, nabla_grad(u0)))
Fu12 = dot(Du12Dt, v)*dx + \\
    + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v))*dx + Ra*Pr*inner(rho0*therm_inv*k,v)*dx \\
    + dot(p0*n, v)*ds - betav*Pr*(dot(nabla_grad(U)*n, v)*ds + (1.0/3)*dot(div(U)*n,v)*ds)\\
    - (Pr*(1-betav)/We)*dot(tau0*n, v)*ds\\
    + inner(D-Dincomp(u),R)*dx
a1 = lhs(Fu12)
L1 = rhs(Fu12)
a1+= th*DEVSSl_u12
L1+= th*DEVSSr_u12
DEVSSr_u1 = 2*(1-betav)*inner(D12,Dincomp(v))*dx
lhsFus = rho0*((u - u0)/dt + dot(u12, nabla_grad(U)))
Fus = dot(lhsFus, v)*dx + \\
    + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v))*dx + Ra*Pr*inner(rho0*therm_inv*k,v)*dx \\
    + dot(p0*n, v)*ds - betav*Pr*(dot(nabla_grad(U)*n, v)*ds + (1.0/3)*dot(div(U)*n,v)*ds) \\
    - (Pr*(1.0-betav)/We)*dot(tau0*n, v)*ds\\
    + inner(D-Dincomp(u),R)*dx
a2= lhs(Fus)
L2= rhs(Fus)
a2+= th*DEVSSl_u1
L2+= th*DEVSSr_u1
lhs_p_1 = (Ma*Ma/(dt))*p
rhs_p_1 = (Ma*Ma/(dt))*p0 - therm*dot(grad(rho0),us) - therm*rho0*div(us)
lhs_p_2 = therm*0.5*dt*grad(p)
rhs_p_2 = therm*0.5*dt*grad(p0)
a5=inner(lhs_p_1,q)*dx + inner(lhs_p_2,grad(q))*dx
L5=inner(rhs_p_1,q)*dx + inner(rhs_p_2,grad(q))*dx
alpha_supg = h/(magnitude(u12)+DOLFIN_EPS)
SU_rho = inner(dot(u12, grad(rho)), alpha_supg*dot(u12,grad(q)))*dx
rho_eq = (rho - rho0)/dt + dot(u12, grad(rho12)) - rho12*div(u12)
rho_weak = inner(rho_eq,q)*dx
a6 = lhs(rho_weak)
L6 = rhs(rho_weak)
a6+= SU_rho
lhs_u1 = (1./dt)*rho1*u
rhs_u1 = (1./dt)*rho0*us
a7=inner(lhs_u1,v)*dx +inner(D-Dincomp(u),R)*dx
L7=inner(rhs_u1,v)*dx + 0.5*(inner(p1-p0,div(v))*dx - dot((p1-p0)*n, v)*ds)
stress_eq = ((We/dt)+1.0)*tau  +  We*Fdefcom(u1,tau) - (We/dt)*tau0 - Identity(len(u))
A = inner(stress_eq,Rt)*dx
a4 = lhs(A)
L4 = rhs(A)
a4 += LPSl_stress
gamdot = inner(sigmacom(u1, p1, tau1, We, Pr, betav), grad(u1))
lhs_theta1 = (1.0/dt)*rho1*thetal + rho1*dot(u1,grad(thetal))
rhs_theta1 = (1.0/dt)*rho0*thetar + rho1*dot(u1,grad(thetar)) + (1.0/dt)*rho0*theta0 + Vh*gamdot
a8 = inner(lhs_theta1,r)*dx + inner(grad(thetal),grad(r))*dx
L8 = inner(rhs_theta1,r)*dx + inner(grad(thetar),grad(r))*dx + Bi*inner(grad(theta0),n*r)*ds(3) + Bi*inner(grad(theta0),n*r)*ds(4) + inner(We*tau1*grad(thetar),grad(r))*dx
Tdx = inner(grad(theta1),n)
''', ns)
    for name in ("a0", "L0", "a1", "L1", "a2", "L2", "a5", "L5", "a6", "L6", "a7", "L7", "a4", "L4", "a8", "L8"):
        fid = getattr(abi, "FX_FORM_C5_" + name.upper())
        got, want = ff.assemble(ns[name]), dapi.assemble(drv.pr.form(fid))
        a, b = (got.val, want.val) if name.startswith("a") else (got.t, want.t)
        assert lc.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-13, name
        assert float(b.abs().max()) > 0.0, name
    for expr, fid in ((ns["rho0"] + (ns["Ma"]*ns["Ma"])*ns["therm_inv"]*(ns["p1"]-ns["p0"]), abi.FX_FORM_C5_L_RHO1),
                      ((ns["T0"]-ns["T_0"])/(ns["T_h"]-ns["T_0"]), abi.FX_FORM_C5_L_THETA0),
                      ((ns["T1"]-ns["T_0"])/(ns["T_h"]-ns["T_0"]), abi.FX_FORM_C5_L_THETA1),
                      (1.0/ns["therm"], abi.FX_FORM_C5_L_THERM_INV)):
        got = ff.assemble(ff.inner(q, expr) * ff.dx).t.cpu().numpy()
        want = dapi.assemble(drv.pr.form(fid)).t.cpu().numpy()
        assert lc.rel_err(got, want) < 1e-13, fid
    for form, fid in ((0.5*ns["rho1"]*ff.dot(ns["u1"], ns["u1"])*ff.dx, abi.FX_FORM_C5_EK),
                      ((ns["tau1"][0, 0]+ns["tau1"][1, 1]-2.0)*ff.dx, abi.FX_FORM_C5_EE), (-ns["Tdx"]*ff.ds(2), abi.FX_FORM_C5_NUS)):
        got, want = ff.assemble(form), dapi.assemble(drv.pr.form(fid))
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-12), fid


CFG4_FORMS = '''
thetal = (T)/(T_h-T_0)
theta0 = (T0-T_0)/(T_h-T_0)
DEVSSl_T1 = (1.-Di)*inner(grad(thetal), grad(r))*dx
DEVSSr_T1 = inner((1.-Di)*(grad(thetar) + grad(theta0)), grad(r))*dx
DEVSSGl_u1 = 2.0*(1.-betav)*inner(Dincomp(u),Dincomp(v))*dx
LPSl_stress = inner(kapp*h*0.05*grad(tau),grad(Rt))*dx + inner(kapp*h*0.01*div(tau),div(Rt))*dx
DEVSSGr_u1 = (1.-betav)*inner(D0 + D0.T,Dincomp(v))*dx
U = 0.5*(u + u0)
rho_eq = 2.0*(rho - rho0)/dt + dot(u0, grad(rho0)) - rho0*div(u0)
rho_weak = inner(rho_eq,q)*dx
a0 = lhs(rho_weak)
L0 = rhs(rho_weak)
lhsFus = Re*rho0*((u - u0)/dt + conv*dot(u0, nabla_grad(U)))
Fus = dot(lhsFus, v)*dx + \\
    + inner(2.0*betav*phi_s(theta0,A_0)*Dincomp(U), Dincomp(v))*dx + (1.0/3)*inner(betav*phi_s(theta0,A_0)*div(U),div(v))*dx \\
        - ((1.-betav)/(We+DOLFIN_EPS))*inner(div( phi_def(u0, lambda_d)*(fene_func(tau0, b)*tau0 - Identity(len(u)) ) ), v)*dx + inner(grad(p0),v)*dx\\
    + inner(D-grad(u),R)*dx
a2= lhs(Fus)
L2= rhs(Fus)
a2+= th*DEVSSGl_u1
L2+= th*DEVSSGr_u1
lhs_p_1 = (Ma*Ma/(dt))*p
rhs_p_1 = (Ma*Ma/(dt))*p0 - Re*div(rho0*us)
lhs_p_2 = dt*grad(p)
rhs_p_2 = dt*grad(p0)
a5=inner(lhs_p_1,q)*dx + inner(lhs_p_2,grad(q))*dx
L5=inner(rhs_p_1,q)*dx + inner(rhs_p_2,grad(q))*dx
alpha_supg = h/(magnitude(u12)+0.0000000001)
SU_rho = inner(dot(u12, grad(rho)), alpha_supg*dot(u12,grad(q)))*dx
rho_eq = (rho - rho0)/dt + dot(u12, grad(rho12)) - rho12*div(u12)
rho_weak = inner(rho_eq,q)*dx
a6 = lhs(rho_weak)
L6 = rhs(rho_weak)
a6+= SU_rho
lhs_u1 = (Re/dt)*rho1*u
rhs_u1 = (Re/dt)*rho0*us
a7=inner(lhs_u1,v)*dx + inner(D-grad(u),R)*dx
L7=inner(rhs_u1,v)*dx - 0.5*inner(grad(p1-p0),(v))*dx
a7+= th*DEVSSGl_u1
L7+= th*DEVSSGr_u1
lhs_tau1 = (We/dt)*tau + fene_func(tau0, b)*tau + We*FdefG(u1, D1, tau) - We*0.5*(phi_def(u1, lambda_d)-1.)*(tau*Dincomp(u1) + Dincomp(u1)*tau)
rhs_tau1= (We/dt)*tau0  + Identity(len(u))
Astress = inner(lhs_tau1,Rt)*dx - inner(rhs_tau1,Rt)*dx
a4 = lhs(Astress)
L4 = rhs(Astress)
a4 += LPSl_stress
gamdot = inner(fene_sigmacom(u0, p0, tau0, b, lambda_d, betav, We),grad(u0))
lhs_temp1 = (1.0/dt)*rho1*thetal + rho1*dot(u1,grad(thetal))
difflhs_temp1 = Di*grad(thetal) + (Di/We)*tau1*grad(thetal)
rhs_temp1 = (1.0/dt)*rho1*thetar + rho1*dot(u1,grad(thetar)) + (1.0/dt)*rho1*theta0 + Vh*gamdot
diffrhs_temp1 = Di*grad(thetar) + (Di/We)*tau1*grad(thetar)
a9 = inner(lhs_temp1,r)*dx + inner(difflhs_temp1,grad(r))*dx
L9 = inner(rhs_temp1,r)*dx + inner(diffrhs_temp1,grad(r))*dx - Di*Bi*inner(theta0,r)*ds(1)
a9+= 0.0*th*DEVSSl_T1
L9+= 0.0*th*DEVSSr_T1
E_k = 0.5*rho1*dot(u1,u1)*dx
E_e = (fene_func(tau1, b)*(tau1[0,0]+tau1[1,1])-2.)*dx
sigma0 = dot(fene_sigmacom(u1, p1, tau1, b,lambda_d, betav, We), tang)
omegaf0 = dot(fene_sigmacom(u1, p1, tau1, b,lambda_d, betav, We), n)
innerforcex = inner(Constant((1.0, 0.0)), omegaf0)*ds(0)
innerforcey = inner(Constant((0.0, -1.0)), omegaf0)*ds(0)
innertorque = -inner(n, sigma0)*ds(0)
'''


@pytest.mark.parametrize("lam", [0.1, 0.0])
def test_literal_config4_forms_equal_the_driver(lam):
    """fenepmp_jbp.py:489-649: the FENE-P-MP journal-bearing forms written as the script writes them (phi_s, fene_func,
    phi_def inside a divergence, FdefG, the journal torque / load on ds(0)) assemble to what the driver's fx_form ids
    give -- with lambda_d a Parameter (lambda_D = 0.1: the C4M functors) and with the plain 0.0 (the folded C4 ones)."""
    import fenics_b200.shims.eccentric_cylinders.fenics_base as fb
    from fenics_b200.drivers import jbp
    drv = jbp.FenepmpJBP(n_theta=32, n_r=6, lambda_d=lam, We=0.5)
    drv.t = 0.45
    for _ in range(3):
        drv.step()
    ns = {k: getattr(fb, k) for k in dir(fb) if not k.startswith("_")}
    S, pv = drv.S, drv.p
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    P = fb.Parameter
    ns.update({k: P(k, pv[k]) for k in ("Re", "We", "Ma", "dt", "betav", "th", "conv", "Di", "Vh", "Bi")})
    ns.update(T_0=P("T0", pv["T0"]), T_h=P("th_hot", pv["th_hot"]), b=P("b_fene", pv["b_fene"]), A_0=P("a0", pv["a0"]),
              lambda_d=P("lambda_d", lam) if lam else 0.0)
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = fb.trial_functions(Q, Zc, W)
    (v, R_vec) = fb.TestFunctions(W)
    q = fb.TestFunction(Q)
    n = fb.FacetNormal(drv.mesh)
    ns.update(Q=Q, p=p, rho=p, T=p, q=q, r=q, u=u, D=D, tau=tau, v=v, R=m3(R_vec), Rt=m3(fb.TestFunction(Zc)),
              h=fb.CellDiameter(drv.mesh), n=n, tang=fb.as_vector([n[1], -n[0]]))
    for name in ("w0", "w12", "ws", "w1", "p0", "p1", "rho0", "rho12", "rho1", "T0", "T1", "tau0_vec", "tau1_vec", "kapp"):
        ns[name] = getattr(drv, name)
    thetar = dapi.Function(Q, drv.plan)
    thetar.vector().t.fill_(pv["T0"] / (pv["th_hot"] - pv["T0"]))
    ns["thetar"] = thetar
    (ns["u0"], D0_vec), (ns["u12"], _), (ns["us"], _), (ns["u1"], D1_vec) = (drv.w0.split(), drv.w12.split(), drv.ws.split(),
                                                                           drv.w1.split())
    ns.update(D0=m3(D0_vec), D1=m3(D1_vec), tau0=m3(symbolic.as_expr(drv.tau0_vec)), tau1=m3(symbolic.as_expr(drv.tau1_vec)))
    exec(CFG4_FORMS, ns)
    A = abi
    ids = dict(a0=A.FX_FORM_C4_A0, L0=A.FX_FORM_C4_L0, a2=A.FX_FORM_C4_A2, L2=drv.F_L2, a5=A.FX_FORM_C4_A5, L5=A.FX_FORM_C4_L5,
               a6=A.FX_FORM_C4_A6, L6=A.FX_FORM_C4_L6, a7=A.FX_FORM_C4_A7, L7=A.FX_FORM_C4_L7, a4=drv.F_A4, L4=A.FX_FORM_C4_L4,
               a9=A.FX_FORM_C4_A9, L9=drv.F_L9)
    for name, fid in ids.items():
        got, want = fb.assemble(ns[name]), dapi.assemble(drv.pr.form(fid))
        a, b = (got.val, want.val) if name.startswith("a") else (got.t, want.t)
        assert lc.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-13, name
        assert float(b.abs().max()) > 0.0, name
    for name, fid in (("E_k", A.FX_FORM_C4_EK), ("E_e", A.FX_FORM_C4_EE), ("innertorque", drv.F_JOURNAL[0]),
                      ("innerforcex", drv.F_JOURNAL[1]), ("innerforcey", drv.F_JOURNAL[2])):
        got, want = fb.assemble(ns[name]), dapi.assemble(drv.pr.form(fid))
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-10), name


CFG3_FORMS = '''
thetal = (T)/(T_h-T_0)
theta0 = (T0-T_0)/(T_h-T_0)
theta1 = T1-T_0/(T_h-T_0)
DEVSSl_u12 = 2*(1-betav)*inner(Dcomp(u),Dincomp(v))*dx
DEVSSr_u12 = 2*inner(D0,Dincomp(v))*dx
DEVSSl_T1 = (1.-Di)*inner(grad(thetal), grad(r))*dx
DEVSSr_T1 = inner((1.-Di)*(grad(thetar) + grad(theta0)), grad(r))*dx
DEVSSGl_u1 = 2.0*(1.-betav)*inner(Dincomp(u),Dincomp(v))*dx
alpha_supg = h/(magnitude(u1) + 0.0000000001)
SU = inner(dot(u1, grad(tau)), alpha_supg*dot(u1,grad(Rt)))*dx
DEVSSGr_u1 = (1.-betav)*inner(D0 + D0.T,Dincomp(v))*dx
U = 0.5*(u + u0)
rho_eq = 2.0*(rho - rho0)/dt + dot(u0, grad(rho0)) - rho0*div(u0)
rho_weak = inner(rho_eq,q)*dx
a0 = lhs(rho_weak)
L0 = rhs(rho_weak)
lhsv_eq = 2.0*Re*((rho12*u - rho0*u0)/dt + conv*dot(u0, nabla_grad(U)))
v_weak12 = dot(lhsv_eq, v)*dx + \\
    + inner(2.0*betav*phi_s(theta0,A_0)*Dincomp(U), Dincomp(v))*dx + (1.0/3)*inner(betav*phi_s(theta0,A_0)*div(U),div(v))*dx \\
    - ((1.-betav)/(We+DOLFIN_EPS))*inner(div(tau0-rho0*Identity(len(u))), v)*dx + inner(grad(p0),v)*dx\\
    + inner(D-grad(u),R)*dx
a1 = lhs(v_weak12)
L1 = rhs(v_weak12)
a1+= th*DEVSSl_u12
L1+= th*DEVSSr_u12
lhsFus = Re*rho0*((u - u0)/dt + conv*dot(u0, nabla_grad(U)))
Fus = dot(lhsFus, v)*dx + \\
    + inner(2.0*betav*phi_s(theta0,A_0)*Dincomp(U), Dincomp(v))*dx + (1.0/3)*inner(betav*phi_s(theta0,A_0)*div(U),div(v))*dx \\
    - ((1.-betav)/(We+DOLFIN_EPS))*inner(div(tau0-rho0*Identity(len(u))), v)*dx + inner(grad(p0),v)*dx\\
    + inner(D-grad(u),R)*dx
a2= lhs(Fus)
L2= rhs(Fus)
a2+= th*DEVSSGl_u1
L2+= th*DEVSSGr_u1
lhs_p_1 = (Ma*Ma/(dt))*p
rhs_p_1 = (Ma*Ma/(dt))*p0 - Re*(1+al*theta0)*div(rho0*us)
lhs_p_2 = (1+al*theta0)*dt*grad(p)
rhs_p_2 = (1+al*theta0)*dt*grad(p0)
a5=inner(lhs_p_1,q)*dx + inner(lhs_p_2,grad(q))*dx
L5=inner(rhs_p_1,q)*dx + inner(rhs_p_2,grad(q))*dx
alpha_supg = h/(magnitude(u12)+0.0000000001)
SU_rho = inner(dot(u12, grad(rho)), alpha_supg*dot(u12,grad(q)))*dx
rho_eq = (rho - rho0)/dt + dot(u12, grad(rho12)) - rho12*div(u12)
rho_weak = inner(rho_eq,q)*dx
a6 = lhs(rho_weak)
L6 = rhs(rho_weak)
a6+= SU_rho
lhs_u1 = (Re/dt)*rho1*u
rhs_u1 = (Re/dt)*rho0*us
a7=inner(lhs_u1,v)*dx + inner(D-grad(u),R)*dx
L7=inner(rhs_u1,v)*dx - 0.5*inner(grad(p1-p0),(v))*dx
a7+= 0
L7+= 0
gamdot = inner(sigmacon(u0, p0, tau0, betav, We),grad(u0))
lhs_temp1 = (1.0/dt)*rho1*thetal + rho1*dot(u1,grad(thetal))
difflhs_temp1 = Di*grad(thetal)
rhs_temp1 = (1.0/dt)*rho1*thetar + rho1*dot(u1,grad(thetar)) + (1.0/dt)*rho1*theta0 + Vh*phi_ewm(tau1, theta1, k_ewm, B)*gamdot
diffrhs_temp1 = Di*grad(thetar)
a9 = inner(lhs_temp1,r)*dx + inner(difflhs_temp1,grad(r))*dx
L9 = inner(rhs_temp1,r)*dx + inner(diffrhs_temp1,grad(r))*dx - Di*Bi*inner(theta0,r)*ds(1)
a9+= th*DEVSSl_T1
L9+= th*DEVSSr_T1
lhs_tau1 = (We*phi_ewm(tau0, theta1, k_ewm, B)/dt+1.0)*tau  +  We*phi_ewm(tau0, theta1, k_ewm, B)*FdefG(u1, D1, tau)
rhs_tau1= (We*phi_ewm(tau0, theta1, k_ewm, B)/dt)*tau0 + rho0*Identity(len(u))
Ftau = inner(lhs_tau1,Rt)*dx - inner(rhs_tau1,Rt)*dx
a4 = lhs(Ftau)
L4 = rhs(Ftau)
a4_stab = a4 + SU
L4_stab = L4
E_k = 0.5*rho1*dot(u1,u1)*dx
E_e = (tau1_vec[0]+tau1_vec[2]-2.0)*dx
sigma0 = dot(sigmacon(u1, p1, tau1, betav, We), tang)
omegaf0 = dot(sigmacon(u1, p1, tau1, betav, We), n)
innerforcex = inner(Constant((1.0, 0.0)), omegaf0)*ds(0)
innerforcey = -inner(Constant((0.0, 1.0)), omegaf0)*ds(0)
innertorque = -inner(n, sigma0)*ds(0)
'''


@pytest.mark.parametrize("general", [True, False])
def test_literal_config3_forms_equal_the_driver(general):
    """incompressible_benchmarkEWM.py:488-706: the EWM journal-bearing forms written as the script writes them assemble to
    what the driver's fx_form ids give -- with A_0, k_ewm, B wrapped as Parameters (material functions active: the C3E
    functors) and with the script's plain 0.0 (phi_s = phi_ewm = 1: the folded C3 functors)."""
    import fenics_b200.shims.eccentric_cylinders.fenics_base as fb
    from fenics_b200.drivers import ewm
    kw = dict(A_0=0.4, k_ewm=0.7, B=0.2) if general else {}
    drv = ewm.EwmJBP(n_theta=32, n_r=6, prepass=True, We=0.3, Ma=0.05, betav=0.6, dt=0.002, **kw)
    drv.t = 0.45
    for _ in range(3):
        drv.step()
    assert drv.ewm == general
    ns = {k: getattr(fb, k) for k in dir(fb) if not k.startswith("_")}
    S, pv = drv.S, drv.p
    W, Zc, Q = S["W"], S["Zc"], S["Q"]
    P = fb.Parameter
    ns.update({k: P(k, pv[k]) for k in ("Re", "We", "Ma", "dt", "betav", "th", "conv", "Di", "Vh", "Bi")})
    ns.update(T_0=P("T0", pv["T0"]), T_h=P("th_hot", pv["th_hot"]), al=P("al1", pv["al1"]),
              A_0=P("a0", pv["a0"]) if general else 0.0, k_ewm=P("k_ewm", pv["k_ewm"]) if general else 0.0,
              B=P("b_ewm", pv["b_ewm"]) if general else 0.0)
    m3 = lambda v: symbolic.as_matrix([[v[0], v[1]], [v[1], v[2]]])  # noqa: E731
    _, p, _, tau_vec, u, D_vec, D, tau = fb.trial_functions(Q, Zc, W)
    (v, R_vec) = fb.TestFunctions(W)
    q = fb.TestFunction(Q)
    n = fb.FacetNormal(drv.mesh)
    ns.update(Q=Q, p=p, rho=p, T=p, q=q, r=q, u=u, D=D, tau=tau, v=v, R=m3(R_vec), Rt=m3(fb.TestFunction(Zc)),
              h=fb.CellDiameter(drv.mesh), n=n, tang=fb.as_vector([n[1], -n[0]]))
    for name in ("w0", "w12", "ws", "w1", "p0", "p1", "rho0", "rho12", "rho1", "T0", "T1", "tau0_vec", "tau1_vec"):
        ns[name] = getattr(drv, name)
    thetar = dapi.Function(Q, drv.plan)
    thetar.vector().t.fill_(pv["T0"] / (pv["th_hot"] - pv["T0"]))
    ns["thetar"] = thetar
    (ns["u0"], D0_vec), (ns["u12"], _), (ns["us"], _), (ns["u1"], D1_vec) = (drv.w0.split(), drv.w12.split(), drv.ws.split(),
                                                                           drv.w1.split())
    ns.update(D0=m3(D0_vec), D1=m3(D1_vec), tau0=m3(symbolic.as_expr(drv.tau0_vec)), tau1=m3(symbolic.as_expr(drv.tau1_vec)))
    exec(CFG3_FORMS, ns)
    A = abi
    E = (lambda x, y: x) if general else (lambda x, y: y)
    ids = dict(a0=A.FX_FORM_C3_A0, L0=A.FX_FORM_C3_L0, a1=E(A.FX_FORM_C3E_A1, A.FX_FORM_C3_A1), L1=E(A.FX_FORM_C3E_L1, A.FX_FORM_C3_L1),
               a2=A.FX_FORM_C3_A2, L2=E(A.FX_FORM_C3E_L2, A.FX_FORM_C3_L2), a5=A.FX_FORM_C3_A5, L5=A.FX_FORM_C3_L5,
               a6=A.FX_FORM_C3_A6, L6=A.FX_FORM_C3_L6, a7=A.FX_FORM_C3_A7, L7=A.FX_FORM_C3_L7, a9=A.FX_FORM_C3_A9,
               L9=E(A.FX_FORM_C3E_L9, A.FX_FORM_C3_L9), a4_stab=E(A.FX_FORM_C3E_A4, A.FX_FORM_C3_A4),
               L4_stab=E(A.FX_FORM_C3E_L4, A.FX_FORM_C3_L4))
    for name, fid in ids.items():
        drv.pr.set_param(A.FX_P_EPS, drv.SU_RHO_EPS if name == "a6" else drv.SU_EPS)
        got, want = fb.assemble(ns[name]), dapi.assemble(drv.pr.form(fid))
        a, b = (got.val, want.val) if name.startswith("a") else (got.t, want.t)
        assert lc.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-13, name
        assert float(b.abs().max()) > 0.0, name
    for name, fid in (("E_k", A.FX_FORM_C3_EK), ("E_e", A.FX_FORM_C3_EE), ("innertorque", A.FX_FORM_C3_TORQUE),
                      ("innerforcex", A.FX_FORM_C3_FORCE_X), ("innerforcey", A.FX_FORM_C3_FORCE_Y)):
        got, want = fb.assemble(ns[name]), dapi.assemble(drv.pr.form(fid))
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-10), name


def test_literal_lps_indicator_block_equals_the_driver():
    """incomp_LDC.py:298-308 as the script writes it -- four `project` calls (Zd, Zc, Qt, Qt) and numpy's power on a
    Function -- against the driver's lps_indicator() from the same state."""
    import fenics_b200.shims.lid_driven_cavity.fenics_fem as ff
    from fenics_b200.shims.lid_driven_cavity.fenics_fem import (Dcomp, Parameter, as_vector, div, dot, grad, inner,  # noqa: F401
                                                                project, tgrad)
    drv = ldc.IncompLDC(pmesh.Skew_Mesh(12, 1, 1))
    drv.t = 0.45
    drv.step()
    drv.step()
    S, pv = drv.S, drv.p
    Zd, Zc, Qt = S["Zd"], S["Zc"], S["Qt"]
    We, dt, betav = (Parameter(k, pv[k]) for k in ("We", "dt", "betav"))
    tau0_vec, tau1_vec = drv.tau0_vec, drv.tau1_vec
    (u1, D1_vec) = drv.w1.split()
    tau1 = symbolic.as_matrix([[tau1_vec[0], tau1_vec[1]], [tau1_vec[1], tau1_vec[2]]])
    # ---- incomp_LDC.py:298-308 -------------------------------------------------------------------------
    F1R = dot(u1,grad(tau1)) - dot(grad(u1),tau1) - dot(tau1,tgrad(u1)) + div(u1)*tau1
    F1R_vec = as_vector([F1R[0,0], F1R[1,0], F1R[1,1]])
    Dcomp1_vec = as_vector([Dcomp(u1)[0,0], Dcomp(u1)[1,0], Dcomp(u1)[1,1]])
    restau0 = We/dt*(tau1_vec-tau0_vec) + We*F1R_vec + tau1_vec - 2*(1-betav)*Dcomp1_vec
    res_test = project(restau0, Zd)
    res_orth = project(restau0-res_test, Zc)
    res_orth_norm_sq = project(inner(res_orth,res_orth), Qt)
    res_orth_norm = np.power(res_orth_norm_sq, 0.5)
    kapp = project(res_orth_norm, Qt)
    # ---- the driver's own indicator (fx_expr ids), bound back to its own Functions ------------------------
    for slot, fn in ((abi.FX_F_RES_TEST, drv.res_test), (abi.FX_F_RES_ORTH, drv.res_orth),
                     (abi.FX_F_QT_A, drv.res_orth_norm_sq), (abi.FX_F_TAU0, drv.tau0_vec), (abi.FX_F_TAU1, drv.tau1_vec),
                     (abi.FX_F_W1, drv.w1)):
        drv.pr.bind(slot, fn)
    drv.lps_indicator()
    assert lc.rel_err(res_test.vector().get_local(), drv.res_test.vector().get_local()) < 1e-13
    assert lc.rel_err(res_orth.vector().get_local(), drv.res_orth.vector().get_local()) < 1e-9
    assert lc.rel_err(kapp.vector().get_local(), drv.kapp.vector().get_local()) < 1e-9
    assert float(np.abs(kapp.vector().get_local()).max()) > 0.0
    assert ff.project is dapi.project
