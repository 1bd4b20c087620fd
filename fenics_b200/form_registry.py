"""Structural signatures of the reference scripts' forms -> fx_form ids (see fenics_b200/symbolic.py).

Each block below RESTATES the form-building statements of a script (file:line cited) on template objects -- template
Functions that only know their space and their fx_field slot, `Parameter`s named like the script's scalars -- and
registers the resulting signature.  A script that builds the same forms (same statements) with its own Functions and
Parameters is then served by the hand-written functors: `assemble(a1)` is literal.

Covered: lid-driven-cavity/incomp_LDC.py (config 1: ten forms, both functionals) and
lid-driven-cavity/comp_LDC_mesh_convergence.py (config 2, the bench configuration: twelve forms through lhs() / rhs(), both
functionals), its sweep sibling comp_LDC.py, and buoyancy-driven-flow/compressible_dgp.py (config 5: sixteen forms, four
projections, three functionals) and "flow between eccentrically rotating cylinders"/fenepmp_jbp.py (config 4: fourteen
forms, five functionals, FENE-P and FENE-P-MP variants) and incompressible_benchmarkEWM.py (config 3: sixteen forms, five
functionals, folded and material-function variants).
"""
from . import _abi as abi
from . import symbolic as S
from .spaces import ELEMENTS

_loaded = False


class _Space:
    def __init__(self, name):
        self.name, self.comps = name, ELEMENTS[name]


class TemplateFunction(S.Operand):
    """stand-in for a Function inside the registry: space + fx_field slot"""

    def __init__(self, space, slot):
        self._space, self.slot = _Space(space), slot

    def function_space(self):
        return self._space

    def split(self):
        return split(self)


def split(fn):
    """(u, D_vec) = w.split() for W Functions (velocity components 0,1 | DEVSS components 2,3,4); scalar components
    otherwise"""
    sp = fn.function_space()
    if sp.name == "W":
        return S.coefficient(fn, (0, 1)), S.coefficient(fn, (2, 3, 4))
    return tuple(S.coefficient(fn, (i,)) for i in range(len(sp.comps)))


def arguments(space, number):
    """TestFunction(s) (number 0) / TrialFunction(s) (number 1) of a space, split like dolfin's TestFunctions(W)"""
    name, obj = (space, None) if isinstance(space, str) else (space.name, space)
    if name == "W":
        return S.argument(number, "W", (0, 1), obj), S.argument(number, "W", (2, 3, 4), obj)
    return S.argument(number, name, tuple(range(len(ELEMENTS[name]))), obj)


def _sym3(v):
    return S.as_matrix([[v[0], v[1]], [v[1], v[2]]])


def _register_cfg1():
    """lid-driven-cavity/incomp_LDC.py: trial/test functions :142-156, helpers fenics_fem.py:289-314, DEVSS :259-262,
    loop :296-421"""
    from .symbolic import Identity, Parameter, as_vector, div, dot, dx, grad, inner, transpose  # noqa: F401
    A = abi
    Re, We, dt, betav, th, conv, c1, c2 = (Parameter(n, 1.0) for n in ("Re", "We", "dt", "betav", "th", "conv", "c1", "c2"))
    T = TemplateFunction
    w0, w12, ws, w1 = T("W", A.FX_F_W0), T("W", A.FX_F_W12), T("W", A.FX_F_WS), T("W", A.FX_F_W1)
    p0, p1 = T("Q", A.FX_F_P0), T("Q", A.FX_F_P1)
    tau0_vec, tau1_vec, kapp = T("Zc", A.FX_F_TAU0), T("Zc", A.FX_F_TAU1), T("Qt", A.FX_F_KAPP)
    h = S.CellDiameter()

    # the tensor helpers are the shim's own (fenics_b200/shims/_common.py = fenics_fem.py:289-314): a script that
    # imports them from the drop-in module builds exactly these trees
    from .shims._common import Dcomp, Dincomp, tgrad

    p, q = arguments("Q", 1), arguments("Q", 0)                     # incomp_LDC.py:142-156
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, D12_vec), (us, Ds_vec), (u1, D1_vec) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D12, tau0, tau1 = _sym3(D0_vec), _sym3(D12_vec), _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))

    DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx            # :259
    DEVSSl_u1 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx             # :261
    LPSl_stress = inner(kapp * h * c1 * grad(tau), grad(Rt)) * dx + inner(kapp * h * c2 * div(tau), div(Rt)) * dx  # :309
    DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx                  # :313
    # velocity half step :317-325
    visc_u12 = betav * grad(u)
    lhs_u12 = (Re / (dt / 2.0)) * u
    rhs_u12 = (Re / (dt / 2.0)) * u0 - Re * conv * grad(u0) * u0
    a1 = inner(lhs_u12, v) * dx + inner(visc_u12, grad(v)) * dx + (inner(D - Dincomp(u), R) * dx)
    L1 = inner(rhs_u12, v) * dx + inner(p0, div(v)) * dx - inner(tau0, grad(v)) * dx
    a1 += th * DEVSSl_u12
    L1 += th * DEVSSr_u12
    _reg(a1, A.FX_FORM_C1_A1, "a1 (incomp_LDC.py:321,325)")
    _reg(L1, A.FX_FORM_C1_L1, "L1 (incomp_LDC.py:322,326)")
    DEVSSr_u1 = 2 * (1 - betav) * inner(D12, Dincomp(v)) * dx                  # :335
    # predictor :355-359
    lhs_us = (Re / dt) * u
    rhs_us = (Re / dt) * u0 - Re * conv * grad(u12) * u12
    a2 = inner(lhs_us, v) * dx + inner(D - Dincomp(u), R) * dx
    L2 = inner(rhs_us, v) * dx - 0.5 * betav * (inner(grad(u0), grad(v)) * dx) + inner(p0, div(v)) * dx \
        - inner(tau0, grad(v)) * dx
    _reg(a2, A.FX_FORM_C1_A2, "a2 (incomp_LDC.py:358)")
    _reg(L2, A.FX_FORM_C1_L2, "L2 (incomp_LDC.py:359)")
    # pressure correction :370-371
    a5 = inner(grad(p), grad(q)) * dx
    L5 = inner(grad(p0), grad(q)) * dx + (Re / dt) * inner(us, grad(q)) * dx
    _reg(a5, A.FX_FORM_C1_A5, "a5 (incomp_LDC.py:370)")
    _reg(L5, A.FX_FORM_C1_L5, "L5 (incomp_LDC.py:371)")
    # velocity update :379-385 (the script adds the DEVSS terms of this step to a1 / L1, :388-389: a7 / L7 go without)
    visc_u1 = 0.5 * betav * grad(u)
    lhs_u1 = (Re / dt) * u
    rhs_u1 = (Re / dt) * us
    a7 = inner(lhs_u1, v) * dx + inner(visc_u1, grad(v)) * dx + inner(D - Dincomp(u), R) * dx
    L7 = inner(rhs_u1, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx
    _reg(a7, A.FX_FORM_C1_A7, "a7 (incomp_LDC.py:384)")
    _reg(L7, A.FX_FORM_C1_L7, "L7 (incomp_LDC.py:385)")
    # stress full step :400-409
    F1 = dot(u1, grad(tau)) - dot(grad(u1), tau) - dot(tau, tgrad(u1))
    lhs_tau1 = (We / dt + 1.0) * tau + We * F1
    rhs_tau1 = (We / dt) * tau0 + 2.0 * (1.0 - betav) * Dincomp(u12)
    a4 = inner(lhs_tau1, Rt) * dx
    L4 = inner(rhs_tau1, Rt) * dx
    a4 += LPSl_stress
    L4 += 0
    _reg(a4, A.FX_FORM_C1_A4, "a4 (incomp_LDC.py:404,409)")
    _reg(L4, A.FX_FORM_C1_L4, "L4 (incomp_LDC.py:405,410)")
    # energies :420-421
    _reg(0.5 * dot(u1, u1) * dx, A.FX_FORM_C1_EK, "E_k (incomp_LDC.py:420)")
    _reg((tau1[0, 0] + tau1[1, 1]) * dx, A.FX_FORM_C1_EE, "E_e (incomp_LDC.py:421)")
    del DEVSSl_u1, DEVSSr_u1


def _register_cfg2():
    """lid-driven-cavity/comp_LDC_mesh_convergence.py (config 2): trial/test functions :203-217, DEVSS :343-349, loop
    :392-549.  The script writes the two momentum steps and the stress step as ONE residual each and splits it with
    lhs() / rhs(); symbolic.lhs / rhs do the same split, so the registered trees are the script's."""
    from .shims._common import Dcomp, Dincomp, Fdef_incompressible as Fdef, ldc_sigma as sigma
    from .symbolic import CellDiameter, FacetNormal, Parameter, div, dot, ds, dx, grad, inner, lhs, nabla_grad, rhs
    A = abi
    Re, We, Ma, dt, betav, th, conv, c1, c2 = (Parameter(n_, 1.0) for n_ in
                                               ("Re", "We", "Ma", "dt", "betav", "th", "conv", "c1", "c2"))
    T = TemplateFunction
    w0, w12, ws, w1 = T("W", A.FX_F_W0), T("W", A.FX_F_W12), T("W", A.FX_F_WS), T("W", A.FX_F_W1)
    p0, p1, rho0, rho1 = T("Q", A.FX_F_P0), T("Q", A.FX_F_P1), T("Q", A.FX_F_RHO0), T("Q", A.FX_F_RHO1)
    tau0_vec, tau1_vec, kapp = T("Zc", A.FX_F_TAU0), T("Zc", A.FX_F_TAU1), T("Qt", A.FX_F_KAPP)
    h, n = CellDiameter(), FacetNormal()
    rho = p = arguments("Q", 1)
    q = arguments("Q", 0)
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, D12_vec), (us, _), (u1, _) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D12, tau0 = _sym3(D0_vec), _sym3(D12_vec), _sym3(S.as_expr(tau0_vec))

    DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx            # :346
    DEVSSl_u1 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx             # :348
    LPSl_stress = inner(kapp * h * c1 * grad(tau), grad(Rt)) * dx + inner(kapp * h * c2 * div(tau), div(Rt)) * dx  # :402
    DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx                  # :412
    U = 0.5 * (u + u0)                                                         # :414
    # density half step :417-421
    rho_eq = 2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0)
    rho_weak = inner(rho_eq, q) * dx
    _reg(lhs(rho_weak), A.FX_FORM_C2_A0, "a0 = lhs(rho_weak) (comp_LDC_mesh_convergence.py:420)")
    _reg(rhs(rho_weak), A.FX_FORM_C2_L0, "L0 = rhs(rho_weak) (comp_LDC_mesh_convergence.py:421)")
    # velocity half step :430-443
    lhsFu12 = Re * rho0 * (2.0 * (u - u0) / dt + conv * dot(u0, nabla_grad(u0)))
    Fu12 = dot(lhsFu12, v) * dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
        + dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
        - dot(tau0 * n, v) * ds \
        + inner(D - Dcomp(u), R) * dx
    a1, L1 = lhs(Fu12), rhs(Fu12)
    a1 += th * DEVSSl_u12
    L1 += th * DEVSSr_u12
    _reg(a1, A.FX_FORM_C2_A1, "a1 = lhs(Fu12) + th*DEVSSl_u12 (comp_LDC_mesh_convergence.py:439,442)")
    _reg(L1, A.FX_FORM_C2_L1, "L1 = rhs(Fu12) + th*DEVSSr_u12 (comp_LDC_mesh_convergence.py:440,443)")
    # the sweep script comp_LDC.py writes the same step with `Re*conv*dot(u0, nabla_grad(u0))` (:418): same functor, its
    # `conv` slot holding Re*conv
    re_conv = dict(conv=lambda P: P["Re"] * P["conv"])
    Fu12s = dot(Re * rho0 * (2.0 * (u - u0) / dt + Re * conv * dot(u0, nabla_grad(u0))), v) * dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
        + dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
        - dot(tau0 * n, v) * ds \
        + inner(D - Dcomp(u), R) * dx
    _reg(rhs(Fu12s) + th * DEVSSr_u12, A.FX_FORM_C2_L1, "L1 = rhs(Fu12) + th*DEVSSr_u12 (comp_LDC.py:418-431)", re_conv)
    DEVSSr_u1 = 2 * (1 - betav) * inner(D12, Dincomp(v)) * dx                  # :454
    # predictor :458-473
    lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u12, nabla_grad(u12)))
    Fus = dot(lhsFus, v) * dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
        + 0.5 * dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
        - dot(tau0 * n, v) * ds \
        + inner(D - Dcomp(u), R) * dx
    a2, L2 = lhs(Fus), rhs(Fus)
    a2 += th * DEVSSl_u1
    L2 += th * DEVSSr_u1
    _reg(a2, A.FX_FORM_C2_A2, "a2 = lhs(Fus) + th*DEVSSl_u1 (comp_LDC_mesh_convergence.py:469,472)")
    _reg(L2, A.FX_FORM_C2_L2, "L2 = rhs(Fus) + th*DEVSSr_u1 (comp_LDC_mesh_convergence.py:470,473)")
    Fuss = dot(Re * rho0 * ((u - u0) / dt + Re * conv * dot(u12, nabla_grad(u12))), v) * dx + \
        + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
        + 0.5 * dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
        - dot(tau0 * n, v) * ds \
        + inner(D - Dcomp(u), R) * dx
    _reg(rhs(Fuss) + th * DEVSSr_u1, A.FX_FORM_C2_L2, "L2 = rhs(Fus) + th*DEVSSr_u1 (comp_LDC.py:468-481)", re_conv)
    # pressure :483-490
    lhs_p_1 = (Ma * Ma / (dt * Re)) * p
    rhs_p_1 = (Ma * Ma / (dt * Re)) * p0 - rho0 * div(us) + dot(grad(rho0), us)
    lhs_p_2 = 0.5 * dt * grad(p)
    rhs_p_2 = 0.5 * dt * grad(p0)
    _reg(inner(lhs_p_1, q) * dx + inner(lhs_p_2, grad(q)) * dx, A.FX_FORM_C2_A5, "a5 (comp_LDC_mesh_convergence.py:489)")
    _reg(inner(rhs_p_1, q) * dx + inner(rhs_p_2, grad(q)) * dx, A.FX_FORM_C2_L5, "L5 (comp_LDC_mesh_convergence.py:490)")
    # equation of state :501-502: rho1 = project(rho0 + (Ma*Ma/Re)*(p1-p0), Q) -- dolfin.project's mass matrix and
    # right-hand side (dolfin_api.project builds inner(w, Pv)*dx and inner(w, v)*dx)
    _reg(inner(q, p) * dx, A.FX_FORM_MASS_Q, "project(., Q): mass matrix")
    _reg(inner(q, rho0 + (Ma * Ma / Re) * (p1 - p0)) * dx, A.FX_FORM_C2_L_RHO1,
         "project(rho0 + (Ma*Ma/Re)*(p1-p0), Q) (comp_LDC_mesh_convergence.py:501-502)")
    # velocity update :506-512 (rho1 is the PROJECTED density, a Function of Q, :502)
    lhs_u1 = (Re / dt) * rho1 * u
    rhs_u1 = (Re / dt) * rho1 * us
    a7 = inner(lhs_u1, v) * dx + inner(D - Dcomp(u), R) * dx
    L7 = inner(rhs_u1, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx - 0.5 * dot(p1 * n, v) * ds
    a7 += 0
    L7 += 0
    _reg(a7, A.FX_FORM_C2_A7, "a7 (comp_LDC_mesh_convergence.py:509)")
    _reg(L7, A.FX_FORM_C2_L7, "L7 (comp_LDC_mesh_convergence.py:510)")
    # stress :528-537
    lhs_tau1 = (We / dt + 1.0) * tau + We * Fdef(u1, tau)
    rhs_tau1 = (We / dt) * tau0 + 2.0 * (1.0 - betav) * Dcomp(u1)
    Ares = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
    a4, L4 = lhs(Ares), rhs(Ares)
    a4 += LPSl_stress
    L4 += 0
    _reg(a4, A.FX_FORM_C2_A4, "a4 = lhs(A) + LPSl_stress (comp_LDC_mesh_convergence.py:533,536)")
    _reg(L4, A.FX_FORM_C2_L4, "L4 = rhs(A) (comp_LDC_mesh_convergence.py:534,537)")
    # energies :548-549
    _reg(0.5 * rho1 * dot(u1, u1) * dx, A.FX_FORM_C2_EK, "E_k (comp_LDC_mesh_convergence.py:548)")
    _reg((tau1_vec[0] + tau1_vec[2]) * dx, A.FX_FORM_C2_EE, "E_e (comp_LDC_mesh_convergence.py:549)")


def _register_cfg5():
    """buoyancy-driven-flow/compressible_dgp.py (config 5): trial/test functions :100-118, theta :294-296, DEVSS :336-338,
    LPS :391, loop :404-583.  Parameters named like the parameter slots (T_0 -> "T0", T_h -> "th_hot", al_1 -> "al1",
    al_2 -> "al2"); DOLFIN_EPS of the SU weight (:510) is a number in the script: the functor reads it from FX_P_EPS."""
    from .shims._common import (DOLFIN_EPS, Dcomp, Dincomp, Fdef_compressible as Fdefcom, dgp_sigmacom as sigmacom,
                                magnitude)
    from .symbolic import (CellDiameter, FacetNormal, Identity, Parameter, as_vector, div, dot, ds, dx, grad, inner, lhs,
                           nabla_grad, rhs)
    A = abi
    Ra, Pr, We, Ma, dt, betav, th, Vh, Bi, c1, c2, T_0, T_h, al_1, al_2 = (
        Parameter(n_, 1.0) for n_ in ("Ra", "Pr", "We", "Ma", "dt", "betav", "th", "Vh", "Bi", "c1", "c2", "T0", "th_hot",
                                      "al1", "al2"))
    eps = dict(eps=lambda P: DOLFIN_EPS)
    T_ = TemplateFunction
    w0, w12, ws, w1 = T_("W", A.FX_F_W0), T_("W", A.FX_F_W12), T_("W", A.FX_F_WS), T_("W", A.FX_F_W1)
    p0, p1 = T_("Q", A.FX_F_P0), T_("Q", A.FX_F_P1)
    rho0, rho12, rho1 = T_("Q", A.FX_F_RHO0), T_("Q", A.FX_F_RHO12), T_("Q", A.FX_F_RHO1)
    T0, T1 = T_("Q", A.FX_F_T0), T_("Q", A.FX_F_T1)
    theta0, therm_inv, theta1 = T_("Q", A.FX_F_THETA0), T_("Q", A.FX_F_THERM_INV), T_("Q", A.FX_F_Q_A)
    tau0_vec, tau1_vec, kapp = T_("Zc", A.FX_F_TAU0), T_("Zc", A.FX_F_TAU1), T_("Qt", A.FX_F_KAPP)
    h, n = CellDiameter(), FacetNormal()
    k = as_vector([0.0, 1.0])                                                   # Expression(('0','1'), degree=2) :177
    rho = p = T = arguments("Q", 1)
    q = r = arguments("Q", 0)
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, D12_vec), (us, _), (u1, _) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D12 = _sym3(D0_vec), _sym3(D12_vec)
    tau0, tau1 = _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))
    thetal = (T) / (T_h - T_0)                                                  # :294
    thetar = T_("Q", -1)   # :295-296 project(T_0/(T_h-T_0), Q): a Function in the script, a constant inside the functors
    #                        (slot -1: recognised as a coefficient of the form, not handed to the kernel)
    DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx             # :336
    DEVSSl_u1 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx              # :338
    LPSl_stress = inner(kapp * h * c1 * grad(tau), grad(Rt)) * dx + inner(kapp * h * c2 * div(tau), div(Rt)) * dx  # :391
    # per-step projections :405-408
    _reg(inner(q, (T0 - T_0) / (T_h - T_0)) * dx, A.FX_FORM_C5_L_THETA0, "project((T0-T_0)/(T_h-T_0), Q) (compressible_dgp.py:405)")
    therm = (al_1 + al_2 * theta0)
    _reg(inner(q, 1.0 / therm) * dx, A.FX_FORM_C5_L_THERM_INV, "project(1.0/therm, Q) (compressible_dgp.py:408)")
    DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx                   # :413
    U = 0.5 * (u + u0)                                                          # :416
    # density half step :418-421 (same statements as config 2: the first registration of that signature stands)
    rho_eq = 2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0)
    rho_weak = inner(rho_eq, q) * dx
    _reg(lhs(rho_weak), A.FX_FORM_C5_A0, "a0 (compressible_dgp.py:420)")
    _reg(rhs(rho_weak), A.FX_FORM_C5_L0, "L0 (compressible_dgp.py:421)")
    # velocity half step :432-442
    Du12Dt = rho0 * (2.0 * (u - u0) / dt + dot(u0, nabla_grad(u0)))
    Fu12 = dot(Du12Dt, v) * dx + \
        + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(rho0 * therm_inv * k, v) * dx \
        + dot(p0 * n, v) * ds - betav * Pr * (dot(nabla_grad(U) * n, v) * ds + (1.0 / 3) * dot(div(U) * n, v) * ds) \
        - (Pr * (1 - betav) / We) * dot(tau0 * n, v) * ds \
        + inner(D - Dincomp(u), R) * dx
    a1, L1 = lhs(Fu12), rhs(Fu12)
    a1 += th * DEVSSl_u12
    L1 += th * DEVSSr_u12
    _reg(a1, A.FX_FORM_C5_A1, "a1 (compressible_dgp.py:439,441)")
    _reg(L1, A.FX_FORM_C5_L1, "L1 (compressible_dgp.py:440,442)")
    DEVSSr_u1 = 2 * (1 - betav) * inner(D12, Dincomp(v)) * dx                   # :453
    # predictor :473-483
    lhsFus = rho0 * ((u - u0) / dt + dot(u12, nabla_grad(U)))
    Fus = dot(lhsFus, v) * dx + \
        + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(rho0 * therm_inv * k, v) * dx \
        + dot(p0 * n, v) * ds - betav * Pr * (dot(nabla_grad(U) * n, v) * ds + (1.0 / 3) * dot(div(U) * n, v) * ds) \
        - (Pr * (1.0 - betav) / We) * dot(tau0 * n, v) * ds \
        + inner(D - Dincomp(u), R) * dx
    a2, L2 = lhs(Fus), rhs(Fus)
    a2 += th * DEVSSl_u1
    L2 += th * DEVSSr_u1
    _reg(a2, A.FX_FORM_C5_A2, "a2 (compressible_dgp.py:480,482)")
    _reg(L2, A.FX_FORM_C5_L2, "L2 (compressible_dgp.py:481,483)")
    # pressure :493-500
    lhs_p_1 = (Ma * Ma / (dt)) * p
    rhs_p_1 = (Ma * Ma / (dt)) * p0 - therm * dot(grad(rho0), us) - therm * rho0 * div(us)
    lhs_p_2 = therm * 0.5 * dt * grad(p)
    rhs_p_2 = therm * 0.5 * dt * grad(p0)
    _reg(inner(lhs_p_1, q) * dx + inner(lhs_p_2, grad(q)) * dx, A.FX_FORM_C5_A5, "a5 (compressible_dgp.py:499)")
    _reg(inner(rhs_p_1, q) * dx + inner(rhs_p_2, grad(q)) * dx, A.FX_FORM_C5_L5, "L5 (compressible_dgp.py:500)")
    # SU density update :510-517
    alpha_supg = h / (magnitude(u12) + DOLFIN_EPS)
    SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
    rho_eq = (rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12)
    rho_weak = inner(rho_eq, q) * dx
    a6, L6 = lhs(rho_weak), rhs(rho_weak)
    a6 += SU_rho
    _reg(a6, A.FX_FORM_C5_A6, "a6 (compressible_dgp.py:515,517)", eps)
    _reg(L6, A.FX_FORM_C5_L6, "L6 (compressible_dgp.py:516)")
    # equation of state :529-530
    _reg(inner(q, rho0 + (Ma * Ma) * therm_inv * (p1 - p0)) * dx, A.FX_FORM_C5_L_RHO1,
         "project(rho0 + (Ma*Ma)*therm_inv*(p1-p0), Q) (compressible_dgp.py:529-530)")
    # velocity update :533-537
    lhs_u1 = (1. / dt) * rho1 * u
    rhs_u1 = (1. / dt) * rho0 * us
    a7 = inner(lhs_u1, v) * dx + inner(D - Dincomp(u), R) * dx
    L7 = inner(rhs_u1, v) * dx + 0.5 * (inner(p1 - p0, div(v)) * dx - dot((p1 - p0) * n, v) * ds)
    a7 += 0
    L7 += 0
    _reg(a7, A.FX_FORM_C5_A7, "a7 (compressible_dgp.py:535)")
    _reg(L7, A.FX_FORM_C5_L7, "L7 (compressible_dgp.py:536)")
    # stress :550-555
    stress_eq = ((We / dt) + 1.0) * tau + We * Fdefcom(u1, tau) - (We / dt) * tau0 - Identity(len(u))
    Ast = inner(stress_eq, Rt) * dx
    a4, L4 = lhs(Ast), rhs(Ast)
    a4 += LPSl_stress
    L4 += 0
    _reg(a4, A.FX_FORM_C5_A4, "a4 (compressible_dgp.py:552,554)")
    _reg(L4, A.FX_FORM_C5_L4, "L4 (compressible_dgp.py:553,555)")
    # temperature :563-567
    gamdot = inner(sigmacom(u1, p1, tau1, We, Pr, betav), grad(u1))
    lhs_theta1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
    rhs_theta1 = (1.0 / dt) * rho0 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho0 * theta0 + Vh * gamdot
    a8 = inner(lhs_theta1, r) * dx + inner(grad(thetal), grad(r)) * dx
    L8 = inner(rhs_theta1, r) * dx + inner(grad(thetar), grad(r)) * dx + Bi * inner(grad(theta0), n * r) * ds(3) \
        + Bi * inner(grad(theta0), n * r) * ds(4) + inner(We * tau1 * grad(thetar), grad(r)) * dx
    _reg(a8, A.FX_FORM_C5_A8, "a8 (compressible_dgp.py:566)")
    _reg(L8, A.FX_FORM_C5_L8, "L8 (compressible_dgp.py:567)")
    # energies, Nusselt number :577-583
    _reg(0.5 * rho1 * dot(u1, u1) * dx, A.FX_FORM_C5_EK, "E_k (compressible_dgp.py:577)")
    _reg((tau1[0, 0] + tau1[1, 1] - 2.0) * dx, A.FX_FORM_C5_EE, "E_e (compressible_dgp.py:578)")
    _reg(inner(q, (T1 - T_0) / (T_h - T_0)) * dx, A.FX_FORM_C5_L_THETA1, "project((T1-T_0)/(T_h-T_0), Q) (compressible_dgp.py:581)")
    Tdx = inner(grad(theta1), n)
    _reg(-Tdx * ds(2), A.FX_FORM_C5_NUS, "Nus (compressible_dgp.py:582-583)")


def _register_cfg4():
    """"flow between eccentrically rotating cylinders"/fenepmp_jbp.py (config 4): theta :391-394, DEVSS :406-416, LPS :473,
    loop :489-649.  The forms that carry phi_def(u, lambda_d) are registered twice: with lambda_d a `Parameter` (FENE-P-MP,
    FX_FORM_C4M_* functors, UFL degrees 13 / 13 / 14) and with the plain number 0.0 the CSV default gives (UFL folds
    phi_def to 1: the FX_FORM_C4_* functors)."""
    from .shims._common import (DOLFIN_EPS, Dincomp, FdefG, fene_func, fene_sigmacom, magnitude, phi_def, phi_s)
    from .symbolic import (CellDiameter, Constant, FacetNormal, Identity, Parameter, as_vector, div, dot, ds, dx, grad,
                           inner, lhs, nabla_grad, rhs)
    A = abi
    Re, We, Ma, dt, betav, th, conv, b, A_0, Di, Vh, Bi, T_0, T_h = (
        Parameter(n_, 1.0) for n_ in ("Re", "We", "Ma", "dt", "betav", "th", "conv", "b_fene", "a0", "Di", "Vh", "Bi", "T0",
                                      "th_hot"))
    su_eps = dict(eps=lambda P: 0.0000000001)
    T_ = TemplateFunction
    w0, w12, ws, w1 = T_("W", A.FX_F_W0), T_("W", A.FX_F_W12), T_("W", A.FX_F_WS), T_("W", A.FX_F_W1)
    p0, p1 = T_("Q", A.FX_F_P0), T_("Q", A.FX_F_P1)
    rho0, rho12, rho1 = T_("Q", A.FX_F_RHO0), T_("Q", A.FX_F_RHO12), T_("Q", A.FX_F_RHO1)
    T0, T1, thetar = T_("Q", A.FX_F_T0), T_("Q", A.FX_F_T1), T_("Q", -1)
    tau0_vec, tau1_vec, kapp = T_("Zc", A.FX_F_TAU0), T_("Zc", A.FX_F_TAU1), T_("Qt", A.FX_F_KAPP)
    h, n = CellDiameter(), FacetNormal()
    tang = as_vector([n[1], -n[0]])                                             # :182
    rho = p = T = arguments("Q", 1)
    q = r = arguments("Q", 0)
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, _), (us, _), (u1, D1_vec) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D1 = _sym3(D0_vec), _sym3(D1_vec)
    tau0, tau1 = _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))
    thetal = (T) / (T_h - T_0)                                                  # :391
    theta0 = (T0 - T_0) / (T_h - T_0)                                           # :394
    DEVSSl_T1 = (1. - Di) * inner(grad(thetal), grad(r)) * dx                   # :406
    DEVSSr_T1 = inner((1. - Di) * (grad(thetar) + grad(theta0)), grad(r)) * dx  # :407
    DEVSSGl_u1 = 2.0 * (1. - betav) * inner(Dincomp(u), Dincomp(v)) * dx        # :415
    LPSl_stress = inner(kapp * h * 0.05 * grad(tau), grad(Rt)) * dx + inner(kapp * h * 0.01 * div(tau), div(Rt)) * dx  # :473
    DEVSSGr_u1 = (1. - betav) * inner(D0 + D0.T, Dincomp(v)) * dx               # :495
    U = 0.5 * (u + u0)                                                          # :496
    # density half step :499-503 (config 2's statements)
    rho_weak = inner(2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0), q) * dx
    _reg(lhs(rho_weak), A.FX_FORM_C4_A0, "a0 (fenepmp_jbp.py:502)")
    _reg(rhs(rho_weak), A.FX_FORM_C4_L0, "L0 (fenepmp_jbp.py:503)")
    # pressure :536-543
    lhs_p_1 = (Ma * Ma / (dt)) * p
    rhs_p_1 = (Ma * Ma / (dt)) * p0 - Re * div(rho0 * us)
    lhs_p_2 = dt * grad(p)
    rhs_p_2 = dt * grad(p0)
    _reg(inner(lhs_p_1, q) * dx + inner(lhs_p_2, grad(q)) * dx, A.FX_FORM_C4_A5, "a5 (fenepmp_jbp.py:542)")
    _reg(inner(rhs_p_1, q) * dx + inner(rhs_p_2, grad(q)) * dx, A.FX_FORM_C4_L5, "L5 (fenepmp_jbp.py:543)")
    # density update :553-562
    alpha_supg = h / (magnitude(u12) + 0.0000000001)
    SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
    rho_weak = inner((rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12), q) * dx
    a6, L6 = lhs(rho_weak), rhs(rho_weak)
    a6 += SU_rho
    _reg(a6, A.FX_FORM_C4_A6, "a6 (fenepmp_jbp.py:559,561)", su_eps)
    _reg(L6, A.FX_FORM_C4_L6, "L6 (fenepmp_jbp.py:560)")
    # velocity update :573-580
    lhs_u1 = (Re / dt) * rho1 * u
    rhs_u1 = (Re / dt) * rho0 * us
    a7 = inner(lhs_u1, v) * dx + inner(D - grad(u), R) * dx
    L7 = inner(rhs_u1, v) * dx - 0.5 * inner(grad(p1 - p0), (v)) * dx
    a7 += th * DEVSSGl_u1
    L7 += th * DEVSSGr_u1
    _reg(a7, A.FX_FORM_C4_A7, "a7 (fenepmp_jbp.py:577,579)")
    _reg(L7, A.FX_FORM_C4_L7, "L7 (fenepmp_jbp.py:578,580)")
    _reg(0.5 * rho1 * dot(u1, u1) * dx, A.FX_FORM_C4_EK, "E_k (fenepmp_jbp.py:635)")
    _reg((fene_func(tau1, b) * (tau1[0, 0] + tau1[1, 1]) - 2.) * dx, A.FX_FORM_C4_EE, "E_e (fenepmp_jbp.py:636)")
    for lambda_d, mp in ((Parameter("lambda_d", 1.0), True), (0.0, False)):
        ids = (dict(L2=A.FX_FORM_C4M_L2, A4=A.FX_FORM_C4M_A4, L9=A.FX_FORM_C4M_L9, TQ=A.FX_FORM_C4M_TORQUE,
                    FX=A.FX_FORM_C4M_FORCE_X, FY=A.FX_FORM_C4M_FORCE_Y) if mp else
               dict(L2=A.FX_FORM_C4_L2, A4=A.FX_FORM_C4_A4, L9=A.FX_FORM_C4_L9, TQ=A.FX_FORM_C4_TORQUE,
                    FX=A.FX_FORM_C4_FORCE_X, FY=A.FX_FORM_C4_FORCE_Y))
        tag = " [lambda_D != 0]" if mp else " [lambda_D = 0.0]"
        # predictor :513-524
        lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u0, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx + \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(
                div(phi_def(u0, lambda_d) * (fene_func(tau0, b) * tau0 - Identity(len(u)))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSGl_u1
        L2 += th * DEVSSGr_u1
        _reg(a2, A.FX_FORM_C4_A2, "a2 (fenepmp_jbp.py:521,523)")
        _reg(L2, ids["L2"], "L2 (fenepmp_jbp.py:522,524)" + tag)
        # stress :595-606
        lhs_tau1 = (We / dt) * tau + fene_func(tau0, b) * tau + We * FdefG(u1, D1, tau) \
            - We * 0.5 * (phi_def(u1, lambda_d) - 1.) * (tau * Dincomp(u1) + Dincomp(u1) * tau)
        rhs_tau1 = (We / dt) * tau0 + Identity(len(u))
        Astress = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
        a4, L4 = lhs(Astress), rhs(Astress)
        a4 += LPSl_stress
        L4 += 0
        _reg(a4, ids["A4"], "a4 (fenepmp_jbp.py:601,604)" + tag)
        _reg(L4, A.FX_FORM_C4_L4, "L4 (fenepmp_jbp.py:602,605)")
        # temperature :615-624
        gamdot = inner(fene_sigmacom(u0, p0, tau0, b, lambda_d, betav, We), grad(u0))
        lhs_temp1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
        difflhs_temp1 = Di * grad(thetal) + (Di / We) * tau1 * grad(thetal)
        rhs_temp1 = (1.0 / dt) * rho1 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho1 * theta0 + Vh * gamdot
        diffrhs_temp1 = Di * grad(thetar) + (Di / We) * tau1 * grad(thetar)
        a9 = inner(lhs_temp1, r) * dx + inner(difflhs_temp1, grad(r)) * dx
        L9 = inner(rhs_temp1, r) * dx + inner(diffrhs_temp1, grad(r)) * dx - Di * Bi * inner(theta0, r) * ds(1)
        a9 += 0.0 * th * DEVSSl_T1
        L9 += 0.0 * th * DEVSSr_T1
        _reg(a9, A.FX_FORM_C4_A9, "a9 (fenepmp_jbp.py:620,623)")
        _reg(L9, ids["L9"], "L9 (fenepmp_jbp.py:621,624)" + tag)
        # journal torque / load :639-649
        sigma0 = dot(fene_sigmacom(u1, p1, tau1, b, lambda_d, betav, We), tang)
        omegaf0 = dot(fene_sigmacom(u1, p1, tau1, b, lambda_d, betav, We), n)
        _reg(inner(Constant((1.0, 0.0)), omegaf0) * ds(0), ids["FX"], "innerforcex (fenepmp_jbp.py:647)" + tag)
        _reg(inner(Constant((0.0, -1.0)), omegaf0) * ds(0), ids["FY"], "innerforcey (fenepmp_jbp.py:648)" + tag)
        _reg(-inner(n, sigma0) * ds(0), ids["TQ"], "innertorque (fenepmp_jbp.py:649)" + tag)


def _register_cfg3():
    """"flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py (config 3): theta :398-402, DEVSS
    :410-422, SU :474-475, loop :488-706.  The forms that carry phi_s(theta0, A_0) / phi_ewm(tau, theta1, k_ewm, B) are
    registered twice: with A_0, k_ewm, B as `Parameter`s (the FX_FORM_C3E_* functors) and with the plain numbers 0.0 the
    script sets (:81-84; UFL folds both functions to 1: the FX_FORM_C3_* functors).  The functors keep the temperature
    DEVSS weight in FX_P_TH_T and the SU epsilon in FX_P_EPS (derived parameters)."""
    from .shims._common import (DOLFIN_EPS, Dcomp, Dincomp, FdefG, jbp_sigmacon as sigmacon, magnitude, phi_ewm, phi_s)
    from .symbolic import (CellDiameter, Constant, FacetNormal, Identity, Parameter, as_vector, div, dot, ds, dx, grad,
                           inner, lhs, nabla_grad, rhs)
    A = abi
    Re, We, Ma, dt, betav, th, conv, al, Di, Vh, Bi, T_0, T_h = (
        Parameter(n_, 1.0) for n_ in ("Re", "We", "Ma", "dt", "betav", "th", "conv", "al1", "Di", "Vh", "Bi", "T0", "th_hot"))
    su_eps = dict(eps=lambda P: 0.0000000001)
    th_t = dict(th_t=lambda P: P["th"])
    T_ = TemplateFunction
    w0, w12, ws, w1 = T_("W", A.FX_F_W0), T_("W", A.FX_F_W12), T_("W", A.FX_F_WS), T_("W", A.FX_F_W1)
    p0, p1 = T_("Q", A.FX_F_P0), T_("Q", A.FX_F_P1)
    rho0, rho12, rho1 = T_("Q", A.FX_F_RHO0), T_("Q", A.FX_F_RHO12), T_("Q", A.FX_F_RHO1)
    T0, T1, thetar = T_("Q", A.FX_F_T0), T_("Q", A.FX_F_T1), T_("Q", -1)
    tau0_vec, tau1_vec = T_("Zc", A.FX_F_TAU0), T_("Zc", A.FX_F_TAU1)
    h, n = CellDiameter(), FacetNormal()
    tang = as_vector([n[1], -n[0]])                                             # :168
    rho = p = T = arguments("Q", 1)
    q = r = arguments("Q", 0)
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, _), (us, _), (u1, D1_vec) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D1 = _sym3(D0_vec), _sym3(D1_vec)
    tau0, tau1 = _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))
    thetal = (T) / (T_h - T_0)                                                  # :398
    theta0 = (T0 - T_0) / (T_h - T_0)                                           # :401
    theta1 = S.as_expr(T1) - T_0 / (T_h - T_0)                                  # :402, :658 (sic)
    DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx             # :410
    DEVSSr_u12 = 2 * inner(D0, Dincomp(v)) * dx                                 # :411
    DEVSSl_T1 = (1. - Di) * inner(grad(thetal), grad(r)) * dx                   # :415
    DEVSSr_T1 = inner((1. - Di) * (grad(thetar) + grad(theta0)), grad(r)) * dx  # :416
    DEVSSGl_u1 = 2.0 * (1. - betav) * inner(Dincomp(u), Dincomp(v)) * dx        # :422
    alpha_supg = h / (magnitude(u1) + 0.0000000001)                             # :474
    SU = inner(dot(u1, grad(tau)), alpha_supg * dot(u1, grad(Rt))) * dx         # :475
    DEVSSGr_u1 = (1. - betav) * inner(D0 + D0.T, Dincomp(v)) * dx               # :491
    U = 0.5 * (u + u0)                                                          # :492
    # density half step :497-503 (config 2's statements)
    rho_weak = inner(2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0), q) * dx
    _reg(lhs(rho_weak), A.FX_FORM_C3_A0, "a0 (incompressible_benchmarkEWM.py:502)")
    _reg(rhs(rho_weak), A.FX_FORM_C3_L0, "L0 (incompressible_benchmarkEWM.py:503)")
    # pressure :586-593
    lhs_p_1 = (Ma * Ma / (dt)) * p
    rhs_p_1 = (Ma * Ma / (dt)) * p0 - Re * (1 + al * theta0) * div(rho0 * us)
    lhs_p_2 = (1 + al * theta0) * dt * grad(p)
    rhs_p_2 = (1 + al * theta0) * dt * grad(p0)
    _reg(inner(lhs_p_1, q) * dx + inner(lhs_p_2, grad(q)) * dx, A.FX_FORM_C3_A5, "a5 (incompressible_benchmarkEWM.py:592)")
    _reg(inner(rhs_p_1, q) * dx + inner(rhs_p_2, grad(q)) * dx, A.FX_FORM_C3_L5, "L5 (incompressible_benchmarkEWM.py:593)")
    # density update :598-608
    alpha_supg = h / (magnitude(u12) + 0.0000000001)
    SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
    rho_weak = inner((rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12), q) * dx
    a6, L6 = lhs(rho_weak), rhs(rho_weak)
    a6 += SU_rho
    _reg(a6, A.FX_FORM_C3_A6, "a6 (incompressible_benchmarkEWM.py:605,607)", su_eps)
    _reg(L6, A.FX_FORM_C3_L6, "L6 (incompressible_benchmarkEWM.py:606)")
    # velocity update :617-624 (`a7 += 0`, `L7 += 0`)
    lhs_u1 = (Re / dt) * rho1 * u
    rhs_u1 = (Re / dt) * rho0 * us
    a7 = inner(lhs_u1, v) * dx + inner(D - grad(u), R) * dx
    L7 = inner(rhs_u1, v) * dx - 0.5 * inner(grad(p1 - p0), (v)) * dx
    a7 += 0
    L7 += 0
    _reg(a7, A.FX_FORM_C3_A7, "a7 (incompressible_benchmarkEWM.py:620)")
    _reg(L7, A.FX_FORM_C3_L7, "L7 (incompressible_benchmarkEWM.py:621)")
    # energies, journal torque / load :684-706
    _reg(0.5 * rho1 * dot(u1, u1) * dx, A.FX_FORM_C3_EK, "E_k (incompressible_benchmarkEWM.py:684)")
    _reg((tau1_vec[0] + tau1_vec[2] - 2.0) * dx, A.FX_FORM_C3_EE, "E_e (incompressible_benchmarkEWM.py:685)")
    sigma0 = dot(sigmacon(u1, p1, tau1, betav, We), tang)
    omegaf0 = dot(sigmacon(u1, p1, tau1, betav, We), n)
    _reg(inner(Constant((1.0, 0.0)), omegaf0) * ds(0), A.FX_FORM_C3_FORCE_X, "innerforcex (incompressible_benchmarkEWM.py:703)")
    _reg(-inner(Constant((0.0, 1.0)), omegaf0) * ds(0), A.FX_FORM_C3_FORCE_Y, "innerforcey (incompressible_benchmarkEWM.py:704)")
    _reg(-inner(n, sigma0) * ds(0), A.FX_FORM_C3_TORQUE, "innertorque (incompressible_benchmarkEWM.py:706)")
    _reg(inner(Constant((0.0, -1.0)), omegaf0) * ds(0), A.FX_FORM_C3_FORCE_Y, "innerforcey (comp_ewm_jpb.py:653)")
    for (A_0, k_ewm, B), gen in (((Parameter("a0", 1.0), Parameter("k_ewm", 1.0), Parameter("b_ewm", 1.0)), True),
                                  ((0.0, 0.0, 0.0), False)):
        ids = (dict(A1=A.FX_FORM_C3E_A1, L1=A.FX_FORM_C3E_L1, L2=A.FX_FORM_C3E_L2, A4=A.FX_FORM_C3E_A4, L4=A.FX_FORM_C3E_L4,
                    L9=A.FX_FORM_C3E_L9) if gen else
               dict(A1=A.FX_FORM_C3_A1, L1=A.FX_FORM_C3_L1, L2=A.FX_FORM_C3_L2, A4=A.FX_FORM_C3_A4, L4=A.FX_FORM_C3_L4,
                    L9=A.FX_FORM_C3_L9))
        tag = " [material functions]" if gen else " [A_0 = k_ewm = B = 0.0]"
        # velocity half step :517-531
        lhsv_eq = 2.0 * Re * ((rho12 * u - rho0 * u0) / dt + conv * dot(u0, nabla_grad(U)))
        v_weak12 = dot(lhsv_eq, v) * dx + \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(div(tau0 - rho0 * Identity(len(u))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a1, L1 = lhs(v_weak12), rhs(v_weak12)
        a1 += th * DEVSSl_u12
        L1 += th * DEVSSr_u12
        _reg(a1, ids["A1"], "a1 (incompressible_benchmarkEWM.py:527,530)" + tag)
        _reg(L1, ids["L1"], "L1 (incompressible_benchmarkEWM.py:528,531)" + tag)
        # predictor :563-575
        lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u0, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx + \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(div(tau0 - rho0 * Identity(len(u))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSGl_u1
        L2 += th * DEVSSGr_u1
        _reg(a2, A.FX_FORM_C3_A2, "a2 (incompressible_benchmarkEWM.py:571,574)")
        _reg(L2, ids["L2"], "L2 (incompressible_benchmarkEWM.py:572,575)" + tag)
        # temperature :641-651
        gamdot = inner(sigmacon(u0, p0, tau0, betav, We), grad(u0))
        lhs_temp1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
        difflhs_temp1 = Di * grad(thetal)
        rhs_temp1 = (1.0 / dt) * rho1 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho1 * theta0 \
            + Vh * phi_ewm(tau1, theta1, k_ewm, B) * gamdot
        diffrhs_temp1 = Di * grad(thetar)
        a9 = inner(lhs_temp1, r) * dx + inner(difflhs_temp1, grad(r)) * dx
        L9 = inner(rhs_temp1, r) * dx + inner(diffrhs_temp1, grad(r)) * dx - Di * Bi * inner(theta0, r) * ds(1)
        a9 += th * DEVSSl_T1
        L9 += th * DEVSSr_T1
        _reg(a9, A.FX_FORM_C3_A9, "a9 (incompressible_benchmarkEWM.py:646,650)", th_t)
        _reg(L9, ids["L9"], "L9 (incompressible_benchmarkEWM.py:647,651)" + tag, th_t)
        # comp_ewm_jpb.py:601-605 switches the temperature DEVSS terms off by writing 0.0*th*DEVSSl_T1:
        # the same functors with the weight FX_P_TH_T = 0.  Every other statement of its loop (:426-659) is written like
        # incompressible_benchmarkEWM.py's, fenepmp_jbp.py's or compressible_dgp.py's and is registered above.
        a9 = inner(lhs_temp1, r) * dx + inner(difflhs_temp1, grad(r)) * dx
        L9 = inner(rhs_temp1, r) * dx + inner(diffrhs_temp1, grad(r)) * dx - Di * Bi * inner(theta0, r) * ds(1)
        a9 += 0.0 * th * DEVSSl_T1
        L9 += 0.0 * th * DEVSSr_T1
        off = dict(th_t=lambda P: 0.0)
        _reg(a9, A.FX_FORM_C3_A9, "a9 (comp_ewm_jpb.py:601,604)", off)
        _reg(L9, ids["L9"], "L9 (comp_ewm_jpb.py:602,605)" + tag, off)
        # stress :663-673
        lhs_tau1 = (We * phi_ewm(tau0, theta1, k_ewm, B) / dt + 1.0) * tau + We * phi_ewm(tau0, theta1, k_ewm, B) * FdefG(u1, D1, tau)
        rhs_tau1 = (We * phi_ewm(tau0, theta1, k_ewm, B) / dt) * tau0 + rho0 * Identity(len(u))
        Ftau = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
        a4, L4 = lhs(Ftau), rhs(Ftau)
        _reg(a4 + SU, ids["A4"], "a4_stab = a4 + SU (incompressible_benchmarkEWM.py:668,672)" + tag, su_eps)
        _reg(L4, ids["L4"], "L4_stab = L4 (incompressible_benchmarkEWM.py:669,673)" + tag)


def _register_incomp_dgp():
    """buoyancy-driven-flow/incompressible_dgp.py: theta :285-288, DEVSS :326-329 (+ :421 inside the loop),
    loop :369-523.  The script's DOLFIN_EPS in the DEVSS weights is a number in the form and in the functors
    (fx_forms_dgp.h C6_A12 / C6_L12); its pressure right-hand side carries 1/dt where config 1's carries Re/dt: the shared
    functor reads FX_P_RE, handed over as the derived parameter re = 1."""
    from .shims._common import DOLFIN_EPS, Dincomp, Fdef_incompressible as Fdef, dgp_sigma as sigma, dgp_sigma_n as sigma_n
    from .symbolic import (CellDiameter, FacetNormal, Identity, Parameter, as_vector, div, dot, ds, dx, grad, inner, lhs,
                           nabla_grad, rhs)
    A = abi
    Ra, Pr, We, dt, th, Vh, Bi, c1, c2, T_0, T_h = (
        Parameter(n_, 1.0) for n_ in ("Ra", "Pr", "We", "dt", "th", "Vh", "Bi", "c1", "c2", "T0", "th_hot"))
    betav = Parameter("betav", 0.5)          # sigma / sigma_n branch on betav < 1 (fenics_fem.py:69, :75): the polymeric branch
    T_ = TemplateFunction
    w0, w12, ws, w1 = T_("W", A.FX_F_W0), T_("W", A.FX_F_W12), T_("W", A.FX_F_WS), T_("W", A.FX_F_W1)
    p0, p1, T0, T1 = T_("Q", A.FX_F_P0), T_("Q", A.FX_F_P1), T_("Q", A.FX_F_T0), T_("Q", A.FX_F_T1)
    theta1 = T_("Q", A.FX_F_Q_A)
    tau0_vec, tau1_vec, kapp = T_("Zc", A.FX_F_TAU0), T_("Zc", A.FX_F_TAU1), T_("Qt", A.FX_F_KAPP)
    h, n = CellDiameter(), FacetNormal()
    f = as_vector([0.0, -1.0])                                                  # Expression(('0','-1'), degree=2) :155
    p = T = arguments("Q", 1)
    q = r = arguments("Q", 0)
    tau_vec, Rt_vec = arguments("Zc", 1), arguments("Zc", 0)
    (u, D_vec), (v, R_vec) = arguments("W", 1), arguments("W", 0)
    D, tau, R, Rt = _sym3(D_vec), _sym3(tau_vec), _sym3(R_vec), _sym3(Rt_vec)
    (u0, D0_vec), (u12, D12_vec), (us, _), (u1, _) = w0.split(), w12.split(), ws.split(), w1.split()
    D0, D12 = _sym3(D0_vec), _sym3(D12_vec)
    tau0, tau1 = _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))
    thetal = (T) / (T_h - T_0)                                                  # :285
    thetar = T_("Q", -1)                                                        # :286-287 project(T_0/(T_h-T_0), Q)
    theta0 = (T0 - T_0) / (T_h - T_0)                                           # :288 (an expression of T0 here)
    DEVSSl_u12 = 2 * (1 - betav + DOLFIN_EPS) * inner(Dincomp(u), Dincomp(v)) * dx   # :326
    DEVSSr_u12 = 2 * inner(D0, Dincomp(v)) * dx                                 # :327
    DEVSSl_u1 = 2 * (1 - betav + DOLFIN_EPS) * inner(Dincomp(u), Dincomp(v)) * dx    # :328
    LPSl_stress = inner(kapp * h * c1 * grad(tau), grad(Rt)) * dx + inner(kapp * h * c2 * div(tau), div(Rt)) * dx  # :380
    U = 0.5 * (u + u0)                                                          # :395
    # velocity half step :397-408
    Du12Dt = (2.0 * (u - u0) / dt + dot(u0, nabla_grad(u0)))
    Fu12 = dot(Du12Dt, v) * dx + \
        + inner(sigma(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(theta0 * f, v) * dx \
        + inner(D - Dincomp(u), R) * dx \
        - dot(sigma_n(U, p0, tau0, n, We, Pr, betav), v) * ds
    a1, L1 = lhs(Fu12), rhs(Fu12)
    a1 += th * DEVSSl_u12
    L1 += th * DEVSSr_u12
    _reg(a1, A.FX_FORM_C6_A1, "a1 (incompressible_dgp.py:403,407)")
    _reg(L1, A.FX_FORM_C6_L1, "L1 (incompressible_dgp.py:404,408)")
    DEVSSr_u1 = 2 * Pr * (1. - betav + DOLFIN_EPS) * inner(D12, Dincomp(v)) * dx     # :421
    # predictor :439-450
    lhsFus = ((u - u0) / dt + dot(u12, nabla_grad(U)))
    Fus = dot(lhsFus, v) * dx + inner(D - Dincomp(u), R) * dx + \
        + inner(sigma(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(theta0 * f, v) * dx \
        - dot(sigma_n(U, p0, tau0, n, We, Pr, betav), v) * ds
    a2, L2 = lhs(Fus), rhs(Fus)
    _reg(a2 + th * DEVSSl_u1, A.FX_FORM_C6_A2, "a2_stab (incompressible_dgp.py:447,449)")
    _reg(L2 + th * DEVSSr_u1, A.FX_FORM_C6_L2, "L2_stab (incompressible_dgp.py:448,450)")
    # pressure correction :459-460 (a5 is config 1's Laplacian, word for word)
    _reg(inner(grad(p), grad(q)) * dx, A.FX_FORM_C6_A5, "a5 (incompressible_dgp.py:459)")
    _reg(inner(grad(p0), grad(q)) * dx + (1.0 / dt) * inner(us, grad(q)) * dx, A.FX_FORM_C6_L5,
         "L5 (incompressible_dgp.py:460)", dict(re=lambda P: 1.0))
    # velocity update :469-472
    lhs_u1 = (1. / dt) * u
    rhs_u1 = (1. / dt) * us
    _reg(inner(lhs_u1, v) * dx + inner(D - Dincomp(u), R) * dx, A.FX_FORM_C6_A7, "a7 (incompressible_dgp.py:471)")
    _reg(inner(rhs_u1, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx - 0.5 * dot((p1 - p0) * n, v) * ds, A.FX_FORM_C6_L7,
         "L7 (incompressible_dgp.py:472)")
    # stress :487-492
    stress_eq = (We / dt + 1.0) * tau + We * Fdef(u1, tau) - (We / dt) * tau0 - Identity(len(u))
    Ast = inner(stress_eq, Rt) * dx
    a4, L4 = lhs(Ast), rhs(Ast)
    L4 += 0
    _reg(a4 + LPSl_stress, A.FX_FORM_C6_A4, "a4_stab (incompressible_dgp.py:489,491)")
    _reg(L4, A.FX_FORM_C6_L4, "L4 (incompressible_dgp.py:490,492)")
    # temperature :503-507
    gamdot = inner(sigma(u1, p1, tau1, We, Pr, betav), grad(u1))
    lhs_theta1 = (1.0 / dt) * thetal + dot(u1, grad(thetal))
    rhs_theta1 = (1.0 / dt) * thetar + dot(u1, grad(thetar)) + (1.0 / dt) * theta0 + Vh * gamdot
    a8 = inner(lhs_theta1, r) * dx + inner(grad(thetal), grad(r)) * dx + inner(We * tau1 * grad(thetal), grad(r)) * dx
    L8 = inner(rhs_theta1, r) * dx + inner(grad(thetar), grad(r)) * dx + Bi * inner(grad(theta0), n * r) * ds(1) \
        + Bi * inner(grad(theta0), n * r) * ds(4) + inner(We * tau1 * grad(thetar), grad(r)) * dx
    _reg(a8, A.FX_FORM_C6_A8, "a8 (incompressible_dgp.py:506)")
    _reg(L8, A.FX_FORM_C6_L8, "L8 (incompressible_dgp.py:507)")
    # energies (config 1's E_k, config 5's E_e, word for word), Nusselt number :517-523
    _reg(0.5 * dot(u1, u1) * dx, A.FX_FORM_C6_EK, "E_k (incompressible_dgp.py:517)")
    _reg((tau1[0, 0] + tau1[1, 1] - 2.0) * dx, A.FX_FORM_C6_EE, "E_e (incompressible_dgp.py:518)")
    Tdx = inner(grad(theta1), n)
    _reg(Tdx * ds(2), A.FX_FORM_C6_NUS, "Nus (incompressible_dgp.py:522-523)")


def _register_lps():
    """the LPS indicator block at the top of the loops (incomp_LDC.py:298-308 == comp_LDC_mesh_convergence.py:392-401;
    compressible_dgp.py:382-390; fenepmp_jbp.py:462-472): four project() calls.  A projection is recognised through the
    right-hand side dolfin.project assembles, inner(w, expr)*dx -- onto Zd / Qt (DG0) that is one cell-average kernel
    (("expr", fx_expr id)), onto Zc the consistent mass matrix + the registered right-hand side."""
    from .shims._common import (Dcomp, Dincomp, Fdef_compressible as Fdefcon, fene_func, phi_def)
    from .symbolic import Parameter, as_vector, dx, inner
    A = abi
    We, dt, betav, b = (Parameter(n_, 1.0) for n_ in ("We", "dt", "betav", "b_fene"))
    T_ = TemplateFunction
    w1, tau0_vec, tau1_vec = T_("W", A.FX_F_W1), T_("Zc", A.FX_F_TAU0), T_("Zc", A.FX_F_TAU1)
    res_test, res_orth, qt_a = T_("Zd", A.FX_F_RES_TEST), T_("Zc", A.FX_F_RES_ORTH), T_("Qt", A.FX_F_QT_A)
    (u1, _) = w1.split()
    tau0, tau1 = _sym3(S.as_expr(tau0_vec)), _sym3(S.as_expr(tau1_vec))
    wd, wc, wq = arguments("Zd", 0), arguments("Zc", 0), arguments("Qt", 0)
    I_vec = as_vector([1.0, 0.0, 1.0])                    # Expression(('1.0','0.0','1.0'), degree=2)

    def block(restau0, expr_id, orth_id, what):
        _reg(inner(wd, restau0) * dx, ("expr", expr_id), "project(restau0, Zd) " + what)
        _reg(inner(wc, restau0 - res_test) * dx, orth_id, "project(restau0-res_test, Zc) " + what)
    _reg(inner(wc, arguments("Zc", 1)) * dx, A.FX_FORM_MASS_ZC, "project(., Zc): mass matrix")
    _reg(inner(wq, inner(res_orth, res_orth)) * dx, ("expr", A.FX_EXPR_RES_ORTH_SQ), "project(inner(res_orth,res_orth), Qt)")
    _reg(inner(wq, S.as_expr(qt_a) ** 0.5) * dx, ("expr", A.FX_EXPR_SQRT_QT_A), "project(np.power(res_orth_norm_sq, 0.5), Qt)")
    kapp = T_("Qt", A.FX_F_KAPP)
    _reg(inner(wq, S.CellDiameter() * kapp) * dx, ("expr", A.FX_EXPR_H_KAPP), "err = project(h*kapp, Qt) (incomp_LDC.py:434)")
    # lid-driven cavity (configs 1, 2): F1R written out with + div(u1)*tau1, deviatoric rate of strain
    from .symbolic import div, dot, grad
    from .shims._common import tgrad
    F1R = dot(u1, grad(tau1)) - dot(grad(u1), tau1) - dot(tau1, tgrad(u1)) + div(u1) * tau1
    F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
    Dcomp1_vec = as_vector([Dcomp(u1)[0, 0], Dcomp(u1)[1, 0], Dcomp(u1)[1, 1]])
    block(We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + tau1_vec - 2 * (1 - betav) * Dcomp1_vec,
          A.FX_EXPR_LDC_RESTAU, A.FX_FORM_LDC_L_RES_ORTH, "(incomp_LDC.py:300-305)")
    # buoyancy-driven flow (config 5): Fdefcom, - I_vec
    F1R = Fdefcon(u1, tau1)
    F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
    block(We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + tau1_vec - I_vec,
          A.FX_EXPR_C5_RESTAU, A.FX_FORM_C5_L_RES_ORTH, "(compressible_dgp.py:382-387)")
    # incompressible double glazing (incompressible_dgp.py:369-378): Fdef without the div(u) term
    from .shims._common import Fdef_incompressible as Fdef
    F1R = Fdef(u1, tau1)
    F1R_vec_i = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
    block(We / dt * (tau1_vec - tau0_vec) + We * F1R_vec_i + tau1_vec - I_vec,
          A.FX_EXPR_C6_RESTAU, A.FX_FORM_C6_L_RES_ORTH, "(incompressible_dgp.py:369-374)")
    # journal bearing (config 4): FENE function and the dissipation term; lambda_d Parameter / plain 0.0 as in _register_cfg4
    for lambda_d, eid, oid in ((Parameter("lambda_d", 1.0), A.FX_EXPR_C4M_RESTAU, A.FX_FORM_C4M_L_RES_ORTH),
                               (0.0, A.FX_EXPR_C4_RESTAU, A.FX_FORM_C4_L_RES_ORTH)):
        dissipation = We * 0.5 * (phi_def(u1, lambda_d) - 1.) * (tau1 * Dincomp(u1) + Dincomp(u1) * tau1)
        diss_vec = as_vector([dissipation[0, 0], dissipation[1, 0], dissipation[1, 1]])
        block(We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + fene_func(tau0, b) * tau1_vec - diss_vec - I_vec,
              eid, oid, "(fenepmp_jbp.py:462-468)")


def _reg(form, fid, what, derived=None):
    _, ctx = form.signature()
    S.register(form, fid, [c.slot for c in ctx.coefs], what, derived)


def load():
    global _loaded
    if not _loaded:
        _loaded = True
        _register_cfg1()
        _register_cfg2()
        _register_cfg5()
        _register_cfg4()
        _register_cfg3()
        _register_incomp_dgp()
        _register_lps()
