"""Structural signatures of the reference scripts' forms -> fx_form ids (see fenics_b200/symbolic.py).

Each block below RESTATES the form-building statements of a script (file:line cited) on template objects -- template
Functions that only know their space and their fx_field slot, `Parameter`s named like the script's scalars -- and
registers the resulting signature.  A script that builds the same forms (same statements) with its own Functions and
Parameters is then served by the hand-written functors: `assemble(a1)` is literal.

Covered: lid-driven-cavity/incomp_LDC.py (config 1), all ten forms and both functionals of the time loop.  The other
loops are driven through `fenics_b200.drivers.*` (fx_form ids directly); their forms can be registered the same way.
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


def _reg(form, fid, what):
    _, ctx = form.signature()
    S.register(form, fid, [c.slot for c in ctx.coefs], what)


def load():
    global _loaded
    if not _loaded:
        _loaded = True
        _register_cfg1()
