"""ORACLE (test infrastructure, NOT product code) -- the reference's time-step bodies restated with
oracle/ufl_lite + oracle/fem, statement order preserved (forms see coefficient values at
assemble time; project() is eager -- SURVEY.md Appendix A "Evaluation-order / aliasing rule").

parity unpinned (see oracle/ufl_lite.py).  Each class cites the reference loop it follows.
Linear solves default to exact LU so the oracle is the converged limit of the reference's
BiCGStab solves; the singular pure-Neumann pressure system of config 1 is
solved in the zero-sum gauge.
"""
import numpy as np
import scipy.sparse as sp

from . import fem
from .ufl_lite import (as_matrix, as_vector, dot, inner, grad, nabla_grad, div, dx, ds, lhs, rhs,
                       Identity, CellDiameter, FacetNormal, Const)


# ---- helper algebra (lid-driven-cavity/fenics_fem.py:289-314) -------------------------------------
def tgrad(w):
    return grad(w).T


def Dincomp(w):
    return (grad(w) + tgrad(w)) / 2


def Dcomp(w):
    return ((grad(w) + tgrad(w)) - (2.0 / 3) * div(w) * Identity(len(w))) / 2


def sigma(u, p, Tau, We, betav):
    return 2 * betav * Dcomp(u) - p * Identity(len(u)) + ((1 - betav) / We) * (Tau - Identity(len(u)))


def Fdef(u, Tau):
    return dot(u, grad(Tau)) - dot(grad(u), Tau) - dot(Tau, tgrad(u))


def Fdefcon(u, Tau):
    return dot(u, grad(Tau)) - dot(grad(u), Tau) - dot(Tau, tgrad(u)) + div(u) * Tau


def sym(v):
    """reshape_elements: 3-vector -> symmetric 2x2 (lid-driven-cavity/fenics_fem.py:210-225)."""
    return as_matrix([[v[0], v[1]], [v[1], v[2]]])


def solve_zero_sum(A, b):
    """exact solve of a singular (constant-nullspace) consistent system in the sum(x)=0 gauge."""
    n = A.shape[0]
    one = np.ones((n, 1))
    K = sp.bmat([[A, sp.csr_matrix(one)], [sp.csr_matrix(one.T), None]], format="csc")
    return fem.solve_lu(K, np.concatenate([b, [0.0]]))[:n]


def shear_rate(S, u):
    """shear_rate(u) (lid-driven-cavity/fenics_fem.py:380-398): the stream function's Poisson problem with the
    right-hand side inner(0.5*gamma, gamma), gamma = grad(u) + tgrad(u).  Returns (dof vector, a, L)."""
    Vs = S["Vs"]
    psi, phi = fem.TrialFunction(Vs), fem.TestFunction(Vs)
    a = inner(grad(psi), grad(phi)) * dx
    gamma = grad(u) + tgrad(u)
    g = inner(0.5 * gamma, gamma)
    L = inner(g, phi) * dx
    A, b = fem.assemble(a), fem.assemble(L)
    bc = fem.DirichletBC(Vs, (0,), 0.0, lambda x: np.ones(len(x), dtype=bool))
    bc.apply(A, b)
    return fem.solve_lu(A, b), a, L


def make_spaces(mesh, dofmaps=None):
    """function_spaces(mesh, 2) (lid-driven-cavity/fenics_fem.py:262-287): the spaces the loops use."""
    dm = dofmaps or {}
    comps = {"W": ["P2", "P2", "DG0", "DG0", "DG0"], "V": ["P2", "P2"], "Zc": ["P2"] * 3,
             "Zd": ["DG0"] * 3, "Q": ["P1"], "Qt": ["DG0"], "Vs": ["P2"]}
    return {k: fem.FunctionSpace(mesh, c, dm.get(k), name=k) for k, c in comps.items()}


def stream_function(S, u, rho=None):
    """stream_function(u) / comp_stream_function(rho, u) (lid-driven-cavity/fenics_fem.py:340-378) on
    V.sub(0).collapse() = scalar P2: -lap(psi) = curl(u) [rho-weighted], psi = 0 on DomainBoundary(),
    direct solve.  `u[1].dx(0) - u[0].dx(1)` is written grad(u)[1,0] - grad(u)[0,1] (same UFL degree).
    Returns (psi dof vector, A, b)."""
    Vs = S["Vs"]
    psi, phi = fem.TrialFunction(Vs), fem.TestFunction(Vs)
    gu = grad(u)
    a = inner(grad(psi), grad(phi)) * dx
    if rho is None:
        L = inner(gu[1, 0] - gu[0, 1], phi) * dx
    else:
        L = inner(rho * gu[1, 0] - rho * gu[0, 1], phi) * dx
    A, b = fem.assemble(a), fem.assemble(L)
    bc = fem.DirichletBC(Vs, (0,), 0.0, lambda x: np.ones(len(x), dtype=bool))
    bc.apply(A, b)
    return fem.solve_lu(A, b), a, L


def min_location(vec, space):
    """min_location(u, mesh) (fenics_fem.py:402-417): coordinates of the dofs holding the minimum."""
    xy = space.tabulate_dof_coordinates()
    return xy[np.where(vec == vec.min())]


def lid_boundary_conditions(W, L=1.0):
    """bcu = [noslip, drive1] (lid-driven-cavity/incomp_LDC.py:171-226)."""
    top_bound = 1.0
    noslip = fem.DirichletBC(W, (0, 1), (0.0, 0.0), lambda x: np.ones(len(x), dtype=bool))
    holder = {"t": 0.0}

    def ulidreg(x):
        f = 8 * (1.0 + np.tanh(8 * holder["t"] - 4.0)) * (x[:, 0] * (L - x[:, 0])) ** 2
        return np.stack([f, np.zeros_like(f)], 1)

    drive1 = fem.DirichletBC(W, (0, 1), ulidreg, lambda x: x[:, 1] > top_bound - fem.DOLFIN_EPS)
    return [noslip, drive1], holder


class _LDCBase:
    def _lps_indicator(self):
        """LPS residual indicator kappa (incomp_LDC.py:300-308 == comp_LDC_mesh_convergence.py:393-401)."""
        S, p = self.S, self.p
        u1, tau1, tau1_vec, tau0_vec = self.u1, self.tau1, self.tau1_vec.expr(), self.tau0_vec.expr()
        F1R = Fdefcon(u1, tau1)
        F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
        Dc = Dcomp(u1)
        Dcomp1_vec = as_vector([Dc[0, 0], Dc[1, 0], Dc[1, 1]])
        restau0 = p["We"] / p["dt"] * (tau1_vec - tau0_vec) + p["We"] * F1R_vec + tau1_vec \
            - 2 * (1 - p["betav"]) * Dcomp1_vec
        self.restau0 = restau0
        res_test = fem.project(restau0, S["Zd"])
        res_orth = fem.project(restau0 - res_test.expr(), S["Zc"])
        res_orth_norm_sq = fem.project(inner(res_orth.expr(), res_orth.expr()), S["Qt"])
        kapp = fem.project(res_orth_norm_sq.expr() ** 0.5, S["Qt"])
        self.kapp.assign(kapp)
        self.last.update(res_test=res_test.vec, res_orth=res_orth.vec,
                         res_orth_norm_sq=res_orth_norm_sq.vec, kapp=kapp.vec)

    #: "lu" (parity runs: exact solves) or (method, pc): the reference's own Krylov path, e.g. ("bicgstab", "ilu") =
    #: solve(A, x, b, "bicgstab", "default") at PETSc's rtol 1e-5 from the previous solution (bench.py's CPU arm)
    solver = "lu"
    solver_rtol = 1e-5

    def _solve(self, A, x, b):
        if self.solver == "lu":
            x[:] = fem.solve_lu(A, b)
        else:
            it = fem.solve_krylov_c(A, x, b, self.solver[0], self.solver[1], rtol=self.solver_rtol)[0]
            self.last.setdefault("iterations", []).append(it)


class IncompLDC(_LDCBase):
    """Config 1: lid-driven-cavity/incomp_LDC.py time loop :280-447 (Oldroyd-B, incompressible)."""

    def __init__(self, mesh, dofmaps=None, dt=0.005, Re=0.5, We=0.25, betav=0.5, c1=0.05, c2=0.01,
                 th=1.0, conv=None):
        if conv is None:
            conv = 0.0 if Re < 1 else 1  # incomp_LDC.py:49,93-94
        self.p = dict(dt=dt, Re=Re, We=We, betav=betav, c1=c1, c2=c2, th=th, conv=conv)
        self.mesh, self.S = mesh, make_spaces(mesh, dofmaps)
        S = self.S
        F = fem.Function
        self.p0, self.p1 = F(S["Q"]), F(S["Q"])
        self.w0, self.w12, self.ws, self.w1 = F(S["W"]), F(S["W"]), F(S["W"]), F(S["W"])
        self.tau0_vec, self.tau1_vec = F(S["Zc"]), F(S["Zc"])
        self.kapp = F(S["Qt"])
        grp = [(0, 1), (2, 3, 4)]
        self.u0, self.D0_vec = self.w0.split_comps(grp)
        self.u12, self.D12_vec = self.w12.split_comps(grp)
        self.us, self.Ds_vec = self.ws.split_comps(grp)
        self.u1, self.D1_vec = self.w1.split_comps(grp)
        self.tau0, self.tau1 = sym(self.tau0_vec.expr()), sym(self.tau1_vec.expr())
        self.bcu, self._bct = lid_boundary_conditions(S["W"])
        self.t = dt  # :277
        self.err_array = []
        self.last = {}
        self._forms()

    def _forms(self):
        S, p = self.S, self.p
        dt, Re, We, betav, th, conv = p["dt"], p["Re"], p["We"], p["betav"], p["th"], p["conv"]
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        pt, q = fem.TrialFunction(Q), fem.TestFunction(Q)
        tau, Rt = sym(fem.TrialFunction(Zc)), sym(fem.TestFunction(Zc))
        h = CellDiameter()
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        p0, p1 = self.p0.expr(), self.p1.expr()
        tau0 = self.tau0
        D0, D12 = sym(self.D0_vec), sym(self.D12_vec)
        kapp = self.kapp.expr()

        DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx  # :260
        DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx  # :313
        # velocity half step :317-326
        a1 = inner((Re / (dt / 2.0)) * u, v) * dx + inner(betav * grad(u), grad(v)) * dx \
            + (inner(D - Dincomp(u), R) * dx)
        L1 = inner((Re / (dt / 2.0)) * u0 - Re * conv * grad(u0) * u0, v) * dx + inner(p0, div(v)) * dx \
            - inner(tau0, grad(v)) * dx
        a1 += th * DEVSSl_u12
        L1 += th * DEVSSr_u12
        # predictor :355-359
        a2 = inner((Re / dt) * u, v) * dx + inner(D - Dincomp(u), R) * dx
        L2 = inner((Re / dt) * u0 - Re * conv * grad(u12) * u12, v) * dx \
            - 0.5 * betav * (inner(grad(u0), grad(v)) * dx) + inner(p0, div(v)) * dx - inner(tau0, grad(v)) * dx
        # pressure :370-371
        a5 = inner(grad(pt), grad(q)) * dx
        L5 = inner(grad(p0), grad(q)) * dx + (Re / dt) * inner(us, grad(q)) * dx
        # velocity update :379-385
        a7 = inner((Re / dt) * u, v) * dx + inner(0.5 * betav * grad(u), grad(v)) * dx \
            + inner(D - Dincomp(u), R) * dx
        L7 = inner((Re / dt) * us, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx
        # stress :400-409 with LPS :309
        LPSl_stress = inner(kapp * h * p["c1"] * grad(tau), grad(Rt)) * dx \
            + inner(kapp * h * p["c2"] * div(tau), div(Rt)) * dx
        F1 = dot(u1, grad(tau)) - dot(grad(u1), tau) - dot(tau, tgrad(u1))
        a4 = inner((We / dt + 1.0) * tau + We * F1, Rt) * dx + LPSl_stress
        L4 = inner((We / dt) * tau0 + 2.0 * (1.0 - betav) * Dincomp(u12), Rt) * dx
        self.forms = dict(a1=a1, L1=L1, a2=a2, L2=L2, a5=a5, L5=L5, a7=a7, L7=L7, a4=a4, L4=L4)
        self.E_k_form = 0.5 * dot(u1, u1) * dx  # :420
        self.E_e_form = (self.tau1[0, 0] + self.tau1[1, 1]) * dx  # :421
        self.err_expr = h * kapp  # :434

    def step(self):
        f, last = self.forms, self.last
        self._bct["t"] = self.t  # ulidreg.t = t :293
        self._lps_indicator()
        for name, a, L, x in (("1", "a1", "L1", self.w12), ("2", "a2", "L2", self.ws)):
            A, b = fem.assemble(f[a]), fem.assemble(f[L])
            [bc.apply(A, b) for bc in self.bcu]
            self._solve(A, x.vec, b)
            last["A" + name], last["b" + name] = A, b
        A5, b5 = fem.assemble(f["a5"]), fem.assemble(f["L5"])
        self.p1.vec[:] = solve_zero_sum(A5, b5)  # bcp = [] -> singular, SURVEY.md 7/H3
        A7, b7 = fem.assemble(f["a7"]), fem.assemble(f["L7"])
        [bc.apply(A7, b7) for bc in self.bcu]
        self._solve(A7, self.w1.vec, b7)
        A4, b4 = fem.assemble(f["a4"]), fem.assemble(f["L4"])
        self._solve(A4, self.tau1_vec.vec, b4)
        last.update(A5=A5, b5=b5, A7=A7, b7=b7, A4=A4, b4=b4)
        E_k, E_e = fem.assemble(self.E_k_form), fem.assemble(self.E_e_form)
        t_rec = self.t
        self.t += self.p["dt"]  # :429
        err = fem.project(self.err_expr, self.S["Qt"])
        self.err_array.append(fem.norm(err.vec, "l2"))
        diverged = False
        if self.t > 0.5 and len(self.err_array) > 1 and self.err_array[-2] != 0.0:
            diverged = self.err_array[-1] / self.err_array[-2] > 1.1  # :437-441
        self.w0.assign(self.w1)  # :444-447
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        return dict(t=t_rec, E_k=E_k, E_e=E_e, err=self.err_array[-1], diverged=diverged)

    def fields(self):
        return dict(w1=self.w1.vec, p1=self.p1.vec - self.p1.vec.mean(), tau1=self.tau1_vec.vec,
                    kapp=self.kapp.vec)


class CompLDC(_LDCBase):
    """Config 2: lid-driven-cavity/comp_LDC_mesh_convergence.py time loop :376-565."""
    RE_CONV = False

    def __init__(self, mesh, dofmaps=None, dt=0.001, Re=10.0, We=0.25, Ma=0.001, betav=0.5,
                 c1=0.1, c2=0.01, th=1.0, conv=1):
        self.p = dict(dt=dt, Re=Re, We=We, Ma=Ma, betav=betav, c1=c1, c2=c2, th=th, conv=conv)
        self.mesh, self.S = mesh, make_spaces(mesh, dofmaps)
        S = self.S
        F = fem.Function
        self.rho0, self.rho12, self.rho1 = F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.p0, self.p1 = F(S["Q"]), F(S["Q"])
        self.w0, self.w12, self.ws, self.w1 = F(S["W"]), F(S["W"]), F(S["W"]), F(S["W"])
        self.tau0_vec, self.tau1_vec = F(S["Zc"]), F(S["Zc"])
        self.kapp = F(S["Qt"])
        grp = [(0, 1), (2, 3, 4)]
        self.u0, self.D0_vec = self.w0.split_comps(grp)
        self.u12, self.D12_vec = self.w12.split_comps(grp)
        self.us, self.Ds_vec = self.ws.split_comps(grp)
        self.u1, self.D1_vec = self.w1.split_comps(grp)
        self.tau0, self.tau1 = sym(self.tau0_vec.expr()), sym(self.tau1_vec.expr())
        self.bcu, self._bct = lid_boundary_conditions(S["W"])
        self.rho0.assign(fem.project(1.0, S["Q"]))  # :285-286
        self.t = dt  # :373
        self.last = {}
        self._forms()

    def _forms(self):
        S, p = self.S, self.p
        dt, Re, We, Ma, betav, th, conv = (p[k] for k in ("dt", "Re", "We", "Ma", "betav", "th", "conv"))
        if self.RE_CONV:  # comp_LDC.py:418,468 write Re*conv*dot(u, nabla_grad(u)) where the mesh-convergence script has conv*...
            conv = Re * conv
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        rho = pt = fem.TrialFunction(Q)
        q = fem.TestFunction(Q)
        tau, Rt = sym(fem.TrialFunction(Zc)), sym(fem.TestFunction(Zc))
        h, n = CellDiameter(), FacetNormal()
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        rho0, rho1 = self.rho0.expr(), self.rho1.expr()
        p0, p1 = self.p0.expr(), self.p1.expr()
        tau0 = self.tau0
        D0, D12 = sym(self.D0_vec), sym(self.D12_vec)
        kapp = self.kapp.expr()

        LPSl_stress = inner(kapp * h * p["c1"] * grad(tau), grad(Rt)) * dx \
            + inner(kapp * h * p["c2"] * div(tau), div(Rt)) * dx  # :402
        DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx  # :346
        DEVSSl_u1 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx  # :348
        DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx  # :412
        DEVSSr_u1 = 2 * (1 - betav) * inner(D12, Dincomp(v)) * dx  # :454
        U = 0.5 * (u + u0)  # :414
        # density half step :417-421
        rho_eq = 2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0)
        rho_weak = inner(rho_eq, q) * dx
        a0, L0 = lhs(rho_weak), rhs(rho_weak)
        # velocity half step :430-441
        lhsFu12 = Re * rho0 * (2.0 * (u - u0) / dt + conv * dot(u0, nabla_grad(u0)))
        Fu12 = dot(lhsFu12, v) * dx \
            + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
            + dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds \
            - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
            - dot(tau0 * n, v) * ds \
            + inner(D - Dcomp(u), R) * dx
        a1, L1 = lhs(Fu12), rhs(Fu12)
        a1 += th * DEVSSl_u12
        L1 += th * DEVSSr_u12
        # predictor :458-470
        lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u12, nabla_grad(u12)))
        Fus = dot(lhsFus, v) * dx \
            + inner(sigma(U, p0, tau0, We, betav), Dincomp(v)) * dx \
            + 0.5 * dot(p0 * n, v) * ds - dot(betav * nabla_grad(U) * n, v) * ds \
            - (1.0 / 3) * betav * dot(div(U) * n, v) * ds \
            - dot(tau0 * n, v) * ds \
            + inner(D - Dcomp(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSl_u1
        L2 += th * DEVSSr_u1
        # pressure :483-490
        lhs_p_1 = (Ma * Ma / (dt * Re)) * pt
        rhs_p_1 = (Ma * Ma / (dt * Re)) * p0 - rho0 * div(us) + dot(grad(rho0), us)
        a5 = inner(lhs_p_1, q) * dx + inner(0.5 * dt * grad(pt), grad(q)) * dx
        L5 = inner(rhs_p_1, q) * dx + inner(0.5 * dt * grad(p0), grad(q)) * dx
        # equation of state :501-502
        self.rho1_expr = rho0 + (Ma * Ma / Re) * (p1 - p0)
        # velocity update :506-510
        a7 = inner((Re / dt) * rho1 * u, v) * dx + inner(D - Dcomp(u), R) * dx
        L7 = inner((Re / dt) * rho1 * us, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx \
            - 0.5 * dot(p1 * n, v) * ds
        # stress :528-536
        lhs_tau1 = (We / dt + 1.0) * tau + We * Fdef(u1, tau)
        rhs_tau1 = (We / dt) * tau0 + 2.0 * (1.0 - betav) * Dcomp(u1)
        A = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
        a4, L4 = lhs(A), rhs(A)
        a4 += LPSl_stress
        self.forms = dict(a0=a0, L0=L0, a1=a1, L1=L1, a2=a2, L2=L2, a5=a5, L5=L5, a7=a7, L7=L7,
                          a4=a4, L4=L4)
        self.E_k_form = 0.5 * rho1 * dot(u1, u1) * dx  # :548
        tv = self.tau1_vec.expr()
        self.E_e_form = (tv[0] + tv[2]) * dx  # :549

    def step(self):
        f, last, S = self.forms, self.last, self.S
        self._bct["t"] = self.t
        self._lps_indicator()
        A0, b0 = fem.assemble(f["a0"]), fem.assemble(f["L0"])
        self._solve(A0, self.rho12.vec, b0)
        last.update(A0=A0, b0=b0)
        for name, a, L, x in (("1", "a1", "L1", self.w12), ("2", "a2", "L2", self.ws)):
            A, b = fem.assemble(f[a]), fem.assemble(f[L])
            [bc.apply(A, b) for bc in self.bcu]
            self._solve(A, x.vec, b)
            last["A" + name], last["b" + name] = A, b
        A5, b5 = fem.assemble(f["a5"]), fem.assemble(f["L5"])
        self._solve(A5, self.p1.vec, b5)
        self.rho1.assign(fem.project(self.rho1_expr, S["Q"]))
        A7, b7 = fem.assemble(f["a7"]), fem.assemble(f["L7"])
        [bc.apply(A7, b7) for bc in self.bcu]
        self._solve(A7, self.w1.vec, b7)
        A4, b4 = fem.assemble(f["a4"]), fem.assemble(f["L4"])
        self._solve(A4, self.tau1_vec.vec, b4)
        last.update(A5=A5, b5=b5, A7=A7, b7=b7, A4=A4, b4=b4)
        E_k, E_e = fem.assemble(self.E_k_form), fem.assemble(self.E_e_form)
        t_rec = self.t
        self.w0.assign(self.w1)  # :558-562
        self.rho0.assign(self.rho1)
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        self.t += self.p["dt"]
        return dict(t=t_rec, E_k=E_k, E_e=E_e)

    def fields(self):
        return dict(w1=self.w1.vec, p1=self.p1.vec, tau1=self.tau1_vec.vec, rho1=self.rho1.vec,
                    rho12=self.rho12.vec, kapp=self.kapp.vec)


class CompLDCSweep(CompLDC):
    """lid-driven-cavity/comp_LDC.py time loop :364-594 (the Re/We/Ma sweep script).  Statement for statement
    the loop of comp_LDC_mesh_convergence.py (the commented-out stress / temperature half steps aside), except
    that the convective terms of the two velocity predictor steps carry `Re*conv` (:418, :468)."""
    RE_CONV = True
