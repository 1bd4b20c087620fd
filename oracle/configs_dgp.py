"""ORACLE (test infrastructure, NOT product code) -- config 5: compressible buoyancy-driven flow,
buoyancy-driven-flow/compressible_dgp.py time loop :368-595, restated with oracle/ufl_lite + oracle/fem
in the script's statement order.  parity unpinned (see ufl_lite.py).  Helper algebra:
buoyancy-driven-flow/fenics_fem.py:52-87.

Reference defect reproduced on request: `iter` is incremented twice per pass (:369,:377), so the
start-of-step roll (:396-401) already runs in the FIRST step and overwrites the initial state with the
never-initialised (zero) rho1, T1, p1, tau1 -- T0 = 0 makes therm = al_1 + al_2*theta0 = 0 and
therm_inv = project(1/0) is inf.  `faithful_first_roll=False` (default) initialises w1, T1, rho1, p1, tau1
to the initial state instead, which is what the roll evidently intends."""
import numpy as np

from . import fem
from .configs import Dcomp, Dincomp, sym, tgrad
from .ufl_lite import (CellDiameter, Const, FacetNormal, Identity, as_vector, div, dot, ds, dx, grad, inner, lhs,
                       nabla_grad, rhs)


def sigmacom(u, p, Tau, We, Pr, betav):  # fenics_fem.py:80-81
    return 2 * Pr * betav * Dcomp(u) - p * Identity(len(u)) + Pr * ((1 - betav) / (We)) * (Tau - Identity(len(u)))


def Fdefcom(u, Tau):  # fenics_fem.py:86-87
    return dot(u, grad(Tau)) - dot(grad(u), Tau) - dot(Tau, tgrad(u)) + div(u) * Tau


def magnitude(u):  # fenics_fem.py:113
    return (u[0] * u[0] + u[1] * u[1]) ** 0.5


def dgp_mesh(mm, B=1, L=1):
    """DGP_Mesh(mm,B,L) (fenics_fem.py:261-311): RectangleMesh mapped by xskewcavity."""
    return fem.rectangle_mesh(0, 0, B, L, mm * B, mm * L, fem.xskewcavity)


class CompDGP:
    def __init__(self, mesh, dofmaps=None, dt=0.0005, Ra=100.0, We=0.1, Ma=0.05, Pr=2.0, betav=0.1, c1=0.05,
                 c2=0.001, th=0.5, Vh=0.005, Bi=0.2, T_0=300.0, T_h=350.0, al_1=1.0, al_2=1.0 / 6.0,
                 faithful_first_roll=False):
        from .configs import make_spaces
        self.p = dict(dt=dt, Ra=Ra, We=We, Ma=Ma, Pr=Pr, betav=betav, c1=c1, c2=c2, th=th, Vh=Vh, Bi=Bi, T0=T_0,
                      th_hot=T_h, al1=al_1, al2=al_2)
        self.mesh, self.S = mesh, make_spaces(mesh, dofmaps)
        S, F = self.S, fem.Function
        eps = fem.DOLFIN_EPS
        left = lambda x: x[:, 0] < 0.0 + eps  # noqa: E731
        right = lambda x: x[:, 0] > 1.0 - eps  # noqa: E731
        top = lambda x: x[:, 1] > 1.0 - eps  # noqa: E731
        bottom = lambda x: x[:, 1] < 0.0 + eps  # noqa: E731
        everything = lambda x: np.ones(len(x), dtype=bool)  # noqa: E731
        mesh.mark_boundary([(0, everything), (1, left), (2, right), (3, top), (4, bottom)])  # :163-169
        self.rho0, self.rho12, self.rho1 = F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.p0, self.p1, self.T0, self.T1 = F(S["Q"]), F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.theta0, self.therm_inv, self.theta1 = F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.w0, self.w12, self.ws, self.w1 = F(S["W"]), F(S["W"]), F(S["W"]), F(S["W"])
        self.tau0_vec, self.tau1_vec = F(S["Zc"]), F(S["Zc"])
        self.kapp = F(S["Qt"])
        grp = [(0, 1), (2, 3, 4)]
        self.u0, self.D0_vec = self.w0.split_comps(grp)
        self.u12, self.D12_vec = self.w12.split_comps(grp)
        self.us, self.Ds_vec = self.ws.split_comps(grp)
        self.u1, self.D1_vec = self.w1.split_comps(grp)
        self.tau0, self.tau1 = sym(self.tau0_vec.expr()), sym(self.tau1_vec.expr())
        self._tt = {"t": 0.0}

        def ramp(x):  # ramp_function :174
            return np.full(len(x), 0.5 * (1 + np.tanh(4 * (self._tt["t"] - 1.5))) * (T_h - T_0) + T_0)

        noslip = [fem.DirichletBC(S["W"], (0, 1), (0.0, 0.0), f) for f in (everything, left, right, top)]
        self.bcu = noslip  # :193
        self.bcT = [fem.DirichletBC(S["Q"], (0,), ramp, left), fem.DirichletBC(S["Q"], (0,), T_0, right)]  # :195
        self.I_vec = Const(np.array([1.0, 0.0, 1.0]), degree=2)
        self.tau0_vec.assign(fem.project(self.I_vec, S["Zc"]))  # :278-280
        self.rho0.assign(fem.project(1.0, S["Q"]))  # :284-286
        self.T0.assign(fem.project(T_0, S["Q"]))  # :290-291
        self.thetar = fem.project(T_0 / (T_h - T_0), S["Q"])  # :295-296
        if not faithful_first_roll:
            self.T1.assign(self.T0)
            self.rho1.assign(self.rho0)
            self.tau1_vec.assign(self.tau0_vec)
        self.t = dt  # :365
        self.last = {}
        self._forms()

    def _forms(self):
        S, p = self.S, self.p
        dt, Ra, We, Ma, Pr, betav, th, Vh, Bi, T_0, T_h, al_1, al_2 = (
            p[k] for k in ("dt", "Ra", "We", "Ma", "Pr", "betav", "th", "Vh", "Bi", "T0", "th_hot", "al1", "al2"))
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        rho = pt = T = fem.TrialFunction(Q)
        q = r = fem.TestFunction(Q)
        tau, Rt = sym(fem.TrialFunction(Zc)), sym(fem.TestFunction(Zc))
        h, n = CellDiameter(), FacetNormal()
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        rho0, rho12, rho1 = self.rho0.expr(), self.rho12.expr(), self.rho1.expr()
        p0, p1, T0, T1 = self.p0.expr(), self.p1.expr(), self.T0.expr(), self.T1.expr()
        tau0, tau1 = self.tau0, self.tau1
        tau0_vec, tau1_vec = self.tau0_vec.expr(), self.tau1_vec.expr()
        D0, D12 = sym(self.D0_vec), sym(self.D12_vec)
        kapp = self.kapp.expr()
        theta0, therm_inv, theta1 = self.theta0.expr(), self.therm_inv.expr(), self.theta1.expr()
        k = Const(np.array([0.0, 1.0]), degree=2)  # Expression(('0','1'), degree=2) :177
        thetal = (T) / (T_h - T_0)  # :293
        thetar = self.thetar.expr()
        # LPS residual :382-385
        F1R = Fdefcom(u1, tau1)
        F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
        self.restau0 = We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + tau1_vec - self.I_vec
        LPSl_stress = inner(kapp * h * p["c1"] * grad(tau), grad(Rt)) * dx \
            + inner(kapp * h * p["c2"] * div(tau), div(Rt)) * dx  # :391
        # per-step projections :404-408
        self.theta0_expr = (T0 - T_0) / (T_h - T_0)
        therm = (al_1 + al_2 * theta0)
        self.therm_inv_expr = 1.0 / therm
        DEVSSl = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx  # :333,:335
        DEVSSr_u12 = 2 * (1 - betav) * inner(D0, Dincomp(v)) * dx  # :413
        DEVSSr_u1 = 2 * (1 - betav) * inner(D12, Dincomp(v)) * dx  # :453
        U = 0.5 * (u + u0)  # :416
        rho_weak = inner(2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0), q) * dx  # :420-421
        a0, L0 = lhs(rho_weak), rhs(rho_weak)
        # u^{n+1/2} :432-442
        Du12Dt = rho0 * (2.0 * (u - u0) / dt + dot(u0, nabla_grad(u0)))
        Fu12 = dot(Du12Dt, v) * dx \
            + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(rho0 * therm_inv * k, v) * dx \
            + dot(p0 * n, v) * ds - betav * Pr * (dot(nabla_grad(U) * n, v) * ds + (1.0 / 3) * dot(div(U) * n, v) * ds) \
            - (Pr * (1 - betav) / We) * dot(tau0 * n, v) * ds \
            + inner(D - Dincomp(u), R) * dx
        a1, L1 = lhs(Fu12), rhs(Fu12)
        a1 += th * DEVSSl
        L1 += th * DEVSSr_u12
        # u* :473-483
        lhsFus = rho0 * ((u - u0) / dt + dot(u12, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx \
            + inner(sigmacom(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(rho0 * therm_inv * k, v) * dx \
            + dot(p0 * n, v) * ds - betav * Pr * (dot(nabla_grad(U) * n, v) * ds + (1.0 / 3) * dot(div(U) * n, v) * ds) \
            - (Pr * (1.0 - betav) / We) * dot(tau0 * n, v) * ds \
            + inner(D - Dincomp(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSl
        L2 += th * DEVSSr_u1
        # pressure :493-500
        lhs_p_1 = (Ma * Ma / (dt)) * pt
        rhs_p_1 = (Ma * Ma / (dt)) * p0 - therm * dot(grad(rho0), us) - therm * rho0 * div(us)
        a5 = inner(lhs_p_1, q) * dx + inner(therm * 0.5 * dt * grad(pt), grad(q)) * dx
        L5 = inner(rhs_p_1, q) * dx + inner(therm * 0.5 * dt * grad(p0), grad(q)) * dx
        # SU density update :510-521 (result overwritten by the projection below -- dead work, kept)
        alpha_supg = h / (magnitude(u12) + fem.DOLFIN_EPS)
        SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
        rho_weak = inner((rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12), q) * dx
        a6, L6 = lhs(rho_weak), rhs(rho_weak)
        a6 += SU_rho
        self.rho1_expr = rho0 + (Ma * Ma) * therm_inv * (p1 - p0)  # :529
        # u^{n+1} :533-537
        a7 = inner((1. / dt) * rho1 * u, v) * dx + inner(D - Dincomp(u), R) * dx
        L7 = inner((1. / dt) * rho0 * us, v) * dx + 0.5 * (inner(p1 - p0, div(v)) * dx - dot((p1 - p0) * n, v) * ds)
        # stress :550-555
        stress_eq = ((We / dt) + 1.0) * tau + We * Fdefcom(u1, tau) - (We / dt) * tau0 - Identity(len(u))
        A = inner(stress_eq, Rt) * dx
        a4, L4 = lhs(A), rhs(A)
        a4 += LPSl_stress
        # temperature :563-567
        gamdot = inner(sigmacom(u1, p1, tau1, We, Pr, betav), grad(u1))
        lhs_theta1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
        rhs_theta1 = (1.0 / dt) * rho0 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho0 * theta0 + Vh * gamdot
        a8 = inner(lhs_theta1, r) * dx + inner(grad(thetal), grad(r)) * dx
        L8 = inner(rhs_theta1, r) * dx + inner(grad(thetar), grad(r)) * dx + Bi * inner(grad(theta0), n * r) * ds(3) \
            + Bi * inner(grad(theta0), n * r) * ds(4) + inner(We * tau1 * grad(thetar), grad(r)) * dx
        self.forms = dict(a0=a0, L0=L0, a1=a1, L1=L1, a2=a2, L2=L2, a5=a5, L5=L5, a6=a6, L6=L6, a7=a7, L7=L7,
                          a4=a4, L4=L4, a8=a8, L8=L8)
        self.E_k_form = 0.5 * rho1 * dot(u1, u1) * dx  # :577
        self.E_e_form = (tau1[0, 0] + tau1[1, 1] - 2.0) * dx  # :578
        self.theta1_expr = (T1 - T_0) / (T_h - T_0)  # :581
        self.nus_form = -inner(grad(theta1), n) * ds(2)  # :582-583

    def lps_indicator(self):
        S = self.S
        res_test = fem.project(self.restau0, S["Zd"])
        res_orth = fem.project(self.restau0 - res_test.expr(), S["Zc"])
        res_orth_norm_sq = fem.project(inner(res_orth.expr(), res_orth.expr()), S["Qt"])
        kapp = fem.project(res_orth_norm_sq.expr() ** 0.5, S["Qt"])
        self.kapp.assign(kapp)
        self.last.update(res_test=res_test.vec, res_orth=res_orth.vec, kapp=kapp.vec)

    def step(self):
        f, last, S = self.forms, self.last, self.S
        self._tt["t"] = self.t  # :379
        self.lps_indicator()  # :382-391
        self.w0.assign(self.w1)  # :396-401 (`iter > 1` holds from the first pass)
        self.T0.assign(self.T1)
        self.rho0.assign(self.rho1)
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        self.theta0.assign(fem.project(self.theta0_expr, S["Q"]))  # :405
        self.therm_inv.assign(fem.project(self.therm_inv_expr, S["Q"]))  # :408

        def system(a, L, x, bcs):
            A, b = fem.assemble(f[a]), fem.assemble(f[L])
            [bc.apply(A, b) for bc in bcs]
            x.vec[:] = fem.solve_lu(A, b)
            last["A" + a[1:]], last["b" + a[1:]] = A, b

        system("a0", "L0", self.rho12, [])
        system("a1", "L1", self.w12, self.bcu)
        system("a2", "L2", self.ws, self.bcu)
        system("a5", "L5", self.p1, [])
        system("a6", "L6", self.rho1, [])
        self.rho1.assign(fem.project(self.rho1_expr, S["Q"]))  # :529-530
        system("a7", "L7", self.w1, self.bcu)
        system("a4", "L4", self.tau1_vec, [])
        system("a8", "L8", self.T1, self.bcT)
        E_k, E_e = fem.assemble(self.E_k_form), fem.assemble(self.E_e_form)
        self.theta1.assign(fem.project(self.theta1_expr, S["Q"]))  # :581
        Nus = fem.assemble(self.nus_form)
        t_rec = self.t
        self.t += self.p["dt"]
        return dict(t=t_rec, E_k=E_k, E_e=E_e, Nus=Nus)

    def fields(self):
        return dict(w1=self.w1.vec, p1=self.p1.vec, tau1=self.tau1_vec.vec, rho1=self.rho1.vec, T1=self.T1.vec,
                    kapp=self.kapp.vec)


# ---------------------------------------------------------------------------------------------------------------
def sigma(u, p, Tau, We, Pr, betav):  # fenics_fem.py:68-72
    if betav < 1:
        return 2 * Pr * betav * Dincomp(u) - p * Identity(len(u)) + Pr * ((1 - betav) / (We)) * (Tau - Identity(len(u)))
    return 2 * Pr * betav * Dincomp(u) - p * Identity(len(u))


def sigma_n(U, p, tau, n, We, Pr, betav):  # fenics_fem.py:74-78
    if betav < 1:
        return Pr * betav * nabla_grad(U) * n + (Pr * (1 - betav) / We) * tau * n - p * n
    return Pr * betav * nabla_grad(U) * n - p * n


def Fdef(u, Tau):  # fenics_fem.py:83-84
    return dot(u, grad(Tau)) - dot(grad(u), Tau) - dot(Tau, tgrad(u))


def solve_weighted_gauge(A, b, x0):
    """A is singular with the constants as null space (pure-Neumann pressure matrix, bcp = []) and b consistent.
    The reference hands it to BiCGStab (incompressible_dgp.py:454-459), whose answer depends on the initial guess
    and the preconditioner: with LEFT Jacobi preconditioning x - x0 stays in D^-1 range(A), i.e. sum_i D_ii (x - x0)_i
    = 0.  That member of the solution family is computed here exactly (bordered system), because the temperature
    step reads p1 through Vh*inner(sigma(u1, p1, tau1), grad(u1)) (:493), which is NOT invariant under p -> p + c when
    div(u1) != 0 discretely."""
    import scipy.sparse as sp
    n = A.shape[0]
    w = A.diagonal()
    K = sp.bmat([[A, sp.csr_matrix(np.ones((n, 1)))], [sp.csr_matrix(w[None, :]), None]], format="csc")
    return fem.solve_lu(K, np.concatenate([b, [w @ x0]]))[:n]


class IncompDGP:
    """buoyancy-driven-flow/incompressible_dgp.py time loop :357-535 (incompressible Oldroyd-B double-glazing problem),
    restated in the script's statement order.  Boundary markers :141-151: 0 everything, then 1 bottom, 2 left,
    3 right, 4 top.  No start-of-loop defect here: `iter` is incremented once (:358), so the state roll starts with
    the second step (:385-390)."""

    def __init__(self, mesh, dofmaps=None, dt=0.001, Ra=1000.0, We=0.1, Pr=1.0, betav=0.5, c1=0.05, c2=0.001, th=0.5,
                 Vh=0.005, Bi=0.0 + fem.DOLFIN_EPS, T_0=300.0, T_h=350.0):
        from .configs import make_spaces
        self.p = dict(dt=dt, Ra=Ra, We=We, Pr=Pr, betav=betav, c1=c1, c2=c2, th=th, Vh=Vh, Bi=Bi, T0=T_0, th_hot=T_h)
        self.mesh, self.S = mesh, make_spaces(mesh, dofmaps)
        S, F = self.S, fem.Function
        eps = fem.DOLFIN_EPS
        left = lambda x: x[:, 0] < 0.0 + eps  # noqa: E731
        right = lambda x: x[:, 0] > 1.0 - eps  # noqa: E731
        top = lambda x: x[:, 1] > 1.0 - eps  # noqa: E731
        bottom = lambda x: x[:, 1] < 0.0 + eps  # noqa: E731
        everything = lambda x: np.ones(len(x), dtype=bool)  # noqa: E731
        mesh.mark_boundary([(0, everything), (1, bottom), (2, left), (3, right), (4, top)])  # :147-151
        self.p0, self.p1, self.T0, self.T1, self.theta1 = F(S["Q"]), F(S["Q"]), F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.rho0, self.rho1 = F(S["Q"]), F(S["Q"])  # rolled (:388) but used by no form
        self.w0, self.w12, self.ws, self.w1 = F(S["W"]), F(S["W"]), F(S["W"]), F(S["W"])
        self.tau0_vec, self.tau1_vec = F(S["Zc"]), F(S["Zc"])
        self.kapp = F(S["Qt"])
        grp = [(0, 1), (2, 3, 4)]
        self.u0, self.D0_vec = self.w0.split_comps(grp)
        self.u12, self.D12_vec = self.w12.split_comps(grp)
        self.us, self.Ds_vec = self.ws.split_comps(grp)
        self.u1, self.D1_vec = self.w1.split_comps(grp)
        self.tau0, self.tau1 = sym(self.tau0_vec.expr()), sym(self.tau1_vec.expr())
        self._tt = {"t": 0.0}

        def ramp(x):  # ramp_function :156
            return np.full(len(x), 0.5 * (1 + np.tanh(4 * (self._tt["t"] - 1.5))) * (T_h - T_0) + T_0)

        self.bcu = [fem.DirichletBC(S["W"], (0, 1), (0.0, 0.0), f) for f in (everything, left, right, top)]  # :166-173
        self.bcT = [fem.DirichletBC(S["Q"], (0,), ramp, left), fem.DirichletBC(S["Q"], (0,), T_0, right)]  # :170-175
        self.I_vec = Const(np.array([1.0, 0.0, 1.0]), degree=2)
        self.tau0_vec.assign(fem.project(self.I_vec, S["Zc"]))  # :262-264
        self.T0.assign(fem.project(T_0, S["Q"]))  # :266-267
        self.thetar = fem.project(T_0 / (T_h - T_0), S["Q"])  # :274-275
        self.t, self.iter = 0.0, 0  # :350, :355
        self.last = {}
        self._forms()

    def _forms(self):
        S, p = self.S, self.p
        dt, Ra, We, Pr, betav, th, Vh, Bi, T_0, T_h = (
            p[k] for k in ("dt", "Ra", "We", "Pr", "betav", "th", "Vh", "Bi", "T0", "th_hot"))
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        pt = T = fem.TrialFunction(Q)
        q = r = fem.TestFunction(Q)
        tau, Rt = sym(fem.TrialFunction(Zc)), sym(fem.TestFunction(Zc))
        h, n = CellDiameter(), FacetNormal()
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        p0, p1, T0, T1 = self.p0.expr(), self.p1.expr(), self.T0.expr(), self.T1.expr()
        tau0, tau1 = self.tau0, self.tau1
        tau0_vec, tau1_vec = self.tau0_vec.expr(), self.tau1_vec.expr()
        D0, D12 = sym(self.D0_vec), sym(self.D12_vec)
        kapp = self.kapp.expr()
        DOLFIN_EPS = fem.DOLFIN_EPS
        f = Const(np.array([0.0, -1.0]), degree=2)  # Expression(('0','-1'), degree=2) :157
        thetal = (T) / (T_h - T_0)  # :272
        thetar = self.thetar.expr()
        theta0 = (T0 - T_0) / (T_h - T_0)  # :276
        # LPS residual :369-379
        F1R = Fdef(u1, tau1)
        F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
        self.restau0 = We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + tau1_vec - self.I_vec
        LPSl_stress = inner(kapp * h * p["c1"] * grad(tau), grad(Rt)) * dx \
            + inner(kapp * h * p["c2"] * div(tau), div(Rt)) * dx
        DEVSSl_u12 = 2 * (1 - betav + DOLFIN_EPS) * inner(Dincomp(u), Dincomp(v)) * dx  # :331
        DEVSSr_u12 = 2 * inner(D0, Dincomp(v)) * dx  # :332
        DEVSSl_u1 = 2 * (1 - betav + DOLFIN_EPS) * inner(Dincomp(u), Dincomp(v)) * dx  # :335
        DEVSSr_u1 = 2 * Pr * (1. - betav + DOLFIN_EPS) * inner(D12, Dincomp(v)) * dx  # :421
        U = 0.5 * (u + u0)  # :395
        # u^{n+1/2} :397-411
        Du12Dt = (2.0 * (u - u0) / dt + dot(u0, nabla_grad(u0)))
        Fu12 = dot(Du12Dt, v) * dx \
            + inner(sigma(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(theta0 * f, v) * dx \
            + inner(D - Dincomp(u), R) * dx \
            - dot(sigma_n(U, p0, tau0, n, We, Pr, betav), v) * ds
        a1, L1 = lhs(Fu12), rhs(Fu12)
        a1 += th * DEVSSl_u12
        L1 += th * DEVSSr_u12
        # u* :439-451
        lhsFus = ((u - u0) / dt + dot(u12, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx + inner(D - Dincomp(u), R) * dx \
            + inner(sigma(U, p0, tau0, We, Pr, betav), Dincomp(v)) * dx + Ra * Pr * inner(theta0 * f, v) * dx \
            - dot(sigma_n(U, p0, tau0, n, We, Pr, betav), v) * ds
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 = a2 + th * DEVSSl_u1
        L2 = L2 + th * DEVSSr_u1
        # pressure correction :454-460 (bcp = []: pure Neumann)
        a5 = inner(grad(pt), grad(q)) * dx
        L5 = inner(grad(p0), grad(q)) * dx + (1.0 / dt) * inner(us, grad(q)) * dx
        # u^{n+1} :463-481
        a7 = inner((1. / dt) * u, v) * dx + inner(D - Dincomp(u), R) * dx
        L7 = inner((1. / dt) * us, v) * dx + 0.5 * inner(p1 - p0, div(v)) * dx - 0.5 * dot((p1 - p0) * n, v) * ds
        # stress :486-499
        stress_eq = (We / dt + 1.0) * tau + We * Fdef(u1, tau) - (We / dt) * tau0 - Identity(len(u))
        A = inner(stress_eq, Rt) * dx
        a4, L4 = lhs(A), rhs(A)
        a4 = a4 + LPSl_stress
        # temperature :503-513
        gamdot = inner(sigma(u1, p1, tau1, We, Pr, betav), grad(u1))
        lhs_theta1 = (1.0 / dt) * thetal + dot(u1, grad(thetal))
        rhs_theta1 = (1.0 / dt) * thetar + dot(u1, grad(thetar)) + (1.0 / dt) * theta0 + Vh * gamdot
        a8 = inner(lhs_theta1, r) * dx + inner(grad(thetal), grad(r)) * dx + inner(We * tau1 * grad(thetal), grad(r)) * dx
        L8 = inner(rhs_theta1, r) * dx + inner(grad(thetar), grad(r)) * dx + Bi * inner(grad(theta0), n * r) * ds(1) \
            + Bi * inner(grad(theta0), n * r) * ds(4) + inner(We * tau1 * grad(thetar), grad(r)) * dx
        self.forms = dict(a1=a1, L1=L1, a2=a2, L2=L2, a5=a5, L5=L5, a7=a7, L7=L7, a4=a4, L4=L4, a8=a8, L8=L8)
        self.E_k_form = 0.5 * dot(u1, u1) * dx  # :517
        self.E_e_form = (tau1[0, 0] + tau1[1, 1] - 2.0) * dx  # :518
        self.theta1_expr = (T1 - T_0) / (T_h - T_0)  # :521
        self.nus_form = inner(grad(self.theta1.expr()), n) * ds(2)  # :522-523

    def lps_indicator(self):
        S = self.S
        res_test = fem.project(self.restau0, S["Zd"])
        res_orth = fem.project(self.restau0 - res_test.expr(), S["Zc"])
        res_orth_norm_sq = fem.project(inner(res_orth.expr(), res_orth.expr()), S["Qt"])
        kapp = fem.project(res_orth_norm_sq.expr() ** 0.5, S["Qt"])
        kapp.vec[:] = np.absolute(kapp.vec)  # absolute(kapp) :377
        self.kapp.assign(kapp)
        self.last.update(res_test=res_test.vec, res_orth=res_orth.vec, kapp=kapp.vec)

    def step(self):
        f, last, S = self.forms, self.last, self.S
        self.iter += 1
        self._tt["t"] = self.t  # :367
        self.lps_indicator()  # :369-379
        if self.iter > 1:  # :384-390
            self.w0.assign(self.w1)
            self.T0.assign(self.T1)
            self.rho0.assign(self.rho1)
            self.p0.assign(self.p1)
            self.tau0_vec.assign(self.tau1_vec)

        def system(a, L, x, bcs):
            A, b = fem.assemble(f[a]), fem.assemble(f[L])
            [bc.apply(A, b) for bc in bcs]
            x.vec[:] = fem.solve_lu(A, b)
            last["A" + a[1:]], last["b" + a[1:]] = A, b

        system("a1", "L1", self.w12, self.bcu)
        system("a2", "L2", self.ws, self.bcu)
        A5, b5 = fem.assemble(f["a5"]), fem.assemble(f["L5"])
        self.p1.vec[:] = solve_weighted_gauge(A5, b5, self.p1.vec)  # nonzero_initial_guess: x0 = previous p1
        last.update(A5=A5, b5=b5)
        system("a7", "L7", self.w1, self.bcu)
        if self.p["betav"] < 0.99:  # :485
            system("a4", "L4", self.tau1_vec, [])
        system("a8", "L8", self.T1, self.bcT)
        E_k, E_e = fem.assemble(self.E_k_form), fem.assemble(self.E_e_form)
        self.theta1.assign(fem.project(self.theta1_expr, S["Q"]))  # :521
        Nus = fem.assemble(self.nus_form)
        t_rec = self.t
        self.t += self.p["dt"]
        return dict(t=t_rec, E_k=E_k, E_e=E_e, Nus=Nus)

    def fields(self):
        p = self.p1.vec
        return dict(w1=self.w1.vec, p1=p - p.mean(), tau1=self.tau1_vec.vec, T1=self.T1.vec, kapp=self.kapp.vec)
