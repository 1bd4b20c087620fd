"""ORACLE (test infrastructure, NOT product code) -- config 4: FENE-P-MP journal bearing,
"flow between eccentrically rotating cylinders"/fenepmp_jbp.py time loop :454-690, restated with
oracle/ufl_lite + oracle/fem in the script's statement order.  parity unpinned (see ufl_lite.py).
Helper algebra follows fenics_base.py:158-294 (note: this module's Fdef INCLUDES + div(u)*Tau)."""
import numpy as np

from . import fem
from .configs import Dcomp, Dincomp, sym, tgrad
from .ufl_lite import (CellDiameter, Const, FacetNormal, Identity, as_matrix, as_vector, div, dot, ds, dx, grad,
                       inner, lhs, nabla_grad, rhs)


def Fdef(u, Tau):  # fenics_base.py:189-190
    return dot(u, grad(Tau)) - dot(grad(u), Tau) - dot(Tau, tgrad(u)) + div(u) * Tau


def FdefG(u, G, Tau):  # fenics_base.py:192-193
    return dot(u, grad(Tau)) - dot(G, Tau) - dot(Tau, G.T)


def Tr(tau):  # fenics_base.py:246-247
    return abs(tau[0, 0] + tau[1, 1])


def fene_func(tau, b):  # fenics_base.py:255-257
    return 1.0 / (1. - (Tr(tau) - 2.) / (b * b))


def I_3(A):
    return A[0, 0] * A[1, 1] - A[1, 0] * A[0, 1]


def I_2(A):
    return 0.5 * (A[0, 0] * A[0, 0] + A[1, 1] * A[1, 1] + 2 * A[1, 0] * A[0, 1])


def phi_def(u, lambda_d):  # fenics_base.py:287-291
    D_u = as_matrix([[Dincomp(u)[0, 0], Dincomp(u)[0, 1]], [Dincomp(u)[1, 0], Dincomp(u)[1, 1]]])
    return 1. + (lambda_d * 3 * I_3(D_u) / (I_2(D_u) + 0.0000001)) ** 2


def phi_s(T, A_0):  # fenics_base.py:227-238
    T_0, T_h, k_b = 300.0, 350.0, 8.3144598
    A = A_0 / (k_b * (T_h - T_0))
    C = T_h / (T_h - T_0)
    return 1. - A * (1. / (C + T))


def fene_sigmacom(u, p, Tau, b, lambda_d, betav, We):  # fenics_base.py:186-187
    return 2.0 * betav * Dcomp(u) - p * Identity(len(u)) + ((1. - betav) / We) * (
        phi_def(u, lambda_d) * (fene_func(Tau, b) * Tau - Identity(len(u))))


def magnitude(u):  # fenics_base.py:204-205 (np.power -> UFL Power)
    return (u[0] * u[0] + u[1] * u[1]) ** 0.5


class FenepmpJBP:
    """Config 4.  Geometry: journal (x_1,y_1,r_a) in bearing (x_2,y_2,r_b) (:141-151); markers
    omega0 (journal) = 0, omega1 (bearing) = 1 (:212-237)."""

    def __init__(self, mesh, dofmaps=None, dt=0.00125, Re=25.0, We=0.1, Ma=0.01, lambda_d=0.0, betav=0.5, b=20.0,
                 A_0=1.0, Di=0.005, Vh=0.0000069, Bi=0.2, T_0=300.0, T_h=350.0, th=1.0, conv=1,
                 x_1=-0.80, y_1=0.0, x_2=0.0, y_2=0.0, r_a=1.0, r_b=2.0):
        from .configs import make_spaces
        self.p = dict(dt=dt, Re=Re, We=We, Ma=Ma, lambda_d=lambda_d, betav=betav, b_fene=b, a0=A_0, Di=Di, Vh=Vh,
                      Bi=Bi, T0=T_0, th_hot=T_h, th=th, conv=conv)
        self.geo = dict(x_1=x_1, y_1=y_1, x_2=x_2, y_2=y_2, r_a=r_a, r_b=r_b)
        self.mesh, self.S = mesh, make_spaces(mesh, dofmaps)
        S, F = self.S, fem.Function
        g = self.geo
        self.in_omega0 = lambda x: (x[:, 0] - g["x_1"]) ** 2 + (x[:, 1] - g["y_1"]) ** 2 < (0.9 * r_a ** 2 + 0.1 * r_b ** 2)
        self.in_omega1 = lambda x: (x[:, 0] - g["x_2"]) ** 2 + (x[:, 1] - g["y_2"]) ** 2 > (0.1 * r_a ** 2 + 0.9 * r_b ** 2)
        mesh.mark_boundary([(0, self.in_omega0), (1, self.in_omega1)], default=0)
        self.rho0, self.rho12, self.rho1 = F(S["Q"]), F(S["Q"]), F(S["Q"])
        self.p0, self.p1, self.T0, self.T1 = F(S["Q"]), F(S["Q"]), F(S["Q"]), F(S["Q"])
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

        def wspin(x):  # Expression w (:249)
            f = 0.5 * (1.0 + np.tanh(8 * (self._tt["t"] - 0.5))) if getattr(self, "ramp", True) else 1.0
            return np.stack([f * (x[:, 1] - y_1) / r_a, -f * (x[:, 0] - x_1) / r_a], 1)

        spin = fem.DirichletBC(S["W"], (0, 1), wspin, self.in_omega0)
        noslip = fem.DirichletBC(S["W"], (0, 1), (0.0, 0.0), self.in_omega1)
        self.bcu = [noslip, spin]  # :255
        self.bcT = [fem.DirichletBC(S["Q"], (0,), T_h, self.in_omega0)]  # :253,257
        # initial state :367-382
        self.rho0.assign(fem.project(1.0, S["Q"]))
        self.I_vec = Const(np.array([1.0, 0.0, 1.0]), degree=2)  # Expression(('1.0','0.0','1.0'), degree=2)
        self.tau0_vec.assign(fem.project(self.I_vec, S["Zc"]))
        self.T0.assign(fem.project(T_0, S["Q"]))
        self.thetar = fem.project(T_0 / (T_h - T_0), S["Q"])  # :388-389
        self.t, self.iter = 0.0, 0  # :448-449
        self.last = {}
        self._forms()

    def _forms(self):
        S, p = self.S, self.p
        dt, Re, We, Ma, lam, betav, b, A_0, Di, Vh, Bi, T_0, T_h, th, conv = (
            p[k] for k in ("dt", "Re", "We", "Ma", "lambda_d", "betav", "b_fene", "a0", "Di", "Vh", "Bi", "T0",
                           "th_hot", "th", "conv"))
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        rho = pt = T = fem.TrialFunction(Q)
        q = r = fem.TestFunction(Q)
        tau_vec_t = fem.TrialFunction(Zc)
        tau, Rt = sym(tau_vec_t), sym(fem.TestFunction(Zc))
        h, n = CellDiameter(), FacetNormal()
        tang = as_vector([n[1], -n[0]])  # :182
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        rho0, rho12, rho1 = self.rho0.expr(), self.rho12.expr(), self.rho1.expr()
        p0, p1, T0 = self.p0.expr(), self.p1.expr(), self.T0.expr()
        tau0, tau1 = self.tau0, self.tau1
        tau0_vec, tau1_vec = self.tau0_vec.expr(), self.tau1_vec.expr()
        D0, D1 = sym(self.D0_vec), sym(self.D1_vec)
        kapp = self.kapp.expr()
        DOLFIN_EPS = fem.DOLFIN_EPS
        thetal = (T) / (T_h - T_0)  # :387
        thetar = self.thetar.expr()
        theta0 = (T0 - T_0) / (T_h - T_0)  # :390
        # LPS residual :462-468
        F1R = Fdef(u1, tau1)
        F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
        dissipation = We * 0.5 * (phi_def(u1, lam) - 1.) * (tau1 * Dincomp(u1) + Dincomp(u1) * tau1)
        diss_vec = as_vector([dissipation[0, 0], dissipation[1, 0], dissipation[1, 1]])
        self.restau0 = We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + fene_func(tau0, b) * tau1_vec - diss_vec \
            - self.I_vec
        LPSl_stress = inner(kapp * h * 0.05 * grad(tau), grad(Rt)) * dx \
            + inner(kapp * h * 0.01 * div(tau), div(Rt)) * dx  # :473
        DEVSSGl_u1 = 2.0 * (1. - betav) * inner(Dincomp(u), Dincomp(v)) * dx  # :416
        DEVSSGr_u1 = (1. - betav) * inner(D0 + D0.T, Dincomp(v)) * dx  # :495
        U = 0.5 * (u + u0)  # :496
        # density half step :501-505
        rho_weak = inner(2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0), q) * dx
        a0, L0 = lhs(rho_weak), rhs(rho_weak)
        # predictor :513-524
        lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u0, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(
                div(phi_def(u0, lam) * (fene_func(tau0, b) * tau0 - Identity(len(u)))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSGl_u1
        L2 += th * DEVSSGr_u1
        # pressure :536-543
        a5 = inner((Ma * Ma / (dt)) * pt, q) * dx + inner(dt * grad(pt), grad(q)) * dx
        L5 = inner((Ma * Ma / (dt)) * p0 - Re * div(rho0 * us), q) * dx + inner(dt * grad(p0), grad(q)) * dx
        # density update :554-564
        alpha_supg = h / (magnitude(u12) + 0.0000000001)
        SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
        rho_weak = inner((rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12), q) * dx
        a6, L6 = lhs(rho_weak), rhs(rho_weak)
        a6 += SU_rho
        # velocity update :573-580
        a7 = inner((Re / dt) * rho1 * u, v) * dx + inner(D - grad(u), R) * dx
        L7 = inner((Re / dt) * rho0 * us, v) * dx - 0.5 * inner(grad(p1 - p0), (v)) * dx
        a7 += th * DEVSSGl_u1
        L7 += th * DEVSSGr_u1
        # stress :595-606
        lhs_tau1 = (We / dt) * tau + fene_func(tau0, b) * tau + We * FdefG(u1, D1, tau) \
            - We * 0.5 * (phi_def(u1, lam) - 1.) * (tau * Dincomp(u1) + Dincomp(u1) * tau)
        rhs_tau1 = (We / dt) * tau0 + Identity(len(u))
        Astress = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
        a4, L4 = lhs(Astress), rhs(Astress)
        a4 += LPSl_stress
        # temperature :615-624
        gamdot = inner(fene_sigmacom(u0, p0, tau0, b, lam, betav, We), grad(u0))
        lhs_temp1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
        difflhs_temp1 = Di * grad(thetal) + (Di / We) * tau1 * grad(thetal)
        rhs_temp1 = (1.0 / dt) * rho1 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho1 * theta0 + Vh * gamdot
        diffrhs_temp1 = Di * grad(thetar) + (Di / We) * tau1 * grad(thetar)
        a9 = inner(lhs_temp1, r) * dx + inner(difflhs_temp1, grad(r)) * dx
        L9 = inner(rhs_temp1, r) * dx + inner(diffrhs_temp1, grad(r)) * dx - Di * Bi * inner(theta0, r) * ds(1)
        self.forms = dict(a0=a0, L0=L0, a2=a2, L2=L2, a5=a5, L5=L5, a6=a6, L6=L6, a7=a7, L7=L7, a4=a4, L4=L4,
                          a9=a9, L9=L9)
        # functionals :635-649
        self.E_k_form = 0.5 * rho1 * dot(u1, u1) * dx
        self.E_e_form = (fene_func(tau1, b) * (tau1[0, 0] + tau1[1, 1]) - 2.) * dx
        sig = fene_sigmacom(u1, p1, tau1, b, lam, betav, We)
        sigma0, omegaf0 = dot(sig, tang), dot(sig, n)
        self.force_x_form = inner(Const(np.array([1.0, 0.0])), omegaf0) * ds(0)
        self.force_y_form = inner(Const(np.array([0.0, -1.0])), omegaf0) * ds(0)
        self.torque_form = -inner(n, sigma0) * ds(0)

    def lps_indicator(self):
        S = self.S
        res_test = fem.project(self.restau0, S["Zd"])
        res_orth = fem.project(self.restau0 - res_test.expr(), S["Zc"])
        res_orth_norm_sq = fem.project(inner(res_orth.expr(), res_orth.expr()), S["Qt"])
        kapp = fem.project(res_orth_norm_sq.expr() ** 0.5, S["Qt"])
        kapp.vec[:] = np.absolute(kapp.vec)  # absolute(kapp) :472
        self.kapp.assign(kapp)
        self.last.update(res_test=res_test.vec, res_orth=res_orth.vec, res_orth_norm_sq=res_orth_norm_sq.vec,
                         kapp=kapp.vec)

    def step(self):
        f, last, S = self.forms, self.last, self.S
        self.iter += 1
        self._tt["t"] = self.t  # w.t = t :458
        self.lps_indicator()
        if self.iter > 1:  # :481-486
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

        system("a0", "L0", self.rho12, [])
        system("a2", "L2", self.ws, self.bcu)
        system("a5", "L5", self.p1, [])
        system("a6", "L6", self.rho1, [])
        system("a7", "L7", self.w1, self.bcu)
        system("a4", "L4", self.tau1_vec, [])
        system("a9", "L9", self.T1, self.bcT)
        out = dict(t=self.t, E_k=fem.assemble(self.E_k_form), E_e=fem.assemble(self.E_e_form),
                   torque=fem.assemble(self.torque_form), F_x=fem.assemble(self.force_x_form),
                   F_y=fem.assemble(self.force_y_form))
        self.t += self.p["dt"]
        return out

    REFINE_RATIO, REFINE_EPS, MAX_REFINE = 0.3, 0.000001, 1  # fenepmp_jbp.py:701,704,708

    def refinement_indicator(self, err_count=0):
        """post-loop adaptive-refinement indicator, fenepmp_jbp.py:691-711 (same statements at
        incompressible_benchmarkEWM.py:903-924, comp_ewm_jpb.py:735-753) + adaptive_refinement's cell marking
        (fenics_base.py:143-155)."""
        S, p = self.S, self.p
        We, dt = p["We"], p["dt"]
        tau0_vec, tau1_vec = self.tau0_vec.expr(), self.tau1_vec.expr()
        F1R = Fdef(self.u1, self.tau1)
        F1R_vec = as_vector([F1R[0, 0], F1R[1, 0], F1R[1, 1]])
        restau0 = We / dt * (tau1_vec - tau0_vec) + We * F1R_vec + tau1_vec - self.I_vec
        res_test = inner(restau0, restau0)
        kapp = fem.project(res_test, S["Qt"])
        kapp.vec[:] = kapp.vec / np.max(np.abs(kapp.vec))  # normalize_solution (in place: norm_kapp is kapp)
        ratio = self.REFINE_RATIO / (1 * err_count + 1.0)
        tau_average = fem.project((tau1_vec[0] + tau1_vec[1] + tau1_vec[2]) / 3.0, S["Qt"])
        error_rat = fem.project(kapp.expr() / (tau_average.expr() + self.REFINE_EPS), S["Qt"])
        error_rat.vec[:] = np.absolute(error_rat.vec)
        level = np.percentile(kapp.vec, (1 - ratio) * 100)
        markers = kapp.vec[np.asarray(S["Qt"].cell_dofs)[:, 0]] > level  # kapp(cell midpoint) > kapp_level
        self.res_sq_form = res_test
        return dict(kapp=kapp.vec.copy(), error_rat=error_rat.vec.copy(), error_rat_max=float(error_rat.vec.max()),
                    ratio=ratio, refine=bool(error_rat.vec.max() > 0.01 and err_count < self.MAX_REFINE), markers=markers)

    def fields(self):
        return dict(w1=self.w1.vec, p1=self.p1.vec, tau1=self.tau1_vec.vec, rho1=self.rho1.vec, T1=self.T1.vec,
                    kapp=self.kapp.vec)
