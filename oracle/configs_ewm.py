"""ORACLE (test infrastructure, NOT product code) -- config 3: extended White-Metzner journal bearing,
"flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py time loop :465-782,
restated with oracle/ufl_lite + oracle/fem in the script's statement order.  parity unpinned (see
ufl_lite.py).  Helper algebra follows fenics_base.py:158-294.

As shipped the script cannot reach its loop: it uses names it never defines,
and mshr's generate_mesh is not available offline.  The restatement keeps every form of the loop
verbatim and takes the mesh from oracle.fem.annulus_mesh.

The script runs the loop in two modes (:291-296, :459-462, :660):
  pre-pass  (adaptive_loop False, err_count < max_refine): un-ramped spin BC, dt = hmin^2, 15 steps,
            stress solve executed                                   -> EwmJBP(..., prepass=True)
  main pass (adaptive_loop True): tanh-ramped spin BC, dt = 10 hmin^2, stress solve skipped because
            We = 1e-6 <= 1e-5                                       -> EwmJBP(..., prepass=False)"""
import numpy as np

from . import fem
from .configs import Dcomp, Dincomp, sym
from .configs_jbp import FdefG, FenepmpJBP, Tr, magnitude, phi_s
from .ufl_lite import (CellDiameter, Const, FacetNormal, Identity, as_vector, div, dot, ds, dx, grad, inner, lhs,
                       nabla_grad, rhs)


def sigmacon(u, p, Tau, betav, We):  # fenics_base.py:180-181
    return 2 * betav * Dcomp(u) - p * Identity(len(u)) + ((1.0 - betav) / We) * (Tau - Identity(len(u)))


def phi_ewm(tau, T, k, B):  # fenics_base.py:240-241
    return (1. - B * (T / (1. + T))) * ((0.5 * Tr(tau) + 0.00001) ** k)


class EwmJBP(FenepmpJBP):
    """Config 3.  Geometry / markers / BC objects as config 4 (:262-302); parameters :56-100, :304-307."""
    REFINE_RATIO, REFINE_EPS, MAX_REFINE = 0.2, fem.DOLFIN_EPS, 2  # incompressible_benchmarkEWM.py:912,915; max_refine :52
    SU_RHO_EPS = 0.0000000001  # alpha_supg = h/(magnitude(u12)+0.0000000001) :603
    TEMP_DEVSS = True          # a9 += th*DEVSSl_T1 :650
    ALWAYS_STRESS = False      # stress solve only `if We > 0.00001 or adaptive_loop == False ...` :660

    def __init__(self, mesh, dofmaps=None, dt=None, prepass=False, Re=10, We=0.000001, Ma=0.000001,
                 betav=1.0 - fem.DOLFIN_EPS, A_0=0.0, k_ewm=0.0, B=0.0, al=0.001, Di=0.0025, Vh=0.0000069, Bi=0.2,
                 T_0=300, T_h=350, th=1.0, conv=1, x_1=-0.90, **geo):
        self.prepass, self.ramp = prepass, not prepass  # :291-295
        if dt is None:
            dt = (1.0 if prepass else 10.0) * mesh.hmin() ** 2  # :164, :461-462
        self.k_ewm, self.B, self.al = k_ewm, B, al
        super().__init__(mesh, dofmaps, dt=dt, Re=Re, We=We, Ma=Ma, lambda_d=0.0, betav=betav, A_0=A_0, Di=Di, Vh=Vh,
                         Bi=Bi, T_0=T_0, T_h=T_h, th=th, conv=conv, x_1=x_1, **geo)
        self.p["al1"] = al
        self.p.update(k_ewm=k_ewm, b_ewm=B, th_t=(th if self.TEMP_DEVSS else 0.0))

    def _forms(self):
        S, p = self.S, self.p
        dt, Re, We, Ma, betav, A_0, Di, Vh, Bi, T_0, T_h, th, conv = (
            p[k] for k in ("dt", "Re", "We", "Ma", "betav", "a0", "Di", "Vh", "Bi", "T0", "th_hot", "th", "conv"))
        al, k_ewm, B = self.al, self.k_ewm, self.B
        W, Q, Zc = S["W"], S["Q"], S["Zc"]
        u, D_vec = fem.TrialFunction(W, (0, 1)), fem.TrialFunction(W, (2, 3, 4))
        v, R_vec = fem.TestFunction(W, (0, 1)), fem.TestFunction(W, (2, 3, 4))
        D, R = sym(D_vec), sym(R_vec)
        rho = pt = T = fem.TrialFunction(Q)
        q = r = fem.TestFunction(Q)
        tau, Rt = sym(fem.TrialFunction(Zc)), sym(fem.TestFunction(Zc))
        h, n = CellDiameter(), FacetNormal()
        tang = as_vector([n[1], -n[0]])  # :168
        u0, u12, us, u1 = self.u0, self.u12, self.us, self.u1
        rho0, rho12, rho1 = self.rho0.expr(), self.rho12.expr(), self.rho1.expr()
        p0, p1, T0, T1 = self.p0.expr(), self.p1.expr(), self.T0.expr(), self.T1.expr()
        tau0, tau1 = self.tau0, self.tau1
        D0, D1 = sym(self.D0_vec), sym(self.D1_vec)
        DOLFIN_EPS = fem.DOLFIN_EPS
        thetal = (T) / (T_h - T_0)  # :403
        thetar = self.thetar.expr()  # :404-405
        theta0 = (T0 - T_0) / (T_h - T_0)  # :406
        theta1 = T1 - T_0 / (T_h - T_0)  # :407, :658
        DEVSSl_u12 = 2 * (1 - betav) * inner(Dcomp(u), Dincomp(v)) * dx  # :415
        DEVSSr_u12 = 2 * inner(D0, Dincomp(v)) * dx  # :416
        DEVSSl_T1 = (1. - Di) * inner(grad(thetal), grad(r)) * dx  # :420
        DEVSSr_T1 = inner((1. - Di) * (grad(thetar) + grad(theta0)), grad(r)) * dx  # :421
        DEVSSGl_u1 = 2.0 * (1. - betav) * inner(Dincomp(u), Dincomp(v)) * dx  # :427
        # loop top :473-474, :491-492
        alpha_supg = h / (magnitude(u1) + 0.0000000001)
        SU = inner(dot(u1, grad(tau)), alpha_supg * dot(u1, grad(Rt))) * dx
        DEVSSGr_u1 = (1. - betav) * inner(D0 + D0.T, Dincomp(v)) * dx
        U = 0.5 * (u + u0)
        # density half step :497-505
        rho_weak = inner(2.0 * (rho - rho0) / dt + dot(u0, grad(rho0)) - rho0 * div(u0), q) * dx
        a0, L0 = lhs(rho_weak), rhs(rho_weak)
        # velocity half step :517-537
        lhsv_eq = 2.0 * Re * ((rho12 * u - rho0 * u0) / dt + conv * dot(u0, nabla_grad(U)))
        v_weak12 = dot(lhsv_eq, v) * dx \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(div(tau0 - rho0 * Identity(len(u))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a1, L1 = lhs(v_weak12), rhs(v_weak12)
        a1 += th * DEVSSl_u12
        L1 += th * DEVSSr_u12
        # predictor :563-581
        lhsFus = Re * rho0 * ((u - u0) / dt + conv * dot(u0, nabla_grad(U)))
        Fus = dot(lhsFus, v) * dx \
            + inner(2.0 * betav * phi_s(theta0, A_0) * Dincomp(U), Dincomp(v)) * dx \
            + (1.0 / 3) * inner(betav * phi_s(theta0, A_0) * div(U), div(v)) * dx \
            - ((1. - betav) / (We + DOLFIN_EPS)) * inner(div(tau0 - rho0 * Identity(len(u))), v) * dx \
            + inner(grad(p0), v) * dx \
            + inner(D - grad(u), R) * dx
        a2, L2 = lhs(Fus), rhs(Fus)
        a2 += th * DEVSSGl_u1
        L2 += th * DEVSSGr_u1
        # pressure :586-598
        lhs_p_1 = (Ma * Ma / (dt)) * pt
        rhs_p_1 = (Ma * Ma / (dt)) * p0 - Re * (1 + al * theta0) * div(rho0 * us)
        lhs_p_2 = (1 + al * theta0) * dt * grad(pt)
        rhs_p_2 = (1 + al * theta0) * dt * grad(p0)
        a5 = inner(lhs_p_1, q) * dx + inner(lhs_p_2, grad(q)) * dx
        L5 = inner(rhs_p_1, q) * dx + inner(rhs_p_2, grad(q)) * dx
        # density update :603-615
        alpha_supg = h / (magnitude(u12) + self.SU_RHO_EPS)
        SU_rho = inner(dot(u12, grad(rho)), alpha_supg * dot(u12, grad(q))) * dx
        rho_weak = inner((rho - rho0) / dt + dot(u12, grad(rho12)) - rho12 * div(u12), q) * dx
        a6, L6 = lhs(rho_weak), rhs(rho_weak)
        a6 += SU_rho
        # velocity update :617-629 (`a7 += 0`, `L7 += 0`)
        a7 = inner((Re / dt) * rho1 * u, v) * dx + inner(D - grad(u), R) * dx
        L7 = inner((Re / dt) * rho0 * us, v) * dx - 0.5 * inner(grad(p1 - p0), (v)) * dx
        # temperature :641-656
        gamdot = inner(sigmacon(u0, p0, tau0, betav, We), grad(u0))
        lhs_temp1 = (1.0 / dt) * rho1 * thetal + rho1 * dot(u1, grad(thetal))
        difflhs_temp1 = Di * grad(thetal)
        rhs_temp1 = (1.0 / dt) * rho1 * thetar + rho1 * dot(u1, grad(thetar)) + (1.0 / dt) * rho1 * theta0 \
            + Vh * phi_ewm(tau1, theta1, k_ewm, B) * gamdot
        diffrhs_temp1 = Di * grad(thetar)
        a9 = inner(lhs_temp1, r) * dx + inner(difflhs_temp1, grad(r)) * dx
        L9 = inner(rhs_temp1, r) * dx + inner(diffrhs_temp1, grad(r)) * dx - Di * Bi * inner(theta0, r) * ds(1)
        if self.TEMP_DEVSS:
            a9 += th * DEVSSl_T1
            L9 += th * DEVSSr_T1
        else:  # comp_ewm_jpb.py: `a9+= 0.0*th*DEVSSl_T1` -- UFL folds the product to Zero
            a9 += 0.0 * th * DEVSSl_T1
            L9 += 0.0 * th * DEVSSr_T1
        # stress :660-677
        lhs_tau1 = (We * phi_ewm(tau0, theta1, k_ewm, B) / dt + 1.0) * tau \
            + We * phi_ewm(tau0, theta1, k_ewm, B) * FdefG(u1, D1, tau)
        rhs_tau1 = (We * phi_ewm(tau0, theta1, k_ewm, B) / dt) * tau0 + rho0 * Identity(len(u))
        Ftau = inner(lhs_tau1, Rt) * dx - inner(rhs_tau1, Rt) * dx
        a4, L4 = lhs(Ftau), rhs(Ftau)
        a4 = a4 + SU
        self.forms = dict(a0=a0, L0=L0, a1=a1, L1=L1, a2=a2, L2=L2, a5=a5, L5=L5, a6=a6, L6=L6, a7=a7, L7=L7,
                          a4=a4, L4=L4, a9=a9, L9=L9)
        # functionals :684-707
        self.E_k_form = 0.5 * rho1 * dot(u1, u1) * dx
        tau1_vec = self.tau1_vec.expr()
        self.E_e_form = (tau1_vec[0] + tau1_vec[2] - 2.0) * dx
        sig = sigmacon(u1, p1, tau1, betav, We)
        sigma0, omegaf0 = dot(sig, tang), dot(sig, n)
        self.force_x_form = inner(Const(np.array([1.0, 0.0])), omegaf0) * ds(0)
        self.force_y_form = -inner(Const(np.array([0.0, 1.0])), omegaf0) * ds(0)
        self.torque_form = -inner(n, sigma0) * ds(0)

    def step(self):
        f, last = self.forms, self.last
        self.iter += 1
        self._tt["t"] = self.t  # w.t = t :468
        if self.iter > 1:  # :477-482
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
        system("a1", "L1", self.w12, self.bcu)
        system("a2", "L2", self.ws, self.bcu)
        system("a5", "L5", self.p1, [])
        system("a6", "L6", self.rho1, [])
        system("a7", "L7", self.w1, self.bcu)
        system("a9", "L9", self.T1, self.bcT)
        if self.p["We"] > 0.00001 or self.prepass or self.ALWAYS_STRESS:  # :660
            system("a4", "L4", self.tau1_vec, [])
        out = dict(t=self.t, E_k=fem.assemble(self.E_k_form), E_e=fem.assemble(self.E_e_form),
                   torque=fem.assemble(self.torque_form), F_x=fem.assemble(self.force_x_form),
                   F_y=fem.assemble(self.force_y_form))
        self.t += self.p["dt"]
        return out


class CompEwmJBP(EwmJBP):
    """"flow between eccentrically rotating cylinders"/comp_ewm_jpb.py time loop :421-732 -- the TRUE extended
    White-Metzner model: A_0 = 120, k_ewm = -0.7, B = 0.1 (:88-90), so phi_s(theta0) and phi_ewm(tau, theta1) are
    genuine coefficient functions.  Statement for statement the loop of incompressible_benchmarkEWM.py except:
    the SU weight of the density update uses DOLFIN_EPS (`h/(magnitude(u12)+DOLFIN_EPS)`), the temperature DEVSS
    terms are multiplied by the literal 0.0 (`a9+= 0.0*th*DEVSSl_T1`), and the stress solve is unconditional."""
    SU_RHO_EPS = fem.DOLFIN_EPS
    TEMP_DEVSS = False
    ALWAYS_STRESS = True

    def __init__(self, mesh, dofmaps=None, dt=None, Re=50.0, We=0.5, Ma=0.01, betav=0.5, A_0=120.0, k_ewm=-0.7, B=0.1,
                 x_1=-0.80, **kw):
        super().__init__(mesh, dofmaps, dt=dt, prepass=False, Re=Re, We=We, Ma=Ma, betav=betav, A_0=A_0, k_ewm=k_ewm,
                         B=B, x_1=x_1, **kw)
