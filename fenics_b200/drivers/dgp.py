"""Buoyancy-driven compressible viscoelastic flow on the GPU:

  CompDGP  config 5  buoyancy-driven-flow/compressible_dgp.py:368-595

Same call sequence as the script's loop body.  One documented deviation: the script increments `iter`
twice per pass (:369,:377), so its start-of-step roll (:396-401) runs in the very first step and replaces
the initial rho0, T0, tau0 by the never-initialised (zero) rho1, T1, tau1 -- T0 = 0 gives therm = 0 and
therm_inv = project(1/0) = inf.  With `faithful_first_roll=False` (default) rho1, T1, tau1 start from the
initial state, which makes the first roll the no-op it is evidently meant to be; `True` reproduces the
script literally (and fails in the first solve, like the reference).
"""
import numpy as np

from .. import _abi as abi
from ..dolfin_api import ConsistentProjection, DirichletBC, Function, assemble_scalar_async, project
from ..mesh import DOLFIN_EPS, RectangleMesh, _mapped, xskewcavity
from .ldc import _LDC


def DGP_Mesh(mm, B, L):
    """DGP_Mesh (buoyancy-driven-flow/fenics_fem.py:261-311): x -> (1-cos(pi x))/2."""
    return _mapped(RectangleMesh((0, 0), (B, L), mm * B, mm * L), xskewcavity)


class CompDGP(_LDC):
    STATE_FIELDS = ("w0", "p0", "tau0_vec", "rho0", "T0", "w1", "tau1_vec")
    RESULT_FIELDS = ("w1", "p1", "tau1_vec", "rho1", "T1")

    def __init__(self, mesh=None, dofmaps=None, dt=0.0005, Ra=100.0, We=0.1, Ma=0.05, Pr=2.0, betav=0.1, c1=0.05,
                 c2=0.001, th=0.5, Vh=0.005, Bi=0.2, T_0=300.0, T_h=350.0, al_1=1.0, al_2=1.0 / 6.0,
                 mesh_resolution=50, faithful_first_roll=False, device=None):
        if mesh is None:
            mesh = DGP_Mesh(mesh_resolution, 1, 1)
        left = lambda x, on_b: (x[0] < 0.0 + DOLFIN_EPS) & on_b  # noqa: E731
        right = lambda x, on_b: (x[0] > 1.0 - DOLFIN_EPS) & on_b  # noqa: E731
        top = lambda x, on_b: (x[1] > 1.0 - DOLFIN_EPS) & on_b  # noqa: E731
        bottom = lambda x, on_b: (x[1] < 0.0 + DOLFIN_EPS) & on_b  # noqa: E731
        everything = lambda x, on_b: np.ones(np.shape(x[0]), dtype=bool) & on_b  # noqa: E731
        mesh.mark_exterior_facets([(0, everything), (1, left), (2, right), (3, top), (4, bottom)])  # :163-169
        self.p = dict(dt=dt, Ra=Ra, We=We, Ma=Ma, Pr=Pr, betav=betav, c1=c1, c2=c2, th=th, Vh=Vh, Bi=Bi, T0=T_0,
                      th_hot=T_h, al1=al_1, al2=al_2, eps=DOLFIN_EPS)
        self._setup(mesh, dofmaps, device)
        b_, F = self.pr.bind, lambda s: Function(self.S[s], self.plan)  # noqa: E731
        A = abi
        self.rho0, self.rho12, self.rho1 = (b_(s, F("Q")) for s in (A.FX_F_RHO0, A.FX_F_RHO12, A.FX_F_RHO1))
        self.T0, self.T1 = b_(A.FX_F_T0, F("Q")), b_(A.FX_F_T1, F("Q"))
        self.theta0, self.therm_inv = b_(A.FX_F_THETA0, F("Q")), b_(A.FX_F_THERM_INV, F("Q"))
        self.theta1 = b_(A.FX_F_Q_A, F("Q"))
        f = self.pr.form
        mq = f(A.FX_FORM_MASS_Q)
        self.proj_res_orth = ConsistentProjection(f(A.FX_FORM_MASS_ZC), f(A.FX_FORM_C5_L_RES_ORTH))
        self.proj_theta0 = ConsistentProjection(mq, f(A.FX_FORM_C5_L_THETA0))
        self.proj_theta1 = ConsistentProjection(mq, f(A.FX_FORM_C5_L_THETA1))
        self.proj_therm_inv = ConsistentProjection(mq, f(A.FX_FORM_C5_L_THERM_INV))
        self.proj_rho1 = ConsistentProjection(mq, f(A.FX_FORM_C5_L_RHO1))
        W, Q = self.S["W"], self.S["Q"]
        tt = self._bct = {"t": 0.0}
        ramp = lambda x: np.full(len(x), 0.5 * (1 + np.tanh(4 * (tt["t"] - 1.5))) * (T_h - T_0) + T_0)  # noqa: E731  :174
        self.bcu = [DirichletBC(W.sub(0), (0.0, 0.0), sd) for sd in (everything, left, right, top)]  # :185-193
        self.bcT = [DirichletBC(Q, ramp, left), DirichletBC(Q, T_0, right)]                            # :189-195
        # initial state :269-291 (projections of constants are exact)
        tv = np.zeros(self.S["Zc"].dim())
        comp_of = np.zeros(self.S["Zc"].dim(), dtype=np.int64)
        for c in range(3):
            comp_of[self.S["Zc"].cell_dofs[:, 6 * c:6 * c + 6].ravel()] = c
        tv[comp_of != 1] = 1.0
        self.tau0_vec.vector().set_local(tv)
        self.rho0.vector().t.fill_(1.0)
        self.T0.vector().t.fill_(T_0)
        if not faithful_first_roll:
            self.tau1_vec.assign(self.tau0_vec)
            self.rho1.assign(self.rho0)
            self.T1.assign(self.T0)
        self.t = dt  # :365

    def lps_indicator(self):
        e = self.pr.expr
        with self._phase("project_dg0:restau"):
            project(e(abi.FX_EXPR_C5_RESTAU), self.S["Zd"], function=self.res_test)
        self._project_consistent(self.proj_res_orth, self.res_orth, self.AZ, self.bZ, "res_orth")
        with self._phase("project_dg0:kappa"):
            project(e(abi.FX_EXPR_RES_ORTH_SQ), self.S["Qt"], function=self.res_orth_norm_sq)
            project(e(abi.FX_EXPR_SQRT_QT_A), self.S["Qt"], function=self.kapp)

    def step(self):
        self._begin_step()
        A = abi
        self._bct["t"] = self.t                         # ramp_function.t = t :379
        self.lps_indicator()                            # :382-391
        self.w0.assign(self.w1)                         # :396-401 (`iter > 1` holds from the first pass)
        self.T0.assign(self.T1)
        self.rho0.assign(self.rho1)
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        self._project_consistent(self.proj_theta0, self.theta0, self.AQ, self.bQ, "theta0")        # :405
        self._project_consistent(self.proj_therm_inv, self.therm_inv, self.AQ, self.bQ, "therm_inv")  # :408
        self._system("rho12", A.FX_FORM_C5_A0, A.FX_FORM_C5_L0, self.AQ, self.bQ, self.rho12, ())     # :420-428
        self._system("u12", A.FX_FORM_C5_A1, A.FX_FORM_C5_L1, self.AW, self.bW, self.w12, self.bcu)   # :432-447
        self._system("us", A.FX_FORM_C5_A2, A.FX_FORM_C5_L2, self.AW, self.bW, self.ws, self.bcu)     # :473-489
        self._system("p", A.FX_FORM_C5_A5, A.FX_FORM_C5_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method, pc=self._pressure_pc())                                                      # :493-506
        self._system("rho1su", A.FX_FORM_C5_A6, A.FX_FORM_C5_L6, self.AQ, self.bQ, self.rho1, ())     # :510-525 (dead)
        self._project_consistent(self.proj_rho1, self.rho1, self.AQ, self.bQ, "rho1")                 # :529-530
        self._system("u1", A.FX_FORM_C5_A7, A.FX_FORM_C5_L7, self.AW, self.bW, self.w1, self.bcu)     # :533-543
        self._system("tau", A.FX_FORM_C5_A4, A.FX_FORM_C5_L4, self.AZ, self.bZ, self.tau1_vec, ())    # :550-559
        self._system("T", A.FX_FORM_C5_A8, A.FX_FORM_C5_L8, self.AQ, self.bQ, self.T1, self.bcT)      # :563-572
        f = self.pr.form
        with self._phase("functionals"):
            assemble_scalar_async(f(A.FX_FORM_C5_EK), self.scal, 0)                                   # :577
            assemble_scalar_async(f(A.FX_FORM_C5_EE), self.scal, 1)                                   # :578
        self._project_consistent(self.proj_theta1, self.theta1, self.AQ, self.bQ, "theta1")           # :581
        assemble_scalar_async(f(A.FX_FORM_C5_NUS), self.scal, 2)                                      # :582-583
        vals = self._finish_step()
        t_rec = self.t
        self.t += self.p["dt"]                          # :591
        return self._result(t_rec, vals, dict(E_k=0, E_e=1, Nus=2))

    def fields(self):
        g = lambda fn: fn.vector().get_local()  # noqa: E731
        return dict(w1=g(self.w1), p1=g(self.p1), tau1=g(self.tau1_vec), rho1=g(self.rho1), T1=g(self.T1),
                    kapp=g(self.kapp))


class IncompDGP(_LDC):
    """buoyancy-driven-flow/incompressible_dgp.py:357-535 -- the incompressible double-glazing problem.

    Same call sequence as the script's loop body: LPS indicator, the start-of-step roll from the second pass on
    (`iter > 1`, :384-390), u^{n+1/2}, u*, the pure-Neumann pressure Poisson problem (bcp = []), u^{n+1}, the stress
    update when betav < 0.99, the temperature update, energies and the Nusselt number on ds(2) (the left wall).
    `solve(A7, ...)` and `solve(A8, ...)` are direct (LU) solves in the reference: BiCGStab driven to `direct_rtol`.
    The pressure matrix is singular (constants): the Jacobi-preconditioned Krylov kernels keep sum_i A_ii (p - p_start)_i
    = 0, which is the member of the solution family the oracle computes too -- it matters because the temperature
    source Vh*inner(sigma(u1, p1, tau1), grad(u1)) (:503) is not invariant under p -> p + c when div(u1) != 0."""
    STATE_FIELDS = ("w0", "p0", "tau0_vec", "T0", "w1", "tau1_vec", "T1", "p1")
    RESULT_FIELDS = ("w1", "p1", "tau1_vec", "T1")
    direct_rtol = 1e-12

    def __init__(self, mesh=None, dofmaps=None, dt=0.001, Ra=1000.0, We=0.1, Pr=1.0, betav=0.5, c1=0.05, c2=0.001, th=0.5,
                 Vh=0.005, Bi=0.0 + DOLFIN_EPS, T_0=300.0, T_h=350.0, mesh_resolution=40, device=None):
        if mesh is None:
            mesh = DGP_Mesh(mesh_resolution, 1, 1)
        left = lambda x, on_b: (x[0] < 0.0 + DOLFIN_EPS) & on_b  # noqa: E731
        right = lambda x, on_b: (x[0] > 1.0 - DOLFIN_EPS) & on_b  # noqa: E731
        top = lambda x, on_b: (x[1] > 1.0 - DOLFIN_EPS) & on_b  # noqa: E731
        bottom = lambda x, on_b: (x[1] < 0.0 + DOLFIN_EPS) & on_b  # noqa: E731
        everything = lambda x, on_b: np.ones(np.shape(x[0]), dtype=bool) & on_b  # noqa: E731
        mesh.mark_exterior_facets([(0, everything), (1, bottom), (2, left), (3, right), (4, top)])  # :147-151
        self.p = dict(dt=dt, Ra=Ra, We=We, Pr=Pr, betav=betav, c1=c1, c2=c2, th=th, Vh=Vh, Bi=Bi, T0=T_0, th_hot=T_h,
                      re=1.0)  # re: the shared L5 functor carries Re/dt, this loop 1/dt
        self._setup(mesh, dofmaps, device)
        b_, F = self.pr.bind, lambda s: Function(self.S[s], self.plan)  # noqa: E731
        A = abi
        self.T0, self.T1 = b_(A.FX_F_T0, F("Q")), b_(A.FX_F_T1, F("Q"))
        self.theta1 = b_(A.FX_F_Q_A, F("Q"))
        f = self.pr.form
        self.proj_res_orth = ConsistentProjection(f(A.FX_FORM_MASS_ZC), f(A.FX_FORM_C6_L_RES_ORTH))
        self.proj_theta1 = ConsistentProjection(f(A.FX_FORM_MASS_Q), f(A.FX_FORM_C5_L_THETA1))
        W, Q = self.S["W"], self.S["Q"]
        tt = self._bct = {"t": 0.0}
        ramp = lambda x: np.full(len(x), 0.5 * (1 + np.tanh(4 * (tt["t"] - 1.5))) * (T_h - T_0) + T_0)  # noqa: E731  :156
        self.bcu = [DirichletBC(W.sub(0), (0.0, 0.0), sd) for sd in (everything, left, right, top)]  # :166-173
        self.bcT = [DirichletBC(Q, ramp, left), DirichletBC(Q, T_0, right)]                            # :170-175
        tv = np.zeros(self.S["Zc"].dim())                                                              # :262-267
        comp_of = np.zeros(self.S["Zc"].dim(), dtype=np.int64)
        for c in range(3):
            comp_of[self.S["Zc"].cell_dofs[:, 6 * c:6 * c + 6].ravel()] = c
        tv[comp_of != 1] = 1.0
        self.tau0_vec.vector().set_local(tv)
        self.T0.vector().t.fill_(T_0)
        self.t, self.iter = 0.0, 0                                                                     # :350, :355

    def lps_indicator(self):
        e = self.pr.expr
        with self._phase("project_dg0:restau"):
            project(e(abi.FX_EXPR_C6_RESTAU), self.S["Zd"], function=self.res_test)                  # :373
        self._project_consistent(self.proj_res_orth, self.res_orth, self.AZ, self.bZ, "res_orth")    # :374
        with self._phase("project_dg0:kappa"):
            project(e(abi.FX_EXPR_RES_ORTH_SQ), self.S["Qt"], function=self.res_orth_norm_sq)        # :375
            project(e(abi.FX_EXPR_SQRT_QT_A), self.S["Qt"], function=self.kapp)                      # :376-378 (|.| of a sqrt)

    def step(self):
        self._begin_step()
        A = abi
        self.iter += 1                                  # :358
        self._bct["t"] = self.t                         # ramp_function.t = t :367
        self.lps_indicator()                            # :369-379
        if self.iter > 1:                               # :384-390
            self.w0.assign(self.w1)
            self.T0.assign(self.T1)
            self.p0.assign(self.p1)
            self.tau0_vec.assign(self.tau1_vec)
        tight = self.direct_rtol if self.solve_rtol is None else min(self.direct_rtol, self.solve_rtol)
        self._system("u12", A.FX_FORM_C6_A1, A.FX_FORM_C6_L1, self.AW, self.bW, self.w12, self.bcu)   # :397-414
        self._system("us", A.FX_FORM_C6_A2, A.FX_FORM_C6_L2, self.AW, self.bW, self.ws, self.bcu)     # :439-451
        self._system("p", A.FX_FORM_C6_A5, A.FX_FORM_C6_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method)                                                      # :454-460
        self._system("u1", A.FX_FORM_C6_A7, A.FX_FORM_C6_L7, self.AW, self.bW, self.w1, self.bcu, rtol=tight)  # :463-481
        if self.p["betav"] < 0.99:                                                                     # :485
            self._system("tau", A.FX_FORM_C6_A4, A.FX_FORM_C6_L4, self.AZ, self.bZ, self.tau1_vec, ())  # :486-499
        self._system("T", A.FX_FORM_C6_A8, A.FX_FORM_C6_L8, self.AQ, self.bQ, self.T1, self.bcT, rtol=tight)  # :503-513
        f = self.pr.form
        with self._phase("functionals"):
            assemble_scalar_async(f(A.FX_FORM_C6_EK), self.scal, 0)                                   # :517
            assemble_scalar_async(f(A.FX_FORM_C6_EE), self.scal, 1)                                   # :518
        self._project_consistent(self.proj_theta1, self.theta1, self.AQ, self.bQ, "theta1")           # :521
        assemble_scalar_async(f(A.FX_FORM_C6_NUS), self.scal, 2)                                      # :522-523
        vals = self._finish_step()
        t_rec = self.t
        self.t += self.p["dt"]                          # :532
        return self._result(t_rec, vals, dict(E_k=0, E_e=1, Nus=2))

    def fields(self):
        g = lambda fn: fn.vector().get_local()  # noqa: E731
        p = g(self.p1)
        return dict(w1=g(self.w1), p1=p - p.mean(), tau1=g(self.tau1_vec), T1=g(self.T1), kapp=g(self.kapp))
