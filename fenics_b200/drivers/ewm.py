"""Journal-bearing time loop on the GPU:

  EwmJBP      config 3  "flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py:465-782
              (extended White-Metzner at the script's A_0 = k_ewm = B = 0, so phi_s = phi_ewm = 1)
  CompEwmJBP  the same directory's comp_ewm_jpb.py:421-732: the true EWM model (A_0 = 120, k_ewm = -0.7, B = 0.1),
              non-constant phi_s / phi_ewm through the FX_FORM_C3E_* functors

Same call sequence as the script's loop body: start-of-step state roll (`iter > 1`), density half
step, u^{n+1/2}, u*, pressure, SU density update, u^{n+1}, temperature, the stress update when
`We > 1e-5` or in the pre-pass (:660), then the energies and the journal torque / load on ds(0).

The script runs the loop in two modes (:291-296, :459-462): the 15-step pre-pass before mesh adaption
(un-ramped spin, dt = hmin^2, stress solved) and the main pass (tanh-ramped spin, dt = 10 hmin^2).
`prepass` selects between them.  The mesh adaption between the passes is outside the hot path.

`solve(A7, w1.vector(), b7)` (:627) is a direct LU solve in the reference; here it is BiCGStab+Jacobi
driven to `direct_rtol` (ILU/LU factorisations are not part of this library).
"""
from .. import _abi as abi
from ..dolfin_api import assemble_scalar_async
from ..mesh import annulus_mesh
from .jbp import FenepmpJBP

DOLFIN_EPS = 3.0e-16


class EwmJBP(FenepmpJBP):
    direct_rtol = 1e-12
    REFINE_RATIO, REFINE_EPS, MAX_REFINE = 0.2, DOLFIN_EPS, 2   # incompressible_benchmarkEWM.py:912,915; max_refine :52
    SU_RHO_EPS = 0.0000000001   # alpha_supg = h/(magnitude(u12)+0.0000000001)  :603
    TEMP_DEVSS = True           # a9 += th*DEVSSl_T1  :650
    ALWAYS_STRESS = False       # stress solve only `if We > 0.00001 or adaptive_loop == False ...`  :660

    def __init__(self, mesh=None, dofmaps=None, dt=None, prepass=False, Re=10.0, We=0.000001, Ma=0.000001,
                 betav=1.0 - DOLFIN_EPS, A_0=0.0, k_ewm=0.0, B=0.0, al=0.001, Di=0.0025, Vh=0.0000069, Bi=0.2,
                 T_0=300.0, T_h=350.0, th=1.0, conv=1.0, x_1=-0.90, y_1=0.0, x_2=0.0, y_2=0.0, r_a=1.0, r_b=2.0,
                 n_theta=96, n_r=24, device=None):
        # non-zero A_0 / k_ewm / B: phi_s and phi_ewm are coefficient functions -> the C3E functors (UFL folds
        # them to the literal 1 only when the parameters are the literal 0.0)
        self.ewm = A_0 != 0.0 or k_ewm != 0.0 or B != 0.0
        if mesh is None:
            mesh = annulus_mesh(n_theta, n_r, x_1, y_1, r_a, x_2, y_2, r_b)
        if dt is None:
            dt = (1.0 if prepass else 10.0) * mesh.hmin() ** 2                      # :164, :461-462
        self.prepass, self.ramp = prepass, not prepass   # un-ramped spin in the pre-pass  (:291-295)
        super().__init__(mesh, dofmaps, dt=dt, Re=Re, We=We, Ma=Ma, lambda_d=0.0, betav=betav, A_0=A_0, Di=Di, Vh=Vh,
                         Bi=Bi, T_0=T_0, T_h=T_h, th=th, conv=conv, x_1=x_1, y_1=y_1, x_2=x_2, y_2=y_2, r_a=r_a,
                         r_b=r_b, device=device)
        self.p["al1"] = al
        self.p.update(k_ewm=k_ewm, b_ewm=B, th_t=(th if self.TEMP_DEVSS else 0.0))
        for slot, v in ((abi.FX_P_AL1, al), (abi.FX_P_K_EWM, k_ewm), (abi.FX_P_B_EWM, B), (abi.FX_P_TH_T, self.p["th_t"])):
            self.pr.set_param(slot, v)

    def lps_indicator(self):
        raise NotImplementedError("config 3 has no LPS indicator")

    def step(self):
        self._begin_step()
        self.iter += 1
        self._bct["t"] = self.t                         # w.t = t  :468
        if self.iter > 1:                               # :477-482
            self.w0.assign(self.w1)
            self.T0.assign(self.T1)
            self.rho0.assign(self.rho1)
            self.p0.assign(self.p1)
            self.tau0_vec.assign(self.tau1_vec)
        A = abi
        e = self.ewm
        a1, L1, L2 = (A.FX_FORM_C3E_A1, A.FX_FORM_C3E_L1, A.FX_FORM_C3E_L2) if e else (A.FX_FORM_C3_A1, A.FX_FORM_C3_L1, A.FX_FORM_C3_L2)
        a4, L4, L9 = (A.FX_FORM_C3E_A4, A.FX_FORM_C3E_L4, A.FX_FORM_C3E_L9) if e else (A.FX_FORM_C3_A4, A.FX_FORM_C3_L4, A.FX_FORM_C3_L9)
        self._system("rho12", A.FX_FORM_C3_A0, A.FX_FORM_C3_L0, self.AQ, self.bQ, self.rho12, ())      # :497-511
        self._system("u12", a1, L1, self.AW, self.bW, self.w12, self.bcu)                               # :517-537
        self._system("us", A.FX_FORM_C3_A2, L2, self.AW, self.bW, self.ws, self.bcu)                    # :563-581
        self._system("p", A.FX_FORM_C3_A5, A.FX_FORM_C3_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method, pc=self._pressure_pc())                                                       # :586-598
        self.pr.set_param(A.FX_P_EPS, self.SU_RHO_EPS)   # the density update's SU weight has its own epsilon
        self._system("rho1", A.FX_FORM_C3_A6, A.FX_FORM_C3_L6, self.AQ, self.bQ, self.rho1, ())        # :603-615
        self.pr.set_param(A.FX_P_EPS, self.SU_EPS)
        self._system("u1", A.FX_FORM_C3_A7, A.FX_FORM_C3_L7, self.AW, self.bW, self.w1, self.bcu,
                     rtol=self.direct_rtol if self.solve_rtol is None else min(self.direct_rtol, self.solve_rtol))                                       # :617-629
        self._system("T", A.FX_FORM_C3_A9, L9, self.AQ, self.bQ, self.T1, self.bcT)                     # :641-656
        if self.p["We"] > 0.00001 or self.prepass or self.ALWAYS_STRESS:                                # :660
            self._system("tau", a4, L4, self.AZ, self.bZ, self.tau1_vec, ())                            # :663-677
        f = self.pr.form
        with self._phase("functionals"):
            for i, fid in enumerate((A.FX_FORM_C3_EK, A.FX_FORM_C3_EE, A.FX_FORM_C3_TORQUE, A.FX_FORM_C3_FORCE_X,
                                     A.FX_FORM_C3_FORCE_Y)):                                            # :684-707
                assemble_scalar_async(f(fid), self.scal, i)
        vals = self._finish_step()
        t_rec = self.t
        self.t += self.p["dt"]                          # :782
        return self._result(t_rec, vals, dict(E_k=0, E_e=1, torque=2, F_x=3, F_y=4))


class CompEwmJBP(EwmJBP):
    """comp_ewm_jpb.py:421-732 -- the true extended White-Metzner loop (oracle: configs_ewm.CompEwmJBP lists the
    three statements that differ from incompressible_benchmarkEWM.py)."""
    SU_RHO_EPS = DOLFIN_EPS     # alpha_supg = h/(magnitude(u12)+DOLFIN_EPS)
    TEMP_DEVSS = False          # a9 += 0.0*th*DEVSSl_T1
    ALWAYS_STRESS = True

    def __init__(self, mesh=None, dofmaps=None, dt=None, Re=50.0, We=0.5, Ma=0.01, betav=0.5, A_0=120.0, k_ewm=-0.7,
                 B=0.1, x_1=-0.80, **kw):
        super().__init__(mesh, dofmaps, dt=dt, prepass=False, Re=Re, We=We, Ma=Ma, betav=betav, A_0=A_0, k_ewm=k_ewm,
                         B=B, x_1=x_1, **kw)
