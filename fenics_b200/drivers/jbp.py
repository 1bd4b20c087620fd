"""Journal-bearing time loop on the GPU:

  FenepmpJBP  config 4  "flow between eccentrically rotating cylinders"/fenepmp_jbp.py:454-690
              (FENE-P-MP, lambda_D = 0 as in parameters-fenep-mp.csv row 4 / fenepmp_jbp.py:297)

Same call sequence as the script's loop body: LPS indicator, start-of-step state roll (`iter > 1`),
density half step, u*, pressure, SU density update, u^{n+1} (DEVSS-G), FENE stress via `solvertau`,
temperature, then the energies and the journal torque / load functionals on ds(0).
"""
import csv

import numpy as np
import torch

from .. import _abi as abi
from .. import la
from .._lib import check, lib
from ..dolfin_api import ConsistentProjection, DirichletBC, Function, assemble_scalar_async, project
from ..mesh import annulus_mesh
from .ldc import _LDC

# "flow between eccentrically rotating cylinders"/parameters-fenep-mp.csv (rows Re, We, Ma, Ld)
CSV_DEFAULT = dict(Re=[50.0, 25.0], We=[0.001, 0.1, 0.5, 1.0], Ma=[0.01, 0.05, 0.1], Ld=[0.0, 0.05, 0.1, 0.15])


def read_parameters(input_csv=None):
    """rows of the parameter CSV as the script indexes them (fenepmp_jbp.py:267-274)."""
    if input_csv is None:
        return CSV_DEFAULT
    with open(input_csv, "r") as f:
        rows = list(csv.reader(f))
    return {r[0]: [float(x) for x in r[1:]] for r in rows}


def sweep(params=None, all_lambda=False):
    """(j, jjj) -> parameters of the script's primary (We) and secondary (Ma) loops
    (fenepmp_jbp.py:153-154, :291-297, :1080-1094): Re = re_row[2], We = we_row[j], Ma = ma_row[jjj+1],
    lambda_D = lambda_row[1].  all_lambda: the sweep once per value of the CSV's `Ld` row (0, 0.05, 0.1, 0.15 -- the
    runs behind plot_data.py:38-47) instead of lambda_row[1] only."""
    p = params or CSV_DEFAULT
    out = []
    for lam in (p["Ld"] if all_lambda else p["Ld"][:1]):
        for jjj in range(len(p["Ma"])):
            for j in range(1, len(p["We"]) + 1):
                out.append(dict(j=j, jjj=jjj, Re=p["Re"][1], We=p["We"][j - 1], Ma=p["Ma"][jjj], lambda_d=lam))
    return out


class FenepmpJBP(_LDC):
    STATE_FIELDS = ("w0", "p0", "tau0_vec", "rho0", "T0", "w1", "tau1_vec")
    RESULT_FIELDS = ("w1", "p1", "tau1_vec", "rho1", "T1")
    SU_EPS = 0.0000000001  # alpha_supg = h/(magnitude(u12)+0.0000000001)  :553

    def __init__(self, mesh=None, dofmaps=None, dt=0.00125, Re=25.0, We=0.1, Ma=0.01, lambda_d=0.0, betav=0.5,
                 b=20.0, A_0=1.0, Di=0.005, Vh=0.0000069, Bi=0.2, T_0=300.0, T_h=350.0, th=1.0, conv=1.0,
                 x_1=-0.80, y_1=0.0, x_2=0.0, y_2=0.0, r_a=1.0, r_b=2.0, n_theta=96, n_r=24, device=None):
        # lambda_D != 0 (FENE-P-MP proper, CSV row `Ld`): the seven forms / one projection that carry phi_def(u) have
        # FX_FORM_C4M_* functors (UFL degrees 13 / 13 / 14, ds 12); lambda_D = 0 folds phi_def to 1 as UFL does
        mp = lambda_d != 0.0
        A = abi
        self.F_A4, self.F_L2, self.F_L9 = ((A.FX_FORM_C4M_A4, A.FX_FORM_C4M_L2, A.FX_FORM_C4M_L9) if mp else
                                           (A.FX_FORM_C4_A4, A.FX_FORM_C4_L2, A.FX_FORM_C4_L9))
        self.F_RES_ORTH = A.FX_FORM_C4M_L_RES_ORTH if mp else A.FX_FORM_C4_L_RES_ORTH
        self.E_RESTAU = A.FX_EXPR_C4M_RESTAU if mp else A.FX_EXPR_C4_RESTAU
        self.F_JOURNAL = ((A.FX_FORM_C4M_TORQUE, A.FX_FORM_C4M_FORCE_X, A.FX_FORM_C4M_FORCE_Y) if mp else
                          (A.FX_FORM_C4_TORQUE, A.FX_FORM_C4_FORCE_X, A.FX_FORM_C4_FORCE_Y))
        if mesh is None:
            mesh = annulus_mesh(n_theta, n_r, x_1, y_1, r_a, x_2, y_2, r_b)
        # boundary_parts: journal omega0 -> 0, bearing omega1 -> 1  (:212-237)
        in0 = lambda x, on_b: ((x[0] - x_1) ** 2 + (x[1] - y_1) ** 2 < (0.9 * r_a ** 2 + 0.1 * r_b ** 2)) & on_b  # noqa: E731
        in1 = lambda x, on_b: ((x[0] - x_2) ** 2 + (x[1] - y_2) ** 2 > (0.1 * r_a ** 2 + 0.9 * r_b ** 2)) & on_b  # noqa: E731
        mesh.mark_exterior_facets([(0, in0), (1, in1)])
        self.p = dict(dt=dt, Re=Re, We=We, Ma=Ma, lambda_d=lambda_d, betav=betav, b_fene=b, a0=A_0, Di=Di, Vh=Vh,
                      Bi=Bi, T0=T_0, th_hot=T_h, th=th, conv=conv, eps=self.SU_EPS)
        self._setup(mesh, dofmaps, device)
        b_, F = self.pr.bind, lambda s: Function(self.S[s], self.plan)  # noqa: E731
        self.rho0, self.rho12, self.rho1 = (b_(s, F("Q")) for s in (abi.FX_F_RHO0, abi.FX_F_RHO12, abi.FX_F_RHO1))
        self.T0, self.T1 = b_(abi.FX_F_T0, F("Q")), b_(abi.FX_F_T1, F("Q"))
        f = self.pr.form
        self.proj_res_orth = ConsistentProjection(f(abi.FX_FORM_MASS_ZC), f(self.F_RES_ORTH))
        W, Q = self.S["W"], self.S["Q"]
        tt = self._bct = {"t": 0.0}

        def wspin(x):  # Expression w, ramped spin of the journal  (:249)
            g = 0.5 * (1.0 + np.tanh(8 * (tt["t"] - 0.5))) if getattr(self, "ramp", True) else 1.0
            return np.column_stack([g * (x[:, 1] - y_1) / r_a, -g * (x[:, 0] - x_1) / r_a])

        spin = DirichletBC(W.sub(0), wspin, in0)
        noslip = DirichletBC(W.sub(0), (0.0, 0.0), in1)
        self.bcu = [noslip, spin]                       # :255
        self.bcT = [DirichletBC(Q, T_h, in0)]           # temp0 :253,257
        # initial state :367-382 (project of constants: exact)
        self.rho0.vector().t.fill_(1.0)
        tv = np.zeros(self.S["Zc"].dim())
        comp_of = np.zeros(self.S["Zc"].dim(), dtype=np.int64)
        for c in range(3):
            comp_of[self.S["Zc"].cell_dofs[:, 6 * c:6 * c + 6].ravel()] = c
        tv[comp_of != 1] = 1.0                          # I_vec = (1, 0, 1)
        self.tau0_vec.vector().set_local(tv)
        self.T0.vector().t.fill_(T_0)
        self.t, self.iter = 0.0, 0                      # :448-449

    def lps_indicator(self):
        """LPS indicator of the FENE residual (fenepmp_jbp.py:462-473); |kapp| = kapp (sqrt >= 0)."""
        e = self.pr.expr
        with self._phase("project_dg0:restau"):
            project(e(self.E_RESTAU), self.S["Zd"], function=self.res_test)
        self._project_consistent(self.proj_res_orth, self.res_orth, self.AZ, self.bZ, "res_orth")
        with self._phase("project_dg0:kappa"):
            project(e(abi.FX_EXPR_RES_ORTH_SQ), self.S["Qt"], function=self.res_orth_norm_sq)
            project(e(abi.FX_EXPR_SQRT_QT_A), self.S["Qt"], function=self.kapp)

    def step(self):
        self._begin_step()
        self.iter += 1
        self._bct["t"] = self.t                         # w.t = t  :458
        self.lps_indicator()                            # :460-473 (mesh_refinement == False)
        if self.iter > 1:                               # :481-486
            self.w0.assign(self.w1)
            self.T0.assign(self.T1)
            self.rho0.assign(self.rho1)
            self.p0.assign(self.p1)
            self.tau0_vec.assign(self.tau1_vec)
        A = abi
        self._system("rho12", A.FX_FORM_C4_A0, A.FX_FORM_C4_L0, self.AQ, self.bQ, self.rho12, ())      # :501-510
        self._system("us", A.FX_FORM_C4_A2, self.F_L2, self.AW, self.bW, self.ws, self.bcu)      # :513-531
        self._system("p", A.FX_FORM_C4_A5, A.FX_FORM_C4_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method, pc=self._pressure_pc())                                                       # :536-549
        self._system("rho1", A.FX_FORM_C4_A6, A.FX_FORM_C4_L6, self.AQ, self.bQ, self.rho1, ())        # :553-570
        self._system("u1", A.FX_FORM_C4_A7, A.FX_FORM_C4_L7, self.AW, self.bW, self.w1, self.bcu)      # :573-586
        self._system("tau", self.F_A4, A.FX_FORM_C4_L4, self.AZ, self.bZ, self.tau1_vec, ())     # :595-611
        self._system("T", A.FX_FORM_C4_A9, self.F_L9, self.AQ, self.bQ, self.T1, self.bcT)       # :615-629
        f = self.pr.form
        with self._phase("functionals"):
            for i, fid in enumerate((A.FX_FORM_C4_EK, A.FX_FORM_C4_EE) + self.F_JOURNAL):               # :635-649
                assemble_scalar_async(f(fid), self.scal, i)
        vals = self._finish_step()
        t_rec = self.t
        self.t += self.p["dt"]                          # :690
        return self._result(t_rec, vals, dict(E_k=0, E_e=1, torque=2, F_x=3, F_y=4))

    # -- post-loop adaptive refinement -------------------------------------------------
    REFINE_RATIO, REFINE_EPS, MAX_REFINE = 0.3, 0.000001, 1   # fenepmp_jbp.py:701,704,708

    def refinement_indicator(self, err_count=0):
        """The stress-residual error indicator the journal-bearing scripts evaluate after a time loop
        (fenepmp_jbp.py:691-711; the same statements at incompressible_benchmarkEWM.py:903-924 and
        comp_ewm_jpb.py:735-753 with other constants):

            kapp = project(inner(restau0, restau0), Qt);  norm_kapp = normalize_solution(kapp)   # in place
            ratio = REFINE_RATIO/(err_count + 1);  tau_average = project(sum(tau1_vec)/3, Qt)
            error_rat = absolute(project(kapp/(tau_average + eps), Qt))
            refine iff error_rat.max() > 0.01 and err_count < MAX_REFINE

        The three projections and the max run on the device; `markers` is what adaptive_refinement
        (fenics_base.py:143-155) hands to dolfin's refine(): kapp at the cell midpoint (= the cell's DG0 value)
        above the (1-ratio) percentile.  refine() itself and the rebuild of the plan on the new mesh stay with
        dolfin / the caller.  Overwrites `kapp` (the next loop's first step recomputes the LPS indicator)."""
        e, Qt = self.pr.expr, self.S["Qt"]
        if not hasattr(self, "_error_rat"):
            self._error_rat = Function(Qt, self.plan)
            self._scale = torch.zeros(3, dtype=torch.float64, device=self.scal.device)
        project(e(abi.FX_EXPR_ADAPT_RES_SQ), Qt, function=self.kapp)
        kv = self.kapp.vector()
        kmax = kv.norm("linf")                                         # normalize_solution: u /= max|u|
        self._scale[0] = 1.0 / kmax
        check(lib.fx_lincomb3_dev(kv.size(), la._ptr(kv.t), la._ptr(self._scale), la._ptr(kv.t), la._ptr(kv.t),
                                  la._ptr(kv.t), la._stream()))
        ratio = self.REFINE_RATIO / (1 * err_count + 1.0)
        project(e(abi.FX_EXPR_TAU_AVG), Qt, function=self.res_orth_norm_sq)   # tau_average in the Qt scratch slot
        eps_loop = self.pr.params[abi.FX_P_EPS]
        self.pr.set_param(abi.FX_P_EPS, self.REFINE_EPS)
        project(e(abi.FX_EXPR_ABS_KAPP_RATIO), Qt, function=self._error_rat)
        self.pr.set_param(abi.FX_P_EPS, eps_loop)
        err_max = self._error_rat.vector().norm("linf")
        kapp = kv.get_local()
        level = np.percentile(kapp, (1 - ratio) * 100)                  # adaptive_refinement(mesh, norm_kapp, ratio)
        markers = kapp[np.asarray(Qt.cell_dofs)[:, 0]] > level
        return dict(kapp=kapp, error_rat=self._error_rat.vector().get_local(), error_rat_max=err_max, ratio=ratio,
                    refine=bool(err_max > 0.01 and err_count < self.MAX_REFINE), markers=markers)

    def fields(self):
        g = lambda fn: fn.vector().get_local()  # noqa: E731
        return dict(w1=g(self.w1), p1=g(self.p1), tau1=g(self.tau1_vec), rho1=g(self.rho1), T1=g(self.T1),
                    kapp=g(self.kapp))
