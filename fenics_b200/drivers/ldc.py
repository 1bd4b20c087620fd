"""Lid-driven-cavity time loops on the GPU:

  IncompLDC  config 1  lid-driven-cavity/incomp_LDC.py:280-447   (incompressible Oldroyd-B)
  CompLDC    config 2  lid-driven-cavity/comp_LDC_mesh_convergence.py:376-565 (compressible Oldroyd-B)

Each `step()` issues the same sequence of assemble / bc.apply / solve / project calls as the reference
loop body, in the same order (forms see the coefficient values present at assemble time).  All fields
stay on the device; per step only the energies (and solver outcomes) reach the host.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from .. import _abi as abi
from .. import la
from .._lib import SolveInfo, check, lib
from ..dolfin_api import (ConsistentProjection, DirichletBC, Function, Problem, assemble,
                          assemble_scalar_async, project)
from ..la import Matrix, Plan, Vector
from ..mesh import DOLFIN_EPS
from ..partition import rcb_partition
from ..spaces import FunctionSpace, function_spaces


def lid_bcs(W, L=1.0):
    """bcu = [noslip, drive1] (incomp_LDC.py:171-226); returns (bcs, time holder)."""
    top_bound = 1.0
    tt = {"t": 0.0}

    def ulidreg(x):  # Expression(('8*(1.0+tanh(8*t-4.0))*(x[0]*(L-x[0]))*(x[0]*(L-x[0]))','0'))
        f = 8 * (1.0 + np.tanh(8 * tt["t"] - 4.0)) * (x[:, 0] * (L - x[:, 0])) ** 2
        return np.column_stack([f, np.zeros_like(f)])

    noslip = DirichletBC(W.sub(0), (0.0, 0.0), lambda x, on_boundary: np.ones(np.shape(x[0]), dtype=bool) & on_boundary)
    drive1 = DirichletBC(W.sub(0), ulidreg, lambda x, on_boundary: (x[1] > top_bound - DOLFIN_EPS) & on_boundary)
    return [noslip, drive1], tt


class PendingStep(dict):
    """Result of a pipelined step: the values are on their way to page-locked host memory.  Reading any
    item waits for that step (not for later ones), raises like dolfin if one of its solves failed, and
    fills the dict."""

    def __init__(self, drv, slot, static):
        super().__init__(static)
        self._drv, self._slot = drv, slot

    def resolve(self):
        if self._drv is not None:
            drv, self._drv = self._drv, None
            try:
                vals = drv._resolve_slot(self._slot)
            except Exception:
                self._drv = drv
                raise
            super().update({k: float(vals[i]) for k, i in self._slot["names"].items()})
        return self

    def __getitem__(self, k):
        if not dict.__contains__(self, k):
            self.resolve()
        return dict.__getitem__(self, k)

    def keys(self):
        return self.resolve() and dict.keys(self)

    def items(self):
        return self.resolve() and dict.items(self)


class _LDC:
    solver = ("bicgstab", "default")   # what the reference passes to solve()
    pressure_method = "bicgstab"       # reference; "cg" is the north_star upgrade (A5 is symmetric)
    pressure_pc = None                 # "jacobi_coarse" after enable_pressure_coarse() (CG only)
    STATE_FIELDS = ()                  # fields carried from step to step (host<->device in e2e mode)
    eliminate_devss = True             # W solves: iterate on the velocity rows, back-substitute the DG0 DEVSS rows
    extrapolate_guess = False          # start every solve from 2 x^n - x^(n-1) instead of x^n (the reference's
                                       # nonzero_initial_guess): same tolerance, fewer iterations; off by default
    pipeline = False                   # with check=False: step() returns a PendingStep instead of syncing, so the
                                       # host enqueues step k+1 while step k runs (solver failures raise on access)

    def _setup(self, mesh, dofmaps, device):
        self.mesh = mesh
        W, V, _, _, Zd, Zc, Q, Qt, _ = function_spaces(mesh, 2, dofmaps)
        Vs = FunctionSpace(mesh, "Vs", (dofmaps or {}).get("Vs"))  # V.sub(0).collapse(): stream function
        self.S = dict(W=W, V=V, Zd=Zd, Zc=Zc, Q=Q, Qt=Qt, Vs=Vs)
        self.plan = Plan(mesh, [W, Zc, Q, Zd, Qt, V, Vs], device)
        self.set_devss_elimination(self.eliminate_devss)
        self.pr = Problem(self.plan)
        self.pr.owner = self  # Functions bound to the tables know their time loop (shims: stream_function(u))
        for k, v in self.p.items():
            self.pr.set_param(getattr(abi, "FX_P_" + k.upper()), v)
        F = lambda s: Function(self.S[s], self.plan)  # noqa: E731
        b = self.pr.bind
        self.w0, self.w12, self.ws, self.w1 = (b(s, F("W")) for s in (abi.FX_F_W0, abi.FX_F_W12, abi.FX_F_WS, abi.FX_F_W1))
        self.p0, self.p1 = b(abi.FX_F_P0, F("Q")), b(abi.FX_F_P1, F("Q"))
        self.tau0_vec, self.tau1_vec = b(abi.FX_F_TAU0, F("Zc")), b(abi.FX_F_TAU1, F("Zc"))
        self.kapp = b(abi.FX_F_KAPP, F("Qt"))
        self.res_test = b(abi.FX_F_RES_TEST, F("Zd"))
        self.res_orth = b(abi.FX_F_RES_ORTH, F("Zc"))
        self.res_orth_norm_sq = b(abi.FX_F_QT_A, F("Qt"))
        self.err = F("Qt")
        self.bcu, self._bct = lid_bcs(W)
        # reusable tensors (the reference allocates new PETSc objects on every assemble() call)
        self.AW, self.AZ, self.AQ = Matrix(self.plan, "W"), Matrix(self.plan, "Zc"), Matrix(self.plan, "Q")
        self.bW, self.bZ, self.bQ = Vector(self.plan, "W"), Vector(self.plan, "Zc"), Vector(self.plan, "Q")
        self.scal = torch.zeros(8, dtype=torch.float64, device="cuda:%d" % self.plan.device)
        self._scal_host = torch.zeros(8, dtype=torch.float64).pin_memory()
        f = self.pr.form
        self.proj_res_orth = ConsistentProjection(f(abi.FX_FORM_MASS_ZC), f(abi.FX_FORM_LDC_L_RES_ORTH))
        self.solve_rtol = None       # None -> parameters["krylov_solver"] (PETSc default 1e-5)
        self.projection_rtol = None  # None -> 1e-13 (dolfin's project uses LU)
        self.check = True            # synchronous convergence check after every solve (dolfin behaviour)
        self.last_info = []
        self.solve_labels = []
        self._ticket = 0
        self.profile = None          # set to a list to record (label, start_event, end_event) per phase

    # -- post-loop diagnostics (lid-driven-cavity/fenics_fem.py:340-417) ------------------------------
    def stream_function(self, compressible=False, rtol=1e-12):
        """stream_function(u1) / comp_stream_function(rho1, u1): -lap(psi) = curl(u1) [rho1-weighted] on the
        scalar P2 space, psi = 0 on the whole boundary.  The reference solves directly (assemble_system +
        LU); here CG + Jacobi to `rtol`.  `project(u1, V)` before the call (incomp_LDC.py:457) is the
        identity for the P2 velocity of w1.  Returns the Function psi."""
        Vs = self.S["Vs"]
        if not hasattr(self, "_psi"):
            self._psi = Function(Vs, self.plan)
            self._Apsi, self._bpsi = Matrix(self.plan, "Vs"), Vector(self.plan, "Vs")
            self._bcpsi = DirichletBC(Vs, 0.0, lambda x, on_boundary: on_boundary)
        f = self.pr.form
        assemble(f(abi.FX_FORM_STREAM_A), tensor=self._Apsi)
        assemble(f(abi.FX_FORM_COMP_STREAM_L if compressible else abi.FX_FORM_STREAM_L), tensor=self._bpsi)
        self._bcpsi.apply(self._Apsi, self._bpsi)
        la.krylov_solve(self._Apsi, self._psi.vector(), self._bpsi, "cg", "jacobi", rtol=rtol, maxit=100000)
        return self._psi

    def shear_rate(self, rtol=1e-12):
        """shear_rate(u1) (lid-driven-cavity/fenics_fem.py:380-398; incomp_LDC.py:466, comp_LDC_mesh_convergence.py:613,
        compressible_dgp.py:643): the stream function's Poisson matrix with the right-hand side
        inner(0.5*gamma, gamma), gamma = grad(u1) + grad(u1)^T; CG + Jacobi to `rtol` where the reference uses LU."""
        Vs = self.S["Vs"]
        if not hasattr(self, "_gam"):
            self._gam = Function(Vs, self.plan)
            self._Agam, self._bgam = Matrix(self.plan, "Vs"), Vector(self.plan, "Vs")
            self._bcgam = DirichletBC(Vs, 0.0, lambda x, on_boundary: on_boundary)
        f = self.pr.form
        assemble(f(abi.FX_FORM_STREAM_A), tensor=self._Agam)
        assemble(f(abi.FX_FORM_SHEAR_L), tensor=self._bgam)
        self._bcgam.apply(self._Agam, self._bgam)
        la.krylov_solve(self._Agam, self._gam.vector(), self._bgam, "cg", "jacobi", rtol=rtol, maxit=100000)
        return self._gam

    def min_location(self, fn):
        """min_location(u, mesh): coordinates of the dofs holding the minimum of `fn`."""
        v = fn.vector().get_local()
        return fn.function_space().tabulate_dof_coordinates()[np.where(v == v.min())]

    def max_location(self, fn):
        v = fn.vector().get_local()
        return fn.function_space().tabulate_dof_coordinates()[np.where(v == v.max())]

    @staticmethod
    def _coarse_size(ndofs, k=None):
        """aggregates of the two-level preconditioner: 256 while the space fits the shared-memory-resident CG kernel
        (<= ~148 k rows), 1024 beyond (the coarse correction then runs inside k_cg, where a finer coarse space pays);
        always >= 8 dofs per aggregate and a multiple of 32."""
        if k is None:
            k = 256 if ndofs <= 150000 else 1024
        return min(int(k), (ndofs // 8) // 32 * 32)

    def enable_pressure_coarse(self, k=None):
        """Two-level preconditioner (Jacobi + aggregate coarse correction, fx_plan_set_aggregates) for the CG
        pressure solve; aggregates = recursive coordinate bisection of the P1 dof coordinates.  Singular
        (pure-Neumann: config 1, incompressible_dgp.py) and nearly singular (Ma -> 0) pressure matrices are handled on
        the device (rank-one shift of the coarse matrix).  Returns whether the library accepted the aggregates."""
        Q = self.S["Q"]
        k = self._coarse_size(Q.dim(), k)
        ok = k >= 32 and self.plan.set_aggregates("Q", rcb_partition(Q.tabulate_dof_coordinates(), k), k)
        self.pressure_pc = "jacobi_coarse" if ok else None
        return bool(ok)

    def _pressure_pc(self):
        rtol = la.parameters["krylov_solver"]["relative_tolerance"] if self.solve_rtol is None else self.solve_rtol
        return self.pressure_pc if (self.pressure_method == "cg" and rtol >= 1e-9) else None

    def set_devss_elimination(self, on):
        """The W systems are block lower triangular: the rows tested with R read (u, D) but no momentum row
        reads D, and D-D is the diagonal DG0 mass (e.g. incomp_LDC.py:317-325).  With elimination on the
        Krylov kernels iterate on the velocity block and recover D by exact back-substitution."""
        self.eliminate_devss = bool(on)
        rows = np.unique(self.S["W"].cell_dofs[:, 12:15]) if on else np.zeros(0, dtype=np.int32)
        self.plan.set_eliminated_dofs("W", rows)

    # -- helpers -------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def _phase(self, label):
        if self.profile is None:
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self.profile.append((label, e0, e1))

    def _solve(self, A, x, b, method=None, pc=None, rtol=None, label=""):
        method = method or self.solver[0]
        pc = pc or self.solver[1]
        rtol = self.solve_rtol if rtol is None else rtol
        self.solve_labels.append("%s:%s:%s" % (label, A.space, method))
        with self._phase("solve:%s:%s:%s" % (label, A.space, method)):
            hist = self._guess_history(label, x) if (self.extrapolate_guess and label) else None
            if hist is not None and hist["n"] >= 2:  # x0 = 2 x^n - x^(n-1)
                xt = x.vector().t
                check(lib.fx_lincomb3_dev(xt.numel(), la._ptr(xt), la._ptr(hist["coef"]), la._ptr(hist["v"][0]),
                                          la._ptr(hist["v"][1]), la._ptr(hist["v"][1]), la._stream()))
            if self.check:
                self.last_info.append(la.krylov_solve(A, x.vector(), b, method, pc, rtol=rtol))
            else:
                la.krylov_solve(A, x.vector(), b, method, pc, rtol=rtol, ticket=self._ticket)
                self._ticket += 1
            if hist is not None:  # keep the two latest solutions of this solve site (device copies, stream ordered)
                hist["v"].reverse()
                hist["v"][0].copy_(x.vector().t)
                hist["n"] += 1

    def _guess_history(self, label, x):
        if not hasattr(self, "_hist"):
            self._hist = {}
        if label not in self._hist:
            t = x.vector().t
            self._hist[label] = dict(n=0, v=[torch.empty_like(t), torch.empty_like(t)],
                                     coef=torch.tensor([2.0, -1.0, 0.0], dtype=torch.float64, device=t.device))
        return self._hist[label]

    def _project_consistent(self, proj, out, A, b, label):
        with self._phase("assemble_matrix:%s:%s" % (label, A.space)):
            assemble(proj.mass_form, tensor=A)
        with self._phase("assemble_vector:%s:%s" % (label, A.space)):
            assemble(proj.rhs_form, tensor=b)
        rtol = 1e-13 if self.projection_rtol is None else self.projection_rtol
        self._solve(A, out, b, "cg", "jacobi", rtol=rtol, label=label)

    def lps_indicator(self):
        """kappa of the LPS stabilisation (incomp_LDC.py:300-308 == comp_LDC_mesh_convergence.py:393-401)."""
        e = self.pr.expr
        with self._phase("project_dg0:restau"):
            project(e(abi.FX_EXPR_LDC_RESTAU), self.S["Zd"], function=self.res_test)
        self._project_consistent(self.proj_res_orth, self.res_orth, self.AZ, self.bZ, "res_orth")
        with self._phase("project_dg0:kappa"):
            project(e(abi.FX_EXPR_RES_ORTH_SQ), self.S["Qt"], function=self.res_orth_norm_sq)
            project(e(abi.FX_EXPR_SQRT_QT_A), self.S["Qt"], function=self.kapp)

    def _system(self, label, a_id, L_id, A, b, x, bcs=(), **kw):
        f = self.pr.form
        with self._phase("assemble_matrix:%s:%s" % (label, A.space)):
            assemble(f(a_id), tensor=A)
        with self._phase("assemble_vector:%s:%s" % (label, A.space)):
            assemble(f(L_id), tensor=b)
        if bcs:
            with self._phase("bc:%s" % label):
                for bc in bcs:
                    bc.apply(A, b)
        self._solve(A, x, b, label=label, **kw)

    def _begin_step(self):
        self.last_info, self.solve_labels, self._ticket = [], [], 0

    def _check_infos(self, infos, labels):
        bad = [i for i, s in enumerate(infos) if s.converged != 1]
        if bad:
            s = infos[bad[0]]
            raise RuntimeError("*** Error: Unable to solve linear system: solve %s of the step: %d iterations, "
                               "residual %.3e, status %d" % (labels[bad[0]], s.iterations, s.residual, s.converged))

    def _resolve_slot(self, slot):
        if slot["pending"]:
            slot["event"].synchronize()
            slot["pending"] = False
            n = slot["count"]
            infos = list((SolveInfo * n).from_buffer_copy(slot["info"].numpy()[:n * C.sizeof(SolveInfo)].tobytes()))
            slot["infos"], slot["vals"] = infos, slot["scal"].numpy().copy()
            self.last_info, self.solve_labels = infos, slot["labels"]
            self._check_infos(infos, slot["labels"])
        return slot["vals"]

    def sync(self):
        """wait for every pipelined step still in flight (raises if one of its solves failed)."""
        for slot in getattr(self, "_slots", []):
            if slot.get("owner") is not None:
                slot.pop("owner").resolve()
            self._resolve_slot(slot)

    def _result(self, t_rec, vals, names):
        """names: result key -> index into the step's scalar buffer."""
        if isinstance(vals, dict):  # pipelined: a ring slot
            vals["names"] = names
            vals["owner"] = PendingStep(self, vals, dict(t=t_rec))
            return vals["owner"]
        out = dict(t=t_rec)
        out.update({k: float(vals[i]) for k, i in names.items()})
        return out

    def _finish_step(self, allow_pipeline=True):
        """one host sync per step: energies (+ solver outcomes when running asynchronously); in pipelined
        mode no sync at all: the copies are enqueued behind the step and an event marks their arrival."""
        if self.pipeline and not self.check and allow_pipeline:
            if not hasattr(self, "_slots"):
                mk = lambda: dict(scal=torch.zeros(8, dtype=torch.float64).pin_memory(),  # noqa: E731
                                  info=torch.zeros(64 * C.sizeof(SolveInfo), dtype=torch.uint8).pin_memory(),
                                  event=torch.cuda.Event(), pending=False, vals=None)
                self._slots, self._slot_i = [mk(), mk()], 0
            slot = self._slots[self._slot_i]
            self._slot_i ^= 1
            if slot.get("owner") is not None:  # two steps old: long finished; fill it in before its buffers are reused
                slot.pop("owner").resolve()
            self._resolve_slot(slot)
            slot["scal"].copy_(self.scal, non_blocking=True)
            if self._ticket:
                check(lib.fx_solve_results_async(self.plan.h, 0, self._ticket, C.c_void_p(slot["info"].data_ptr()),
                                                 la._stream()))
            slot["event"].record()
            slot.update(pending=True, count=self._ticket, labels=list(self.solve_labels))
            return slot
        if not self.check:
            infos = la.solve_results(self.plan, 0, self._ticket)
            self.last_info = infos
            self._check_infos(infos, self.solve_labels)
        self._scal_host.copy_(self.scal, non_blocking=False)
        return self._scal_host.numpy()

    # -- host <-> device state transfer (end-to-end mode of bench.py, I/O hand-off) -----------------
    def state_functions(self):
        return [getattr(self, n) for n in self.STATE_FIELDS]

    def result_functions(self):
        return [getattr(self, n) for n in self.RESULT_FIELDS]


class IncompLDC(_LDC):
    STATE_FIELDS = ("w0", "p0", "tau0_vec", "w1", "tau1_vec")
    RESULT_FIELDS = ("w1", "p1", "tau1_vec")

    def __init__(self, mesh, dofmaps=None, dt=0.005, Re=0.5, We=0.25, betav=0.5, c1=0.05, c2=0.01, th=1.0,
                 conv=None, device=None):
        if conv is None:
            conv = 0.0 if Re < 1 else 1.0  # incomp_LDC.py:49,93-94
        self.p = dict(dt=dt, Re=Re, We=We, betav=betav, c1=c1, c2=c2, th=th, conv=conv)
        self._setup(mesh, dofmaps, device)
        self.t = dt  # :277
        self.err_array = []

    def step(self):
        self._begin_step()
        self._bct["t"] = self.t  # ulidreg.t = t  :293
        self.lps_indicator()  # :300-309
        self._system("u12", abi.FX_FORM_C1_A1, abi.FX_FORM_C1_L1, self.AW, self.bW, self.w12, self.bcu)  # :317-331
        self._system("us", abi.FX_FORM_C1_A2, abi.FX_FORM_C1_L2, self.AW, self.bW, self.ws, self.bcu)  # :355-363
        self._system("p", abi.FX_FORM_C1_A5, abi.FX_FORM_C1_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method, pc=self._pressure_pc())  # :370-375 (bcp = [])
        self._system("u1", abi.FX_FORM_C1_A7, abi.FX_FORM_C1_L7, self.AW, self.bW, self.w1, self.bcu)  # :379-394
        self._system("tau", abi.FX_FORM_C1_A4, abi.FX_FORM_C1_L4, self.AZ, self.bZ, self.tau1_vec, ())  # :400-416
        f = self.pr.form
        with self._phase("functionals"):
            assemble_scalar_async(f(abi.FX_FORM_C1_EK), self.scal, 0)  # :420
            assemble_scalar_async(f(abi.FX_FORM_C1_EE), self.scal, 1)  # :421
            t_rec = self.t
            self.t += self.p["dt"]  # :429
            project(self.pr.expr(abi.FX_EXPR_H_KAPP), self.S["Qt"], function=self.err)  # :434
            ev = self.err.vector()
            check(lib.fx_norm_dev(self.plan.h, ev.size(), la._ptr(ev.t), 0, C.c_void_p(self.scal.data_ptr() + 16),
                                  la._stream()))  # norm(err.vector(),'l2') :435
        vals = self._finish_step(allow_pipeline=False)  # the divergence test below needs this step's error norm
        self.err_array.append(float(vals[2]))
        diverged = False
        if self.t > 0.5 and len(self.err_array) > 1 and self.err_array[-2] != 0.0:
            diverged = self.err_array[-1] / self.err_array[-2] > 1.1  # :437-441
        self.w0.assign(self.w1)  # :444-447
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        return dict(t=t_rec, E_k=float(vals[0]), E_e=float(vals[1]), err=self.err_array[-1], diverged=diverged)

    def fields(self):
        p = self.p1.vector().get_local()
        return dict(w1=self.w1.vector().get_local(), p1=p - p.mean(), tau1=self.tau1_vec.vector().get_local(),
                    kapp=self.kapp.vector().get_local())


class CompLDC(_LDC):
    """lid-driven-cavity/comp_LDC_mesh_convergence.py:376-565 (config 2)."""
    STATE_FIELDS = ("w0", "p0", "tau0_vec", "rho0", "w1", "tau1_vec")
    RESULT_FIELDS = ("w1", "p1", "tau1_vec", "rho1")

    def __init__(self, mesh, dofmaps=None, dt=0.001, Re=10.0, We=0.25, Ma=0.001, betav=0.5, c1=0.1, c2=0.01,
                 th=1.0, conv=1.0, device=None):
        self.p = dict(dt=dt, Re=Re, We=We, Ma=Ma, betav=betav, c1=c1, c2=c2, th=th, conv=conv)
        self._setup(mesh, dofmaps, device)
        b, F = self.pr.bind, lambda s: Function(self.S[s], self.plan)  # noqa: E731
        self.rho0, self.rho12, self.rho1 = b(abi.FX_F_RHO0, F("Q")), b(abi.FX_F_RHO12, F("Q")), b(abi.FX_F_RHO1, F("Q"))
        f = self.pr.form
        self.proj_rho1 = ConsistentProjection(f(abi.FX_FORM_MASS_Q), f(abi.FX_FORM_C2_L_RHO1))
        self.rho0.vector().t.fill_(1.0)  # rho0 = project(1.0, Q)  :285-286
        self.t = dt  # :373

    def step(self):
        self._begin_step()
        self._bct["t"] = self.t
        self.lps_indicator()  # :393-402
        self._system("rho12", abi.FX_FORM_C2_A0, abi.FX_FORM_C2_L0, self.AQ, self.bQ, self.rho12, ())  # :417-426
        self._system("u12", abi.FX_FORM_C2_A1, abi.FX_FORM_C2_L1, self.AW, self.bW, self.w12, self.bcu)  # :430-447
        self._system("us", abi.FX_FORM_C2_A2, abi.FX_FORM_C2_L2, self.AW, self.bW, self.ws, self.bcu)  # :458-476
        self._system("p", abi.FX_FORM_C2_A5, abi.FX_FORM_C2_L5, self.AQ, self.bQ, self.p1, (),
                     method=self.pressure_method, pc=self._pressure_pc())  # :483-497
        self._project_consistent(self.proj_rho1, self.rho1, self.AQ, self.bQ, "rho1")  # :501-502
        self._system("u1", abi.FX_FORM_C2_A7, abi.FX_FORM_C2_L7, self.AW, self.bW, self.w1, self.bcu)  # :506-519
        self._system("tau", abi.FX_FORM_C2_A4, abi.FX_FORM_C2_L4, self.AZ, self.bZ, self.tau1_vec, ())  # :528-544
        f = self.pr.form
        with self._phase("functionals"):
            assemble_scalar_async(f(abi.FX_FORM_C2_EK), self.scal, 0)  # :548
            assemble_scalar_async(f(abi.FX_FORM_C2_EE), self.scal, 1)  # :549
        vals = self._finish_step()
        t_rec = self.t
        self.w0.assign(self.w1)  # :558-562
        self.rho0.assign(self.rho1)
        self.p0.assign(self.p1)
        self.tau0_vec.assign(self.tau1_vec)
        self.t += self.p["dt"]
        return self._result(t_rec, vals, dict(E_k=0, E_e=1))

    def fields(self):
        g = lambda f: f.vector().get_local()  # noqa: E731
        return dict(w1=g(self.w1), p1=g(self.p1), tau1=g(self.tau1_vec), rho1=g(self.rho1), rho12=g(self.rho12),
                    kapp=g(self.kapp))


class CompLDCSweep(CompLDC):
    """lid-driven-cavity/comp_LDC.py:364-594, the Re / We / Ma sweep script.  Its loop is config
    2's loop statement for statement, except that the convective terms of the two velocity predictor steps are
    written `Re*conv*dot(u, nabla_grad(u))` (:418, :468) where the mesh-convergence script has `conv*...`: the
    same form functors with the parameter slot FX_P_CONV holding Re*conv."""

    def __init__(self, mesh, dofmaps=None, Re=10.0, conv=1.0, **kw):
        super().__init__(mesh, dofmaps=dofmaps, Re=Re, conv=Re * conv, **kw)
