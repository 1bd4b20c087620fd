"""GPU-backed versions of the dolfin names the reference's time loops call (Level 1):
assemble, solve, project, KrylovSolver, DirichletBC.apply, Function.assign/.vector(), norm, parameters.

Forms are not parsed from UFL: each UFL form the scripts assemble has a hand-written CUDA functor,
identified by a form id of include/fenics_b200.h; a `Form` object pairs that id with the run's coefficient
("state") and parameter tables, the way a dolfin Form holds references to its Functions and Constants.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi as abi
from . import la
from . import symbolic
from ._lib import check, lib
from .la import Matrix, Plan, Vector, norm, parameters  # noqa: F401  (re-exported)
from .mesh import DOLFIN_EPS, _call_inside, facets_inside  # noqa: F401

_P2_FACET_DOFS = np.array([[1, 2, 3], [0, 2, 4], [0, 1, 5]])
_P1_FACET_DOFS = np.array([[1, 2], [0, 2], [0, 1]])


class Function(symbolic.Operand):
    """dolfin Function stand-in living on the device.  Inside a symbolic form (fenics_b200/symbolic.py) it is a
    coefficient: `inner(p0, div(v))*dx`, `(u0, D0_vec) = w0.split()`."""

    def __init__(self, space, plan):
        self._space, self.plan = space, plan
        self._vec = Vector(plan, space.name)

    def split(self):
        from .form_registry import split
        return split(self)

    def function_space(self):
        return self._space

    def vector(self):
        return self._vec

    def assign(self, other):
        self._vec.t.copy_(other._vec.t)

    def copy(self, deepcopy=True):
        f = Function(self._space, self.plan)
        f.assign(self)
        return f


class Problem:
    """Per-run tables handed to every assembly call: fx_field slot -> device vector, fx_param slot ->
    double.  Equivalent of the Function/Constant references held inside the reference's UFL forms."""

    def __init__(self, plan):
        self.plan = plan
        self.fields = {}
        self.params = np.zeros(abi.NPARAMS, dtype=np.float64)
        self._table = (C.c_void_p * abi.NFIELDS)()

    def bind(self, slot, fn):
        self.fields[slot] = fn
        self._table[slot] = fn.vector().t.data_ptr()
        if getattr(self, "owner", None) is not None:
            fn._driver = self.owner
        return fn

    def set_param(self, slot, value):
        self.params[slot] = float(value)

    def state_ptr(self):
        return self._table

    def params_ptr(self):
        return C.c_void_p(self.params.ctypes.data)

    def form(self, form_id):
        return Form(self, form_id)

    def expr(self, expr_id):
        return DG0Expr(self, expr_id)


class Form:
    def __init__(self, problem, form_id):
        self.problem, self.id = problem, form_id
        rank, space = C.c_int(), C.c_int()
        q = (C.c_int * 2)()
        check(lib.fx_form_info(form_id, C.byref(rank), C.byref(space), q))
        self.rank = rank.value
        self.space = {v: k for k, v in abi.SPACE_ID.items()}.get(space.value)
        self.qdeg = (q[0], q[1])


class DG0Expr:
    def __init__(self, problem, expr_id):
        self.problem, self.id = problem, expr_id


class ConsistentProjection:
    """project(expr, V) onto a continuous space: (mass-matrix form, RHS form)."""

    def __init__(self, mass_form, rhs_form):
        self.mass_form, self.rhs_form = mass_form, rhs_form


def TestFunction(V):
    from .form_registry import arguments
    return arguments(V, 0)


def TrialFunction(V):
    from .form_registry import arguments
    return arguments(V, 1)


TestFunctions, TrialFunctions = TestFunction, TrialFunction  # mixed spaces: arguments() already returns the split


def _problem_of(plan):
    """the field / parameter tables that literal (symbolic) forms of a plan are assembled with"""
    pr = getattr(plan, "_sym_problem", None)
    if pr is None:
        pr = plan._sym_problem = Problem(plan)
    return pr


def _recognised(form):
    """symbolic Form -> Form(problem, fx_form id) with the form's own Functions and Parameters bound (or DG0Expr(problem,
    fx_expr id) for the right-hand side of a projection onto a DG0 space)"""
    fid, coefs, params, what, space_objs = symbolic.recognise(form)
    plan = None
    for _, fn in coefs:
        plan = plan or getattr(fn, "plan", None)
    for sp in space_objs:
        plan = plan or getattr(sp, "_shim_plan", None)
    if plan is None:
        raise RuntimeError("assemble(%s): no device plan reachable from the form (create the Functions with "
                           "solution_functions / Function(V, plan) first)" % what)
    pr = _problem_of(plan)
    for slot, fn in coefs:
        if slot >= 0:   # slot -1: a coefficient the functor replaces by a constant of its own (form_registry.py)
            pr.bind(slot, fn)
    for name, par in params.items():
        pr.set_param(getattr(abi, "FX_P_" + name.upper()), par.value)
    if isinstance(fid, tuple):
        return DG0Expr(pr, fid[1])
    return Form(pr, fid)


def assemble(form, tensor=None):
    """assemble(a) -> Matrix, assemble(L) -> Vector, assemble(M) -> float
    (lid-driven-cavity/incomp_LDC.py:328-329, :420-421).  `tensor` reuses storage.  `form`: a Form of a Problem
    (fx_form id) or a symbolic form built with the scripts' own statements (recognised by structure)."""
    if isinstance(form, symbolic.Form):
        form = _recognised(form)
    pr = form.problem
    st = la._stream()
    if form.rank == 2:
        A = tensor if tensor is not None else Matrix(pr.plan, form.space)
        check(lib.fx_assemble_matrix(pr.plan.h, form.id, pr.state_ptr(), pr.params_ptr(), la._ptr(A.val), st))
        return A
    if form.rank == 1:
        b = tensor if tensor is not None else Vector(pr.plan, form.space)
        check(lib.fx_assemble_vector(pr.plan.h, form.id, pr.state_ptr(), pr.params_ptr(), la._ptr(b.t), st))
        return b
    out = C.c_double()
    check(lib.fx_assemble_scalar(pr.plan.h, form.id, pr.state_ptr(), pr.params_ptr(), C.byref(out), st))
    return out.value


def assemble_scalar_async(form, out_tensor, index=0):
    """functional whose value stays on the device (out_tensor[index]) -- no host sync."""
    pr = form.problem
    ptr = C.c_void_p(out_tensor.data_ptr() + 8 * index)
    check(lib.fx_assemble_scalar_dev(pr.plan.h, form.id, pr.state_ptr(), pr.params_ptr(), ptr, la._stream()))


class DirichletBC:
    """DirichletBC(V.sub(i), value, sub_domain) with dolfin's default 'topological' search: dofs in the
    closure of exterior facets whose vertices and midpoint are inside(x, on_boundary=True).
    value: constants (tuple/float) or a callable x (n,2) -> (n, ncomp) evaluated at the dof coordinates
    at apply() time (time-dependent Expressions: lid-driven-cavity/incomp_LDC.py:214,224,293)."""

    def __init__(self, space, value, sub_domain, plan=None):
        root = space.parent or space
        self.space, self.root, self.value = space, root, value
        mesh = root.mesh
        ok = facets_inside(mesh, sub_domain)
        cells, lf = mesh.bfacet_cell[ok], mesh.bfacet_local[ok]
        dofs, comp = [], []
        for j, c in enumerate(space.sub_comps):
            kind = root.comps[c]
            if kind == "DG0":
                continue
            loc = (_P2_FACET_DOFS if kind == "P2" else _P1_FACET_DOFS)[lf] + root.offsets[c]
            d = root.cell_dofs[cells[:, None], loc].ravel()
            dofs.append(d)
            comp.append(np.full(d.size, j))
        d = np.concatenate(dofs) if dofs else np.zeros(0, dtype=np.int64)
        cj = np.concatenate(comp) if comp else np.zeros(0, dtype=np.int64)
        self.dofs, first = np.unique(d, return_index=True)
        self.dof_comp = cj[first]
        self.x = root.tabulate_dof_coordinates()[self.dofs] if len(self.dofs) else np.zeros((0, 2))
        self._rows = None
        self._g = None
        self._gpinned = None

    def values(self):
        v = self.value
        if callable(v):
            v = np.asarray(v(self.x), dtype=np.float64)
        else:
            v = np.broadcast_to(np.asarray(v, dtype=np.float64), (len(self.dofs),) + np.shape(v))
        if v.ndim == 2:
            v = v[np.arange(len(self.dofs)), self.dof_comp]
        return np.ascontiguousarray(v, dtype=np.float64)

    def apply(self, A=None, b=None):
        plan = (A or b).plan
        n = len(self.dofs)
        if n == 0:
            return
        dev = "cuda:%d" % plan.device
        if self._rows is None:
            self._rows = torch.as_tensor(self.dofs.astype(np.int32), device=dev)
            self._g = torch.empty(n, dtype=torch.float64, device=dev)
            # page-locked staging ring: the host may run ahead of the device (pipelined time loops), so a
            # staging buffer is rewritten only after the copy that last read it has completed
            self._gpinned = [(torch.empty(n, dtype=torch.float64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._gi = 0
        stage, ev = self._gpinned[self._gi]
        self._gi = (self._gi + 1) % len(self._gpinned)
        ev.synchronize()  # no-op until the ring wraps around
        stage.numpy()[:] = self.values()
        self._g.copy_(stage, non_blocking=True)
        ev.record()
        check(lib.fx_apply_bc(plan.h, abi.SPACE_ID[self.root.name], la._ptr(A.val) if A is not None else None,
                              la._ptr(b.t) if b is not None else None, la._ptr(self._rows), la._ptr(self._g), n,
                              la._stream()))


def solve(A, x, b, method="lu", pc="default", **kw):
    """solve(A, x, b, "bicgstab", "default") (incomp_LDC.py:331).  `solve(A, x, b)` (dolfin: LU) is run as
    BiCGStab to a tight tolerance -- there is no sparse direct solver on the device."""
    if method == "lu":
        kw.setdefault("rtol", parameters.get("lu_equivalent_rtol", 1e-13))
        method = "bicgstab"
    return la.krylov_solve(A, x, b, method, pc, **kw)


class KrylovSolver:
    """KrylovSolver(method, preconditioner).solve(A, x, b) ("flow between eccentrically rotating
    cylinders"/fenepmp_jbp.py:443-445,611)."""

    def __init__(self, method="bicgstab", preconditioner="default"):
        self.method, self.pc = method, preconditioner
        self.parameters = dict(parameters["krylov_solver"])

    def solve(self, A, x, b):
        p = self.parameters
        return la.krylov_solve(A, x, b, self.method, self.pc, rtol=p["relative_tolerance"],
                               atol=p["absolute_tolerance"], maxit=p["maximum_iterations"])


def project(v, V, function=None, rtol=None):
    """dolfin.project(v, V): DG0 targets have a diagonal mass matrix (cell average, one kernel); continuous
    targets assemble the consistent mass matrix and RHS and solve with CG+Jacobi to `rtol` (dolfin uses
    LU, i.e. an exact solve; default 1e-13)."""
    if isinstance(v, DG0Expr):
        pr = v.problem
        out = function if function is not None else Function(V, pr.plan)
        check(lib.fx_project_dg0(pr.plan.h, v.id, pr.state_ptr(), pr.params_ptr(), la._ptr(out.vector().t),
                                 la._stream()))
        return out
    if isinstance(v, ConsistentProjection):
        pr = v.rhs_form.problem
        out = function if function is not None else Function(V, pr.plan)
        M = assemble(v.mass_form)
        rhs = assemble(v.rhs_form)
        la.krylov_solve(M, out.vector(), rhs, "cg", "jacobi",
                        rtol=parameters.get("projection_rtol", 1e-13) if rtol is None else rtol)
        return out
    if isinstance(v, (symbolic.E, symbolic.Operand)):
        # the script's own expression (`rho1 = project(rho0 + (Ma*Ma/Re)*(p1-p0), Q)`, comp_LDC_mesh_convergence.py:
        # 501-502): dolfin.project's two forms, recognised like any other literal form
        w, Pv = TestFunction(V), TrialFunction(V)
        target = _recognised(symbolic.inner(w, v) * symbolic.dx)
        if isinstance(target, DG0Expr):   # DG0 target (Zd, Qt): diagonal mass matrix, one cell-average kernel
            return project(target, V, function=function)
        M = assemble(symbolic.inner(w, Pv) * symbolic.dx)
        rhs = assemble(target)
        out = function if function is not None else Function(V, M.plan)
        la.krylov_solve(M, out.vector(), rhs, "cg", "jacobi",
                        rtol=parameters.get("projection_rtol", 1e-13) if rtol is None else rtol)
        return out
    raise TypeError("project(): expected a DG0Expr, a ConsistentProjection or a symbolic expression")
