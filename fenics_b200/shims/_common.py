"""Shared implementation of the three shim modules (see package docstring)."""
import os
import sys

import numpy as np

from .. import mesh as _mesh
from ..dolfin_api import Function

pi = 3.14159265359  # the reference pins this literal (lid-driven-cavity/fenics_fem.py:20)
DOLFIN_EPS = 3.0e-16

# ---- algebra namespace of the symbolic helpers -----------------------------------------------------
_ALG = None


def set_algebra(namespace):
    """namespace providing grad, div, dot, inner, transpose, Identity, as_matrix, nabla_grad (ufl, or
    oracle.ufl_lite in the tests)."""
    global _ALG
    _ALG = namespace


def alg(operand=None):
    """the algebra namespace for an operand: the library's own symbolic module for its own expressions / Functions
    (whatever set_algebra selected), else the configured one"""
    global _ALG
    if operand is not None:
        from .. import symbolic
        if isinstance(operand, (symbolic.E, symbolic.Operand)):
            return symbolic
    if _ALG is None:
        try:
            import ufl
            _ALG = ufl
        except ImportError:
            from .. import symbolic  # the library's own UFL-named algebra: forms built with it are recognised by
            _ALG = symbolic          # structure and assembled by the hand-written functors (fenics_b200/symbolic.py)
    return _ALG


def _T(a):
    A = alg(a)
    return A.transpose(a) if hasattr(A, "transpose") else a.T


# ---- host utilities ---------------------------------------------------------------------------------
def update_progress(job_title, progress):
    """text progress bar (lid-driven-cavity/fenics_fem.py:23-30)."""
    width = 20
    filled = int(round(width * progress))
    line = "\r{0}: [{1}] {2}%".format(job_title, "#" * filled + "-" * (width - filled), round(progress * 100, 4))
    if progress >= 1:
        line += " DONE\r\n"
    sys.stdout.write(line)
    sys.stdout.flush()


def chunks(l, n):
    """successive n-sized pieces of l (fenics_fem.py:444-447)."""
    for start in range(0, len(l), n):
        yield l[start:start + n]


def mesh2triang(mesh):
    import matplotlib.tri as tri
    xy = mesh.coordinates()
    return tri.Triangulation(xy[:, 0], xy[:, 1], mesh.cells())


def mplot(obj):
    """tripcolor of a Function (cell-wise for DG0, vertex values otherwise) or triplot of a Mesh
    (fenics_fem.py:36-51); values are copied from the device once."""
    import matplotlib.pyplot as plt
    plt.gca().set_aspect("equal")
    if isinstance(obj, Function):
        m, S = obj.function_space().mesh, obj.function_space()
        vals = obj.vector().get_local()
        if len(vals) == m.num_cells():
            plt.tripcolor(mesh2triang(m), vals[np.asarray(S.cell_dofs)[:, 0]])
        else:
            plt.tripcolor(mesh2triang(m), compute_vertex_values(obj), shading="gouraud")
    else:
        plt.triplot(mesh2triang(obj), color="k")


def compute_vertex_values(fn):
    """values of a scalar P1/P2 Function at the mesh vertices (local dofs 0..2 of every cell are its vertices)."""
    S = fn.function_space()
    cd, cells = np.asarray(S.cell_dofs), S.mesh.cells()
    out = np.zeros(S.mesh.num_vertices())
    out[cells.ravel()] = fn.vector().get_local()[cd[:, :3].ravel()]
    return out


def _npy(name):
    os.makedirs("data", exist_ok=True)
    return os.path.join("data", name + ".npy")


# ---- Functions living on the device -----------------------------------------------------------------
def normalize_solution(u):
    "u / max|u|, in place (fenics_fem.py:316-324)"
    v = u.vector()
    v.t.mul_(1.0 / v.norm("linf"))
    return u


def absolute(u):
    "|u| entrywise, in place (fenics_fem.py:450-453)"
    u.vector().t.abs_()
    return u


def min_location(u, mesh):
    """coordinates of the dofs where the scalar Function u attains its minimum (fenics_fem.py:402-417)."""
    vals = u.vector().get_local()
    return u.function_space().tabulate_dof_coordinates().reshape((-1, 2))[np.where(vals == vals.min())]


def max_location(u, mesh):
    vals = u.vector().get_local()
    return u.function_space().tabulate_dof_coordinates().reshape((-1, 2))[np.where(vals == vals.max())]


def adaptive_refinement(mesh, kapp, ratio):
    """fenics_fem.py:326-338 / fenics_base.py:143-155.  The cell marking (kapp at the cell midpoint above the
    (1-ratio) percentile) is computed here; refine() is dolfin's (Plaza), so a dolfin mesh is refined and returned,
    while for the library's own Mesh the boolean markers are returned for the caller to hand to dolfin."""
    vals = kapp.vector().get_local()
    level = np.percentile(vals, (1 - ratio) * 100)
    markers = vals[np.asarray(kapp.function_space().cell_dofs)[:, 0]] > level
    if hasattr(mesh, "ufl_cell"):  # a dolfin mesh
        import dolfin
        cd = dolfin.MeshFunction("bool", mesh, 2)
        cd.array()[:] = markers
        return dolfin.refine(mesh, cd, redistribute=True)
    return markers


def _needs_dolfin(name):
    def f(*a, **k):
        raise NotImplementedError("%s: mesh generation by mshr / Plaza refinement stay in dolfin (north_star); without "
                                  "dolfin use the structured meshes of fenics_b200.mesh" % name)
    f.__name__ = name
    return f


# ---- skew maps and structured meshes ------------------------------------------------------------------
skewcavity, xskewcavity, yskewcavity, expskewcavity = (_mesh.skewcavity, _mesh.xskewcavity, _mesh.yskewcavity,
                                                       _mesh.expskewcavity)
Skew_Mesh, DGP_structured_mesh = _mesh.Skew_Mesh, _mesh.DGP_structured_mesh


def LDC_Regular_Mesh(mm, B, L):
    """uniform mm x mm right-diagonal mesh of [0,B]x[0,L] (fenics_fem.py:75-89)."""
    return _mesh.RectangleMesh((0.0, 0.0), (B, L), mm, mm)


# ---- symbolic tensor algebra (shared by all three helper modules) ---------------------------------------
def tgrad(w):
    "transpose of grad(w)"
    return _T(alg(w).grad(w))


def Dincomp(w):
    "rate of strain, incompressible"
    g = alg(w).grad(w)
    return 0.5 * (g + _T(g))


def Dcomp(w):
    "rate of strain, compressible: deviatoric part"
    A = alg(w)
    g = A.grad(w)
    return 0.5 * (g + _T(g)) - (1.0 / 3) * A.div(w) * A.Identity(len(w))


def Fdef_incompressible(u, Tau):
    "upper-convected terms u.grad(Tau) - grad(u) Tau - Tau grad(u)^T"
    A = alg(u)
    return A.dot(u, A.grad(Tau)) - A.dot(A.grad(u), Tau) - A.dot(Tau, tgrad(u))


def Fdef_compressible(u, Tau):
    A = alg(u)
    return Fdef_incompressible(u, Tau) + A.div(u) * Tau


def ldc_sigmain(u, p, Tau, We, betav):
    """lid-driven-cavity/fenics_fem.py:304-305: incompressible total stress 2 betav D(u) - p I + (1-betav)/We (Tau - I)"""
    I = alg(u).Identity(len(u))
    return 2 * betav * Dincomp(u) - p * I + ((1 - betav) / We) * (Tau - I)


def ldc_sigma(u, p, Tau, We, betav):
    """lid-driven-cavity/fenics_fem.py:307-308: compressible total stress, deviatoric rate of strain"""
    I = alg(u).Identity(len(u))
    return 2 * betav * Dcomp(u) - p * I + ((1 - betav) / We) * (Tau - I)


# ---- "flow between eccentrically rotating cylinders"/fenics_base.py:158-294: operand-aware, shared by the drop-in module and
# the form registry (so a script's forms and the registered ones are built by the same code) -----------------------------
def jbp_sigma(u, p, Tau, betav, We):
    return 2 * betav * Dcomp(u) - p * alg(u).Identity(len(u)) + Tau


def jbp_sigmacon(u, p, Tau, betav, We):
    I = alg(u).Identity(len(u))
    return 2 * betav * Dcomp(u) - p * I + ((1.0 - betav) / We) * (Tau - I)


def _fene_polymer(u, Tau, b, lambda_d, betav, We):
    I = alg(u).Identity(len(u))
    return ((1. - betav) / We) * (phi_def(u, lambda_d) * (fene_func(Tau, b) * Tau - I))


def fene_sigma(u, p, Tau, b, lambda_d, betav, We):
    return 2.0 * betav * Dincomp(u) - p * alg(u).Identity(len(u)) + _fene_polymer(u, Tau, b, lambda_d, betav, We)


def fene_sigmacom(u, p, Tau, b, lambda_d, betav, We):
    return 2.0 * betav * Dcomp(u) - p * alg(u).Identity(len(u)) + _fene_polymer(u, Tau, b, lambda_d, betav, We)


def FdefG(u, G, Tau):
    """DEVSS-G upper-convected terms with the projected velocity gradient G"""
    A = alg(u)
    return A.dot(u, A.grad(Tau)) - A.dot(G, Tau) - A.dot(Tau, _T(G))


def Tr(tau):
    return abs(tau[0, 0] + tau[1, 1])


def Trace(tau):
    return tau[0, 0] + tau[1, 1]


def fene_func(tau, b):
    return 1.0 / (1. - (Tr(tau) - 2.) / (b * b))


def phi_s(T, A_0):
    """solvent thinning 1 - A/(C + T), A = A_0/(k_b (T_h - T_0)), C = T_h/(T_h - T_0) with T_0 = 300, T_h = 350"""
    T_0, T_h, k_b = 300.0, 350.0, 8.3144598
    return 1. - (A_0 / (k_b * (T_h - T_0))) * (1. / (T_h / (T_h - T_0) + T))


def phi_ewm(tau, T, k, B):
    return (1. - B * (T / (1. + T))) * ((0.5 * Tr(tau) + 0.00001) ** k)


def I_3(A):
    return A[0, 0] * A[1, 1] - A[1, 0] * A[0, 1]


def I_2(A):
    return 0.5 * (A[0, 0] * A[0, 0] + A[1, 1] * A[1, 1] + 2 * A[1, 0] * A[0, 1])


def phi_def(u, lambda_d):
    D = Dincomp(u)
    D_u = alg(u).as_matrix([[D[0, 0], D[0, 1]], [D[1, 0], D[1, 1]]])
    return 1. + (lambda_d * 3 * I_3(D_u) / (I_2(D_u) + 0.0000001)) ** 2


def end():
    """dolfin's end() closes a begin()/end() log block (every sub-step of the loops calls it): nothing to close here"""


def dgp_sigma(u, p, Tau, We, Pr, betav):
    """buoyancy-driven-flow/fenics_fem.py:68-72: Prandtl-scaled incompressible stress; the polymer part is dropped for
    betav >= 1"""
    I = alg(u).Identity(len(u))
    s = 2 * Pr * betav * Dincomp(u) - p * I
    return s + Pr * ((1 - betav) / We) * (Tau - I) if betav < 1 else s


def dgp_sigma_n(U, p, tau, n, We, Pr, betav):
    """buoyancy-driven-flow/fenics_fem.py:74-78: traction sigma.n as the scripts write it"""
    t = Pr * betav * alg(U).nabla_grad(U) * n - p * n
    return t + (Pr * (1 - betav) / We) * tau * n if betav < 1 else t


def dgp_sigmacom(u, p, Tau, We, Pr, betav):
    """buoyancy-driven-flow/fenics_fem.py:80-81"""
    I = alg(u).Identity(len(u))
    return 2 * Pr * betav * Dcomp(u) - p * I + Pr * ((1 - betav) / We) * (Tau - I)


def reshape_elements(D0_vec, D12_vec, Ds_vec, D1_vec, tau0_vec, tau12_vec, tau1_vec):
    """the seven 3-vectors as symmetric 2x2 tensors [[a,b],[b,c]] (fenics_fem.py:210-225)."""
    A = alg(D0_vec)
    return tuple(A.as_matrix([[v[0], v[1]], [v[1], v[2]]])
                 for v in (D0_vec, D12_vec, Ds_vec, D1_vec, tau0_vec, tau12_vec, tau1_vec))


def magnitude(u):
    return (u[0] * u[0] + u[1] * u[1]) ** 0.5


# ---- spaces / functions ---------------------------------------------------------------------------------
def function_spaces(mesh, order):
    """(W, V, Vd, Z, Zd, Zc, Q, Qt, Qr) as the reference's function_spaces (fenics_fem.py:262-287); only the spaces the
    loops use are built (the others are None)."""
    from ..spaces import function_spaces as fs
    return fs(mesh, order)


def trial_functions(Q, Z, W):
    """(rho, p, T, tau_vec, u, D_vec, D, tau) -- fenics_fem.py:249-260, as symbolic arguments (fenics_b200/symbolic.py):
    forms built from them with the scripts' statements are recognised and assembled by the hand-written functors."""
    from ..dolfin_api import TrialFunction, TrialFunctions
    from ..symbolic import as_matrix
    rho, p, T, tau_vec = TrialFunction(Q), TrialFunction(Q), TrialFunction(Q), TrialFunction(Z)
    (u, D_vec) = TrialFunctions(W)
    D = as_matrix([[D_vec[0], D_vec[1]], [D_vec[1], D_vec[2]]])
    tau = as_matrix([[tau_vec[0], tau_vec[1]], [tau_vec[1], tau_vec[2]]])
    return rho, p, T, tau_vec, u, D_vec, D, tau


def make_solution_functions(plan):
    """solution_functions(Q, W, V, Z) bound to a device plan: the reference's 22-tuple (fenics_fem.py:227-247) of
    GPU Functions (sub-function / tensor views are the Functions themselves: kernels index components)."""
    def solution_functions(Q, W, V, Z):
        F = lambda S: Function(S, plan)  # noqa: E731
        rho0, rho12, rho1, p0, p1, T0, T1 = (F(Q) for _ in range(7))
        w0, w12, ws, w1 = (F(W) for _ in range(4))
        tau0_vec, tau12_vec, tau1_vec = (F(Z) for _ in range(3))
        # reference order: rho0, rho12, rho1, p0, p1, T0, T1, u0, u12, us, u1, D0_vec, D12_vec, Ds_vec, D1_vec,
        #                  w0, w12, ws, w1, tau0_vec, tau12_vec, tau1_vec
        return (rho0, rho12, rho1, p0, p1, T0, T1, w0, w12, ws, w1, w0, w12, ws, w1, w0, w12, ws, w1,
                tau0_vec, tau12_vec, tau1_vec)
    return solution_functions


def solution_functions(Q, W, V, Z):
    """needs the device plan of the run: every driver exposes `drv.plan`; without one, the spaces' own plan is
    built on first use."""
    from ..la import Plan
    plan = getattr(W, "_shim_plan", None)
    if plan is None:
        plan = Plan(W.mesh, [s for s in (W, Z, Q) if s is not None])
        W._shim_plan = plan
    return make_solution_functions(plan)(Q, W, V, Z)

