"""Drop-in for "flow between eccentrically rotating cylinders"/fenics_base.py: same helper names and
signatures.  See fenics_b200/shims/__init__.py."""
from math import cos, fabs, sin, sqrt, tanh  # noqa: F401

import numpy as np  # noqa: F401

from fenics_b200 import mesh as _mesh
from fenics_b200.dolfin_api import (DirichletBC, Function, KrylovSolver, assemble, norm, parameters, project,  # noqa: F401
                           solve)
from fenics_b200.shims._common import end  # noqa: F401
from fenics_b200.shims._common import (FdefG, I_2, I_3, Tr, Trace, fene_func, fene_sigma, fene_sigmacom, phi_def, phi_ewm,  # noqa: F401
                                       phi_s)
from fenics_b200.shims._common import jbp_sigma as sigma, jbp_sigmacon as sigmacon  # noqa: F401
from fenics_b200.symbolic import Constant, lhs, rhs, system  # noqa: F401
from fenics_b200.symbolic import (CellDiameter, FacetNormal, Identity, Parameter, as_matrix, as_vector, div, dot, ds,  # noqa: F401
                                  dx, grad, inner, nabla_grad, outer, sym, tr, transpose)
from fenics_b200.dolfin_api import TestFunction, TestFunctions, TrialFunction, TrialFunctions  # noqa: F401
from fenics_b200.shims._common import (DOLFIN_EPS, Dcomp, Dincomp, _needs_dolfin, _T, absolute, adaptive_refinement, alg,  # noqa: F401
                       chunks, function_spaces, magnitude, max_location, mesh2triang, min_location, mplot,
                       normalize_solution, pi, reshape_elements, set_algebra, solution_functions, tgrad,
                       trial_functions, update_progress)
from fenics_b200.shims._common import Fdef_compressible as Fdef  # noqa: F401  (this module's Fdef carries + div(u) Tau, :189-190)
from fenics_b200.shims.lid_driven_cavity.fenics_fem import (comp_stream_function, l2norm_solution, ramp_function,  # noqa: F401
                                            shear_rate, stream_function)

refine_boundaries = _needs_dolfin("refine_boundaries")
refine_narrow = _needs_dolfin("refine_narrow")


def JBP_mesh(mesh_res, x_1, x_2, y_1, y_2, r_a, r_b):
    """the reference subtracts two mshr circles and calls generate_mesh (:70-91); mshr is not part of this
    library, so the structured O-grid annulus with ~mesh_res cells across the gap stands in (inner circle centre
    (x_1,y_1) radius r_a, outer (x_2,y_2) radius r_b)."""
    n_r = max(2, int(mesh_res) // 4)
    return _mesh.annulus_mesh(8 * n_r, n_r, x_1, y_1, r_a, x_2, y_2, r_b)


def save_solution(u, filename, writer=None):
    """the reference writes dolfin HDF5 (:94-99).  Same layout here for .h5 names (group /f: vector_0, cell_dofs,
    x_cell_dofs, cells; needs h5py -- without it a .h5 name raises), .npy otherwise; with `writer`
    (fenics_b200.io.AsyncWriter) the call returns at once and the device->host copy and the file write overlap the
    following time steps."""
    from fenics_b200.io import AsyncWriter
    if writer is not None:
        writer.save(u, filename)
        return
    with AsyncWriter(depth=1) as w:
        w.save(u, filename)


def load_solution(filename, function_space, plan=None):
    """inverse of save_solution (:102-108); `plan` = the device plan the new Function lives on (default: the one
    solution_functions uses for that space)."""
    from fenics_b200.io import load_array
    vals = load_array(filename)
    if plan is None:
        plan = getattr(function_space, "_shim_plan", None)
    if plan is None:
        raise TypeError("load_solution: pass plan= (the device plan of the run)")
    u = Function(function_space, plan)
    u.vector().set_local(vals)
    return u


# ---- symbolic tensor helpers (:158-294) -------------------------------------------------------------------
def DinG(D):
    return 0.5 * (D + _T(D))


def euc_norm(tau):
    return (tau[0] * tau[0] + tau[1] * tau[1] + tau[2] * tau[2]) ** 0.5


def phi_ewm_inv(tau, T, k, B):
    return 1.0 / phi_ewm(tau, T, k, B)


def phi(u, p, T, A, B, K_0, N, betap):
    """pressure / temperature / shear dependent viscosity function (:215-225); its inner project() needs the
    caller's space Q, as in the reference (a module-level name there)."""
    A_ = alg()
    K = 1. + A * p - B * (T / (1. + T))
    strain = (2 * A_.inner(Dincomp(u), Dincomp(u))) ** 0.5
    ftheta = project(1. / (1. + (K_0 * K * strain) ** N), globals().get("Q"))
    return (betap + (1. - betap) * ftheta) * K


def orthog_proj(u, V, V_d):
    u_d = project(u, V_d)
    return project(u - u_d, V)


def frob_norm(tau):
    return Trace(tau * _T(tau)) ** 0.5


def ernesto_k(u, h, c_a, c_b):
    return c_a * h * magnitude(u) + c_b * h * h * frob_norm(alg().grad(u))


def psi_def(phi):
    return 0.5 * (phi - 1.)
