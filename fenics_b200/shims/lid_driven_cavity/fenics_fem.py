"""Drop-in for lid-driven-cavity/fenics_fem.py: same helper names and signatures, GPU-backed
where the name is on the hot path.  See fenics_b200/shims/__init__.py."""
from math import cos, fabs, sin, sqrt, tanh  # noqa: F401  (the reference re-exports them)

import numpy as np  # noqa: F401

from fenics_b200 import mesh as _mesh
from fenics_b200.dolfin_api import (DirichletBC, Function, KrylovSolver, TestFunction, TestFunctions,  # noqa: F401
                           TrialFunction, TrialFunctions, assemble, norm, parameters, project, solve)
from fenics_b200.symbolic import Constant, lhs, rhs, system  # noqa: F401
from fenics_b200.symbolic import (CellDiameter, FacetNormal, Identity, Parameter, as_matrix, as_vector, div, dot, ds,  # noqa: F401
                                  dx, grad, inner, nabla_grad, outer, sym, tr, transpose)
from fenics_b200.shims._common import (DGP_structured_mesh, DOLFIN_EPS, Dcomp, Dincomp, LDC_Regular_Mesh, Skew_Mesh,  # noqa: F401
                       _needs_dolfin, _npy, absolute, adaptive_refinement, alg, chunks, expskewcavity,
                       function_spaces, max_location, mesh2triang, min_location, mplot, normalize_solution, pi,
                       reshape_elements, set_algebra, skewcavity, solution_functions, tgrad, trial_functions,
                       update_progress, xskewcavity, yskewcavity)
from fenics_b200.shims._common import Fdef_compressible as Fdefcon  # noqa: F401
from fenics_b200.shims._common import ldc_sigma as sigma, ldc_sigmain as sigmain  # noqa: F401
from fenics_b200.shims._common import Fdef_incompressible as Fdef  # noqa: F401

def end():
    """dolfin's end() closes a begin()/end() log block (incomp_LDC.py:332): nothing to close here"""


refine_boundary = _needs_dolfin("refine_boundary")
refine_top = _needs_dolfin("refine_top")


def ramp_function(t):
    """lid ramp 1 + tanh(8(t - 1/2)) (:204-206)"""
    return 1.0 + tanh(8 * (t - 0.5))


def _driver_of(u):
    drv = getattr(u, "_driver", None)
    if drv is None:
        raise TypeError("stream_function(u): pass the velocity Function of a fenics_b200 time loop (drv.w1); the solve "
                        "runs on that run's device plan")
    return drv


def stream_function(u):
    """-lap(psi) = curl(u), psi = 0 on the boundary, on V.sub(0).collapse() (:340-358): CG on the device."""
    return _driver_of(u).stream_function(compressible=False)


def comp_stream_function(rho, u):
    """rho-weighted variant (:360-378)"""
    return _driver_of(u).stream_function(compressible=True)


def shear_rate(u):
    """the stream function's Poisson problem with inner(0.5*gamma, gamma), gamma = grad u + grad u^T, on the right
    (:380-398): CG on the device."""
    return _driver_of(u).shear_rate()


def l2norm_solution(u):
    """u / ||u||_L2 in place (:436-441); the L2 norm comes from the device mass matrix of u's space."""
    from fenics_b200 import _abi as abi
    S = u.function_space()
    fid = {"Zc": abi.FX_FORM_MASS_ZC, "Q": abi.FX_FORM_MASS_Q}.get(S.name)
    drv = getattr(u, "_driver", None)
    if fid is None or drv is None:
        raise NotImplementedError("l2norm_solution: mass matrices exist for the Zc and Q Functions of a time loop")
    M = assemble(drv.pr.form(fid))
    v = u.vector()
    v.t.mul_(1.0 / float(np.sqrt(v.inner(M * v))))
    return u


def save_energy_arrays(t, ek, ee, j, tag):
    """data/{time,ek,ee}-data-<j>-<tag>.npy (:456-459)"""
    for name, arr in (("time", t), ("ek", ek), ("ee", ee)):
        np.save(_npy("%s-data-%s-%s" % (name, j, tag)), arr)


def load_energy_arrays(j, tag):
    return tuple(np.load(_npy("%s-data-%s-%s" % (name, j, tag))) for name in ("time", "ek", "ee"))


def save_err_array(err, j, tag):
    np.save(_npy("err-data-%s-%s" % (j, tag)), err)


def load_err_array(j, tag):
    return np.load(_npy("err-data-%s-%s" % (j, tag)))
