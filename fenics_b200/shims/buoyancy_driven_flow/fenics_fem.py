"""Drop-in for buoyancy-driven-flow/fenics_fem.py: same helper names and signatures.
See fenics_b200/shims/__init__.py."""
from math import cos, fabs, sin, sqrt, tanh  # noqa: F401

import numpy as np  # noqa: F401

from fenics_b200.dolfin_api import (DirichletBC, Function, KrylovSolver, assemble, norm, parameters, project,  # noqa: F401
                           solve)
from fenics_b200.drivers.dgp import DGP_Mesh  # noqa: F401
from fenics_b200.shims._common import (DGP_structured_mesh, DOLFIN_EPS, Dcomp, Dincomp, _needs_dolfin, _npy, absolute,  # noqa: F401
                       adaptive_refinement, alg, chunks, expskewcavity, function_spaces, magnitude, max_location,
                       mesh2triang, min_location, mplot, normalize_solution, pi, reshape_elements, set_algebra,
                       skewcavity, solution_functions, tgrad, trial_functions, update_progress, xskewcavity)
from fenics_b200.shims._common import Fdef_compressible as Fdefcom  # noqa: F401
from fenics_b200.shims._common import dgp_sigma as sigma, dgp_sigma_n as sigma_n, dgp_sigmacom as sigmacom  # noqa: F401
from fenics_b200.symbolic import Constant, lhs, rhs, system  # noqa: F401
from fenics_b200.symbolic import (CellDiameter, FacetNormal, Identity, Parameter, as_matrix, as_vector, div, dot, ds,  # noqa: F401
                                  dx, grad, inner, nabla_grad, outer, sym, tr, transpose)
from fenics_b200.dolfin_api import TestFunction, TestFunctions, TrialFunction, TrialFunctions  # noqa: F401
from fenics_b200.shims._common import Fdef_incompressible as Fdef  # noqa: F401
from fenics_b200.shims._common import end  # noqa: F401
from fenics_b200.shims.lid_driven_cavity.fenics_fem import (comp_stream_function, l2norm_solution, ramp_function,  # noqa: F401
                                            shear_rate, stream_function)

DGP_unstructured_mesh = _needs_dolfin("DGP_unstructured_mesh")
refine_boundary = _needs_dolfin("refine_boundary")
refine_top = _needs_dolfin("refine_top")
refine_walls = _needs_dolfin("refine_walls")


def _tag(loop, subloop, tag):
    return "%s-%s-%s" % (loop, subloop, tag)


def save_energy_arrays(t, ek, ee, loop, subloop, tag):
    """data/{time,ek,ee}-data-<loop>-<subloop>-<tag>.npy (:470-473)"""
    for name, arr in (("time", t), ("ek", ek), ("ee", ee)):
        np.save(_npy("%s-data-%s" % (name, _tag(loop, subloop, tag))), arr)


def load_energy_arrays(loop, subloop, tag):
    return tuple(np.load(_npy("%s-data-%s" % (name, _tag(loop, subloop, tag)))) for name in ("time", "ek", "ee"))


def save_data_array(nus, loop, subloop, tag):
    np.save(_npy("nus-data-" + _tag(loop, subloop, tag)), nus)


def load_data_array(loop, subloop, tag):
    return np.load(_npy("nus-data-" + _tag(loop, subloop, tag)))
