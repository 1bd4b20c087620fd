"""The dolfin A/B harness of SURVEY.md 8c -- runs only where legacy FEniCS (dolfin 2019.1.0) is importable; skipped
otherwise (this image and the GPU box have no dolfin: the oracle is 'parity unpinned').

What it pins the day dolfin is present, on the reference's own first mesh (the reference's lid-driven-cavity helper
module builds the spaces):
  1. mesh ingestion: fenics_b200.mesh.from_dolfin + the dofmaps read from dolfin (V.dofmap().cell_dofs);
  2. sparsity: the oracle's / plan's CSR pattern == dolfin's assembled PETSc matrix pattern, bit-exact;
  3. element conventions (FIAT dof order, FFC quadrature): mass matrices of W, Zc, Q and the config-1 forms a1, L1
     assembled by the oracle on dolfin's numbering == dolfin's, 1e-12;
  4. K steps of config 1 with exact solves: fields within 1e-8.
Set FENICS_REFERENCE to the checkout of ATMackay/fenics (default /root/reference).
"""
import os
import sys

import numpy as np
import pytest

dolfin = pytest.importorskip("dolfin", reason="legacy FEniCS (dolfin) is not installed")

import fxt_common as lc  # noqa: E402
from oracle import configs, fem  # noqa: E402
from oracle.ufl_lite import dx, inner  # noqa: E402

REF = os.environ.get("FENICS_REFERENCE", "/root/reference")
LDC = os.path.join(REF, "lid-driven-cavity")


def _reference_spaces(mm=10):
    if LDC not in sys.path:
        sys.path.insert(0, LDC)
    import fenics_fem as ref  # the reference's own helper module
    mesh = ref.LDC_Mesh(mm) if hasattr(ref, "LDC_Mesh") else dolfin.UnitSquareMesh(mm, mm)
    W, V, _, Z, Zd, Zc, Q, Qt, Qr = ref.function_spaces(mesh, 2)
    return ref, mesh, dict(W=W, V=V, Zd=Zd, Zc=Zc, Q=Q, Qt=Qt)


def _cell_dofs(V, mesh):
    return np.array([V.dofmap().cell_dofs(c) for c in range(mesh.num_cells())], dtype=np.int32)


def _csr(A):
    mat = dolfin.as_backend_type(A).mat()
    ia, ja, va = mat.getValuesCSR()
    order = np.concatenate([np.argsort(ja[ia[r]:ia[r + 1]], kind="stable") + ia[r] for r in range(len(ia) - 1)])
    return ia, ja[order], va[order]


def test_mesh_dofmaps_and_sparsity_are_dolfins():
    ref, dmesh, S = _reference_spaces()
    om = fem.Mesh(dmesh.coordinates(), dmesh.cells())
    dm = {k: _cell_dofs(v, dmesh) for k, v in S.items()}
    sp = configs.make_spaces(om, dm)
    for name in ("W", "Zc", "Q"):
        u, v = dolfin.TrialFunction(S[name]), dolfin.TestFunction(S[name])
        ia, ja, va = _csr(dolfin.assemble(dolfin.inner(u, v) * dolfin.dx))
        A = fem.assemble(inner(fem.TrialFunction(sp[name]), fem.TestFunction(sp[name])) * dx)
        assert np.array_equal(A.indptr, ia) and np.array_equal(A.indices, ja), name   # bit-exact pattern
        assert lc.rel_err(A.data, va) < 1e-12, name                                   # FIAT order + FFC quadrature


def test_k_steps_of_config1_match_dolfin():
    """oracle (exact solves) against the reference's own time loop pieces run through dolfin with LU solves"""
    ref, dmesh, S = _reference_spaces()
    om = fem.Mesh(dmesh.coordinates(), dmesh.cells())
    dm = {k: _cell_dofs(v, dmesh) for k, v in S.items()}
    ocl = configs.IncompLDC(om, dofmaps=dm)
    # a1 / L1 of incomp_LDC.py:317-325 written with the reference's helpers on dolfin Functions holding the oracle's state
    lc.randomize(ocl)
    w0 = dolfin.Function(S["W"])
    w0.vector().set_local(ocl.w0.vec)
    A_o = fem.assemble(ocl.forms["a7"])
    (u, D) = dolfin.TrialFunctions(S["W"])
    (v, R) = dolfin.TestFunctions(S["W"])
    Dv = dolfin.as_matrix([[D[0], D[1]], [D[1], D[2]]])
    Rv = dolfin.as_matrix([[R[0], R[1]], [R[1], R[2]]])
    p = ocl.p
    a7 = dolfin.inner((p["Re"] / p["dt"]) * u, v) * dolfin.dx + dolfin.inner(0.5 * p["betav"] * dolfin.grad(u), dolfin.grad(v)) * dolfin.dx \
        + dolfin.inner(Dv - ref.Dincomp(u), Rv) * dolfin.dx
    ia, ja, va = _csr(dolfin.assemble(a7))
    assert np.array_equal(A_o.indptr, ia) and np.array_equal(A_o.indices, ja)
    assert lc.rel_err(A_o.data, va) < 1e-12
