"""GPU: FX_PC_ILU0 (csrc/fx_ilu.cu) -- `solve(A, x, b, "bicgstab", "default")` as a serial PETSc process runs it
(lid-driven-cavity/incomp_LDC.py:331,363,375,394,416): BiCGStab / CG with natural-ordering ILU(0), left preconditioning.
Checked against the oracle's C restatement (oracle/krylov_c.c) on the same matrices: ||M^-1 b|| (one factorisation and
one pair of triangular solves, summation order identical) to 1e-13, iteration counts within +-2 (only the Krylov dot
products are reduced in a different order), solutions against a direct solve."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

import fxt_common as lc
from fenics_b200 import _abi as abi
from oracle import fem

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from fenics_b200 import dolfin_api as dapi  # noqa: E402
from fenics_b200 import la, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402


def _setup(n=10):
    """config 2 on a small skewed mesh, two time steps in: every field non-trivial, matrices as the loop sees them"""
    drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
    drv.t = 0.45
    drv.step()
    drv.step()
    return drv


@pytest.mark.parametrize("method", ["bicgstab", "cg"])
def test_ilu0_krylov_matches_the_petsc_like_oracle(method):
    drv = _setup()
    rng = np.random.default_rng(5)
    cases = [("Q", abi.FX_FORM_C2_A5, False)]
    if method == "bicgstab":
        cases += [("Zc", abi.FX_FORM_C2_A4, False), ("W", abi.FX_FORM_C2_A1, True), ("W", abi.FX_FORM_C2_A7, True)]
    for name, fid, with_bc in cases:
        A = dapi.assemble(drv.pr.form(fid))
        b = la.Vector(drv.plan, name)
        b.set_local(rng.standard_normal(b.size()))
        if with_bc:
            for bc in drv.bcu:
                bc.apply(A, b)
        As, bv = A.to_scipy(), b.get_local()
        x0 = 0.1 * rng.standard_normal(b.size())             # nonzero initial guess (incomp_LDC.py:267)
        for rtol in (1e-5, 1e-10):
            xo = x0.copy()
            it_o, rn_o, bn_o = fem.solve_krylov_c(As, xo, bv, method, "ilu", rtol=rtol)
            x = la.Vector(drv.plan, name)
            x.set_local(x0)
            info = la.krylov_solve(A, x, b, method, "ilu", rtol=rtol)
            assert info.converged == 1, (name, fid, rtol, info.iterations, info.residual)
            assert abs(info.residual0 - bn_o) <= 1e-13 * bn_o, (name, info.residual0, bn_o)
            assert abs(info.iterations - it_o) <= 2, (name, rtol, info.iterations, it_o)
            assert info.residual <= rtol * info.residual0
            assert lc.rel_err(x.get_local(), xo) < 50 * rtol, (name, rtol)
        ref = spla.splu(As.tocsc()).solve(bv)
        assert lc.rel_err(x.get_local(), ref) < 1e-6, name
        # Jacobi needs several times the iterations (the reason the reference's default matters)
        xj = la.Vector(drv.plan, name)
        xj.set_local(x0)
        ij = la.krylov_solve(A, xj, b, method, "jacobi", rtol=1e-10, maxit=20000)
        assert ij.iterations > info.iterations, (name, ij.iterations, info.iterations)


def test_default_preconditioner_parameter_and_edge_cases():
    import warnings
    drv = _setup(6)
    A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A5))
    b, x = la.Vector(drv.plan, "Q"), la.Vector(drv.plan, "Q")
    info = la.krylov_solve(A, x, b, "bicgstab", "ilu", rtol=1e-10)           # b = 0, x0 = 0: nothing to do
    assert info.converged == 1 and info.iterations == 0 and float(x.t.abs().max()) == 0.0
    b.set_local(np.random.default_rng(0).standard_normal(b.size()))
    with pytest.raises(RuntimeError):                                        # maxit = 0 on a non-trivial system
        la.krylov_solve(A, x, b, "bicgstab", "ilu", rtol=1e-10, maxit=0)
    want = la.krylov_solve(A, x, b, "bicgstab", "ilu").iterations
    la.parameters["default_preconditioner"] = "ilu"
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("error", la.PreconditionerSubstitution)    # "default" IS ILU(0) now: no substitution
            x.zero()
            assert abs(la.krylov_solve(A, x, b, "bicgstab", "default").iterations - want) <= 1   # (atomic dot products)
    finally:
        la.parameters["default_preconditioner"] = "jacobi"
    # a matrix with a structurally present but zero pivot is reported (status -2), not looped on
    A.val.zero_()
    with pytest.raises(RuntimeError):
        la.krylov_solve(A, x, b, "cg", "ilu")
    nan = la.Vector(drv.plan, "Q")
    nan.set_local(np.full(b.size(), np.nan))
    A2 = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A5))
    with pytest.raises(RuntimeError):                                        # NaN right-hand side: reported
        la.krylov_solve(A2, x, nan, "bicgstab", "ilu", maxit=50)
