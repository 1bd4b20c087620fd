"""GPU: edge cases through the C ABI -- smallest meshes, degenerate solves, argument errors (return codes and
messages instead of crashes), a mesh whose cells are not consecutive in memory order."""
import ctypes as C

import numpy as np
import pytest
import torch

import fxt_common as lc
from fenics_b200 import _abi as abi
from oracle import configs, fem

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from fenics_b200 import dolfin_api as dapi  # noqa: E402
from fenics_b200 import la, mesh as pmesh  # noqa: E402
from fenics_b200._lib import FxError, SolveInfo, lib  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402


def _pair(mesh):
    drv = ldc.CompLDC(mesh)
    ocl = configs.CompLDC(fem.Mesh(mesh.coordinates(), mesh.cells()), dofmaps={k: s.cell_dofs for k, s in drv.S.items()})
    return ocl, drv


@pytest.mark.parametrize("n", [1, 2])
def test_smallest_meshes_match_oracle(n):
    """2 and 8 cells: every dof is a boundary dof or nearly so, row blocks are shorter than a warp."""
    ocl, drv = _pair(pmesh.Skew_Mesh(n, 1, 1))
    lc.randomize(ocl)
    for slot, attr in lc.FIELD_ATTR.items():
        fo, fd = getattr(ocl, attr, None), getattr(drv, attr, None)
        if fo is not None and fd is not None:
            fd.vector().set_local(fo.vec)
    for name, (fid, space) in lc.C2_FORMS.items():
        ref = fem.assemble(ocl.forms[name])
        got = dapi.assemble(drv.pr.form(fid))
        err = lc.csr_rel_err(got.to_scipy(), ref) if name.startswith("a") else lc.rel_err(got.get_local(), ref)
        assert err < 1e-12, (n, name, err)
    rp, col = drv.plan.pattern_host("W")
    A = fem.assemble(ocl.forms["a1"])
    assert np.array_equal(rp, A.indptr) and np.array_equal(col, A.indices)
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    drv.t = ocl.t = 0.45
    ro, rd = ocl.step(), drv.step()
    assert abs(ro["E_k"] - rd["E_k"]) <= 1e-8 * abs(ro["E_k"]) + 1e-300


def test_degenerate_solves():
    _, drv = _pair(pmesh.Skew_Mesh(6, 1, 1))
    A = dapi.assemble(drv.pr.form(abi.FX_FORM_MASS_Q))
    b, x = la.Vector(drv.plan, "Q"), la.Vector(drv.plan, "Q")
    for method in ("cg", "bicgstab", "gmres"):
        x.zero()
        info = la.krylov_solve(A, x, b, method, "jacobi", rtol=1e-10)       # b = 0, x0 = 0: nothing to do
        assert info.converged == 1 and info.iterations == 0 and float(x.t.abs().max()) == 0.0
        b.set_local(np.ones(b.size()))
        with pytest.raises(RuntimeError):                                    # maxit = 0 on a non-trivial system
            la.krylov_solve(A, x, b, method, "jacobi", rtol=1e-10, maxit=0)
        b.zero()
    b.set_local(np.full(b.size(), np.nan))
    with pytest.raises(RuntimeError):                                        # NaN right-hand side: reported, not looped on
        la.krylov_solve(A, x, b, "bicgstab", "jacobi", rtol=1e-10, maxit=50)


def test_argument_errors_return_codes():
    _, drv = _pair(pmesh.Skew_Mesh(4, 1, 1))
    h = drv.plan.h
    assert lib.fx_krylov_solve(h, 99, None, None, None, 0, 1, 1e-5, 1e-50, 10, None, None) == abi.FX_ERR_ARG
    A = dapi.assemble(drv.pr.form(abi.FX_FORM_MASS_Q))
    b, x = la.Vector(drv.plan, "Q"), la.Vector(drv.plan, "Q")
    info = SolveInfo()
    args = (h, abi.SPACE_ID["Q"], la._ptr(A.val), la._ptr(x.t), la._ptr(b.t))
    assert lib.fx_krylov_solve(*args, 7, 1, 1e-5, 1e-50, 10, C.byref(info), la._stream()) == abi.FX_ERR_UNSUPPORTED
    assert lib.fx_krylov_solve(*args, abi.FX_KSP_GMRES, abi.FX_PC_ILU0, 1e-5, 1e-50, 10, C.byref(info), la._stream()) \
        == abi.FX_ERR_UNSUPPORTED and b"not implemented" in lib.fx_last_error()
    assert lib.fx_assemble_matrix(h, 424242, drv.pr.state_ptr(), drv.pr.params_ptr(), la._ptr(A.val), la._stream()) \
        == abi.FX_ERR_ARG
    assert lib.fx_plan_set_peer_regions(h, 0, 9, None, 10) == abi.FX_ERR_ARG         # world > 8
    assert lib.fx_plan_set_halo(h, abi.SPACE_ID["Qt"], 0, None, None, None, None) == abi.FX_ERR_ARG  # no CSR pattern
    with pytest.raises(FxError):
        drv.plan.set_eliminated_dofs("W", [0, 0])                                       # repeated row
    # a mesh whose cells violate the UFC vertex order is rejected, not mis-assembled
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    cells = np.array([[2, 0, 1]], dtype=np.int32)
    dm = (C.c_void_p * abi.NSPACES)()
    out = C.c_void_p()
    z = np.zeros(1, dtype=np.int32)
    rc = lib.fx_plan_create(xy.ctypes.data, 3, cells.ctypes.data, 1, dm, z.ctypes.data, z.ctypes.data, z.ctypes.data, 0, 0,
                            C.byref(out))
    assert rc == abi.FX_ERR_ARG and b"UFC" in lib.fx_last_error()


def test_shuffled_cell_order_gives_the_same_matrices():
    """dolfin's cell numbering is arbitrary: permuting the cells (and nothing else) changes the order of the atomic
    scatter only."""
    m = pmesh.Skew_Mesh(6, 1, 1)
    perm = np.random.default_rng(7).permutation(m.num_cells())
    m2 = pmesh.Mesh(m.coordinates(), m.cells()[perm])
    ocl, drv = _pair(m2)
    lc.randomize(ocl)
    for slot, attr in lc.FIELD_ATTR.items():
        fo, fd = getattr(ocl, attr, None), getattr(drv, attr, None)
        if fo is not None and fd is not None:
            fd.vector().set_local(fo.vec)
    for name in ("a1", "a4", "L2", "L4"):
        fid, space = lc.C2_FORMS[name]
        ref = fem.assemble(ocl.forms[name])
        got = dapi.assemble(drv.pr.form(fid))
        err = lc.csr_rel_err(got.to_scipy(), ref) if name.startswith("a") else lc.rel_err(got.get_local(), ref)
        assert err < 1e-12, (name, err)


def test_vector_setitem_honours_the_key_and_pc_names_are_substituted_loudly():
    import warnings
    from fenics_b200 import dolfin_api as dapi, la, mesh as pmesh
    from fenics_b200 import _abi as abi
    from fenics_b200.drivers import ldc
    drv = ldc.CompLDC(pmesh.Skew_Mesh(6, 1, 1))
    v = la.Vector(drv.plan, "Q")
    v[:] = np.arange(v.size(), dtype=float)
    v[3] = -1.0
    v[np.array([0, 5])] = [7.0, 8.0]
    ref = np.arange(v.size(), dtype=float)
    ref[3], ref[0], ref[5] = -1.0, 7.0, 8.0
    assert np.array_equal(v.get_local(), ref)
    v[:] = 2.0
    assert np.all(v.get_local() == 2.0)
    # "default" (PETSc ILU(0)) and "amg" are served by named substitutes, with a warning the first time
    A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A5))
    b = la.Vector(drv.plan, "Q")
    b.set_local(np.random.default_rng(0).standard_normal(b.size()))
    la._warned.clear()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        la.krylov_solve(A, la.Vector(drv.plan, "Q"), b, "bicgstab", "default")
        la.krylov_solve(A, la.Vector(drv.plan, "Q"), b, "cg", "amg")
        la.krylov_solve(A, la.Vector(drv.plan, "Q"), b, "cg", "amg")
    msgs = [str(x.message) for x in w if issubclass(x.category, la.PreconditionerSubstitution)]
    assert len(msgs) == 2 and "'default' is served by 'jacobi'" in msgs[0] and "'amg' is served by 'jacobi'" in msgs[1]
    la.parameters["preconditioner_substitution"] = "error"
    try:
        with pytest.raises(ValueError):
            la.krylov_solve(A, la.Vector(drv.plan, "Q"), b, "bicgstab", "default")
    finally:
        la.parameters["preconditioner_substitution"] = "warn"
    with pytest.raises(ValueError):
        la.krylov_solve(A, la.Vector(drv.plan, "Q"), b, "bicgstab", "nonsense")


def test_deterministic_assembly_is_bitwise_reproducible_and_equals_the_atomic_sum():
    """FX_ASM_DETERMINISTIC (atomic-free row gather): the same matrix / vector bits on every run, equal to the atomic
    scatter-add to rounding (1e-13), facet integrals included (cfg2 L1/L2 carry ds terms; cfg5 has ds(i) matrices)."""
    from fenics_b200 import dolfin_api as dapi, mesh as pmesh
    from fenics_b200 import _abi as abi
    from fenics_b200.drivers import dgp, ldc
    import fxt_common as lc
    for drv, forms in ((ldc.CompLDC(pmesh.Skew_Mesh(14, 1, 1)),
                        (abi.FX_FORM_C2_A1, abi.FX_FORM_C2_A4, abi.FX_FORM_C2_A5, abi.FX_FORM_C2_L1, abi.FX_FORM_C2_L4,
                         abi.FX_FORM_C2_L5)),
                       (dgp.IncompDGP(mesh_resolution=12),
                        (abi.FX_FORM_C6_A1, abi.FX_FORM_C6_A8, abi.FX_FORM_C6_L1, abi.FX_FORM_C6_L8))):
        rng = np.random.default_rng(4)
        for fn in drv.pr.fields.values():
            v = fn.vector()
            v.set_local(rng.standard_normal(v.size()) * 0.3 + 1.0)
        for fid in forms:
            drv.plan.set_assembly_mode("atomic")
            ref = dapi.assemble(drv.pr.form(fid))
            ref = (ref.val if hasattr(ref, "val") else ref.t).clone()
            drv.plan.set_assembly_mode("deterministic")
            runs = []
            for _ in range(3):
                got = dapi.assemble(drv.pr.form(fid))
                runs.append((got.val if hasattr(got, "val") else got.t).clone())
            assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2]), fid
            assert lc.rel_err(runs[0].cpu().numpy(), ref.cpu().numpy()) < 1e-13, fid
        drv.plan.set_assembly_mode("atomic")
    # a whole time step in deterministic mode stays on the oracle-checked trajectory
    d1, d2 = ldc.CompLDC(pmesh.Skew_Mesh(10, 1, 1)), ldc.CompLDC(pmesh.Skew_Mesh(10, 1, 1))
    d2.plan.set_assembly_mode("deterministic")
    for d in (d1, d2):
        d.t = 0.45
        d.solve_rtol = 1e-13
        for _ in range(2):
            d.step()
    f1, f2 = d1.fields(), d2.fields()
    for k in f1:
        assert lc.rel_err(f2[k], f1[k]) < 1e-9, k
