"""Config 5 (buoyancy-driven-flow/compressible_dgp.py) functors through the CPU emulation of the CUDA element
kernels, against the oracle (1e-12 relative; same UFL-estimated quadrature degree asserted)."""
import numpy as np
import pytest

import fxt_common as lc
import fxt_dgp as dg
from fenics_b200 import _abi as abi
from fxt_emu import EmuPlan
from oracle import fem
from oracle.configs_dgp import CompDGP, dgp_mesh
from oracle.ufl_lite import dx, inner

TOL = 1e-12


def _setup(n=6, shuffle=False):
    mesh = dgp_mesh(n)
    cfg = CompDGP(mesh)
    if shuffle:
        rng = np.random.default_rng(7)
        cfg = CompDGP(mesh, dofmaps={k: rng.permutation(s.dim)[s.cell_dofs] for k, s in cfg.S.items()})
    dg.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local,
                   mesh.bfacet_marker)
    return mesh, cfg, plan


@pytest.mark.parametrize("shuffle", [False, True])
def test_cfg5_forms(shuffle):
    mesh, cfg, plan = _setup(shuffle=shuffle)
    assert set(np.unique(mesh.bfacet_marker)) == {1, 2, 3, 4}
    st, par = dg.state_of(cfg), dg.params_of(cfg)
    for name, (fid, space) in dg.C5_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, par), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, par), ref)
        assert err < TOL, (name, err)
        degs = {}
        for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh):
            degs[k] = max(degs.get(k, -1), d)
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)


def test_cfg5_projections_and_functionals():
    mesh, cfg, plan = _setup()
    st, par = dg.state_of(cfg), dg.params_of(cfg)
    Q = cfg.S["Q"]
    w = fem.TestFunction(Q)
    for fid, expr in ((abi.FX_FORM_C5_L_THETA0, cfg.theta0_expr), (abi.FX_FORM_C5_L_THETA1, cfg.theta1_expr),
                      (abi.FX_FORM_C5_L_THERM_INV, cfg.therm_inv_expr), (abi.FX_FORM_C5_L_RHO1, cfg.rho1_expr)):
        L = inner(w, expr) * dx
        assert lc.rel_err(plan.assemble_vector(fid, "Q", st, par), fem.assemble(L)) < TOL, fid
        assert plan.form_qdeg(fid)[0] == fem.quadrature_degrees(L, mesh)[0][2], fid
    cfg.lps_indicator()
    res_test = plan.project_dg0(abi.FX_EXPR_C5_RESTAU, "Zd", st, par)
    assert lc.rel_err(res_test, cfg.last["res_test"]) < TOL
    st[abi.FX_F_RES_TEST] = res_test
    b = plan.assemble_vector(abi.FX_FORM_C5_L_RES_ORTH, "Zc", st, par)
    M = plan.assemble_matrix(abi.FX_FORM_MASS_ZC, "Zc", st, par)
    assert lc.rel_err(fem.solve_lu(M, b), cfg.last["res_orth"]) < 1e-10
    for fid, form in ((abi.FX_FORM_C5_EK, cfg.E_k_form), (abi.FX_FORM_C5_EE, cfg.E_e_form),
                      (abi.FX_FORM_C5_NUS, cfg.nus_form)):
        ref = fem.assemble(form)
        assert abs(plan.assemble_scalar(fid, st, par) - ref) < 1e-11 * max(1.0, abs(ref)), fid
