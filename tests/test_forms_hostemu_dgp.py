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


# ---- incompressible_dgp.py ------------------------------------------------------------------------------------------
def _setup6(n=6, shuffle=False, **kw):
    from oracle.configs_dgp import IncompDGP
    mesh = dgp_mesh(n)
    cfg = IncompDGP(mesh, **kw)
    if shuffle:
        rng = np.random.default_rng(11)
        cfg = IncompDGP(mesh, dofmaps={k: rng.permutation(s.dim)[s.cell_dofs] for k, s in cfg.S.items()}, **kw)
    dg.c6_randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker)
    return mesh, cfg, plan


@pytest.mark.parametrize("shuffle,kw", [(False, {}), (True, dict(Ra=700.0, We=0.4, Pr=1.7, betav=0.3, th=0.8, Vh=0.02, Bi=0.3))])
def test_incompressible_dgp_forms(shuffle, kw):
    mesh, cfg, plan = _setup6(shuffle=shuffle, **kw)
    assert set(np.unique(mesh.bfacet_marker)) == {1, 2, 3, 4}
    st, par = dg.c6_state_of(cfg), dg.c6_params_of(cfg)
    for name, (fid, space) in dg.C6_FORMS.items():
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


def test_incompressible_dgp_projections_and_functionals():
    mesh, cfg, plan = _setup6(We=0.4, Bi=0.3)
    st, par = dg.c6_state_of(cfg), dg.c6_params_of(cfg)
    w = fem.TestFunction(cfg.S["Q"])
    L = inner(w, cfg.theta1_expr) * dx
    assert lc.rel_err(plan.assemble_vector(abi.FX_FORM_C5_L_THETA1, "Q", st, par), fem.assemble(L)) < TOL
    cfg.lps_indicator()
    res_test = plan.project_dg0(abi.FX_EXPR_C6_RESTAU, "Zd", st, par)
    assert lc.rel_err(res_test, cfg.last["res_test"]) < TOL
    st[abi.FX_F_RES_TEST] = res_test
    b = plan.assemble_vector(abi.FX_FORM_C6_L_RES_ORTH, "Zc", st, par)
    M = plan.assemble_matrix(abi.FX_FORM_MASS_ZC, "Zc", st, par)
    assert lc.rel_err(fem.solve_lu(M, b), cfg.last["res_orth"]) < 1e-10
    wz = fem.TestFunction(cfg.S["Zc"])
    Lr = inner(wz, cfg.restau0 - fem.Function(cfg.S["Zd"], cfg.last["res_test"]).expr()) * dx
    assert plan.form_qdeg(abi.FX_FORM_C6_L_RES_ORTH)[0] == fem.quadrature_degrees(Lr, mesh)[0][2]
    for fid, form in ((abi.FX_FORM_C6_EK, cfg.E_k_form), (abi.FX_FORM_C6_EE, cfg.E_e_form), (abi.FX_FORM_C6_NUS, cfg.nus_form)):
        ref = fem.assemble(form)
        assert abs(plan.assemble_scalar(fid, st, par) - ref) < 1e-11 * max(1.0, abs(ref)), fid
        degs = {k: d for k, _, d in fem.quadrature_degrees(form, mesh)}
        q = plan.form_qdeg(fid)
        assert (q[0], q[1]) == (degs.get("dx", -1), degs.get("ds", -1)), (fid, q, degs)
