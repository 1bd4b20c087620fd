"""Config 4 ("flow between eccentrically rotating cylinders"/fenepmp_jbp.py) functors, run through the CPU
emulation of the CUDA element kernels, against the oracle's UFL-semantics evaluation.  1e-12 relative for
polynomial integrands; the FENE / SU forms are non-polynomial but both sides use the same quadrature
rule (the UFL-estimated degree, asserted below), so the same tolerance applies."""
import numpy as np
import pytest

import fxt_common as lc
import fxt_jbp as jb
from fenics_b200 import _abi as abi
from fxt_emu import EmuPlan
from oracle import fem
from oracle.configs_jbp import FenepmpJBP
from oracle.ufl_lite import dx, inner

TOL = 1e-12


def _setup(nt=16, nr=4, shuffle=False, **kw):
    mesh = fem.annulus_mesh(nt, nr)
    cfg = FenepmpJBP(mesh, **kw)
    if shuffle:
        rng = np.random.default_rng(5)
        dm = {k: rng.permutation(s.dim)[s.cell_dofs] for k, s in cfg.S.items()}
        cfg = FenepmpJBP(mesh, dofmaps=dm, **kw)
    jb.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local,
                   mesh.bfacet_marker)
    return mesh, cfg, plan


@pytest.mark.parametrize("shuffle", [False, True])
def test_cfg4_forms(shuffle):
    mesh, cfg, plan = _setup(shuffle=shuffle)
    assert set(np.unique(mesh.bfacet_marker)) == {0, 1}
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    for name, (fid, space) in jb.C4_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, par), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, par), ref)
        assert err < TOL, (name, err)
        degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)


def test_cfg4_su_density_matrix_with_flow():
    """a6's SU term is identically zero in the script (u12 is never solved for); exercise it with u12 != 0."""
    mesh, cfg, plan = _setup()
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    assert np.abs(cfg.w12.vec).max() > 0.1
    ref = fem.assemble(cfg.forms["a6"])
    got = plan.assemble_matrix(abi.FX_FORM_C4_A6, "Q", st, par)
    assert lc.csr_rel_err(got, ref) < TOL
    mass = fem.assemble((1.0 / cfg.p["dt"]) * fem.TrialFunction(cfg.S["Q"]) * fem.TestFunction(cfg.S["Q"]) * dx)
    assert abs(ref - mass).max() > 1e-3 * abs(mass).max()  # the SU part is really there


def test_cfg4_lps_and_functionals():
    mesh, cfg, plan = _setup()
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    cfg.lps_indicator()
    last = cfg.last
    res_test = plan.project_dg0(abi.FX_EXPR_C4_RESTAU, "Zd", st, par)
    assert lc.rel_err(res_test, last["res_test"]) < TOL
    st[abi.FX_F_RES_TEST] = res_test
    b = plan.assemble_vector(abi.FX_FORM_C4_L_RES_ORTH, "Zc", st, par)
    M = plan.assemble_matrix(abi.FX_FORM_MASS_ZC, "Zc", st, par)
    assert lc.rel_err(fem.solve_lu(M, b), last["res_orth"]) < 1e-10
    for fid, form in ((abi.FX_FORM_C4_EK, cfg.E_k_form), (abi.FX_FORM_C4_EE, cfg.E_e_form),
                      (abi.FX_FORM_C4_TORQUE, cfg.torque_form), (abi.FX_FORM_C4_FORCE_X, cfg.force_x_form),
                      (abi.FX_FORM_C4_FORCE_Y, cfg.force_y_form)):
        ref = fem.assemble(form)
        got = plan.assemble_scalar(fid, st, par)
        assert abs(got - ref) < 1e-11 * max(1.0, abs(ref)), (fid, got, ref)
        degs = {k: d for k, _, d in fem.quadrature_degrees(form, mesh)}
        q = plan.form_qdeg(fid)
        assert (q[0], q[1]) == (degs.get("dx", -1), degs.get("ds", -1)), (fid, q, degs)


def test_adaptive_refinement_indicator_exprs():
    """the three DG0 projections of the post-loop refinement indicator (fenepmp_jbp.py:691-705) against the
    oracle, with UFL's quadrature degree for inner(restau0, restau0)."""
    mesh, cfg, plan = _setup()
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    ref = cfg.refinement_indicator(err_count=0)
    qd = [d for k, _, d in fem.quadrature_degrees(cfg.res_sq_form * dx, mesh) if k == "dx"][0]
    assert qd == 6
    kapp = plan.project_dg0(abi.FX_EXPR_ADAPT_RES_SQ, "Qt", st, par)
    kapp = kapp / np.abs(kapp).max()
    assert lc.rel_err(kapp, ref["kapp"]) < TOL
    tau_avg = plan.project_dg0(abi.FX_EXPR_TAU_AVG, "Qt", st, par)
    st2, par2 = dict(st), dict(par)
    st2[abi.FX_F_KAPP], st2[abi.FX_F_QT_A] = kapp, tau_avg
    par2[abi.FX_P_EPS] = cfg.REFINE_EPS
    err = plan.project_dg0(abi.FX_EXPR_ABS_KAPP_RATIO, "Qt", st2, par2)
    assert lc.rel_err(err, ref["error_rat"]) < 1e-11
    assert ref["markers"].sum() in range(int(0.25 * len(kapp)), int(0.35 * len(kapp)) + 1)  # ratio 0.3 of the cells


# ---- FENE-P-MP proper: lambda_D != 0 (fenics_base.py:287-291; fenepmp_jbp.py:464, 519, 597, 615, 639) -------------
C4M_FORMS = dict(a4=(abi.FX_FORM_C4M_A4, "Zc"), L2=(abi.FX_FORM_C4M_L2, "W"), L9=(abi.FX_FORM_C4M_L9, "Q"))


@pytest.mark.parametrize("shuffle,lam", [(False, 0.1), (True, 0.15), (False, 0.05)])
def test_cfg4_forms_with_lambda_d(shuffle, lam):
    """the forms that carry phi_def(u, lambda_D): the predictor right-hand side differentiates it (P2 Hessian of u0),
    the stress matrix and the viscous heating use it pointwise; UFL degrees 13 / 13 / 14 asserted.  Every OTHER form
    of the loop is independent of lambda_D: checked against the lambda_D != 0 oracle with the lambda_D = 0 ids."""
    mesh, cfg, plan = _setup(shuffle=shuffle, lambda_d=lam)
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    assert par[abi.FX_P_LAMBDA_D] == lam
    cfg0 = FenepmpJBP(mesh, dofmaps={k: s.cell_dofs for k, s in cfg.S.items()}, lambda_d=0.0)
    for a in set(jb.FIELD_ATTR.values()):
        getattr(cfg0, a).vec[:] = getattr(cfg, a).vec
    want = dict(a4=13, L2=13, L9=14)
    for name, (fid, space) in jb.C4_FORMS.items():
        if name in C4M_FORMS:
            fid = C4M_FORMS[name][0]
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, par), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, par), ref)
        assert err < TOL, (name, err)
        degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)
        if name in want:
            assert q[0] == want[name]
            ref0 = fem.assemble(cfg0.forms[name])   # ... and lambda_D really changes the form
            a_, b_ = (ref.data, ref0.data) if name.startswith("a") else (ref, ref0)
            assert lc.rel_err(a_, b_) > (1e-6 if name != "L9" else 1e-11), name  # L9: phi_def sits behind Vh = 6.9e-6


def test_cfg4_lps_and_functionals_with_lambda_d():
    mesh, cfg, plan = _setup(lambda_d=0.1)
    st, par = jb.state_of(cfg), jb.params_of(cfg)
    cfg.lps_indicator()
    last = cfg.last
    res_test = plan.project_dg0(abi.FX_EXPR_C4M_RESTAU, "Zd", st, par)
    assert lc.rel_err(res_test, last["res_test"]) < TOL
    st[abi.FX_F_RES_TEST] = res_test
    b = plan.assemble_vector(abi.FX_FORM_C4M_L_RES_ORTH, "Zc", st, par)
    M = plan.assemble_matrix(abi.FX_FORM_MASS_ZC, "Zc", st, par)
    assert lc.rel_err(fem.solve_lu(M, b), last["res_orth"]) < 1e-10
    assert plan.form_qdeg(abi.FX_FORM_C4M_L_RES_ORTH)[0] == 13
    for fid, form in ((abi.FX_FORM_C4M_TORQUE, cfg.torque_form), (abi.FX_FORM_C4M_FORCE_X, cfg.force_x_form),
                      (abi.FX_FORM_C4M_FORCE_Y, cfg.force_y_form)):
        ref = fem.assemble(form)
        got = plan.assemble_scalar(fid, st, par)
        assert abs(got - ref) < 1e-11 * max(1.0, abs(ref)), (fid, got, ref)
        degs = {k: d for k, _, d in fem.quadrature_degrees(form, mesh)}
        q = plan.form_qdeg(fid)
        assert (q[0], q[1]) == (-1, degs["ds"]) and degs["ds"] == 12, (fid, q, degs)
