"""Config 3 ("flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py) functors, run
through the CPU emulation of the CUDA element kernels, against the oracle's UFL-semantics evaluation, at
the script's own parameters and at a general set (fxt_ewm.GENERAL).  Both sides use the UFL-estimated
quadrature degree (asserted), so 1e-12 relative applies to the non-polynomial SU forms too."""
import numpy as np
import pytest

import fxt_common as lc
import fxt_ewm as ew
from fenics_b200 import _abi as abi
from fxt_emu import EmuPlan
from oracle import fem
from oracle.configs_ewm import EwmJBP

TOL = 1e-12


def _setup(nt=16, nr=4, shuffle=False, **kw):
    mesh = fem.annulus_mesh(nt, nr, x_1=-0.9)
    cfg = EwmJBP(mesh, **kw)
    if shuffle:
        rng = np.random.default_rng(5)
        dm = {k: rng.permutation(s.dim)[s.cell_dofs] for k, s in cfg.S.items()}
        cfg = EwmJBP(mesh, dofmaps=dm, **kw)
    ew.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local,
                   mesh.bfacet_marker)
    return mesh, cfg, plan


@pytest.mark.parametrize("kw,shuffle", [({}, False), (ew.GENERAL, False), (ew.GENERAL, True)])
def test_cfg3_forms(kw, shuffle):
    mesh, cfg, plan = _setup(shuffle=shuffle, **kw)
    st, par = ew.state_of(cfg), ew.params_of(cfg)
    for name, (fid, space) in ew.C3_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, par), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, par), ref)
        assert err < TOL, (name, err)
        degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)


def test_cfg3_functionals():
    mesh, cfg, plan = _setup(**ew.GENERAL)
    st, par = ew.state_of(cfg), ew.params_of(cfg)
    for fid, form in ((abi.FX_FORM_C3_EK, cfg.E_k_form), (abi.FX_FORM_C3_EE, cfg.E_e_form),
                      (abi.FX_FORM_C3_TORQUE, cfg.torque_form), (abi.FX_FORM_C3_FORCE_X, cfg.force_x_form),
                      (abi.FX_FORM_C3_FORCE_Y, cfg.force_y_form)):
        ref = fem.assemble(form)
        got = plan.assemble_scalar(fid, st, par)
        assert abs(got - ref) < 1e-11 * max(1.0, abs(ref)), (fid, got, ref)
        degs = {k: d for k, _, d in fem.quadrature_degrees(form, mesh)}
        q = plan.form_qdeg(fid)
        assert (q[0], q[1]) == (degs.get("dx", -1), degs.get("ds", -1)), (fid, q, degs)


@pytest.mark.parametrize("variant", ["benchmark_script_with_ewm_parameters", "comp_ewm_jpb"])
def test_true_ewm_forms(variant):
    """non-constant phi_s(theta0) / phi_ewm(tau, theta1) (A_0 = 120, k_ewm = -0.7, B = 0.1, comp_ewm_jpb.py:88-90):
    the C3E functors against the oracle, with UFL's degrees (L4 and L9 rise to 10 = 36 points)."""
    from oracle.configs_ewm import CompEwmJBP
    mesh = fem.annulus_mesh(16, 4, x_1=-0.9)
    kw = {k: v for k, v in ew.EWM_GENERAL.items()}
    cfg = EwmJBP(mesh, **kw) if variant.startswith("benchmark") else CompEwmJBP(mesh, **kw)
    ew.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker)
    st, par = ew.state_of(cfg), ew.params_of(cfg)
    assert par[abi.FX_P_TH_T] == (0.0 if variant == "comp_ewm_jpb" else ew.GENERAL["th"])
    for name, (fid, space) in ew.C3E_FORMS.items():
        pr = dict(par)
        if name == "a6":
            pr[abi.FX_P_EPS] = cfg.SU_RHO_EPS
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, pr), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, pr), ref)
        assert err < TOL, (name, err)
        degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)
    # the folded functors really differ (phi_s, phi_ewm carry weight at these parameters)
    par0 = dict(par)
    par0[abi.FX_P_A0] = par0[abi.FX_P_K_EWM] = par0[abi.FX_P_B_EWM] = 0.0
    b0 = plan.assemble_vector(abi.FX_FORM_C3_L4, "Zc", st, par0)
    assert lc.rel_err(b0, fem.assemble(cfg.forms["L4"])) > 1e-3
