"""Every hand-written form functor (fenics_b200/csrc/fx_forms_*.h), run through the CPU emulation of
the CUDA element kernels, against the oracle's UFL-semantics evaluation of the reference's forms.
Tolerance: 1e-12 relative to the largest entry (north_star: matrices and vectors within 1e-12)."""
import numpy as np
import pytest

from fenics_b200 import _abi as abi
from oracle import configs, fem
from oracle.ufl_lite import dx, inner
import fxt_common as lc
from fxt_emu import EmuPlan

TOL = 1e-12


def _shuffled_dofmaps(cfg, seed=1):
    """a second, dolfin-like arbitrary numbering: random permutation of every space's dofs."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, s in cfg.S.items():
        perm = rng.permutation(s.dim)
        out[k] = perm[s.cell_dofs]
    return out


def _setup(kind, n=6, shuffle=False, **kw):
    mesh = fem.skew_mesh(n)
    cls = configs.IncompLDC if kind == 1 else configs.CompLDC
    cfg = cls(mesh, **kw)
    if shuffle:
        cfg = cls(mesh, dofmaps=_shuffled_dofmaps(cfg), **kw)
    lc.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local,
                   mesh.bfacet_marker)
    return mesh, cfg, plan


def test_geometry_and_pattern():
    mesh, cfg, plan = _setup(1, n=5)
    g = plan.geometry()
    assert lc.rel_err(g[4], np.abs(mesh.detJ)) < 1e-15
    assert lc.rel_err(g[5], mesh.h) < 1e-15
    K = mesh.K
    for i, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        assert lc.rel_err(g[i], K[:, a, b]) < 1e-14
    for sp_name in ("W", "Zc", "Q"):
        rp, col = plan.pattern(sp_name)
        A = fem.assemble(inner(fem.TrialFunction(cfg.S[sp_name]), fem.TestFunction(cfg.S[sp_name])) * dx)
        assert np.array_equal(rp, A.indptr) and np.array_equal(col, A.indices)  # bit-exact pattern


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("conv", [0.0, 1.0])
def test_cfg1_forms(shuffle, conv):
    mesh, cfg, plan = _setup(1, shuffle=shuffle, Re=0.5, conv=conv)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    for name, (fid, space) in lc.C1_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            got = plan.assemble_matrix(fid, space, st, par)
            err = lc.csr_rel_err(got, ref)
        else:
            got = plan.assemble_vector(fid, space, st, par)
            err = lc.rel_err(got, ref)
        assert err < TOL, (name, err)
        qd = [d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh) if k == "dx"][0]
        assert plan.form_qdeg(fid, conv)[0] == qd, name  # same quadrature rule as UFL's estimate


@pytest.mark.parametrize("shuffle", [False, True])
def test_cfg2_forms(shuffle):
    mesh, cfg, plan = _setup(2, shuffle=shuffle)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    for name, (fid, space) in lc.C2_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            got = plan.assemble_matrix(fid, space, st, par)
            err = lc.csr_rel_err(got, ref)
        else:
            got = plan.assemble_vector(fid, space, st, par)
            err = lc.rel_err(got, ref)
        assert err < TOL, (name, err)
        degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
        q = plan.form_qdeg(fid)
        assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, q, degs)


@pytest.mark.parametrize("kind", [1, 2])
def test_lps_projections_and_functionals(kind):
    mesh, cfg, plan = _setup(kind)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    cfg._lps_indicator()  # oracle: res_test, res_orth, res_orth_norm_sq, kapp
    last = cfg.last
    res_test = plan.project_dg0(abi.FX_EXPR_LDC_RESTAU, "Zd", st, par)
    assert lc.rel_err(res_test, last["res_test"]) < TOL
    st[abi.FX_F_RES_TEST] = res_test
    b = plan.assemble_vector(abi.FX_FORM_LDC_L_RES_ORTH, "Zc", st, par)
    M = plan.assemble_matrix(abi.FX_FORM_MASS_ZC, "Zc", st, par)
    Mref = fem.assemble(inner(fem.TestFunction(cfg.S["Zc"]), fem.TrialFunction(cfg.S["Zc"])) * dx)
    assert lc.csr_rel_err(M, Mref) < TOL
    res_orth = fem.solve_lu(M, b)
    assert lc.rel_err(res_orth, last["res_orth"]) < 1e-10
    st[abi.FX_F_RES_ORTH] = last["res_orth"]
    s = plan.project_dg0(abi.FX_EXPR_RES_ORTH_SQ, "Qt", st, par)
    assert lc.rel_err(s, last["res_orth_norm_sq"]) < TOL
    st[abi.FX_F_QT_A] = last["res_orth_norm_sq"]
    kapp = plan.project_dg0(abi.FX_EXPR_SQRT_QT_A, "Qt", st, par)
    assert lc.rel_err(kapp, last["kapp"]) < TOL
    st[abi.FX_F_KAPP] = cfg.kapp.vec
    hk = plan.project_dg0(abi.FX_EXPR_H_KAPP, "Qt", st, par)
    from oracle.ufl_lite import CellDiameter
    assert lc.rel_err(hk, fem.project(CellDiameter() * cfg.kapp.expr(), cfg.S["Qt"]).vec) < TOL
    MQ = plan.assemble_matrix(abi.FX_FORM_MASS_Q, "Q", st, par)
    MQref = fem.assemble(inner(fem.TestFunction(cfg.S["Q"]), fem.TrialFunction(cfg.S["Q"])) * dx)
    assert lc.csr_rel_err(MQ, MQref) < TOL
    ek = abi.FX_FORM_C1_EK if kind == 1 else abi.FX_FORM_C2_EK
    ee = abi.FX_FORM_C1_EE if kind == 1 else abi.FX_FORM_C2_EE
    assert abs(plan.assemble_scalar(ek, st, par) - fem.assemble(cfg.E_k_form)) < TOL * abs(fem.assemble(cfg.E_k_form))
    assert abs(plan.assemble_scalar(ee, st, par) - fem.assemble(cfg.E_e_form)) < 1e-11 * max(1.0, abs(fem.assemble(cfg.E_e_form)))
    if kind == 2:
        brho = plan.assemble_vector(abi.FX_FORM_C2_L_RHO1, "Q", st, par)
        ref = fem.assemble(inner(fem.TestFunction(cfg.S["Q"]), cfg.rho1_expr) * dx)
        assert lc.rel_err(brho, ref) < TOL


@pytest.mark.parametrize("shuffle", [False, True])
def test_stream_function_forms(shuffle):
    """stream_function / comp_stream_function (lid-driven-cavity/fenics_fem.py:340-378) on the scalar P2
    space V.sub(0).collapse()."""
    mesh, cfg, plan = _setup(2, shuffle=shuffle)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    for rho, fid in ((None, abi.FX_FORM_STREAM_L), (cfg.rho1.expr(), abi.FX_FORM_COMP_STREAM_L)):
        _, a, L = configs.stream_function(cfg.S, cfg.u1, rho)
        assert lc.csr_rel_err(plan.assemble_matrix(abi.FX_FORM_STREAM_A, "Vs", st, par), fem.assemble(a)) < TOL
        assert lc.rel_err(plan.assemble_vector(fid, "Vs", st, par), fem.assemble(L)) < TOL
        for form, f in ((a, abi.FX_FORM_STREAM_A), (L, fid)):
            degs = {k: d for k, _, d in fem.quadrature_degrees(form, mesh)}
            assert plan.form_qdeg(f)[0] == degs["dx"], (f, plan.form_qdeg(f), degs)


def test_comp_ldc_sweep_script_forms():
    """lid-driven-cavity/comp_LDC.py:364-594 = cfg2's loop with Re*conv in the two predictor steps (:418, :468):
    the cfg2 functors with FX_P_CONV = Re*conv reproduce the oracle's transcription of THAT script."""
    mesh = fem.skew_mesh(5)
    cfg = configs.CompLDCSweep(mesh, Re=3.0, conv=0.7)
    lc.randomize(cfg)
    plan = EmuPlan(mesh.coords, mesh.cells, lc.dofmaps_of(cfg), mesh.bfacet_cell, mesh.bfacet_local, mesh.bfacet_marker)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    par[lc.PARAM_SLOT["conv"]] = 3.0 * 0.7  # what drivers.ldc.CompLDCSweep passes
    other = configs.CompLDC(mesh, dofmaps=lc.dofmaps_of(cfg), Re=3.0, conv=0.7)
    for name, (fid, space) in lc.C2_FORMS.items():
        ref = fem.assemble(cfg.forms[name])
        if name.startswith("a"):
            err = lc.csr_rel_err(plan.assemble_matrix(fid, space, st, par), ref)
        else:
            err = lc.rel_err(plan.assemble_vector(fid, space, st, par), ref)
        assert err < TOL, (name, err)
    # and the two scripts really differ there (so the test above is not vacuous)
    for a, b in zip(lc.FIELD_ATTR.values(), lc.FIELD_ATTR.values()):
        if hasattr(cfg, a) and hasattr(other, a) and hasattr(getattr(cfg, a), "vec"):
            getattr(other, a).vec[:] = getattr(cfg, a).vec
    assert lc.rel_err(fem.assemble(other.forms["L2"]), fem.assemble(cfg.forms["L2"])) > 1e-6


def test_shear_rate_form():
    """shear_rate(u1) (lid-driven-cavity/fenics_fem.py:380-398): right-hand side functor against the oracle, UFL degree."""
    mesh, cfg, plan = _setup(2)
    st, par = lc.state_of(cfg), lc.params_of(cfg)
    _, a, L = configs.shear_rate(cfg.S, cfg.u1)
    assert lc.rel_err(plan.assemble_vector(abi.FX_FORM_SHEAR_L, "Vs", st, par), fem.assemble(L)) < TOL
    degs = {k: d for k, _, d in fem.quadrature_degrees(L, mesh)}
    assert plan.form_qdeg(abi.FX_FORM_SHEAR_L)[0] == degs["dx"]
