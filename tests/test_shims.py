"""The Level-1 drop-in surface: the shim modules export every helper name the reference's
fenics_fem.py / fenics_base.py define, with the same positional signatures, and their symbolic tensor helpers
compute the same algebra as the oracle's transcription (run here on oracle.ufl_lite in place of UFL)."""
import inspect
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> positional parameter names, as the reference defines them
COMMON = dict(update_progress="job_title progress", mesh2triang="mesh", mplot="obj", tgrad="w", Dincomp="w", Dcomp="w",
              normalize_solution="u", adaptive_refinement="mesh kapp ratio", stream_function="u",
              comp_stream_function="rho u", shear_rate="u", min_location="u mesh", max_location="u mesh",
              l2norm_solution="u", chunks="l n", absolute="u", ramp_function="t",
              reshape_elements="D0_vec D12_vec Ds_vec D1_vec tau0_vec tau12_vec tau1_vec",
              solution_functions="Q W V Z", trial_functions="Q Z W", function_spaces="mesh order", Fdef="u Tau")
LDC = dict(COMMON, expskewcavity="x y N", skewcavity="x y", xskewcavity="x y", yskewcavity="x y",
           LDC_Regular_Mesh="mm B L", Skew_Mesh="mm B L", DGP_structured_mesh="mm x_0 y_0 x_1 y_1 B L",
           refine_boundary=None, refine_top=None, sigmain="u p Tau We betav", sigma="u p Tau We betav", Fdefcon="u Tau",
           save_energy_arrays="t ek ee j tag", load_energy_arrays="j tag", save_err_array="err j tag",
           load_err_array="j tag")
DGP = dict(COMMON, sigma="u p Tau We Pr betav", sigma_n="U p tau n We Pr betav", sigmacom="u p Tau We Pr betav",
           Fdefcom="u Tau", magnitude="u", DGP_unstructured_mesh=None, DGP_Mesh="mm B L", refine_walls=None,
           expskewcavity="x y N", skewcavity="x y", xskewcavity="x y", DGP_structured_mesh="mm x_0 y_0 x_1 y_1 B L",
           refine_boundary=None, refine_top=None, save_energy_arrays="t ek ee loop subloop tag",
           load_energy_arrays="loop subloop tag", save_data_array="nus loop subloop tag", load_data_array="loop subloop tag")
JBP = dict(COMMON, JBP_mesh="mesh_res x_1 x_2 y_1 y_2 r_a r_b", save_solution="u filename", load_solution="filename function_space",
           refine_boundaries=None, refine_narrow=None, DinG="D", sigma="u p Tau betav We", sigmacon="u p Tau betav We",
           fene_sigma="u p Tau b lambda_d betav We", fene_sigmacom="u p Tau b lambda_d betav We", FdefG="u G Tau",
           magnitude="u", euc_norm="tau", phi="u p T A B K_0 N betap", phi_s="T A_0", phi_ewm="tau T k B",
           phi_ewm_inv="tau T k B", Tr="tau", Trace="tau", fene_func="tau b", orthog_proj="u V V_d", frob_norm="tau",
           ernesto_k="u h c_a c_b", I_3="A", I_2="A", phi_def="u lambda_d", psi_def="phi")
HOT = ("assemble", "solve", "project", "KrylovSolver", "DirichletBC", "Function", "norm", "parameters", "pi")


def _mod(which):
    import importlib
    return importlib.import_module({"ldc": "fenics_b200.shims.lid_driven_cavity.fenics_fem",
                                    "dgp": "fenics_b200.shims.buoyancy_driven_flow.fenics_fem",
                                    "jbp": "fenics_b200.shims.eccentric_cylinders.fenics_base"}[which])


@pytest.mark.parametrize("which,table", [("ldc", LDC), ("dgp", DGP), ("jbp", JBP)])
def test_names_and_signatures(which, table):
    m = _mod(which)
    for name in HOT:
        assert hasattr(m, name), name
    for name, params in table.items():
        fn = getattr(m, name, None)
        assert callable(fn), "%s lacks %s" % (m.__name__, name)
        if params is not None:
            got = [p for p in inspect.signature(fn).parameters.values() if p.default is inspect.Parameter.empty]
            assert [p.name for p in got] == params.split(), (name, [p.name for p in got])


def test_import_as_the_reference_scripts_do():
    """`from fenics_fem import *` with the shim directory on sys.path."""
    d = os.path.join(ROOT, "fenics_b200", "shims", "lid_driven_cavity")
    sys.path.insert(0, d)
    try:
        ns = {}
        exec("from fenics_fem import *", ns)
        assert callable(ns["Skew_Mesh"]) and callable(ns["function_spaces"]) and abs(ns["pi"] - np.pi) < 1e-11
    finally:
        sys.path.remove(d)
        sys.modules.pop("fenics_fem", None)


def test_host_utilities(tmp_path, monkeypatch):
    m = _mod("ldc")
    assert list(m.chunks(list(range(5)), 2)) == [[0, 1], [2, 3], [4]]
    assert m.ramp_function(0.5) == 1.0 and abs(m.ramp_function(2.0) - 2.0) < 1e-9
    monkeypatch.chdir(tmp_path)
    m.save_energy_arrays([0.0, 0.1], [1.0, 2.0], [3.0, 4.0], 2, "x")
    t, ek, ee = m.load_energy_arrays(2, "x")
    assert os.path.exists("data/time-data-2-x.npy") and list(ek) == [1.0, 2.0] and list(ee) == [3.0, 4.0] and list(t) == [0.0, 0.1]
    d = _mod("dgp")
    d.save_data_array([5.0], 1, 2, "y")
    assert os.path.exists("data/nus-data-1-2-y.npy") and list(d.load_data_array(1, 2, "y")) == [5.0]
    mesh = m.Skew_Mesh(4, 1, 1)
    assert mesh.num_cells() == 32 and _mod("jbp").JBP_mesh(16, -0.8, 0.0, 0.0, 0.0, 1.0, 2.0).num_cells() > 0
    W, V, Vd, Z, Zd, Zc, Q, Qt, Qr = m.function_spaces(mesh, 2)
    assert W.dim() == 14 * 16 + 8 * 4 + 2 and Vd is None and Q.dim() == 25
    with pytest.raises(NotImplementedError):
        m.refine_top(0, 0, 1, 1, mesh, 1, 0.1)


def test_symbolic_helpers_match_the_oracle_transcription():
    """same algebra as oracle/configs*.py (which restate the reference's helper bodies): evaluated on random fields."""
    from oracle import configs, configs_ewm, configs_jbp, fem, ufl_lite
    from oracle.ufl_lite import dx, inner
    j = _mod("jbp")
    l = _mod("ldc")
    j.set_algebra(ufl_lite)
    mesh = fem.skew_mesh(3)
    cfg = configs.CompLDC(mesh)
    import fxt_common as lc
    lc.randomize(cfg)
    u, tau, p = cfg.u1, cfg.tau1, cfg.p1.expr()
    T = cfg.rho1.expr()
    R = configs.sym(fem.TestFunction(cfg.S["Zc"]))

    def same(a, b):
        va, vb = fem.assemble(inner(a, R) * dx), fem.assemble(inner(b, R) * dx)
        assert np.abs(va - vb).max() < 1e-13 * max(np.abs(vb).max(), 1e-300)

    same(j.Dincomp(u), configs.Dincomp(u))
    same(j.Dcomp(u), configs.Dcomp(u))
    same(j.Fdef(u, tau), configs_jbp.Fdef(u, tau))
    same(l.Fdef(u, tau), configs.Fdef(u, tau))
    same(l.Fdefcon(u, tau), configs.Fdefcon(u, tau))
    same(j.sigmacon(u, p, tau, 0.6, 0.3), configs_ewm.sigmacon(u, p, tau, 0.6, 0.3))
    same(j.fene_sigmacom(u, p, tau, 20.0, 0.0, 0.6, 0.3), configs_jbp.fene_sigmacom(u, p, tau, 20.0, 0.0, 0.6, 0.3))
    same(j.phi_ewm(tau, T, -0.7, 0.1) * tau, configs_ewm.phi_ewm(tau, T, -0.7, 0.1) * tau)
    same(j.phi_s(T, 120.0) * tau, configs_jbp.phi_s(T, 120.0) * tau)
    same(j.fene_func(tau, 20.0) * tau, configs_jbp.fene_func(tau, 20.0) * tau)
    j.set_algebra(None)
