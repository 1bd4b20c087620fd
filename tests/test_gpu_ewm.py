"""GPU parity tests of config 3 (extended White-Metzner journal bearing) through the C ABI."""
import numpy as np
import pytest
import torch

import fxt_common as lc
import fxt_ewm as ew
from fenics_b200 import _abi as abi
from oracle import fem
from oracle.configs_ewm import EwmJBP as OracleEwm

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from fenics_b200 import dolfin_api as dapi  # noqa: E402
from fenics_b200.drivers import ewm  # noqa: E402

TOL = 1e-12


def _pair(nt=16, nr=4, **kw):
    drv = ewm.EwmJBP(n_theta=nt, n_r=nr, **kw)
    m = drv.mesh
    om = fem.Mesh(m.coordinates(), m.cells())
    ocl = OracleEwm(om, dofmaps={k: s.cell_dofs for k, s in drv.S.items()}, **kw)
    assert np.array_equal(om.bfacet_marker, m.bfacet_marker)
    assert abs(ocl.p["dt"] - drv.p["dt"]) < 1e-15 * drv.p["dt"]
    return ocl, drv


def _push(ocl, drv):
    for slot, attr in ew.state_of.__globals__["FIELD_ATTR"].items():
        getattr(drv, attr).vector().set_local(getattr(ocl, attr).vec)


@pytest.mark.parametrize("kw", [{}, ew.GENERAL])
def test_cfg3_forms_and_functionals_match_oracle(kw):
    ocl, drv = _pair(**kw)
    ew.randomize(ocl)
    _push(ocl, drv)
    for name, (fid, space) in ew.C3_FORMS.items():
        ref = fem.assemble(ocl.forms[name])
        got = dapi.assemble(drv.pr.form(fid))
        err = lc.csr_rel_err(got.to_scipy(), ref) if name.startswith("a") else lc.rel_err(got.get_local(), ref)
        assert err < TOL, (name, err)
    for fid, form in ((abi.FX_FORM_C3_EK, ocl.E_k_form), (abi.FX_FORM_C3_EE, ocl.E_e_form),
                      (abi.FX_FORM_C3_TORQUE, ocl.torque_form), (abi.FX_FORM_C3_FORCE_X, ocl.force_x_form),
                      (abi.FX_FORM_C3_FORCE_Y, ocl.force_y_form)):
        ref = fem.assemble(form)
        assert abs(dapi.assemble(drv.pr.form(fid)) - ref) < 1e-11 * max(1.0, abs(ref)), fid
    for bo, bd in zip(ocl.bcu + ocl.bcT, drv.bcu + drv.bcT):
        assert np.array_equal(bo.dofs, bd.dofs)


def test_cfg3_folded_functors_reject_material_functions_loudly():
    """the FX_FORM_C3_* functors fold phi_s = phi_ewm = 1: with non-zero A_0 / k_ewm / B they refuse and name the
    FX_FORM_C3E_* forms (which the driver picks by itself)."""
    _, drv = _pair(nt=8, nr=2)
    drv.pr.set_param(abi.FX_P_A0, 1.0)
    with pytest.raises(Exception) as e:
        dapi.assemble(drv.pr.form(abi.FX_FORM_C3_A1))
    assert "A_0" in str(e.value) and "C3E" in str(e.value)
    assert ewm.EwmJBP(n_theta=8, n_r=2, k_ewm=0.5).ewm


def _pair_comp(nt=16, nr=4, **kw):
    from oracle.configs_ewm import CompEwmJBP as OracleComp
    drv = ewm.CompEwmJBP(n_theta=nt, n_r=nr, **kw)
    m = drv.mesh
    om = fem.Mesh(m.coordinates(), m.cells())
    ocl = OracleComp(om, dofmaps={k: s.cell_dofs for k, s in drv.S.items()}, **kw)
    assert abs(ocl.p["dt"] - drv.p["dt"]) < 1e-15 * drv.p["dt"]
    return ocl, drv


def test_true_ewm_forms_match_oracle():
    """comp_ewm_jpb.py: non-constant phi_s / phi_ewm through the C3E functors on the device."""
    ocl, drv = _pair_comp(**ew.EWM_GENERAL)
    ew.randomize(ocl)
    _push(ocl, drv)
    for name, (fid, space) in ew.C3E_FORMS.items():
        drv.pr.set_param(abi.FX_P_EPS, drv.SU_RHO_EPS if name == "a6" else drv.SU_EPS)
        ref = fem.assemble(ocl.forms[name])
        got = dapi.assemble(drv.pr.form(fid))
        err = lc.csr_rel_err(got.to_scipy(), ref) if name.startswith("a") else lc.rel_err(got.get_local(), ref)
        assert err < TOL, (name, err)


def test_true_ewm_k_steps_match_oracle():
    """drivers.ewm.CompEwmJBP (comp_ewm_jpb.py:421-732, A_0 = 120, k_ewm = -0.7, B = 0.1) against the oracle's
    transcription of that script: fields and journal load / torque after K steps within 1e-8."""
    K = 3
    ocl, drv = _pair_comp(nt=24, nr=6, **ew.EWM_GENERAL)
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    ocl.t = drv.t = 0.45
    for _ in range(K):
        ro, rd = ocl.step(), drv.step()
        for k in ("E_k", "torque", "F_x", "F_y"):
            assert abs(ro[k] - rd[k]) < 1e-8 * max(abs(ro[k]), 1e-10), (k, ro[k], rd[k])
    fo, fd = ocl.fields(), drv.fields()
    for k in ("w1", "p1", "tau1", "rho1", "T1"):
        assert lc.rel_err(fd[k], fo[k]) < 1e-8, k
    assert len(drv.last_info) == 8 and all(i.converged == 1 for i in drv.last_info)


@pytest.mark.parametrize("prepass", [True, False])
def test_cfg3_k_steps_match_oracle(prepass):
    """fields + journal load/torque after K steps within 1e-8 relative (north_star); pre-pass mode
    exercises the SU-stabilised stress solve, main-pass mode the ramped spin BC."""
    K = 3
    ocl, drv = _pair(nt=24, nr=6, prepass=prepass, **ew.GENERAL)
    drv.solve_rtol = 1e-13
    drv.pressure_method = "cg"
    ocl.t = drv.t = 0.45
    for _ in range(K):
        ro, rd = ocl.step(), drv.step()
        for k in ("E_k", "torque", "F_x", "F_y"):
            assert abs(ro[k] - rd[k]) < 1e-8 * max(abs(ro[k]), 1e-10), (k, ro[k], rd[k])
        assert abs(ro["E_e"] - rd["E_e"]) < 1e-8 * 10.0, (ro["E_e"], rd["E_e"])
    fo, fd = ocl.fields(), drv.fields()
    for k in ("w1", "p1", "tau1", "rho1", "T1"):
        assert lc.rel_err(fd[k], fo[k]) < 1e-8, k
    assert all(i.converged == 1 for i in drv.last_info) and len(drv.last_info) == 8  # We = 0.3 > 1e-5: stress solved in both modes
