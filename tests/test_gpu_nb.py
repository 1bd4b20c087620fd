"""GPU tests of the node-block solver layout (fx_nb.cuh): the stand-alone product through fx_spmv_variant(8) against
scipy on the DOLFIN-pattern matrix, and BiCGStab on it against the scalar compacted-CSR kernel (FX_NB=0 subprocess)."""
import numpy as np
import pytest
import torch

import fxt_common as lc
from fenics_b200 import _abi as abi

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from fenics_b200 import dolfin_api as dapi  # noqa: E402
from fenics_b200 import la  # noqa: E402
from fenics_b200._lib import check, lib  # noqa: E402
from test_gpu_parity import _pair, _push_state  # noqa: E402


def _nb_spmv(plan, space, A, xv, reps=1):
    x, y = la.Vector(plan, space), la.Vector(plan, space)
    x.set_local(xv)
    y.t.fill_(float("nan"))
    check(lib.fx_spmv_variant(plan.h, abi.SPACE_ID[space], la._ptr(A.val), la._ptr(x.t), la._ptr(y.t), 8 + (reps << 8),
                              la._stream()))
    torch.cuda.synchronize()
    return y.get_local()


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("n", [5, 24])
def test_nb_spmv_matches_scipy(n, shuffle):
    ocl, drv = _pair(2, n=n, shuffle=shuffle)
    lc.randomize(ocl)
    _push_state(ocl, drv)
    rng = np.random.default_rng(0)
    for name, fid in (("W", abi.FX_FORM_C2_A1), ("Zc", abi.FX_FORM_C2_A4), ("Q", abi.FX_FORM_C2_A5)):
        A = dapi.assemble(drv.pr.form(fid))
        xv = rng.standard_normal(A.n)
        active = np.ones(A.n, dtype=bool)
        if name == "W":
            dg = np.unique(drv.S["W"].cell_dofs[:, 12:])
            active[dg] = False
            xv[dg] = 0.0  # eliminated entries are identically zero inside the iteration
        ref = A.to_scipy() @ xv
        for reps in (1, 3):
            y = _nb_spmv(drv.plan, name, A, xv, reps)
            assert np.all(np.isnan(y[~active]))
            assert lc.rel_err(y[active], ref[active]) < 1e-13, (name, reps)


def test_nb_layout_needs_the_elimination_for_w():
    ocl, drv = _pair(2, n=6)
    A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A1))
    drv.set_devss_elimination(False)
    x, y = la.Vector(drv.plan, "W"), la.Vector(drv.plan, "W")
    rc = lib.fx_spmv_variant(drv.plan.h, abi.SPACE_ID["W"], la._ptr(A.val), la._ptr(x.t), la._ptr(y.t), 8, la._stream())
    assert rc == abi.FX_ERR_UNSUPPORTED and b"not eliminated" in lib.fx_last_error()
    # the solver falls back to the scalar kernel by itself
    b = la.Vector(drv.plan, "W")
    b.set_local(np.random.default_rng(1).standard_normal(b.size()))
    for bc in drv.bcu:
        bc.apply(A, b)
    assert la.krylov_solve(A, x, b, "bicgstab", "jacobi", rtol=1e-10).converged == 1
    drv.set_devss_elimination(True)
    x2 = la.Vector(drv.plan, "W")
    assert la.krylov_solve(A, x2, b, "bicgstab", "jacobi", rtol=1e-12).converged == 1
    la.krylov_solve(A, x, b, "bicgstab", "jacobi", rtol=1e-12)
    assert lc.rel_err(x2.get_local(), x.get_local()) < 1e-8


def test_nb_bicgstab_equals_scalar_kernel():
    """same systems through FX_NB=0 (scalar compacted CSR) and the default node-block kernel: same solution, iteration
    counts close (BiCGStab is sensitive to the summation order, which differs in the last bits)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for nb in ("0", "1"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "nb_check.py"), "20"], capture_output=True,
                           text=True, timeout=600, env=dict(os.environ, FX_NB=nb))
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs[nb] = json.loads(r.stdout.strip().splitlines()[-1])
    for key, a in outs["0"].items():
        b = outs["1"][key]
        assert a["converged"] == b["converged"] == 1, (key, a, b)
        assert abs(a["iterations"] - b["iterations"]) <= max(3, 0.3 * a["iterations"]), (key, a, b)
        assert b["err_vs_splu"] < 1e-8 and a["err_vs_splu"] < 1e-8, (key, a, b)
