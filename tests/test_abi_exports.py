"""CPU: the C-ABI library builds, loads and exports every symbol include/fenics_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fenics_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fenics_b200.build import build
    so = build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    lib.fx_version.restype = ctypes.c_int
    assert lib.fx_version() == 100
    lib.fx_space_local_dim.restype = ctypes.c_int
    assert [lib.fx_space_local_dim(s) for s in range(6)] == [15, 18, 3, 3, 1, 12]


def test_python_binding_covers_header():
    from fenics_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_form_info_matches_oracle_degrees():
    """fx_form_info is pure host code: the quadrature degree compiled into every kernel equals UFL's
    estimate computed by the oracle for the reference's forms (cfg1 with conv != 0, cfg2)."""
    import fxt_common as lc
    from fenics_b200 import _lib
    from oracle import configs, fem
    mesh = fem.skew_mesh(2)
    for cfg, table in ((configs.IncompLDC(mesh, Re=2.0), lc.C1_FORMS), (configs.CompLDC(mesh), lc.C2_FORMS)):
        for name, (fid, space) in table.items():
            rank, sp = ctypes.c_int(), ctypes.c_int()
            q = (ctypes.c_int * 2)()
            assert _lib.lib.fx_form_info(fid, ctypes.byref(rank), ctypes.byref(sp), q) == 0
            degs = {k: d for k, _, d in fem.quadrature_degrees(cfg.forms[name], mesh)}
            assert q[0] == degs["dx"] and q[1] == degs.get("ds", -1), (name, list(q), degs)
            assert rank.value == (2 if name.startswith("a") else 1)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from fenics_b200 import la, mesh, spaces
    m = mesh.Skew_Mesh(2, 1, 1)
    W, V, _, _, Zd, Zc, Q, Qt, _ = spaces.function_spaces(m)
    with pytest.raises(RuntimeError):
        la.Plan(m, [W, Zc, Q, Zd, Qt, V])
