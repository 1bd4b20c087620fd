"""Node-block solver layout (fenics_b200/csrc/fx_nb_host.h): the host builder and a host replay of the device SpMV
against scipy on the DOLFIN-pattern CSR matrix, for every solver space, under the repo's numbering and under a random
dof permutation (standing in for dolfin's), with the elimination and ghost masks the solvers use."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import fxt_emu
from fenics_b200 import mesh as pmesh
from fenics_b200 import spaces


def _plan(n, permute, seed=0):
    m = pmesh.Skew_Mesh(n, 1, 1)
    rng = np.random.default_rng(seed)
    dm = {}
    for name in ("W", "Zc", "Q", "V", "Vs"):
        cd = spaces.build_cell_dofs(m, spaces.ELEMENTS[name])
        if permute:
            perm = rng.permutation(int(cd.max()) + 1).astype(np.int32)
            cd = perm[cd]
        dm[name] = cd
    z = np.zeros(0, dtype=np.int32)
    return m, dm, fxt_emu.EmuPlan(m.coordinates(), m.cells(), dm, z, z, z)


def _nb_spmv(plan, space, mask, val, x, num_sms=148, ghost_owner=None):
    L = plan.lib
    L.emu_nb_spmv.restype = C.c_int
    y = np.full(len(x), np.nan)
    info = np.zeros(8, dtype=np.int32)
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    rc = L.emu_nb_spmv(plan.h, fxt_emu.SPACES[space], C.c_void_p(mk.ctypes.data if mk is not None else None), num_sms,
                       C.c_void_p(val.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data),
                       C.c_void_p(info.ctypes.data),
                       C.c_void_p(None if ghost_owner is None else np.ascontiguousarray(ghost_owner, dtype=np.uint8).ctypes.data))
    return rc, y, info


@pytest.mark.parametrize("permute", [False, True])
@pytest.mark.parametrize("space", ["W", "Zc", "Q", "V", "Vs"])
@pytest.mark.parametrize("n,num_sms", [(3, 148), (12, 148), (12, 2)])
def test_nb_spmv_matches_csr(space, permute, n, num_sms):
    m, dm, plan = _plan(n, permute)
    rp, col = plan.pattern(space)
    ndof = len(rp) - 1
    rng = np.random.default_rng(1)
    val = rng.standard_normal(len(col))
    val[rng.random(len(col)) < 0.3] = 0.0  # explicit zeros of the cell-block pattern
    mask = None
    active = np.ones(ndof, dtype=bool)
    if space == "W":  # the DG0 DEVSS dofs are eliminated; momentum rows must not couple to them
        dg = np.unique(dm["W"][:, 12:])
        mask = np.zeros(ndof, dtype=np.uint8)
        mask[dg] = 1
        active[dg] = False
        rows = np.repeat(np.arange(ndof), np.diff(rp))
        val[active[rows] & ~active[col]] = 0.0
    A = sp.csr_matrix((val, col, rp), shape=(ndof, ndof))
    x = rng.standard_normal(ndof)
    if mask is not None:
        x[~active] = 0.0  # eliminated entries are identically zero in the iteration
    rc, y, info = _nb_spmv(plan, space, mask, val, x, num_sms)
    assert rc == 0, (rc, info)
    assert info[7] == 0
    ref = A @ x
    assert np.all(np.isnan(y[~active]))
    assert np.allclose(y[active], ref[active], rtol=1e-13, atol=1e-13)
    bs = {"W": 2, "V": 2, "Zc": 3, "Q": 1, "Vs": 1}[space]
    assert info[0] == bs and info[1] * bs == active.sum()


def test_nb_zero_check_and_rejections():
    m, dm, plan = _plan(6, False)
    rp, col = plan.pattern("W")
    ndof = len(rp) - 1
    val = np.ones(len(col))
    x = np.ones(ndof)
    dg = np.unique(dm["W"][:, 12:])
    mask = np.zeros(ndof, dtype=np.uint8)
    mask[dg] = 1
    rc, y, info = _nb_spmv(plan, "W", mask, val, x)
    assert rc == 0 and info[7] == info[6] > 0        # every A[active, eliminated] coupling is reported
    rc, _, _ = _nb_spmv(plan, "W", None, val, x)     # DG0 dofs not eliminated: no node-block structure
    assert rc == -1
    mask2 = mask.copy()
    mask2[dm["W"][0, 0]] = 1                          # a velocity dof eliminated: rejected
    rc, _, _ = _nb_spmv(plan, "W", mask2, val, x)
    assert rc == -1


def test_nb_ghost_rows_are_empty():
    m, dm, plan = _plan(8, False)
    rp, col = plan.pattern("Zc")
    ndof = len(rp) - 1
    rng = np.random.default_rng(3)
    val = rng.standard_normal(len(col))
    x = rng.standard_normal(ndof)
    # ghost = all three components of the nodes of the last cells
    ghost_nodes = np.unique(dm["Zc"][-20:, :6])
    mask = np.zeros(ndof, dtype=np.uint8)
    for k in range(3):
        cols = dm["Zc"][:, 6 * k:6 * k + 6]
        sel = np.isin(dm["Zc"][:, :6], ghost_nodes)
        mask[cols[sel]] = 2
    A = sp.csr_matrix((val, col, rp), shape=(ndof, ndof))
    ref = A @ x
    own = mask == 0
    # (a) ghost rows interleaved in the numbering (chunked as empty rows); (b) partitioned runs: ghost nodes numbered
    # last per owner rank and left out of the chunks altogether
    owner = np.where(own, 0, 1 + (np.arange(ndof) % 3)).astype(np.uint8)
    for k in range(3):  # the components of a node share its owner
        owner[dm["Zc"][:, 6 * k:6 * k + 6]] = owner[dm["Zc"][:, :6]]
    owner[own] = 0
    for go in (None, owner):
        go_arr = None if go is None else np.ascontiguousarray(go)
        rc, y, info = _nb_spmv(plan, "Zc", mask, val, x, ghost_owner=go_arr)
        assert rc == 0
        assert np.allclose(y[own], ref[own], rtol=1e-13, atol=1e-13)
        assert np.all(y[~own] == 0.0)
