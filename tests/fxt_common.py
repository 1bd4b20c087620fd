"""Shared helpers of the LDC parity tests: oracle config objects <-> C-ABI state/param tables."""
import numpy as np

from fenics_b200 import _abi as abi
from oracle import configs, fem

PARAM_SLOT = dict(dt=abi.FX_P_DT, Re=abi.FX_P_RE, We=abi.FX_P_WE, betav=abi.FX_P_BETAV, c1=abi.FX_P_C1,
                  c2=abi.FX_P_C2, th=abi.FX_P_TH, conv=abi.FX_P_CONV, Ma=abi.FX_P_MA)

FIELD_ATTR = {abi.FX_F_W0: "w0", abi.FX_F_W12: "w12", abi.FX_F_WS: "ws", abi.FX_F_W1: "w1",
              abi.FX_F_P0: "p0", abi.FX_F_P1: "p1", abi.FX_F_RHO0: "rho0", abi.FX_F_RHO12: "rho12",
              abi.FX_F_RHO1: "rho1", abi.FX_F_TAU0: "tau0_vec", abi.FX_F_TAU1: "tau1_vec",
              abi.FX_F_KAPP: "kapp"}

# form name in oracle config -> (C-ABI id, space)
C1_FORMS = dict(a1=(abi.FX_FORM_C1_A1, "W"), a2=(abi.FX_FORM_C1_A2, "W"), a5=(abi.FX_FORM_C1_A5, "Q"),
                a7=(abi.FX_FORM_C1_A7, "W"), a4=(abi.FX_FORM_C1_A4, "Zc"),
                L1=(abi.FX_FORM_C1_L1, "W"), L2=(abi.FX_FORM_C1_L2, "W"), L5=(abi.FX_FORM_C1_L5, "Q"),
                L7=(abi.FX_FORM_C1_L7, "W"), L4=(abi.FX_FORM_C1_L4, "Zc"))
C2_FORMS = dict(a0=(abi.FX_FORM_C2_A0, "Q"), a1=(abi.FX_FORM_C2_A1, "W"), a2=(abi.FX_FORM_C2_A2, "W"),
                a5=(abi.FX_FORM_C2_A5, "Q"), a7=(abi.FX_FORM_C2_A7, "W"), a4=(abi.FX_FORM_C2_A4, "Zc"),
                L0=(abi.FX_FORM_C2_L0, "Q"), L1=(abi.FX_FORM_C2_L1, "W"), L2=(abi.FX_FORM_C2_L2, "W"),
                L5=(abi.FX_FORM_C2_L5, "Q"), L7=(abi.FX_FORM_C2_L7, "W"), L4=(abi.FX_FORM_C2_L4, "Zc"))


def randomize(cfg, seed=0):
    """fill every state function of an oracle config with O(1) random data (kapp >= 0)."""
    rng = np.random.default_rng(seed)
    for attr in set(FIELD_ATTR.values()):
        fn = getattr(cfg, attr, None)
        if fn is not None:
            fn.vec[:] = rng.standard_normal(fn.space.dim)
    cfg.kapp.vec[:] = np.abs(cfg.kapp.vec)
    if hasattr(cfg, "rho0"):
        for r in (cfg.rho0, cfg.rho1):
            r.vec[:] = 1.0 + 0.1 * r.vec


def state_of(cfg, extra=None):
    st = {}
    for slot, attr in FIELD_ATTR.items():
        fn = getattr(cfg, attr, None)
        if fn is not None:
            st[slot] = fn.vec
    if extra:
        st.update(extra)
    return st


def params_of(cfg):
    return {PARAM_SLOT[k]: float(v) for k, v in cfg.p.items()}


def dofmaps_of(cfg):
    return {k: s.cell_dofs for k, s in cfg.S.items()}


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a)))


def csr_rel_err(A, B):
    """both CSR with identical pattern (asserted)."""
    A.sort_indices()
    B.sort_indices()
    assert A.shape == B.shape and A.nnz == B.nnz
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    return rel_err(A.data, B.data)
