"""Shared helpers of the config-4 (FENE-P-MP journal bearing) parity tests."""
import numpy as np

from fenics_b200 import _abi as abi

PARAM_SLOT = dict(dt=abi.FX_P_DT, Re=abi.FX_P_RE, We=abi.FX_P_WE, Ma=abi.FX_P_MA, lambda_d=abi.FX_P_LAMBDA_D,
                  betav=abi.FX_P_BETAV, b_fene=abi.FX_P_B_FENE, a0=abi.FX_P_A0, Di=abi.FX_P_DI, Vh=abi.FX_P_VH,
                  Bi=abi.FX_P_BI, T0=abi.FX_P_T0, th_hot=abi.FX_P_TH_HOT, th=abi.FX_P_TH, conv=abi.FX_P_CONV)
FIELD_ATTR = {abi.FX_F_W0: "w0", abi.FX_F_W12: "w12", abi.FX_F_WS: "ws", abi.FX_F_W1: "w1", abi.FX_F_P0: "p0",
              abi.FX_F_P1: "p1", abi.FX_F_RHO0: "rho0", abi.FX_F_RHO12: "rho12", abi.FX_F_RHO1: "rho1",
              abi.FX_F_T0: "T0", abi.FX_F_T1: "T1", abi.FX_F_TAU0: "tau0_vec", abi.FX_F_TAU1: "tau1_vec",
              abi.FX_F_KAPP: "kapp"}
C4_FORMS = dict(a0=(abi.FX_FORM_C4_A0, "Q"), a2=(abi.FX_FORM_C4_A2, "W"), a5=(abi.FX_FORM_C4_A5, "Q"),
                a6=(abi.FX_FORM_C4_A6, "Q"), a7=(abi.FX_FORM_C4_A7, "W"), a4=(abi.FX_FORM_C4_A4, "Zc"),
                a9=(abi.FX_FORM_C4_A9, "Q"),
                L0=(abi.FX_FORM_C4_L0, "Q"), L2=(abi.FX_FORM_C4_L2, "W"), L5=(abi.FX_FORM_C4_L5, "Q"),
                L6=(abi.FX_FORM_C4_L6, "Q"), L7=(abi.FX_FORM_C4_L7, "W"), L4=(abi.FX_FORM_C4_L4, "Zc"),
                L9=(abi.FX_FORM_C4_L9, "Q"))
SU_EPS = 0.0000000001  # fenepmp_jbp.py:553


def randomize(cfg, seed=0):
    """O(1) random state that keeps the FENE function and the densities away from their singularities."""
    rng = np.random.default_rng(seed)
    for attr in set(FIELD_ATTR.values()):
        fn = getattr(cfg, attr)
        fn.vec[:] = rng.standard_normal(fn.space.dim)
    cfg.kapp.vec[:] = np.abs(cfg.kapp.vec)
    for r in (cfg.rho0, cfg.rho12, cfg.rho1):
        r.vec[:] = 1.0 + 0.1 * r.vec
    for T in (cfg.T0, cfg.T1):
        T.vec[:] = 325.0 + 10.0 * T.vec
    for tv in (cfg.tau0_vec, cfg.tau1_vec):  # conformation tensor ~ I + O(0.3)
        comp = tv.space.dof_component()
        tv.vec[:] = 0.3 * tv.vec + np.where(comp == 1, 0.0, 1.0)


def state_of(cfg, extra=None):
    st = {slot: getattr(cfg, attr).vec for slot, attr in FIELD_ATTR.items()}
    if extra:
        st.update(extra)
    return st


def params_of(cfg):
    p = {PARAM_SLOT[k]: float(v) for k, v in cfg.p.items()}
    p[abi.FX_P_EPS] = SU_EPS
    return p
