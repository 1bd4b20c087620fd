"""Shared helpers of the config-5 (compressible buoyancy-driven flow) parity tests."""
import numpy as np

from fenics_b200 import _abi as abi

PARAM_SLOT = dict(dt=abi.FX_P_DT, Ra=abi.FX_P_RA, We=abi.FX_P_WE, Ma=abi.FX_P_MA, Pr=abi.FX_P_PR,
                  betav=abi.FX_P_BETAV, c1=abi.FX_P_C1, c2=abi.FX_P_C2, th=abi.FX_P_TH, Vh=abi.FX_P_VH, Bi=abi.FX_P_BI,
                  T0=abi.FX_P_T0, th_hot=abi.FX_P_TH_HOT, al1=abi.FX_P_AL1, al2=abi.FX_P_AL2)
FIELD_ATTR = {abi.FX_F_W0: "w0", abi.FX_F_W12: "w12", abi.FX_F_WS: "ws", abi.FX_F_W1: "w1", abi.FX_F_P0: "p0",
              abi.FX_F_P1: "p1", abi.FX_F_RHO0: "rho0", abi.FX_F_RHO12: "rho12", abi.FX_F_RHO1: "rho1",
              abi.FX_F_T0: "T0", abi.FX_F_T1: "T1", abi.FX_F_THETA0: "theta0", abi.FX_F_THERM_INV: "therm_inv",
              abi.FX_F_Q_A: "theta1", abi.FX_F_TAU0: "tau0_vec", abi.FX_F_TAU1: "tau1_vec", abi.FX_F_KAPP: "kapp"}
C5_FORMS = dict(a0=(abi.FX_FORM_C5_A0, "Q"), a1=(abi.FX_FORM_C5_A1, "W"), a2=(abi.FX_FORM_C5_A2, "W"),
                a5=(abi.FX_FORM_C5_A5, "Q"), a6=(abi.FX_FORM_C5_A6, "Q"), a7=(abi.FX_FORM_C5_A7, "W"),
                a4=(abi.FX_FORM_C5_A4, "Zc"), a8=(abi.FX_FORM_C5_A8, "Q"),
                L0=(abi.FX_FORM_C5_L0, "Q"), L1=(abi.FX_FORM_C5_L1, "W"), L2=(abi.FX_FORM_C5_L2, "W"),
                L5=(abi.FX_FORM_C5_L5, "Q"), L6=(abi.FX_FORM_C5_L6, "Q"), L7=(abi.FX_FORM_C5_L7, "W"),
                L4=(abi.FX_FORM_C5_L4, "Zc"), L8=(abi.FX_FORM_C5_L8, "Q"))
SU_EPS = 3.0e-16  # DOLFIN_EPS, compressible_dgp.py:510


def randomize(cfg, seed=0):
    rng = np.random.default_rng(seed)
    for attr in set(FIELD_ATTR.values()):
        fn = getattr(cfg, attr)
        fn.vec[:] = rng.standard_normal(fn.space.dim)
    cfg.kapp.vec[:] = np.abs(cfg.kapp.vec)
    for r in (cfg.rho0, cfg.rho12, cfg.rho1):
        r.vec[:] = 1.0 + 0.1 * r.vec
    for T in (cfg.T0, cfg.T1):
        T.vec[:] = 325.0 + 10.0 * T.vec
    cfg.theta0.vec[:] = 0.5 + 0.2 * cfg.theta0.vec   # therm = 1 + theta0/6 stays away from 0
    cfg.therm_inv.vec[:] = 1.0 + 0.1 * cfg.therm_inv.vec


def state_of(cfg):
    return {slot: getattr(cfg, attr).vec for slot, attr in FIELD_ATTR.items()}


def params_of(cfg):
    p = {PARAM_SLOT[k]: float(v) for k, v in cfg.p.items()}
    p[abi.FX_P_EPS] = SU_EPS
    return p


# ---- incompressible_dgp.py ("cfg6") ------------------------------------------------------------------------------
C6_PARAM_SLOT = dict(dt=abi.FX_P_DT, Ra=abi.FX_P_RA, We=abi.FX_P_WE, Pr=abi.FX_P_PR, betav=abi.FX_P_BETAV, c1=abi.FX_P_C1,
                     c2=abi.FX_P_C2, th=abi.FX_P_TH, Vh=abi.FX_P_VH, Bi=abi.FX_P_BI, T0=abi.FX_P_T0, th_hot=abi.FX_P_TH_HOT)
C6_FIELD_ATTR = {abi.FX_F_W0: "w0", abi.FX_F_W12: "w12", abi.FX_F_WS: "ws", abi.FX_F_W1: "w1", abi.FX_F_P0: "p0",
                 abi.FX_F_P1: "p1", abi.FX_F_T0: "T0", abi.FX_F_T1: "T1", abi.FX_F_Q_A: "theta1",
                 abi.FX_F_TAU0: "tau0_vec", abi.FX_F_TAU1: "tau1_vec", abi.FX_F_KAPP: "kapp"}
C6_FORMS = dict(a1=(abi.FX_FORM_C6_A1, "W"), a2=(abi.FX_FORM_C6_A2, "W"), a5=(abi.FX_FORM_C6_A5, "Q"),
                a7=(abi.FX_FORM_C6_A7, "W"), a4=(abi.FX_FORM_C6_A4, "Zc"), a8=(abi.FX_FORM_C6_A8, "Q"),
                L1=(abi.FX_FORM_C6_L1, "W"), L2=(abi.FX_FORM_C6_L2, "W"), L5=(abi.FX_FORM_C6_L5, "Q"),
                L7=(abi.FX_FORM_C6_L7, "W"), L4=(abi.FX_FORM_C6_L4, "Zc"), L8=(abi.FX_FORM_C6_L8, "Q"))


def c6_randomize(cfg, seed=0):
    rng = np.random.default_rng(seed)
    for attr in set(C6_FIELD_ATTR.values()):
        fn = getattr(cfg, attr)
        fn.vec[:] = rng.standard_normal(fn.space.dim)
    cfg.kapp.vec[:] = np.abs(cfg.kapp.vec)
    for T in (cfg.T0, cfg.T1):
        T.vec[:] = 325.0 + 10.0 * T.vec


def c6_state_of(cfg):
    return {slot: getattr(cfg, attr).vec for slot, attr in C6_FIELD_ATTR.items()}


def c6_params_of(cfg):
    p = {C6_PARAM_SLOT[k]: float(v) for k, v in cfg.p.items()}
    p[abi.FX_P_RE] = 1.0  # the shared L5 functor (C1_L5) carries Re/dt; this loop has 1/dt
    return p
