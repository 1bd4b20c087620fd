"""Shared helpers of the config-3 (extended White-Metzner journal bearing) parity tests."""
from fenics_b200 import _abi as abi
from fxt_jbp import PARAM_SLOT as _JBP_SLOT, SU_EPS, randomize, state_of  # noqa: F401

PARAM_SLOT = dict(_JBP_SLOT, al1=abi.FX_P_AL1, k_ewm=abi.FX_P_K_EWM, b_ewm=abi.FX_P_B_EWM, th_t=abi.FX_P_TH_T)
C3_FORMS = dict(a0=(abi.FX_FORM_C3_A0, "Q"), a1=(abi.FX_FORM_C3_A1, "W"), a2=(abi.FX_FORM_C3_A2, "W"),
                a5=(abi.FX_FORM_C3_A5, "Q"), a6=(abi.FX_FORM_C3_A6, "Q"), a7=(abi.FX_FORM_C3_A7, "W"),
                a4=(abi.FX_FORM_C3_A4, "Zc"), a9=(abi.FX_FORM_C3_A9, "Q"),
                L0=(abi.FX_FORM_C3_L0, "Q"), L1=(abi.FX_FORM_C3_L1, "W"), L2=(abi.FX_FORM_C3_L2, "W"),
                L5=(abi.FX_FORM_C3_L5, "Q"), L6=(abi.FX_FORM_C3_L6, "Q"), L7=(abi.FX_FORM_C3_L7, "W"),
                L4=(abi.FX_FORM_C3_L4, "Zc"), L9=(abi.FX_FORM_C3_L9, "Q"))
# the script's own parameters (We = Ma = 1e-6, betav = 1 - DOLFIN_EPS) switch the polymer terms off
# numerically; the parity tests also run a set where every term carries weight
GENERAL = dict(dt=0.002, Re=7.0, We=0.3, Ma=0.05, betav=0.6, al=0.2, Di=0.004, Vh=0.03, Bi=0.2, th=0.7)


def params_of(cfg):
    p = {PARAM_SLOT[k]: float(v) for k, v in cfg.p.items()}
    p[abi.FX_P_EPS] = SU_EPS
    return p

# comp_ewm_jpb.py: the true EWM model -- same loop, the six forms that carry phi_s / phi_ewm go through the C3E functors
C3E_FORMS = dict(C3_FORMS, a1=(abi.FX_FORM_C3E_A1, "W"), L1=(abi.FX_FORM_C3E_L1, "W"), L2=(abi.FX_FORM_C3E_L2, "W"),
                 a4=(abi.FX_FORM_C3E_A4, "Zc"), L4=(abi.FX_FORM_C3E_L4, "Zc"), L9=(abi.FX_FORM_C3E_L9, "Q"))
EWM_GENERAL = dict(GENERAL, A_0=120.0, k_ewm=-0.7, B=0.1)
