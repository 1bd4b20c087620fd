// Form registry: maps the public form ids (include/fenics_b200.h) to the hand-written functors.
// X(id, Functor).  Used by fx_asm.cu (CUDA launchers) and tests/hostemu (CPU emulation).
#pragma once
#include "fx_forms_dgp.h"
#include "fx_forms_ewm.h"
#include "fx_forms_jbp.h"
#include "fx_forms_ldc.h"

// forms whose quadrature degree depends on conv == 0.0 (UFL folds `0.0*expr` to Zero,
// incomp_LDC.py:93-94) are listed with both instantiations: XC(id, FunctorIfConvZero, FunctorElse)
#define FX_MATRIX_FORMS(X)                  \
  X(FX_FORM_MASS_ZC, fx::ldc::MASS_ZC)      \
  X(FX_FORM_MASS_Q, fx::ldc::MASS_Q)        \
  X(FX_FORM_STREAM_A, fx::ldc::STREAM_A)    \
  X(FX_FORM_C1_A1, fx::ldc::C1_A1)          \
  X(FX_FORM_C1_A2, fx::ldc::C1_A2)          \
  X(FX_FORM_C1_A5, fx::ldc::C1_A5)          \
  X(FX_FORM_C1_A7, fx::ldc::C1_A7)          \
  X(FX_FORM_C1_A4, fx::ldc::LDC_A4)         \
  X(FX_FORM_C2_A0, fx::ldc::C2_A0)          \
  X(FX_FORM_C2_A1, fx::ldc::C2_A12<true>)   \
  X(FX_FORM_C2_A2, fx::ldc::C2_A12<false>)  \
  X(FX_FORM_C2_A5, fx::ldc::C2_A5)          \
  X(FX_FORM_C2_A7, fx::ldc::C2_A7)          \
  X(FX_FORM_C2_A4, fx::ldc::LDC_A4)         \
  X(FX_FORM_C3_A0, fx::ldc::C2_A0)          \
  X(FX_FORM_C3_A1, fx::ewm::C3_A1)          \
  X(FX_FORM_C3_A2, fx::jbp::C4_A2)          \
  X(FX_FORM_C3_A5, fx::ewm::C3_A5)          \
  X(FX_FORM_C3_A6, fx::jbp::SU_RHO_A6)      \
  X(FX_FORM_C3_A7, fx::ewm::C3_A7)          \
  X(FX_FORM_C3_A4, fx::ewm::C3_A4)          \
  X(FX_FORM_C3_A9, fx::ewm::C3_A9)          \
  X(FX_FORM_C3E_A1, fx::ewm::C3E_A1)        \
  X(FX_FORM_C3E_A4, fx::ewm::C3E_A4)        \
  X(FX_FORM_C4_A0, fx::ldc::C2_A0)          \
  X(FX_FORM_C4_A2, fx::jbp::C4_A2)          \
  X(FX_FORM_C4_A5, fx::jbp::C4_A5)          \
  X(FX_FORM_C4_A6, fx::jbp::SU_RHO_A6)      \
  X(FX_FORM_C4_A7, fx::jbp::C4_A7)          \
  X(FX_FORM_C4_A4, fx::jbp::C4_A4)          \
  X(FX_FORM_C4_A9, fx::jbp::C4_A9)          \
  X(FX_FORM_C4M_A4, fx::jbp::C4M_A4)        \
  X(FX_FORM_C5_A0, fx::ldc::C2_A0)          \
  X(FX_FORM_C5_A1, fx::dgp::C5_A12<true>)   \
  X(FX_FORM_C5_A2, fx::dgp::C5_A12<false>)  \
  X(FX_FORM_C5_A5, fx::dgp::C5_A5)          \
  X(FX_FORM_C5_A6, fx::jbp::SU_RHO_A6)      \
  X(FX_FORM_C5_A7, fx::dgp::C5_A7)          \
  X(FX_FORM_C5_A4, fx::dgp::C5_A4)          \
  X(FX_FORM_C5_A8, fx::dgp::C5_A8)          \
  X(FX_FORM_C6_A1, fx::dgp::C6_A12<true>)   \
  X(FX_FORM_C6_A2, fx::dgp::C6_A12<false>)  \
  X(FX_FORM_C6_A5, fx::ldc::C1_A5)          \
  X(FX_FORM_C6_A7, fx::dgp::C6_A7)          \
  X(FX_FORM_C6_A4, fx::ldc::LDC_A4)         \
  X(FX_FORM_C6_A8, fx::dgp::C6_A8)

#define FX_VECTOR_FORMS(X, XC)                              \
  XC(FX_FORM_C1_L1, fx::ldc::C1_L1<4>, fx::ldc::C1_L1<5>)   \
  XC(FX_FORM_C1_L2, fx::ldc::C1_L2<4>, fx::ldc::C1_L2<5>)   \
  X(FX_FORM_STREAM_L, fx::ldc::STREAM_L<false>)             \
  X(FX_FORM_COMP_STREAM_L, fx::ldc::STREAM_L<true>)         \
  X(FX_FORM_SHEAR_L, fx::ldc::SHEAR_L)                      \
  X(FX_FORM_C1_L5, fx::ldc::C1_L5)                          \
  X(FX_FORM_C1_L7, fx::ldc::C1_L7)                          \
  X(FX_FORM_C1_L4, fx::ldc::LDC_L4<false>)                  \
  X(FX_FORM_LDC_L_RES_ORTH, fx::ldc::LDC_L_RES_ORTH)        \
  X(FX_FORM_C2_L0, fx::ldc::C2_L0)                          \
  X(FX_FORM_C2_L1, fx::ldc::C2_L12<true>)                   \
  X(FX_FORM_C2_L2, fx::ldc::C2_L12<false>)                  \
  X(FX_FORM_C2_L5, fx::ldc::C2_L5)                          \
  X(FX_FORM_C2_L7, fx::ldc::C2_L7)                          \
  X(FX_FORM_C2_L4, fx::ldc::LDC_L4<true>)                   \
  X(FX_FORM_C2_L_RHO1, fx::ldc::C2_L_RHO1)                  \
  X(FX_FORM_C3_L0, fx::ldc::C2_L0)                          \
  X(FX_FORM_C3_L1, fx::ewm::C3_L1)                          \
  X(FX_FORM_C3_L2, fx::ewm::C3_L2)                          \
  X(FX_FORM_C3_L5, fx::ewm::C3_L5)                          \
  X(FX_FORM_C3_L6, fx::jbp::SU_RHO_L6)                      \
  X(FX_FORM_C3_L7, fx::ewm::C3_L7)                          \
  X(FX_FORM_C3_L4, fx::ewm::C3_L4)                          \
  X(FX_FORM_C3_L9, fx::ewm::C3_L9)                          \
  X(FX_FORM_C3E_L1, fx::ewm::C3E_L1)                        \
  X(FX_FORM_C3E_L2, fx::ewm::C3E_L2)                        \
  X(FX_FORM_C3E_L4, fx::ewm::C3E_L4)                        \
  X(FX_FORM_C3E_L9, fx::ewm::C3E_L9)                        \
  X(FX_FORM_C4_L0, fx::ldc::C2_L0)                          \
  X(FX_FORM_C4_L2, fx::jbp::C4_L2)                          \
  X(FX_FORM_C4_L5, fx::jbp::C4_L5)                          \
  X(FX_FORM_C4_L6, fx::jbp::SU_RHO_L6)                      \
  X(FX_FORM_C4_L7, fx::jbp::C4_L7)                          \
  X(FX_FORM_C4_L4, fx::jbp::C4_L4)                          \
  X(FX_FORM_C4_L9, fx::jbp::C4_L9)                          \
  X(FX_FORM_C4_L_RES_ORTH, fx::jbp::C4_L_RES_ORTH)          \
  X(FX_FORM_C4M_L2, fx::jbp::C4M_L2)                        \
  X(FX_FORM_C4M_L9, fx::jbp::C4M_L9)                        \
  X(FX_FORM_C4M_L_RES_ORTH, fx::jbp::C4M_L_RES_ORTH)        \
  X(FX_FORM_C5_L0, fx::ldc::C2_L0)                          \
  X(FX_FORM_C5_L1, fx::dgp::C5_L12<true>)                   \
  X(FX_FORM_C5_L2, fx::dgp::C5_L12<false>)                  \
  X(FX_FORM_C5_L5, fx::dgp::C5_L5)                          \
  X(FX_FORM_C5_L6, fx::jbp::SU_RHO_L6)                      \
  X(FX_FORM_C5_L7, fx::dgp::C5_L7)                          \
  X(FX_FORM_C5_L4, fx::jbp::C4_L4)                          \
  X(FX_FORM_C5_L8, fx::dgp::C5_L8)                          \
  X(FX_FORM_C5_L_RES_ORTH, fx::dgp::C5_L_RES_ORTH)          \
  X(FX_FORM_C5_L_THETA0, fx::dgp::C5_L_THETA<fx::ldc::F_T0>) \
  X(FX_FORM_C5_L_THETA1, fx::dgp::C5_L_THETA<fx::ldc::F_T1>) \
  X(FX_FORM_C5_L_THERM_INV, fx::dgp::C5_L_THERM_INV)        \
  X(FX_FORM_C5_L_RHO1, fx::dgp::C5_L_RHO1)                  \
  X(FX_FORM_C6_L1, fx::dgp::C6_L12<true>)                   \
  X(FX_FORM_C6_L2, fx::dgp::C6_L12<false>)                  \
  X(FX_FORM_C6_L5, fx::ldc::C1_L5)                          \
  X(FX_FORM_C6_L7, fx::dgp::C6_L7)                          \
  X(FX_FORM_C6_L4, fx::jbp::C4_L4)                          \
  X(FX_FORM_C6_L8, fx::dgp::C6_L8)                          \
  X(FX_FORM_C6_L_RES_ORTH, fx::dgp::C6_L_RES_ORTH)

#define FX_SCALAR_FORMS(X)              \
  X(FX_FORM_C1_EK, fx::ldc::C1_EK)      \
  X(FX_FORM_C1_EE, fx::ldc::LDC_EE)     \
  X(FX_FORM_C2_EK, fx::ldc::C2_EK)      \
  X(FX_FORM_C2_EE, fx::ldc::LDC_EE)     \
  X(FX_FORM_C3_EK, fx::ldc::C2_EK)      \
  X(FX_FORM_C3_EE, fx::dgp::C5_EE)      \
  X(FX_FORM_C4_EK, fx::ldc::C2_EK)      \
  X(FX_FORM_C4_EE, fx::jbp::C4_EE)      \
  X(FX_FORM_C5_EK, fx::ldc::C2_EK)      \
  X(FX_FORM_C5_EE, fx::dgp::C5_EE)      \
  X(FX_FORM_C6_EK, fx::ldc::C1_EK)      \
  X(FX_FORM_C6_EE, fx::dgp::C5_EE)

#define FX_FACET_SCALAR_FORMS(X)                    \
  X(FX_FORM_C3_TORQUE, fx::ewm::C3_JOURNAL<0>)      \
  X(FX_FORM_C3_FORCE_X, fx::ewm::C3_JOURNAL<1>)     \
  X(FX_FORM_C3_FORCE_Y, fx::ewm::C3_JOURNAL<2>)     \
  X(FX_FORM_C4_TORQUE, fx::jbp::C4_JOURNAL<0>)      \
  X(FX_FORM_C4_FORCE_X, fx::jbp::C4_JOURNAL<1>)     \
  X(FX_FORM_C4_FORCE_Y, fx::jbp::C4_JOURNAL<2>)     \
  X(FX_FORM_C4M_TORQUE, fx::jbp::C4M_TORQUE)        \
  X(FX_FORM_C4M_FORCE_X, fx::jbp::C4M_FORCE_X)      \
  X(FX_FORM_C4M_FORCE_Y, fx::jbp::C4M_FORCE_Y)      \
  X(FX_FORM_C5_NUS, fx::dgp::C5_NUS)                \
  X(FX_FORM_C6_NUS, fx::dgp::C6_NUS)

#define FX_DG0_EXPRS(X)                           \
  X(FX_EXPR_LDC_RESTAU, fx::ldc::E_LDC_RESTAU)    \
  X(FX_EXPR_RES_ORTH_SQ, fx::ldc::E_RES_ORTH_SQ)  \
  X(FX_EXPR_SQRT_QT_A, fx::ldc::E_SQRT_QT_A)      \
  X(FX_EXPR_H_KAPP, fx::ldc::E_H_KAPP)            \
  X(FX_EXPR_C4_RESTAU, fx::jbp::E_C4_RESTAU)      \
  X(FX_EXPR_C5_RESTAU, fx::dgp::E_C5_RESTAU)      \
  X(FX_EXPR_C6_RESTAU, fx::dgp::E_C6_RESTAU)      \
  X(FX_EXPR_ADAPT_RES_SQ, fx::jbp::E_ADAPT_RES_SQ) \
  X(FX_EXPR_TAU_AVG, fx::jbp::E_TAU_AVG)          \
  X(FX_EXPR_ABS_KAPP_RATIO, fx::jbp::E_ABS_KAPP_RATIO) \
  X(FX_EXPR_C4M_RESTAU, fx::jbp::E_C4M_RESTAU)

namespace fx {
template <class F> constexpr bool has_facet() { return F::Facet::present; }
template <class F> constexpr int facet_qdeg() {
  if constexpr (F::Facet::present) return F::Facet::QDEG; else return -1;
}
}  // namespace fx
