// Form registry: maps the public form ids (include/fenics_b200.h) to the hand-written functors.
// X(id, Functor).  Used by fx_kernels.cu (CUDA launchers) and tests/hostemu (CPU emulation).
#pragma once
#include "fx_forms_ldc.h"

// forms whose quadrature degree depends on conv == 0.0 (UFL folds `0.0*expr` to Zero,
// incomp_LDC.py:93-94) are listed with both instantiations: XC(id, FunctorIfConvZero, FunctorElse)
#define FX_MATRIX_FORMS(X)                  \
  X(FX_FORM_MASS_ZC, fx::ldc::MASS_ZC)      \
  X(FX_FORM_MASS_Q, fx::ldc::MASS_Q)        \
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
  X(FX_FORM_C2_A4, fx::ldc::LDC_A4)

#define FX_VECTOR_FORMS(X, XC)                              \
  XC(FX_FORM_C1_L1, fx::ldc::C1_L1<4>, fx::ldc::C1_L1<5>)   \
  XC(FX_FORM_C1_L2, fx::ldc::C1_L2<4>, fx::ldc::C1_L2<5>)   \
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
  X(FX_FORM_C2_L_RHO1, fx::ldc::C2_L_RHO1)

#define FX_SCALAR_FORMS(X)              \
  X(FX_FORM_C1_EK, fx::ldc::C1_EK)      \
  X(FX_FORM_C1_EE, fx::ldc::LDC_EE)     \
  X(FX_FORM_C2_EK, fx::ldc::C2_EK)      \
  X(FX_FORM_C2_EE, fx::ldc::LDC_EE)

#define FX_DG0_EXPRS(X)                           \
  X(FX_EXPR_LDC_RESTAU, fx::ldc::E_LDC_RESTAU)    \
  X(FX_EXPR_RES_ORTH_SQ, fx::ldc::E_RES_ORTH_SQ)  \
  X(FX_EXPR_SQRT_QT_A, fx::ldc::E_SQRT_QT_A)      \
  X(FX_EXPR_H_KAPP, fx::ldc::E_H_KAPP)

namespace fx {
template <class F> constexpr bool has_facet() { return F::Facet::present; }
template <class F> constexpr int facet_qdeg() {
  if constexpr (F::Facet::present) return F::Facet::QDEG; else return -1;
}
}  // namespace fx
