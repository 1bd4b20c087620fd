// Hand-written form functors for config 3:
//   "flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py  (time loop :465-782)
// Weakly compressible extended-White-Metzner journal bearing at the script's own settings
// A_0 = k_ewm = B = 0 (:81-83): UFL folds `0.0*expr` to Zero, so phi_s == 1 and phi_ewm == 1 are
// LITERAL constants and do not raise the estimated quadrature degree: the C3_*
// functors.  The true extended White-Metzner model of comp_ewm_jpb.py (A_0 = 120, k_ewm = -0.7, B = 0.1,
// :88-90; the same loop statements, :421-732) has the non-constant
//   phi_s(theta0)        = 1 - A/(C + theta0)                                  (fenics_base.py:227-238)
//   phi_ewm(tau, theta1) = (1 - B theta1/(1 + theta1)) (|tr tau|/2 + 1e-5)^k   (fenics_base.py:240-247)
//   theta1 = T1 - T_0/(T_h - T_0)   (as written, :407 / comp_ewm_jpb.py)
// in the viscous terms of a1/L1/a2/L2, the heating term of L9 and the relaxation terms of a4/L4: the
// C3E_* functors (same structs, EWM = true), with UFL's higher degrees for L4 and L9 (10: 36 points).
// Forms shared with config 4 (same algebra once phi_s == 1): a0/L0 (C2_A0/L0), a2 (C4_A2 with A_0 = 0),
// a6/L6 (SU_RHO_*).  Helper algebra: fenics_base.py:158-294.
//   sigmacon(u,p,tau) = 2 betav Dcomp(u) - p I + ((1-betav)/We)(S(tau) - I)          (fenics_base.py:180-181)
#pragma once
#include "fx_forms_jbp.h"

namespace fx {
namespace ewm {

using namespace jbp;

FX_HD double theta1_of(double T1, const double* p) { return T1 - p[P_T0] / (p[P_TH_HOT] - p[P_T0]); }
FX_HD double phi_ewm_of(const Tau& t, double theta1, const double* p) {
  return (1.0 - p[P_B_EWM] * (theta1 / (1.0 + theta1))) * pow(0.5 * fabs(t.t0.v + t.t2.v) + 0.00001, p[P_K_EWM]);
}

FX_HD SigmaF sigma_con(const Vel& u, double p, const Tau& t, const double* par) {
  const double bv = par[P_BETAV], ce = (1.0 - bv) / par[P_WE];
  const double dv3 = (u.g00 + u.g11) / 3.0;
  SigmaF s;
  s.s00 = 2.0 * bv * (u.g00 - dv3) - p + ce * (t.t0.v - 1.0);
  s.s01 = bv * (u.g01 + u.g10) + ce * t.t1.v;
  s.s11 = 2.0 * bv * (u.g11 - dv3) - p + ce * (t.t2.v - 1.0);
  return s;
}

// a1 = lhs(v_weak12) + th DEVSSl_u12                            incompressible_benchmarkEWM.py:517-531, :415
//  (2Re rho12/dt)(u,v) + Re conv ((u0 . nabla_grad) u, v) + betav (D(u), D(v)) + (betav/6)(div u, div v)
//  + (D - grad u, R) + th 2(1-betav)(Dcomp(u), D(v))
template <bool EWM>
struct C3_A1T {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 4;
  enum { RHO, U0, U1, PHIS };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    d[RHO] = eval_q(x.m, x.s.f[F_RHO12], x.cell, b1).v;
    d[U0] = u.u0; d[U1] = u.u1;
    d[PHIS] = EWM ? phi_s_of(theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, x.p.v), x.p.v[P_A0]) : 1.0;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double bv = p.v[P_BETAV];
    w_mass(a, 2.0 * p.v[P_RE] * d[RHO] / p.v[P_DT]);
    const double c = p.v[P_RE] * p.v[P_CONV];
    a.template add<UXV, UXX>(c * d[U0]);
    a.template add<UXV, UXY>(c * d[U1]);
    a.template add<UYV, UYX>(c * d[U0]);
    a.template add<UYV, UYY>(c * d[U1]);
    w_d_d(a, bv * d[PHIS]);
    const double g6 = bv * d[PHIS] / 6.0;
    a.template add<UXX, UXX>(g6);
    a.template add<UXX, UYY>(g6);
    a.template add<UYY, UXX>(g6);
    a.template add<UYY, UYY>(g6);
    w_devss_proj<false>(a);
    w_dcomp_d(a, p.v[P_TH] * 2.0 * (1.0 - bv));
  }
};
using C3_A1 = C3_A1T<false>;
using C3E_A1 = C3_A1T<true>;

// shared point data of L1 / L2: momentum source F, stress-like tensor S tested against grad(v)
//   F = m (rho0 u0/dt - c (conv/2)(u0 . nabla_grad) u0) + ce div(S(tau0) - rho0 I) - grad p0
//       L1: m = 2Re, c = 1 (the half-step convection carries no density, :517);  L2: m = Re, c = rho0
//   S = -betav D(u0) - (betav/6) div(u0) I + gam S(D0),   gam = 2 th | 2 th (1-betav)
template <bool RHO_CONV, bool EWM>
FX_HD void ewm_momentum_rhs(const PtArgs& x, double m, double gam, double* d) {
  P2Phys b; P1Phys b1;
  p2_phys(x.xi, x.et, x.g, b);
  p1_phys(x.xi, x.et, x.g, b1);
  const double* p = x.p.v;
  const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
  const Sym3 D0 = eval_devss(x.m, x.s.f[F_W0], x.cell);
  const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
  const VG p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
  const VG rho = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1);
  const double ce = (1.0 - p[P_BETAV]) / (p[P_WE] + 3.0e-16);
  // viscous part betav phi_s(theta0) (EWM) | betav
  const double bv = p[P_BETAV] * (EWM ? phi_s_of(theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p), p[P_A0]) : 1.0);
  const double rr = rho.v / p[P_DT], hc = 0.5 * p[P_CONV] * (RHO_CONV ? rho.v : 1.0);
  d[0] = m * (rr * u.u0 - hc * (u.g00 * u.u0 + u.g01 * u.u1)) + ce * (t.t0.x + t.t1.y - rho.x) - p0.x;
  d[1] = m * (rr * u.u1 - hc * (u.g10 * u.u0 + u.g11 * u.u1)) + ce * (t.t1.x + t.t2.y - rho.y) - p0.y;
  const double dv6 = bv * (u.g00 + u.g11) / 6.0;
  d[2] = -bv * u.g00 - dv6 + gam * D0.a;
  d[3] = -bv * 0.5 * (u.g01 + u.g10) + gam * D0.b;
  d[4] = -bv * u.g11 - dv6 + gam * D0.c;
}
template <class A> FX_HDC void ewm_momentum_vterms(A& a, const double* d) {
  a.template add<UXV>(d[0]);
  a.template add<UYV>(d[1]);
  a.template add<UXX>(d[2]);
  a.template add<UXY>(d[3]);
  a.template add<UYX>(d[3]);
  a.template add<UYY>(d[4]);
}
// L1 = rhs(v_weak12) + th DEVSSr_u12,  DEVSSr_u12 = 2 (S(D0), D(v))      :517-532, :416
template <bool EWM>
struct C3_L1T {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 5;
  FX_HD static void point(const PtArgs& x, double* d) {
    ewm_momentum_rhs<false, EWM>(x, 2.0 * x.p.v[P_RE], 2.0 * x.p.v[P_TH], d);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { ewm_momentum_vterms(a, d); }
};
using C3_L1 = C3_L1T<false>;
using C3E_L1 = C3_L1T<true>;
// L2 = rhs(Fus) + th DEVSSGr_u1,  DEVSSGr_u1 = (1-betav)(D0 + D0^T, D(v))  :563-575, :491
template <bool EWM>
struct C3_L2T {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 6, NP = 5;
  FX_HD static void point(const PtArgs& x, double* d) {
    ewm_momentum_rhs<true, EWM>(x, x.p.v[P_RE], x.p.v[P_TH] * 2.0 * (1.0 - x.p.v[P_BETAV]), d);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { ewm_momentum_vterms(a, d); }
};
using C3_L2 = C3_L2T<false>;
using C3E_L2 = C3_L2T<true>;

// a5 = (Ma^2/dt)(p,q) + ((1 + al theta0) dt grad p, grad q)               :586-592
struct C3_A5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = (1.0 + x.p.v[P_AL1] * theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, x.p.v)) * x.p.v[P_DT];
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    a.template add<QV, QV>(p.v[P_MA] * p.v[P_MA] / p.v[P_DT]);
    a.template add<QX, QX>(d[0]);
    a.template add<QY, QY>(d[0]);
  }
};
// L5 = ((Ma^2/dt) p0 - Re (1 + al theta0) div(rho0 us), q) + ((1 + al theta0) dt grad p0, grad q)   :587-593
struct C3_L5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const VG p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const VG r = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1);
    const double k = 1.0 + p[P_AL1] * theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p);
    d[0] = (p[P_MA] * p[P_MA] / p[P_DT]) * p0.v
           - p[P_RE] * k * (r.x * us.u0 + r.y * us.u1 + r.v * (us.g00 + us.g11));
    d[1] = k * p[P_DT] * p0.x;
    d[2] = k * p[P_DT] * p0.y;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QV>(d[0]);
    a.template add<QX>(d[1]);
    a.template add<QY>(d[2]);
  }
};

// a7 = ((Re/dt) rho1 u, v) + (D - grad u, R)     (`a7 += 0`)               :617-623
struct C3_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, p.v[P_RE] * d[0] / p.v[P_DT]);
    w_devss_proj<false>(a);
  }
};
// L7 = ((Re/dt) rho0 us, v) - (1/2)(grad(p1 - p0), v)                      :618-624
struct C3_L7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const VG p1 = eval_q(x.m, x.s.f[F_P1], x.cell, b1), p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const double m = p[P_RE] / p[P_DT] * eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    d[0] = m * us.u0 - 0.5 * (p1.x - p0.x);
    d[1] = m * us.u1 - 0.5 * (p1.y - p0.y);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};

// a9 = ((1/dt) rho1 thetal + rho1 dot(u1, grad thetal), r) + ((Di + th (1-Di)) grad thetal, grad r)
//      thetal = T/(T_h - T_0)                                              :642-650, :420
struct C3_A9 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 4;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const double rho = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    const double idT = 1.0 / (p[P_TH_HOT] - p[P_T0]);
    d[0] = rho / p[P_DT] * idT;
    d[1] = rho * u.u0 * idT;
    d[2] = rho * u.u1 * idT;
    d[3] = (p[P_DI] + p[P_TH_T] * (1.0 - p[P_DI])) * idT;  // th_T = th (:650) | 0 (comp_ewm_jpb.py: `0.0*th*DEVSSl_T1`)
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par&) {
    a.template add<QV, QV>(d[0]);
    a.template add<QV, QX>(d[1]);
    a.template add<QV, QY>(d[2]);
    a.template add<QX, QX>(d[3]);
    a.template add<QY, QY>(d[3]);
  }
};
// L9 = ((1/dt) rho1 (thetar + theta0) + Vh inner(sigmacon(u0,p0,tau0), grad u0), r)
//      + th (1-Di)(grad theta0, grad r) - Di Bi <theta0, r>_{ds(1)}         :641-651, :421
// thetar = project(T_0/(T_h-T_0), Q) is constant: its gradient terms vanish.
template <bool EWM>
struct C3_L9T {
  using Space = SpQ; using Facet = C4_L9_Facet;
  static constexpr int QDEG = EWM ? 10 : 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u0 = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double rho1 = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    const VG T0 = eval_q(x.m, x.s.f[F_T0], x.cell, b1);
    const double idT = 1.0 / (p[P_TH_HOT] - p[P_T0]);
    const double thr = p[P_T0] * idT;
    const SigmaF s = sigma_con(u0, p0, t0, p);
    const double gamdot = s.s00 * u0.g00 + s.s01 * (u0.g01 + u0.g10) + s.s11 * u0.g11;
    const double kd = p[P_TH_T] * (1.0 - p[P_DI]) * idT;
    double heat = p[P_VH] * gamdot;
    if (EWM)  // Vh*phi_ewm(tau1, theta1, k_ewm, B)*gamdot
      heat *= phi_ewm_of(eval_tau(x.m, x.s.f[F_TAU1], x.cell, b), theta1_of(eval_q(x.m, x.s.f[F_T1], x.cell, b1).v, p), p);
    d[0] = rho1 / p[P_DT] * (thr + theta_of(T0.v, p)) + heat;
    d[1] = kd * T0.x;
    d[2] = kd * T0.y;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QV>(d[0]);
    a.template add<QX>(d[1]);
    a.template add<QY>(d[2]);
  }
};
using C3_L9 = C3_L9T<false>;
using C3E_L9 = C3_L9T<true>;

// a4 = ((We/dt + 1) tau + We FdefG(u1, S(D1), tau), Rt) + SU,                :660-672, :473-474
//      SU = (dot(u1, grad tau), alpha dot(u1, grad Rt)),  alpha = h/(|u1| + 1e-10)
// (UFL degree 12 -> 49-point collapsed Gauss-Jacobi rule)
// EWM: We -> We phi_ewm(tau0, theta1) in the relaxation / deformation terms (comp_ewm_jpb.py; the SU term is not scaled)
template <bool EWM>
struct C3_A4T {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 12, NP = 7;
  enum { U0, U1, GA, GB, GC, AL, PHE };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Sym3 D1 = eval_devss(x.m, x.s.f[F_W1], x.cell);
    d[U0] = u.u0; d[U1] = u.u1; d[GA] = D1.a; d[GB] = D1.b; d[GC] = D1.c;
    d[AL] = x.g.h / (sqrt(u.u0 * u.u0 + u.u1 * u.u1) + x.p.v[P_EPS]);
    d[PHE] = 1.0;
    if (EWM) {
      P1Phys b1;
      p1_phys(x.xi, x.et, x.g, b1);
      d[PHE] = phi_ewm_of(eval_tau(x.m, x.s.f[F_TAU0], x.cell, b), theta1_of(eval_q(x.m, x.s.f[F_T1], x.cell, b1).v, x.p.v), x.p.v);
    }
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double We = p.v[P_WE] * d[PHE], m = We / p.v[P_DT] + 1.0;
    a.template add<T0V, T0V>(m);
    a.template add<T1V, T1V>(2.0 * m);
    a.template add<T2V, T2V>(m);
    // We (M - G tau - tau G^T, S(Rt)), G = S(D1), M_jk = u_i d_k tau_ij  (as fx::jbp::C4_A4)
    const double wu0 = We * d[U0], wu1 = We * d[U1];
    a.template add<T0V, T0X>(wu0);
    a.template add<T0V, T1X>(wu1);
    a.template add<T0V, T0V>(-2.0 * We * d[GA]);
    a.template add<T0V, T1V>(-2.0 * We * d[GB]);
    a.template add<T1V, T0Y>(wu0);
    a.template add<T1V, T1Y>(wu1);
    a.template add<T1V, T1X>(wu0);
    a.template add<T1V, T2X>(wu1);
    a.template add<T1V, T0V>(-2.0 * We * d[GB]);
    a.template add<T1V, T1V>(-2.0 * We * (d[GA] + d[GC]));
    a.template add<T1V, T2V>(-2.0 * We * d[GB]);
    a.template add<T2V, T1Y>(wu0);
    a.template add<T2V, T2Y>(wu1);
    a.template add<T2V, T1V>(-2.0 * We * d[GB]);
    a.template add<T2V, T2V>(-2.0 * We * d[GC]);
    // SU: sum_jk N_jk(Rt) alpha M_jk(tau),  M_00 = u0 t0,x + u1 t1,x   M_01 = u0 t0,y + u1 t1,y
    //                                       M_10 = u0 t1,x + u1 t2,x   M_11 = u0 t1,y + u1 t2,y
    const double a00 = d[AL] * d[U0] * d[U0], a01 = d[AL] * d[U0] * d[U1], a11 = d[AL] * d[U1] * d[U1];
    a.template add<T0X, T0X>(a00);
    a.template add<T0X, T1X>(a01);
    a.template add<T1X, T0X>(a01);
    a.template add<T1X, T1X>(a11 + a00);
    a.template add<T1X, T2X>(a01);
    a.template add<T2X, T1X>(a01);
    a.template add<T2X, T2X>(a11);
    a.template add<T0Y, T0Y>(a00);
    a.template add<T0Y, T1Y>(a01);
    a.template add<T1Y, T0Y>(a01);
    a.template add<T1Y, T1Y>(a11 + a00);
    a.template add<T1Y, T2Y>(a01);
    a.template add<T2Y, T1Y>(a01);
    a.template add<T2Y, T2Y>(a11);
  }
};
using C3_A4 = C3_A4T<false>;
using C3E_A4 = C3_A4T<true>;
// L4 = ((We/dt) S(tau0) + rho0 I, S(Rt))                                     :663-667
template <bool EWM>
struct C3_L4T {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = EWM ? 10 : 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double rho = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    double wd = x.p.v[P_WE] / x.p.v[P_DT];
    if (EWM) wd *= phi_ewm_of(t, theta1_of(eval_q(x.m, x.s.f[F_T1], x.cell, b1).v, x.p.v), x.p.v);
    d[0] = wd * t.t0.v + rho;
    d[1] = 2.0 * wd * t.t1.v;
    d[2] = wd * t.t2.v + rho;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};
using C3_L4 = C3_L4T<false>;
using C3E_L4 = C3_L4T<true>;

// journal load and torque on ds(0) with sigmacon(u1,p1,tau1) (:690-707); WHICH as fx::jbp::C4_JOURNAL
template <int WHICH>
struct C3_JOURNAL {
  static constexpr int QDEG = 2, MARKERS = 1 << 0;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const SigmaF s = sigma_con(u, eval_q(x.m, x.s.f[F_P1], x.cell, b1).v, t, x.p.v);
    const double n0 = x.nx, n1 = x.ny;
    if (WHICH == 0) return -(n0 * (s.s00 * n1 - s.s01 * n0) + n1 * (s.s01 * n1 - s.s11 * n0));
    if (WHICH == 1) return s.s00 * n0 + s.s01 * n1;
    return -(s.s01 * n0 + s.s11 * n1);
  }
};

}  // namespace ewm
}  // namespace fx
