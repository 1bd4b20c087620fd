// Hand-written form functors for the lid-driven-cavity scripts:
//   cfg1 = lid-driven-cavity/incomp_LDC.py            (time loop :280-447)
//   cfg2 = lid-driven-cavity/comp_LDC_mesh_convergence.py (time loop :376-565)
// Each functor states the UFL form it implements and the reference lines.  Notation:
//   G = grad(u): Gij = d u_i / d x_j;  S(a) = [[a0,a1],[a1,a2]] (reshape_elements, fenics_fem.py:210-225)
//   Dincomp(u) = (G+G^T)/2;  Dcomp(u) = Dincomp(u) - (1/3) div(u) I   (fenics_fem.py:297-302)
//   UFL dot(u, grad(tau))[j,k] = sum_i u_i d tau_ij / d x_k  (contracts u with the FIRST index of
//   tau -- reproduced as written in the reference, fenics_fem.py:310-314).
#pragma once
#include "fx_assemble.h"

namespace fx {
namespace ldc {

// parameter / field slots (values must match include/fenics_b200.h)
enum : int { P_DT = 0, P_RE, P_WE, P_BETAV, P_C1, P_C2, P_TH, P_CONV, P_MA };
enum : int { F_W0 = 0, F_W12, F_WS, F_W1, F_P0, F_P1, F_RHO0, F_RHO12, F_RHO1, F_T0, F_T1, F_THETA0,
             F_THERM_INV, F_TAU0, F_TAU12, F_TAU1, F_KAPP, F_RES_TEST, F_RES_ORTH, F_QT_A, F_Q_A };

// W slots: ux {value, d/dx, d/dy}, uy {...}, D0, D1, D2
enum : int { UXV = 0, UXX, UXY, UYV, UYX, UYY, WD0, WD1, WD2 };
// Zc slots: tau0 {v,x,y}, tau1 {v,x,y}, tau2 {v,x,y}
enum : int { T0V = 0, T0X, T0Y, T1V, T1X, T1Y, T2V, T2X, T2Y };
// Q slots
enum : int { QV = 0, QX, QY };

// ---- shared bilinear building blocks on W x W ---------------------------------------------------
template <class A> FX_HDC void w_mass(A& a, double m) {  // m (u, v)
  a.template add<UXV, UXV>(m);
  a.template add<UYV, UYV>(m);
}
template <class A> FX_HDC void w_gradgrad(A& a, double b) {  // b (grad u, grad v)
  a.template add<UXX, UXX>(b);
  a.template add<UXY, UXY>(b);
  a.template add<UYX, UYX>(b);
  a.template add<UYY, UYY>(b);
}
// (S(D) - Dincomp(u), S(R))   [COMP=false]   or   (S(D) - Dcomp(u), S(R))   [COMP=true]
template <bool COMP, class A> FX_HDC void w_devss_proj(A& a) {
  a.template add<WD0, WD0>(1.0);
  a.template add<WD1, WD1>(2.0);
  a.template add<WD2, WD2>(1.0);
  a.template add<WD1, UXY>(-1.0);
  a.template add<WD1, UYX>(-1.0);
  if constexpr (COMP) {
    a.template add<WD0, UXX>(-2.0 / 3);
    a.template add<WD0, UYY>(1.0 / 3);
    a.template add<WD2, UXX>(1.0 / 3);
    a.template add<WD2, UYY>(-2.0 / 3);
  } else {
    a.template add<WD0, UXX>(-1.0);
    a.template add<WD2, UYY>(-1.0);
  }
}
// c (Dcomp(u), Dincomp(v))
template <class A> FX_HDC void w_dcomp_d(A& a, double c) {
  a.template add<UXX, UXX>(c * (2.0 / 3));
  a.template add<UXX, UYY>(c * (-1.0 / 3));
  a.template add<UYY, UYY>(c * (2.0 / 3));
  a.template add<UYY, UXX>(c * (-1.0 / 3));
  a.template add<UXY, UXY>(0.5 * c);
  a.template add<UXY, UYX>(0.5 * c);
  a.template add<UYX, UXY>(0.5 * c);
  a.template add<UYX, UYX>(0.5 * c);
}

// ---- cfg1 bilinear forms on W (constant coefficients) -------------------------------------------
// a1 = (2Re/dt)(u,v) + betav (grad u, grad v) + (D - Dincomp(u), R) + th 2(1-betav)(Dcomp(u), Dincomp(v))
//      incomp_LDC.py:317-325, :260
struct C1_A1 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    w_mass(a, 2.0 * p.v[P_RE] / p.v[P_DT]);
    w_gradgrad(a, p.v[P_BETAV]);
    w_devss_proj<false>(a);
    w_dcomp_d(a, p.v[P_TH] * 2.0 * (1.0 - p.v[P_BETAV]));
  }
};
// a2 = (Re/dt)(u,v) + (D - Dincomp(u), R)       incomp_LDC.py:355-358
struct C1_A2 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    w_mass(a, p.v[P_RE] / p.v[P_DT]);
    w_devss_proj<false>(a);
  }
};
// a7 = (Re/dt)(u,v) + (betav/2)(grad u, grad v) + (D - Dincomp(u), R)     incomp_LDC.py:379-384
struct C1_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    w_mass(a, p.v[P_RE] / p.v[P_DT]);
    w_gradgrad(a, 0.5 * p.v[P_BETAV]);
    w_devss_proj<false>(a);
  }
};
// a5 = (grad p, grad q)     incomp_LDC.py:370
struct C1_A5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 0, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par&) {
    a.template add<QX, QX>(1.0);
    a.template add<QY, QY>(1.0);
  }
};

// ---- stress matrix (cfg1 a4 incomp_LDC.py:400-409 + LPS :309; cfg2 a4 :528-536 + :402) ----------
// a4 = ((We/dt+1) tau + We F(u1,tau), Rt) + c1 (kapp h grad tau, grad Rt) + c2 (kapp h div tau, div Rt)
// F(u,tau) = dot(u,grad(tau)) - dot(grad(u),tau) - dot(tau,grad(u)^T)
struct LDC_A4 {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 7;
  enum { U0, U1, G00, G01, G10, G11, KH };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    d[U0] = u.u0; d[U1] = u.u1; d[G00] = u.g00; d[G01] = u.g01; d[G10] = u.g10; d[G11] = u.g11;
    d[KH] = eval_dg0(x.s.f[F_KAPP], x.m.dm[S_QT], x.m.nc, x.cell, 0) * x.g.h;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double We = p.v[P_WE], m = We / p.v[P_DT] + 1.0;
    // (m S(tau), S(Rt)) = m (t0 r0 + 2 t1 r1 + t2 r2)
    a.template add<T0V, T0V>(m);
    a.template add<T1V, T1V>(2.0 * m);
    a.template add<T2V, T2V>(m);
    // We (F, S(Rt)) = We [F00 r0 + (F01+F10) r1 + F11 r2]
    //   M_jk = u_i d_k tau_ij:  M00 = u0 t0,x + u1 t1,x   M01 = u0 t0,y + u1 t1,y
    //                           M10 = u0 t1,x + u1 t2,x   M11 = u0 t1,y + u1 t2,y
    //   F00 = M00 - 2(G00 t0 + G01 t1);  F11 = M11 - 2(G10 t1 + G11 t2)
    //   F01+F10 = M01 + M10 - 2(G10 t0 + (G00+G11) t1 + G01 t2)
    const double wu0 = We * d[U0], wu1 = We * d[U1];
    a.template add<T0V, T0X>(wu0);
    a.template add<T0V, T1X>(wu1);
    a.template add<T0V, T0V>(-2.0 * We * d[G00]);
    a.template add<T0V, T1V>(-2.0 * We * d[G01]);
    a.template add<T1V, T0Y>(wu0);
    a.template add<T1V, T1Y>(wu1);
    a.template add<T1V, T1X>(wu0);
    a.template add<T1V, T2X>(wu1);
    a.template add<T1V, T0V>(-2.0 * We * d[G10]);
    a.template add<T1V, T1V>(-2.0 * We * (d[G00] + d[G11]));
    a.template add<T1V, T2V>(-2.0 * We * d[G01]);
    a.template add<T2V, T1Y>(wu0);
    a.template add<T2V, T2Y>(wu1);
    a.template add<T2V, T1V>(-2.0 * We * d[G10]);
    a.template add<T2V, T2V>(-2.0 * We * d[G11]);
    // LPS: c1 kh (grad S(tau), grad S(Rt)) -- off-diagonal component counted twice
    const double c1 = p.v[P_C1] * d[KH], c2 = p.v[P_C2] * d[KH];
    a.template add<T0X, T0X>(c1);
    a.template add<T0Y, T0Y>(c1);
    a.template add<T1X, T1X>(2.0 * c1);
    a.template add<T1Y, T1Y>(2.0 * c1);
    a.template add<T2X, T2X>(c1);
    a.template add<T2Y, T2Y>(c1);
    // c2 kh (div S(tau), div S(Rt)), div S(a) = (a0,x + a1,y, a1,x + a2,y)
    a.template add<T0X, T0X>(c2);
    a.template add<T0X, T1Y>(c2);
    a.template add<T1Y, T0X>(c2);
    a.template add<T1Y, T1Y>(c2);
    a.template add<T1X, T1X>(c2);
    a.template add<T1X, T2Y>(c2);
    a.template add<T2Y, T1X>(c2);
    a.template add<T2Y, T2Y>(c2);
  }
};

// ---- mass matrices of dolfin.project -----------------------------------------------------------
struct MASS_ZC {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par&) {
    a.template add<T0V, T0V>(1.0);
    a.template add<T1V, T1V>(1.0);
    a.template add<T2V, T2V>(1.0);
  }
};
struct MASS_Q {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par&) { a.template add<QV, QV>(1.0); }
};

// ---- cfg1 linear forms --------------------------------------------------------------------------
// L1 = ((2Re/dt) u0 - Re conv grad(u0) u0, v) + (p0, div v) - (S(tau0), grad v)
//      + th 2(1-betav)(S(D0), Dincomp(v))                       incomp_LDC.py:319-326, :313
// QDEG 4 when conv == 0.0 (the term folds to Zero, incomp_LDC.py:93-94) else 5.
template <int Q>
struct C1_L1 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = Q, NP = 5;
  enum { F0, F1, SXX, SXY, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Sym3 D0 = eval_devss(x.m, x.s.f[F_W0], x.cell);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double m = 2.0 * p[P_RE] / p[P_DT], rc = p[P_RE] * p[P_CONV];
    const double gam = p[P_TH] * 2.0 * (1.0 - p[P_BETAV]);
    d[F0] = m * u.u0 - rc * (u.g00 * u.u0 + u.g01 * u.u1);
    d[F1] = m * u.u1 - rc * (u.g10 * u.u0 + u.g11 * u.u1);
    d[SXX] = p0 - t.t0.v + gam * D0.a;
    d[SXY] = -t.t1.v + gam * D0.b;
    d[SYY] = p0 - t.t2.v + gam * D0.c;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[F0]);
    a.template add<UYV>(d[F1]);
    a.template add<UXX>(d[SXX]);
    a.template add<UXY>(d[SXY]);
    a.template add<UYX>(d[SXY]);
    a.template add<UYY>(d[SYY]);
  }
};
// L2 = ((Re/dt) u0 - Re conv grad(u12) u12, v) - (betav/2)(grad u0, grad v) + (p0, div v)
//      - (S(tau0), grad v)                                       incomp_LDC.py:356-359
template <int Q>
struct C1_L2 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = Q, NP = 6;
  enum { F0, F1, SXX, SXY, SYX, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Vel h = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double m = p[P_RE] / p[P_DT], rc = p[P_RE] * p[P_CONV], hb = 0.5 * p[P_BETAV];
    d[F0] = m * u.u0 - rc * (h.g00 * h.u0 + h.g01 * h.u1);
    d[F1] = m * u.u1 - rc * (h.g10 * h.u0 + h.g11 * h.u1);
    d[SXX] = -hb * u.g00 + p0 - t.t0.v;
    d[SXY] = -hb * u.g01 - t.t1.v;
    d[SYX] = -hb * u.g10 - t.t1.v;
    d[SYY] = -hb * u.g11 + p0 - t.t2.v;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[F0]);
    a.template add<UYV>(d[F1]);
    a.template add<UXX>(d[SXX]);
    a.template add<UXY>(d[SXY]);
    a.template add<UYX>(d[SYX]);
    a.template add<UYY>(d[SYY]);
  }
};
// L5 = (grad p0, grad q) + (Re/dt)(us, grad q)                    incomp_LDC.py:371
struct C1_L5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const VG p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const double m = x.p.v[P_RE] / x.p.v[P_DT];
    d[0] = p0.x + m * us.u0;
    d[1] = p0.y + m * us.u1;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QX>(d[0]);
    a.template add<QY>(d[1]);
  }
};
// L7 = ((Re/dt) us, v) + (1/2)(p1 - p0, div v)                    incomp_LDC.py:382-385
struct C1_L7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const double m = x.p.v[P_RE] / x.p.v[P_DT];
    d[0] = m * us.u0;
    d[1] = m * us.u1;
    d[2] = 0.5 * (eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
    a.template add<UXX>(d[2]);
    a.template add<UYY>(d[2]);
  }
};
// L4 = ((We/dt) S(tau0) + 2(1-betav) E(u), S(Rt)), E = Dincomp(u12) in cfg1 (incomp_LDC.py:402-405),
//      E = Dcomp(u1) in cfg2 (comp_LDC_mesh_convergence.py:529-533)
template <bool COMP>
struct LDC_L4 {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[COMP ? F_W1 : F_W12], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double wd = p[P_WE] / p[P_DT], c = 2.0 * (1.0 - p[P_BETAV]);
    const double tr3 = COMP ? (u.g00 + u.g11) / 3.0 : 0.0;
    d[0] = wd * t.t0.v + c * (u.g00 - tr3);
    d[1] = 2.0 * wd * t.t1.v + c * (u.g01 + u.g10);
    d[2] = wd * t.t2.v + c * (u.g11 - tr3);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};

// restau0 = (We/dt)(tau1v - tau0v) + We [F00, F10, F11] + tau1v - 2(1-betav)[Dc00, Dc10, Dc11],
//   F = Fdefcon(u1, tau1) = dot(u1,grad(tau1)) - grad(u1) tau1 - tau1 grad(u1)^T + div(u1) tau1
//   incomp_LDC.py:300-303 == comp_LDC_mesh_convergence.py:393-396
FX_HD void ldc_restau(const PtArgs& x, double* r) {
  P2Phys b;
  p2_phys(x.xi, x.et, x.g, b);
  const double* p = x.p.v;
  const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
  const Tau t1 = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
  const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
  const double We = p[P_WE], wd = We / p[P_DT], c = 2.0 * (1.0 - p[P_BETAV]);
  const double dv = u.g00 + u.g11;
  const double a0 = t1.t0.v, a1 = t1.t1.v, a2 = t1.t2.v;
  const double F00 = u.u0 * t1.t0.x + u.u1 * t1.t1.x - 2.0 * (u.g00 * a0 + u.g01 * a1) + dv * a0;
  const double F10 = u.u0 * t1.t1.x + u.u1 * t1.t2.x - (u.g10 * a0 + u.g11 * a1) - (a1 * u.g00 + a2 * u.g01)
                     + dv * a1;
  const double F11 = u.u0 * t1.t1.y + u.u1 * t1.t2.y - 2.0 * (u.g10 * a1 + u.g11 * a2) + dv * a2;
  r[0] = wd * (a0 - t0.t0.v) + We * F00 + a0 - c * (u.g00 - dv / 3.0);
  r[1] = wd * (a1 - t0.t1.v) + We * F10 + a1 - c * 0.5 * (u.g01 + u.g10);
  r[2] = wd * (a2 - t0.t2.v) + We * F11 + a2 - c * (u.g11 - dv / 3.0);
}
// L = inner(w, restau0 - res_test) on Zc (RHS of project(restau0-res_test, Zc), incomp_LDC.py:305)
struct LDC_L_RES_ORTH {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    ldc_restau(x, d);
    for (int j = 0; j < 3; ++j) d[j] -= eval_dg0(x.s.f[F_RES_TEST], x.m.dm[S_ZD], x.m.nc, x.cell, j);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};

// ---- DG0 projections -----------------------------------------------------------------------------
struct E_LDC_RESTAU {  // project(restau0, Zd)  incomp_LDC.py:304
  static constexpr int QDEG = 3, NOUT = 3;
  FX_HD static void value(const PtArgs& x, double* v) { ldc_restau(x, v); }
};
struct E_RES_ORTH_SQ {  // project(inner(res_orth,res_orth), Qt)  :306
  static constexpr int QDEG = 4, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_RES_ORTH], x.cell, b);
    v[0] = t.t0.v * t.t0.v + t.t1.v * t.t1.v + t.t2.v * t.t2.v;
  }
};
struct E_SQRT_QT_A {  // project(np.power(res_orth_norm_sq, 0.5), Qt)  :307-308
  static constexpr int QDEG = 2, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
    v[0] = sqrt(eval_dg0(x.s.f[F_QT_A], x.m.dm[S_QT], x.m.nc, x.cell, 0));
  }
};
struct E_H_KAPP {  // project(h*kapp, Qt)  :434
  static constexpr int QDEG = 0, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
    v[0] = x.g.h * eval_dg0(x.s.f[F_KAPP], x.m.dm[S_QT], x.m.nc, x.cell, 0);
  }
};

// ---- functionals -----------------------------------------------------------------------------------
struct C1_EK {  // 0.5*dot(u1,u1)*dx  incomp_LDC.py:420
  static constexpr int QDEG = 4;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    return 0.5 * (u.u0 * u.u0 + u.u1 * u.u1);
  }
};
struct LDC_EE {  // (tau1[0,0]+tau1[1,1])*dx  incomp_LDC.py:421, comp_LDC_mesh_convergence.py:549
  static constexpr int QDEG = 2;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    return t.t0.v + t.t2.v;
  }
};
struct C2_EK {  // 0.5*rho1*dot(u1,u1)*dx  comp_LDC_mesh_convergence.py:548
  static constexpr int QDEG = 5;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    return 0.5 * eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v * (u.u0 * u.u0 + u.u1 * u.u1);
  }
};

// ---- cfg2 ---------------------------------------------------------------------------------------
// a0 = (2/dt)(rho, q)                                             comp_LDC_mesh_convergence.py:417-420
struct C2_A0 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    a.template add<QV, QV>(2.0 / p.v[P_DT]);
  }
};
// L0 = ((2/dt) rho0 - dot(u0, grad rho0) + rho0 div u0, q)         :417-421
struct C2_L0 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 3, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const VG r = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1);
    d[0] = (2.0 / x.p.v[P_DT]) * r.v - (u.u0 * r.x + u.u1 * r.y) + r.v * (u.g00 + u.g11);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};

// Facet part shared by a1/a2:  -betav <nabla_grad(u/2) n, v> - (betav/3) <div(u/2) n, v>
struct C2_A12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int MARKERS = -1;  // un-indexed ds: every exterior facet
  static constexpr int QDEG = 3, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) { d[0] = x.nx; d[1] = x.ny; }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double hb = -0.5 * p.v[P_BETAV], sb = -p.v[P_BETAV] / 6.0;
    // (nabla_grad(u) n)_k = sum_i d_k u_i n_i, tested with v_k
    a.template add<UXV, UXX>(hb * d[0]);
    a.template add<UXV, UYX>(hb * d[1]);
    a.template add<UYV, UXY>(hb * d[0]);
    a.template add<UYV, UYY>(hb * d[1]);
    // div(u) n . v
    a.template add<UXV, UXX>(sb * d[0]);
    a.template add<UXV, UYY>(sb * d[0]);
    a.template add<UYV, UXX>(sb * d[1]);
    a.template add<UYV, UYY>(sb * d[1]);
  }
};
// a1 = lhs(Fu12) + th DEVSSl_u12  (:430-440):  (2 Re rho0/dt)(u,v) + betav (Dcomp(u), Dincomp(v))
//      + (D - Dcomp(u), R) + th 2(1-betav)(Dcomp(u), Dincomp(v)) + facet part
// a2 = lhs(Fus) + th DEVSSl_u1    (:458-469):  same with (Re rho0/dt)
template <bool HALF>
struct C2_A12 {
  using Space = SpW; using Facet = C2_A12_Facet;
  static constexpr int QDEG = 5, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, (HALF ? 2.0 : 1.0) * p.v[P_RE] * d[0] / p.v[P_DT]);
    w_dcomp_d(a, p.v[P_BETAV] + p.v[P_TH] * 2.0 * (1.0 - p.v[P_BETAV]));
    w_devss_proj<true>(a);
  }
};
// Facet part of L1/L2: -c_p <p0 n, v> + (betav/2)<nabla_grad(u0) n, v> + (betav/6)<div(u0) n, v> + <S(tau0) n, v>
//   c_p = 1 (L1, :433) or 1/2 (L2, :461)
template <bool HALF>
struct C2_L12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int MARKERS = -1;
  static constexpr int QDEG = 4, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double cp = HALF ? 1.0 : 0.5, hb = 0.5 * p[P_BETAV], sb = p[P_BETAV] / 6.0;
    const double dv = u.g00 + u.g11;
    d[0] = -cp * p0 * x.nx + hb * (u.g00 * x.nx + u.g10 * x.ny) + sb * dv * x.nx + (t.t0.v * x.nx + t.t1.v * x.ny);
    d[1] = -cp * p0 * x.ny + hb * (u.g01 * x.nx + u.g11 * x.ny) + sb * dv * x.ny + (t.t1.v * x.nx + t.t2.v * x.ny);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};
// L1 = rhs(Fu12) + th DEVSSr_u12 (:430-441, :412);  L2 = rhs(Fus) + th DEVSSr_u1 (:458-470, :454)
//  cell part: (Re rho0 [c u0/dt - conv (w . nabla_grad) w], v) + (Sigma, Dincomp(v)),  c=2,w=u0 | c=1,w=u12
//  Sigma = -betav Dcomp(u0) + p0 I - ((1-betav)/We)(S(tau0) - I) + th 2(1-betav) S(D0 | D12)
template <bool HALF>
struct C2_L12 {
  using Space = SpW; using Facet = C2_L12_Facet<HALF>;
  static constexpr int QDEG = 6, NP = 5;
  enum { F0, F1, SXX, SXY, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Vel w = HALF ? u : eval_vel(x.m, x.s.f[F_W12], x.cell, b);
    const Sym3 D = eval_devss(x.m, x.s.f[HALF ? F_W0 : F_W12], x.cell);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double rho = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    const double m = (HALF ? 2.0 : 1.0) / p[P_DT], conv = p[P_CONV];
    const double bv = p[P_BETAV], ce = (1.0 - bv) / p[P_WE], gam = p[P_TH] * 2.0 * (1.0 - bv);
    // dot(w, nabla_grad(w))_i = w_k d_k w_i
    d[F0] = p[P_RE] * rho * (m * u.u0 - conv * (w.g00 * w.u0 + w.g01 * w.u1));
    d[F1] = p[P_RE] * rho * (m * u.u1 - conv * (w.g10 * w.u0 + w.g11 * w.u1));
    const double dv3 = (u.g00 + u.g11) / 3.0;
    d[SXX] = -bv * (u.g00 - dv3) + p0 - ce * (t.t0.v - 1.0) + gam * D.a;
    d[SXY] = -bv * 0.5 * (u.g01 + u.g10) - ce * t.t1.v + gam * D.b;
    d[SYY] = -bv * (u.g11 - dv3) + p0 - ce * (t.t2.v - 1.0) + gam * D.c;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[F0]);
    a.template add<UYV>(d[F1]);
    a.template add<UXX>(d[SXX]);
    a.template add<UXY>(d[SXY]);
    a.template add<UYX>(d[SXY]);
    a.template add<UYY>(d[SYY]);
  }
};
// a5 = (Ma^2/(dt Re))(p,q) + (dt/2)(grad p, grad q)                :483-489
struct C2_A5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    a.template add<QV, QV>(p.v[P_MA] * p.v[P_MA] / (p.v[P_DT] * p.v[P_RE]));
    a.template add<QX, QX>(0.5 * p.v[P_DT]);
    a.template add<QY, QY>(0.5 * p.v[P_DT]);
  }
};
// L5 = ((Ma^2/(dt Re)) p0 - rho0 div us + dot(grad rho0, us), q) + (dt/2)(grad p0, grad q)   :484-490
struct C2_L5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 3, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const VG p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const VG r = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1);
    d[0] = (p[P_MA] * p[P_MA] / (p[P_DT] * p[P_RE])) * p0.v - r.v * (us.g00 + us.g11) + (r.x * us.u0 + r.y * us.u1);
    d[1] = 0.5 * p[P_DT] * p0.x;
    d[2] = 0.5 * p[P_DT] * p0.y;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QV>(d[0]);
    a.template add<QX>(d[1]);
    a.template add<QY>(d[2]);
  }
};
// RHS of rho1 = project(rho0 + (Ma^2/Re)(p1-p0), Q)                :501-502
struct C2_L_RHO1 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    d[0] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v
           + (p[P_MA] * p[P_MA] / p[P_RE]) * (eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// a7 = ((Re/dt) rho1 u, v) + (D - Dcomp(u), R)                     :506-509
struct C2_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, p.v[P_RE] * d[0] / p.v[P_DT]);
    w_devss_proj<true>(a);
  }
};
// L7 = ((Re/dt) rho1 us, v) + (1/2)(p1-p0, div v) - (1/2) <p1 n, v>   :507-510
struct C2_L7_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int MARKERS = -1;  // un-indexed ds: every exterior facet
  static constexpr int QDEG = 3, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const double p1 = eval_q(x.m, x.s.f[F_P1], x.cell, b1).v;
    d[0] = -0.5 * p1 * x.nx;
    d[1] = -0.5 * p1 * x.ny;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};
struct C2_L7 {
  using Space = SpW; using Facet = C2_L7_Facet;
  static constexpr int QDEG = 5, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const double m = x.p.v[P_RE] / x.p.v[P_DT] * eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    d[0] = m * us.u0;
    d[1] = m * us.u1;
    d[2] = 0.5 * (eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
    a.template add<UXX>(d[2]);
    a.template add<UYY>(d[2]);
  }
};

// ---- post-loop diagnostics: stream function on the scalar P2 space V.sub(0).collapse() -------------------
// (lid-driven-cavity/fenics_fem.py:340-378; slots of the single P2 component: QV, QX, QY)
struct STREAM_A {  // a = inner(grad(psi), grad(phi))*dx
  using Space = SpVs; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par&) {
    a.template add<QX, QX>(1.0);
    a.template add<QY, QY>(1.0);
  }
};
// L = inner(u[1].dx(0) - u[0].dx(1), phi)*dx   [COMP: rho1 times each derivative, fenics_fem.py:371]
template <bool COMP>
struct STREAM_L {
  using Space = SpVs; using Facet = NoFacet;
  static constexpr int QDEG = COMP ? 4 : 3, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    double rho = 1.0;
    if (COMP) {
      P1Phys b1;
      p1_phys(x.xi, x.et, x.g, b1);
      rho = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    }
    d[0] = rho * (u.g10 - u.g01);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};

// shear_rate(u) (lid-driven-cavity/fenics_fem.py:380-398): L = inner(g, phi)*dx, g = inner(0.5*gamma, gamma),
// gamma = grad(u) + grad(u)^T  ->  g = 2 u0,x^2 + (u0,y + u1,x)^2 + 2 u1,y^2;  same matrix as stream_function
struct SHEAR_L {
  using Space = SpVs; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const double s = u.g01 + u.g10;
    d[0] = 2.0 * u.g00 * u.g00 + s * s + 2.0 * u.g11 * u.g11;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};

}  // namespace ldc
}  // namespace fx
