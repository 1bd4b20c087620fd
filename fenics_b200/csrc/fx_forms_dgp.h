// Hand-written form functors for config 5:
//   buoyancy-driven-flow/compressible_dgp.py  (time loop :368-595)
// Compressible buoyancy-driven Oldroyd-B flow with temperature coupling.  Helper algebra:
// buoyancy-driven-flow/fenics_fem.py:52-87 (sigmacom is Prandtl-scaled, Fdefcom has + div(u) Tau).
//   sigmacom(u,p,tau) = 2 Pr betav Dcomp(u) - p I + Pr ((1-betav)/We)(S(tau) - I)
//   therm = al_1 + al_2 theta0,  theta0 = project((T0-T_0)/(T_h-T_0), Q),  therm_inv = project(1/therm, Q)
// Boundary markers (:163-169): 1 left, 2 right, 3 top, 4 bottom.
#pragma once
#include "fx_forms_jbp.h"

namespace fx {
namespace dgp {

using namespace jbp;  // slot / parameter enums and shared building blocks

FX_HD double therm_of(double theta0, const double* p) { return p[P_AL1] + p[P_AL2] * theta0; }

// Facet part shared by a1/a2:  -betav Pr (<nabla_grad(u/2) n, v> + (1/3) <div(u/2) n, v>)     :435, :476
struct C5_A12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int QDEG = 3, NP = 2, MARKERS = -1;
  FX_HD static void point(const PtArgs& x, double* d) { d[0] = x.nx; d[1] = x.ny; }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double hb = -0.5 * p.v[P_BETAV] * p.v[P_PR], sb = -p.v[P_BETAV] * p.v[P_PR] / 6.0;
    a.template add<UXV, UXX>(hb * d[0]);
    a.template add<UXV, UYX>(hb * d[1]);
    a.template add<UYV, UXY>(hb * d[0]);
    a.template add<UYV, UYY>(hb * d[1]);
    a.template add<UXV, UXX>(sb * d[0]);
    a.template add<UXV, UYY>(sb * d[0]);
    a.template add<UYV, UXX>(sb * d[1]);
    a.template add<UYV, UYY>(sb * d[1]);
  }
};
// a1 = lhs(Fu12) + th DEVSSl_u12 (:432-441):  (2 rho0/dt)(u,v) + Pr betav (Dcomp(u), D(v)) + (D - Dincomp(u), R)
//      + th 2(1-betav)(Dcomp(u), D(v)) + facet part
// a2 = lhs(Fus) + th DEVSSl_u1 (:473-482):   (rho0/dt)(u,v) + (rho0/2)((u12 . nabla_grad) u, v) + the same
template <bool HALF>
struct C5_A12 {
  using Space = SpW; using Facet = C5_A12_Facet;
  static constexpr int QDEG = HALF ? 5 : 6, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    d[1] = 0.0; d[2] = 0.0;
    if (!HALF) {
      const Vel u = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
      d[1] = u.u0; d[2] = u.u1;
    }
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, (HALF ? 2.0 : 1.0) * d[0] / p.v[P_DT]);
    if constexpr (!HALF) {
      const double c = 0.5 * d[0];
      a.template add<UXV, UXX>(c * d[1]);
      a.template add<UXV, UXY>(c * d[2]);
      a.template add<UYV, UYX>(c * d[1]);
      a.template add<UYV, UYY>(c * d[2]);
    }
    w_dcomp_d(a, p.v[P_PR] * p.v[P_BETAV] + p.v[P_TH] * 2.0 * (1.0 - p.v[P_BETAV]));
    w_devss_proj<false>(a);
  }
};
// Facet part of L1/L2: -<p0 n, v> + (betav Pr/2)<nabla_grad(u0) n, v> + (betav Pr/6)<div(u0) n, v>
//                      + (Pr (1-betav)/We) <S(tau0) n, v>
struct C5_L12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int QDEG = 4, NP = 2, MARKERS = -1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double bp = p[P_BETAV] * p[P_PR], hb = 0.5 * bp, sb = bp / 6.0, ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    const double dv = u.g00 + u.g11;
    d[0] = -p0 * x.nx + hb * (u.g00 * x.nx + u.g10 * x.ny) + sb * dv * x.nx + ce * (t.t0.v * x.nx + t.t1.v * x.ny);
    d[1] = -p0 * x.ny + hb * (u.g01 * x.nx + u.g11 * x.ny) + sb * dv * x.ny + ce * (t.t1.v * x.nx + t.t2.v * x.ny);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};
// L1 = rhs(Fu12) + th DEVSSr_u12 (:432-442, :413);  L2 = rhs(Fus) + th DEVSSr_u1 (:473-483, :453)
//  (rho0 [c u0/dt - w], v) - Ra Pr (rho0 therm_inv k, v) + (Sigma, Dincomp(v))
//    L1: c = 2, w = (u0 . nabla_grad) u0;   L2: c = 1, w = (1/2)(u12 . nabla_grad) u0;   k = (0,1)
//  Sigma = -Pr betav Dcomp(u0) + p0 I - Pr ((1-betav)/We)(S(tau0) - I) + th 2(1-betav) S(D0 | D12)
template <bool HALF>
struct C5_L12 {
  using Space = SpW; using Facet = C5_L12_Facet;
  static constexpr int QDEG = 6, NP = 5;
  enum { F0, F1, SXX, SXY, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Sym3 D = eval_devss(x.m, x.s.f[HALF ? F_W0 : F_W12], x.cell);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double rho = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    const double ti = eval_q(x.m, x.s.f[F_THERM_INV], x.cell, b1).v;
    double w0, w1;
    if (HALF) {
      w0 = u.g00 * u.u0 + u.g01 * u.u1;
      w1 = u.g10 * u.u0 + u.g11 * u.u1;
    } else {
      const Vel h = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
      w0 = 0.5 * (u.g00 * h.u0 + u.g01 * h.u1);
      w1 = 0.5 * (u.g10 * h.u0 + u.g11 * h.u1);
    }
    const double m = (HALF ? 2.0 : 1.0) / p[P_DT];
    const double bp = p[P_PR] * p[P_BETAV], ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    const double gam = p[P_TH] * 2.0 * (1.0 - p[P_BETAV]);
    d[F0] = rho * (m * u.u0 - w0);
    d[F1] = rho * (m * u.u1 - w1) - p[P_RA] * p[P_PR] * rho * ti;
    const double dv3 = (u.g00 + u.g11) / 3.0;
    d[SXX] = -bp * (u.g00 - dv3) + p0 - ce * (t.t0.v - 1.0) + gam * D.a;
    d[SXY] = -bp * 0.5 * (u.g01 + u.g10) - ce * t.t1.v + gam * D.b;
    d[SYY] = -bp * (u.g11 - dv3) + p0 - ce * (t.t2.v - 1.0) + gam * D.c;
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
// a5 = (Ma^2/dt)(p,q) + (therm dt/2 grad p, grad q)                      :493-499
struct C5_A5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = therm_of(eval_q(x.m, x.s.f[F_THETA0], x.cell, b1).v, x.p.v);
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    a.template add<QV, QV>(p.v[P_MA] * p.v[P_MA] / p.v[P_DT]);
    a.template add<QX, QX>(0.5 * p.v[P_DT] * d[0]);
    a.template add<QY, QY>(0.5 * p.v[P_DT] * d[0]);
  }
};
// L5 = ((Ma^2/dt) p0 - therm dot(grad rho0, us) - therm rho0 div us, q) + (therm dt/2 grad p0, grad q)   :494-500
struct C5_L5 {
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
    const double th = therm_of(eval_q(x.m, x.s.f[F_THETA0], x.cell, b1).v, p);
    d[0] = (p[P_MA] * p[P_MA] / p[P_DT]) * p0.v - th * (r.x * us.u0 + r.y * us.u1) - th * r.v * (us.g00 + us.g11);
    d[1] = 0.5 * p[P_DT] * th * p0.x;
    d[2] = 0.5 * p[P_DT] * th * p0.y;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QV>(d[0]);
    a.template add<QX>(d[1]);
    a.template add<QY>(d[2]);
  }
};
// RHS of theta0 = project((T0-T_0)/(T_h-T_0), Q) (:405) [FIELD = F_T0] and theta1 (:581) [FIELD = F_T1]
template <int FIELD>
struct C5_L_THETA {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = theta_of(eval_q(x.m, x.s.f[FIELD], x.cell, b1).v, x.p.v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// RHS of therm_inv = project(1.0/therm, Q)                                :407-408
struct C5_L_THERM_INV {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = 1.0 / therm_of(eval_q(x.m, x.s.f[F_THETA0], x.cell, b1).v, x.p.v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// RHS of rho1 = project(rho0 + Ma^2 therm_inv (p1-p0), Q)                 :529-530
struct C5_L_RHO1 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 3, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    d[0] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v
           + p[P_MA] * p[P_MA] * eval_q(x.m, x.s.f[F_THERM_INV], x.cell, b1).v
                 * (eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// a7 = ((1/dt) rho1 u, v) + (D - Dincomp(u), R)                           :533-536
struct C5_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, d[0] / p.v[P_DT]);
    w_devss_proj<false>(a);
  }
};
// L7 = ((1/dt) rho0 us, v) + (1/2)[(p1-p0, div v) - <(p1-p0) n, v>]        :534-537
struct C5_L7_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int QDEG = 3, NP = 2, MARKERS = -1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const double dp = eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    d[0] = -0.5 * dp * x.nx;
    d[1] = -0.5 * dp * x.ny;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};
struct C5_L7 {
  using Space = SpW; using Facet = C5_L7_Facet;
  static constexpr int QDEG = 5, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const double m = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v / x.p.v[P_DT];
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
// a4 = (((We/dt)+1) tau + We Fdefcom(u1,tau), Rt) + c1 (kapp h grad tau, grad Rt) + c2 (kapp h div tau, div Rt)
//      Fdefcom = Fdef + div(u1) tau                                        :550-555, :391
struct C5_A4 {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 7;
  enum { U0, U1, G00, G01, G10, G11, KH };
  FX_HD static void point(const PtArgs& x, double* d) { LDC_A4::point(x, d); }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    LDC_A4::terms(a, d, p);
    const double wdv = p.v[P_WE] * (d[G00] + d[G11]);
    a.template add<T0V, T0V>(wdv);
    a.template add<T1V, T1V>(2.0 * wdv);
    a.template add<T2V, T2V>(wdv);
  }
};
// a8 = ((1/dt) rho1 thetal + rho1 dot(u1, grad thetal), r) + (grad thetal, grad r),  thetal = T/(T_h-T_0)   :564-566
struct C5_A8 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 4;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const double rho = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    const double idT = 1.0 / (x.p.v[P_TH_HOT] - x.p.v[P_T0]);  // (not in terms(): the sparsity mask is
    d[0] = rho / x.p.v[P_DT] * idT;                            //  evaluated with all-ones parameters)
    d[1] = rho * u.u0 * idT;
    d[2] = rho * u.u1 * idT;
    d[3] = idT;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par&) {
    a.template add<QV, QV>(d[0]);
    a.template add<QV, QX>(d[1]);
    a.template add<QV, QY>(d[2]);
    a.template add<QX, QX>(d[3]);
    a.template add<QY, QY>(d[3]);
  }
};
// L8 = ((1/dt) rho0 (thetar + theta0) + rho1 dot(u1, grad thetar) + Vh inner(sigmacom(u1,p1,tau1), grad u1), r)
//      + (grad thetar, grad r) + (We S(tau1) grad thetar, grad r) + Bi <grad theta0 . n, r>_{ds(3)+ds(4)}   :563-567
// thetar is the constant T_0/(T_h-T_0): its gradient terms vanish.
struct C5_L8_Facet {
  using Space = SpQ;
  static constexpr bool present = true;
  static constexpr int QDEG = 1, NP = 1, MARKERS = (1 << 3) | (1 << 4);
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const VG th = eval_q(x.m, x.s.f[F_THETA0], x.cell, b1);
    d[0] = x.p.v[P_BI] * (th.x * x.nx + th.y * x.ny);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
struct C5_L8 {
  using Space = SpQ; using Facet = C5_L8_Facet;
  static constexpr int QDEG = 4, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const double p1 = eval_q(x.m, x.s.f[F_P1], x.cell, b1).v;
    const double rho0 = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    const double th0 = eval_q(x.m, x.s.f[F_THETA0], x.cell, b1).v;
    const double thr = p[P_T0] / (p[P_TH_HOT] - p[P_T0]);
    const double bp = p[P_PR] * p[P_BETAV], ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    const double dv3 = (u.g00 + u.g11) / 3.0;
    const double s00 = 2.0 * bp * (u.g00 - dv3) - p1 + ce * (t.t0.v - 1.0);
    const double s01 = bp * (u.g01 + u.g10) + ce * t.t1.v;
    const double s11 = 2.0 * bp * (u.g11 - dv3) - p1 + ce * (t.t2.v - 1.0);
    const double gamdot = s00 * u.g00 + s01 * (u.g01 + u.g10) + s11 * u.g11;
    d[0] = rho0 / p[P_DT] * (thr + th0) + p[P_VH] * gamdot;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};

// restau0 = (We/dt)(tau1v - tau0v) + We [F00,F10,F11] + tau1v - I_vec,  F = Fdefcom(u1,tau1)   :382-385
FX_HD void dgp_restau(const PtArgs& x, double* r) {
  P2Phys b;
  p2_phys(x.xi, x.et, x.g, b);
  const double* p = x.p.v;
  const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
  const Tau t1 = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
  const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
  const double We = p[P_WE], wd = We / p[P_DT];
  const double dv = u.g00 + u.g11;
  const double a0 = t1.t0.v, a1 = t1.t1.v, a2 = t1.t2.v;
  const double F00 = u.u0 * t1.t0.x + u.u1 * t1.t1.x - 2.0 * (u.g00 * a0 + u.g01 * a1) + dv * a0;
  const double F10 = u.u0 * t1.t1.x + u.u1 * t1.t2.x - (u.g10 * a0 + u.g11 * a1) - (a1 * u.g00 + a2 * u.g01)
                     + dv * a1;
  const double F11 = u.u0 * t1.t1.y + u.u1 * t1.t2.y - 2.0 * (u.g10 * a1 + u.g11 * a2) + dv * a2;
  r[0] = wd * (a0 - t0.t0.v) + We * F00 + a0 - 1.0;
  r[1] = wd * (a1 - t0.t1.v) + We * F10 + a1;
  r[2] = wd * (a2 - t0.t2.v) + We * F11 + a2 - 1.0;
}
struct C5_L_RES_ORTH {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    dgp_restau(x, d);
    for (int j = 0; j < 3; ++j) d[j] -= eval_dg0(x.s.f[F_RES_TEST], x.m.dm[S_ZD], x.m.nc, x.cell, j);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};
struct E_C5_RESTAU {
  static constexpr int QDEG = 3, NOUT = 3;
  FX_HD static void value(const PtArgs& x, double* v) { dgp_restau(x, v); }
};
struct C5_EE {  // (tau1[0,0]+tau1[1,1]-2.0)*dx   :578
  static constexpr int QDEG = 2;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    return t.t0.v + t.t2.v - 2.0;
  }
};
struct C5_NUS {  // Nusselt number: assemble(-inner(grad(theta1), n)*ds(2)), theta1 in the Q scratch field   :581-583
  static constexpr int QDEG = 0, MARKERS = 1 << 2;
  FX_HD static double integrand(const PtArgs& x) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const VG th = eval_q(x.m, x.s.f[F_Q_A], x.cell, b1);
    return -(th.x * x.nx + th.y * x.ny);
  }
};


// ==================================================================================================================
// buoyancy-driven-flow/incompressible_dgp.py (time loop :357-535): the incompressible double-glazing problem.
// Helper algebra fenics_fem.py:68-84:
//   sigma(u,p,tau)   = 2 Pr betav Dincomp(u) - p I + Pr ((1-betav)/We)(S(tau) - I)          (betav < 1)
//   sigma_n(U,p,tau) = Pr betav nabla_grad(U) n + (Pr (1-betav)/We) S(tau) n - p n
//   Fdef = dot(u,grad(Tau)) - grad(u) Tau - Tau grad(u)^T                                   (no div(u) Tau)
// Boundary markers (:147-151): 1 bottom, 2 left, 3 right, 4 top.  theta0 = (T0 - T_0)/(T_h - T_0) is a UFL
// expression of T0 here (not a projected field).  Shared with other loops: a5 = C1_A5, L5 = C1_L5 (Re slot = 1),
// a4 = LDC_A4, L4 = C4_L4, E_k = C1_EK, E_e = C5_EE, theta1 projection = C5_L_THETA<F_T1>.
// ==================================================================================================================
// Facet part of a1/a2: -<sigma_n(u/2, 0, 0), v> = -(Pr betav/2) <nabla_grad(u) n, v>                  :401, :443
struct C6_A12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int QDEG = 3, NP = 2, MARKERS = -1;
  FX_HD static void point(const PtArgs& x, double* d) { d[0] = x.nx; d[1] = x.ny; }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double hb = -0.5 * p.v[P_BETAV] * p.v[P_PR];
    a.template add<UXV, UXX>(hb * d[0]);
    a.template add<UXV, UYX>(hb * d[1]);
    a.template add<UYV, UXY>(hb * d[0]);
    a.template add<UYV, UYY>(hb * d[1]);
  }
};
// a1 = lhs(Fu12) + th DEVSSl_u12 (:397-408):  (2/dt)(u,v) + Pr betav (D(u), D(v)) + (D - Dincomp(u), R)
//      + th 2(1-betav+eps)(D(u), D(v)) + facet part
// a2 = lhs(Fus) + th DEVSSl_u1 (:439-448):    (1/dt)(u,v) + (1/2)((u12 . nabla_grad) u, v) + the same
template <bool HALF>
struct C6_A12 {
  using Space = SpW; using Facet = C6_A12_Facet;
  static constexpr int QDEG = HALF ? 4 : 5, NP = 2;
  FX_HD static void point(const PtArgs& x, double* d) {
    d[0] = 0.0; d[1] = 0.0;
    if (!HALF) {
      P2Phys b;
      p2_phys(x.xi, x.et, x.g, b);
      const Vel u = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
      d[0] = u.u0; d[1] = u.u1;
    }
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, (HALF ? 2.0 : 1.0) / p.v[P_DT]);
    if constexpr (!HALF) {
      a.template add<UXV, UXX>(0.5 * d[0]);
      a.template add<UXV, UXY>(0.5 * d[1]);
      a.template add<UYV, UYX>(0.5 * d[0]);
      a.template add<UYV, UYY>(0.5 * d[1]);
    }
    w_d_d(a, p.v[P_PR] * p.v[P_BETAV] + p.v[P_TH] * 2.0 * (1.0 - p.v[P_BETAV] + 3.0e-16));
    w_devss_proj<false>(a);
  }
};
// Facet part of L1/L2: +<sigma_n(u0/2, p0, tau0), v> = (Pr betav/2)<nabla_grad(u0) n, v> + (Pr(1-betav)/We)<S(tau0) n, v> - <p0 n, v>
struct C6_L12_Facet {
  using Space = SpW;
  static constexpr bool present = true;
  static constexpr int QDEG = 4, NP = 2, MARKERS = -1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double hb = 0.5 * p[P_BETAV] * p[P_PR], ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    d[0] = -p0 * x.nx + hb * (u.g00 * x.nx + u.g10 * x.ny) + ce * (t.t0.v * x.nx + t.t1.v * x.ny);
    d[1] = -p0 * x.ny + hb * (u.g01 * x.nx + u.g11 * x.ny) + ce * (t.t1.v * x.nx + t.t2.v * x.ny);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
  }
};
// L1 = rhs(Fu12) + th DEVSSr_u12 (:397-409, :332);  L2 = rhs(Fus) + th DEVSSr_u1 (:439-449, :421)
//  ([c u0/dt - w], v) - Ra Pr (theta0 f, v) + (Sigma, Dincomp(v)),  f = (0,-1)
//    L1: c = 2, w = (u0 . nabla_grad) u0;   L2: c = 1, w = (1/2)(u12 . nabla_grad) u0
//  Sigma = -Pr betav Dincomp(u0) + p0 I - Pr ((1-betav)/We)(S(tau0) - I) + gam S(D0 | D12),
//    gam = 2 th (DEVSSr_u12 carries no (1-betav), :332) | 2 th Pr (1-betav+eps) (:421)
template <bool HALF>
struct C6_L12 {
  using Space = SpW; using Facet = C6_L12_Facet;
  static constexpr int QDEG = 5, NP = 5;
  enum { F0, F1, SXX, SXY, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Sym3 D = eval_devss(x.m, x.s.f[HALF ? F_W0 : F_W12], x.cell);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double th0 = theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p);
    double w0, w1;
    if (HALF) {
      w0 = u.g00 * u.u0 + u.g01 * u.u1;
      w1 = u.g10 * u.u0 + u.g11 * u.u1;
    } else {
      const Vel h = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
      w0 = 0.5 * (u.g00 * h.u0 + u.g01 * h.u1);
      w1 = 0.5 * (u.g10 * h.u0 + u.g11 * h.u1);
    }
    const double m = (HALF ? 2.0 : 1.0) / p[P_DT];
    const double bp = p[P_PR] * p[P_BETAV], ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    const double gam = HALF ? 2.0 * p[P_TH] : 2.0 * p[P_TH] * p[P_PR] * (1.0 - p[P_BETAV] + 3.0e-16);
    d[F0] = m * u.u0 - w0;
    d[F1] = m * u.u1 - w1 + p[P_RA] * p[P_PR] * th0;
    d[SXX] = -bp * u.g00 + p0 - ce * (t.t0.v - 1.0) + gam * D.a;
    d[SXY] = -bp * 0.5 * (u.g01 + u.g10) - ce * t.t1.v + gam * D.b;
    d[SYY] = -bp * u.g11 + p0 - ce * (t.t2.v - 1.0) + gam * D.c;
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
// a7 = ((1/dt) u, v) + (D - Dincomp(u), R)                                 :463-466
struct C6_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    w_mass(a, 1.0 / p.v[P_DT]);
    w_devss_proj<false>(a);
  }
};
// L7 = ((1/dt) us, v) + (1/2)(p1-p0, div v) - (1/2)<(p1-p0) n, v>            :464-467
struct C6_L7 {
  using Space = SpW; using Facet = C5_L7_Facet;
  static constexpr int QDEG = 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    d[0] = us.u0 / x.p.v[P_DT];
    d[1] = us.u1 / x.p.v[P_DT];
    d[2] = 0.5 * (eval_q(x.m, x.s.f[F_P1], x.cell, b1).v - eval_q(x.m, x.s.f[F_P0], x.cell, b1).v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
    a.template add<UXX>(d[2]);
    a.template add<UYY>(d[2]);
  }
};
// a8 = ((1/dt) thetal + dot(u1, grad thetal), r) + ((I + We S(tau1)) grad thetal, grad r),  thetal = T/(T_h-T_0)   :504-507
struct C6_A8 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 3, NP = 6;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const double idT = 1.0 / (x.p.v[P_TH_HOT] - x.p.v[P_T0]), We = x.p.v[P_WE];
    d[0] = idT / x.p.v[P_DT];
    d[1] = u.u0 * idT;
    d[2] = u.u1 * idT;
    d[3] = idT * (1.0 + We * t.t0.v);
    d[4] = idT * We * t.t1.v;
    d[5] = idT * (1.0 + We * t.t2.v);
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par&) {
    a.template add<QV, QV>(d[0]);
    a.template add<QV, QX>(d[1]);
    a.template add<QV, QY>(d[2]);
    a.template add<QX, QX>(d[3]);
    a.template add<QX, QY>(d[4]);
    a.template add<QY, QX>(d[4]);
    a.template add<QY, QY>(d[5]);
  }
};
// L8 = ((1/dt)(thetar + theta0) + Vh inner(sigma(u1,p1,tau1), grad u1), r) + Bi <grad theta0 . n, r>_{ds(1)+ds(4)}   :503-508
// thetar is the constant T_0/(T_h-T_0): its gradient terms (incl. We S(tau1) grad thetar) vanish.
struct C6_L8_Facet {
  using Space = SpQ;
  static constexpr bool present = true;
  static constexpr int QDEG = 1, NP = 1, MARKERS = (1 << 1) | (1 << 4);
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const VG T0 = eval_q(x.m, x.s.f[F_T0], x.cell, b1);
    d[0] = x.p.v[P_BI] * (T0.x * x.nx + T0.y * x.ny) / (x.p.v[P_TH_HOT] - x.p.v[P_T0]);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
struct C6_L8 {
  using Space = SpQ; using Facet = C6_L8_Facet;
  static constexpr int QDEG = 4, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const double p1 = eval_q(x.m, x.s.f[F_P1], x.cell, b1).v;
    const double th0 = theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p);
    const double thr = p[P_T0] / (p[P_TH_HOT] - p[P_T0]);
    const double bp = p[P_PR] * p[P_BETAV], ce = p[P_PR] * (1.0 - p[P_BETAV]) / p[P_WE];
    const double s00 = 2.0 * bp * u.g00 - p1 + ce * (t.t0.v - 1.0);
    const double s01 = bp * (u.g01 + u.g10) + ce * t.t1.v;
    const double s11 = 2.0 * bp * u.g11 - p1 + ce * (t.t2.v - 1.0);
    const double gamdot = s00 * u.g00 + s01 * (u.g01 + u.g10) + s11 * u.g11;
    d[0] = (thr + th0) / p[P_DT] + p[P_VH] * gamdot;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// restau0 = (We/dt)(tau1v - tau0v) + We [F00,F10,F11] + tau1v - I_vec,  F = Fdef(u1,tau1)      :369-372
FX_HD void dgp_restau_incomp(const PtArgs& x, double* r) {
  P2Phys b;
  p2_phys(x.xi, x.et, x.g, b);
  const double* p = x.p.v;
  const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
  const Tau t1 = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
  const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
  const double We = p[P_WE], wd = We / p[P_DT];
  const double a0 = t1.t0.v, a1 = t1.t1.v, a2 = t1.t2.v;
  const double F00 = u.u0 * t1.t0.x + u.u1 * t1.t1.x - 2.0 * (u.g00 * a0 + u.g01 * a1);
  const double F10 = u.u0 * t1.t1.x + u.u1 * t1.t2.x - (u.g10 * a0 + u.g11 * a1) - (a1 * u.g00 + a2 * u.g01);
  const double F11 = u.u0 * t1.t1.y + u.u1 * t1.t2.y - 2.0 * (u.g10 * a1 + u.g11 * a2);
  r[0] = wd * (a0 - t0.t0.v) + We * F00 + a0 - 1.0;
  r[1] = wd * (a1 - t0.t1.v) + We * F10 + a1;
  r[2] = wd * (a2 - t0.t2.v) + We * F11 + a2 - 1.0;
}
struct C6_L_RES_ORTH {  // RHS of project(restau0 - res_test, Zc)  :374
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    dgp_restau_incomp(x, d);
    for (int j = 0; j < 3; ++j) d[j] -= eval_dg0(x.s.f[F_RES_TEST], x.m.dm[S_ZD], x.m.nc, x.cell, j);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};
struct E_C6_RESTAU {  // project(restau0, Zd)  :373
  static constexpr int QDEG = 3, NOUT = 3;
  FX_HD static void value(const PtArgs& x, double* v) { dgp_restau_incomp(x, v); }
};
struct C6_NUS {  // Nusselt number: assemble(inner(grad(theta1), n)*ds(2)), theta1 in the Q scratch field   :521-523
  static constexpr int QDEG = 0, MARKERS = 1 << 2;
  FX_HD static double integrand(const PtArgs& x) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    const VG th = eval_q(x.m, x.s.f[F_Q_A], x.cell, b1);
    return th.x * x.nx + th.y * x.ny;
  }
};

}  // namespace dgp
}  // namespace fx
