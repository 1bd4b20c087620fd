// Hand-written form functors for config 4:
//   "flow between eccentrically rotating cylinders"/fenepmp_jbp.py  (time loop :454-690)
// FENE-P-MP journal bearing.  At the CSV default lambda_D = 0 (fenepmp_jbp.py:297, parameters-fenep-mp.csv row 4)
// phi_def(u, lambda_D) folds to the literal 1 and the MP dissipation term to Zero: the
// functors instantiated with MP = false.  lambda_D != 0 (CSV row `Ld`: 0.05, 0.1, 0.15; plot_data.py:45-47) keeps
//   phi_def(u) = 1 + (3 lambda_D I_3(D(u)) / (I_2(D(u)) + 1e-7))^2                       (fenics_base.py:287-291)
// in five places -- the predictor's div(phi_def(u0)(f tau0 - I)) (needs the P2 Hessian of u0), the stress matrix and
// the LPS residual (dissipation term), the viscous heating and the journal functionals (sigma_f) -- and raises UFL's
// quadrature degrees to 13 / 13 / 14 (dx) and 12 (ds): the MP = true instantiations, form ids FX_FORM_C4M_*.
// The MP = false ids reject lambda_D != 0 (FX_ERR_UNSUPPORTED) and name their C4M counterparts.
// Helper algebra: fenics_base.py:158-294.  Notation as in fx_forms_ldc.h; additionally
//   f(tau) = fene_func(tau, b) = 1 / (1 - (|tau00 + tau11| - 2)/b^2)            (fenics_base.py:246-257)
//   phi_s(theta) = 1 - A/(C + theta),  A = A_0/(8.3144598*50), C = 7              (fenics_base.py:227-238)
//   sigma_f(u,p,tau) = 2 betav Dcomp(u) - p I + ((1-betav)/We)(f(tau) S(tau) - I)  (fene_sigmacom :186-187)
#pragma once
#include "fx_forms_ldc.h"

namespace fx {
namespace jbp {

using namespace ldc;  // slot enums, shared building blocks

enum : int { P_PR = 9, P_RA, P_DI, P_VH, P_BI, P_T0, P_TH_HOT, P_AL1, P_AL2, P_LAMBDA_D, P_B_FENE, P_A0, P_EPS, P_T,
             P_K_EWM, P_B_EWM, P_TH_T };

constexpr int QD_MP_RESTAU = 11, QD_MP_RES_ORTH = 13;  // UFL's degrees of the two LPS projections with lambda_D != 0

FX_HD double fene_f(double t0, double t2, double b) { return 1.0 / (1.0 - (fabs(t0 + t2) - 2.0) / (b * b)); }
FX_HD double sgn(double x) { return (x > 0.0) - (x < 0.0); }
FX_HD double theta_of(double T, const double* p) { return (T - p[P_T0]) / (p[P_TH_HOT] - p[P_T0]); }
FX_HD double phi_s_of(double theta, double A0) {
  const double A = A0 / (8.3144598 * (350.0 - 300.0)), C = 350.0 / (350.0 - 300.0);
  return 1.0 - A * (1.0 / (C + theta));
}
// sigma_f : grad(w) and the components of sigma_f for (u,p,tau) given at a point
struct SigmaF {
  double s00, s01, s11;
};
// phi_def(u, lambda_D) and (with the Hessian) its gradient
FX_HD double phi_def_of(const Vel& u, double lam) {
  const double d00 = u.g00, d01 = 0.5 * (u.g01 + u.g10), d11 = u.g11;
  const double i3 = d00 * d11 - d01 * d01, i2 = 0.5 * (d00 * d00 + d11 * d11 + 2.0 * d01 * d01);
  const double q = lam * 3.0 * i3 / (i2 + 0.0000001);
  return 1.0 + q * q;
}
FX_HD void phi_def_grad(const Vel& u, const VelH& h, double lam, double& px, double& py) {
  const double d00 = u.g00, d01 = 0.5 * (u.g01 + u.g10), d11 = u.g11;
  const double i3 = d00 * d11 - d01 * d01, i2e = 0.5 * (d00 * d00 + d11 * d11 + 2.0 * d01 * d01) + 0.0000001;
  const double q = lam * 3.0 * i3 / i2e;
  // d_j D00 = d_j d_0 u_0, d_j D11 = d_j d_1 u_1, d_j D01 = (d_j d_1 u_0 + d_j d_0 u_1) / 2
  const double d00j[2] = {h.a_xx, h.a_xy}, d11j[2] = {h.c_xy, h.c_yy};
  const double d01j[2] = {0.5 * (h.a_xy + h.c_xx), 0.5 * (h.a_yy + h.c_xy)};
  double out[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double di3 = d00j[j] * d11 + d00 * d11j[j] - 2.0 * d01 * d01j[j];
    const double di2 = d00 * d00j[j] + d11 * d11j[j] + 2.0 * d01 * d01j[j];
    out[j] = 2.0 * q * lam * 3.0 * (di3 * i2e - i3 * di2) / (i2e * i2e);
  }
  px = out[0]; py = out[1];
}
FX_HD SigmaF sigma_f(const Vel& u, double p, const Tau& t, const double* par, double phi = 1.0) {
  const double bv = par[P_BETAV], ce = phi * (1.0 - bv) / par[P_WE];
  const double f = fene_f(t.t0.v, t.t2.v, par[P_B_FENE]);
  const double dv3 = (u.g00 + u.g11) / 3.0;
  SigmaF s;
  s.s00 = 2.0 * bv * (u.g00 - dv3) - p + ce * (f * t.t0.v - 1.0);
  s.s01 = bv * (u.g01 + u.g10) + ce * f * t.t1.v;
  s.s11 = 2.0 * bv * (u.g11 - dv3) - p + ce * (f * t.t2.v - 1.0);
  return s;
}

// c (Dincomp(u), Dincomp(v))
template <class A> FX_HDC void w_d_d(A& a, double c) {
  a.template add<UXX, UXX>(c);
  a.template add<UYY, UYY>(c);
  a.template add<UXY, UXY>(0.5 * c);
  a.template add<UXY, UYX>(0.5 * c);
  a.template add<UYX, UXY>(0.5 * c);
  a.template add<UYX, UYX>(0.5 * c);
}

// a2 = lhs(Fus) + th DEVSSGl_u1                                        fenepmp_jbp.py:513-527, :416
//  (Re rho0/dt)(u,v) + (Re rho0 conv/2)((u0 . nabla_grad) u, v) + betav phi_s (D(u), D(v))
//  + (betav phi_s/6)(div u, div v) + (D - grad u, R) + th 2(1-betav)(D(u), D(v))
struct C4_A2 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 6, NP = 4;
  enum { RHO, U0, U1, PHIS };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    d[RHO] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    d[U0] = u.u0; d[U1] = u.u1;
    d[PHIS] = phi_s_of(theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, x.p.v), x.p.v[P_A0]);
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double bv = p.v[P_BETAV];
    w_mass(a, p.v[P_RE] * d[RHO] / p.v[P_DT]);
    const double c = 0.5 * p.v[P_RE] * d[RHO] * p.v[P_CONV];
    a.template add<UXV, UXX>(c * d[U0]);
    a.template add<UXV, UXY>(c * d[U1]);
    a.template add<UYV, UYX>(c * d[U0]);
    a.template add<UYV, UYY>(c * d[U1]);
    w_d_d(a, bv * d[PHIS] + p.v[P_TH] * 2.0 * (1.0 - bv));
    const double g6 = bv * d[PHIS] / 6.0;
    a.template add<UXX, UXX>(g6);
    a.template add<UXX, UYY>(g6);
    a.template add<UYY, UXX>(g6);
    a.template add<UYY, UYY>(g6);
    w_devss_proj<false>(a);
  }
};
// L2 = rhs(Fus) + th DEVSSGr_u1                                        :513-528, :495
// MP: div(phi (f S - I))_i = phi div(f S - I)_i + (f S - I)_ij d_j phi
template <bool MP>
struct C4_L2T {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = MP ? 13 : 6, NP = 5;
  enum { F0, F1, SXX, SXY, SYY };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Sym3 D0 = eval_devss(x.m, x.s.f[F_W0], x.cell);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const VG p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const double rho = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    const double phis = phi_s_of(theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p), p[P_A0]);
    const double bv = p[P_BETAV], ce = (1.0 - bv) / (p[P_WE] + 3.0e-16), gam = p[P_TH] * 2.0 * (1.0 - bv);
    // div(f(tau0) S(tau0) - I)_i = f d_j tau_ij + tau_ij d_j f,  d_j f = f^2 sgn(tr) d_j tr / b^2
    const double bb = p[P_B_FENE] * p[P_B_FENE];
    const double f = fene_f(t.t0.v, t.t2.v, p[P_B_FENE]);
    const double k = f * f * sgn(t.t0.v + t.t2.v) / bb;
    const double trx = t.t0.x + t.t2.x, try_ = t.t0.y + t.t2.y;
    double dm0 = f * (t.t0.x + t.t1.y) + k * (t.t0.v * trx + t.t1.v * try_);
    double dm1 = f * (t.t1.x + t.t2.y) + k * (t.t1.v * trx + t.t2.v * try_);
    if constexpr (MP) {
      const double phi = phi_def_of(u, p[P_LAMBDA_D]);
      double px, py;
      phi_def_grad(u, eval_vel_hess(x.m, x.s.f[F_W0], x.cell, x.g), p[P_LAMBDA_D], px, py);
      dm0 = phi * dm0 + (f * t.t0.v - 1.0) * px + f * t.t1.v * py;
      dm1 = phi * dm1 + f * t.t1.v * px + (f * t.t2.v - 1.0) * py;
    }
    const double rr = p[P_RE] * rho, hc = 0.5 * p[P_CONV];
    d[F0] = rr * (u.u0 / p[P_DT] - hc * (u.g00 * u.u0 + u.g01 * u.u1)) + ce * dm0 - p0.x;
    d[F1] = rr * (u.u1 / p[P_DT] - hc * (u.g10 * u.u0 + u.g11 * u.u1)) + ce * dm1 - p0.y;
    const double g = bv * phis, dv6 = g * (u.g00 + u.g11) / 6.0;
    d[SXX] = -g * u.g00 - dv6 + gam * D0.a;
    d[SXY] = -g * 0.5 * (u.g01 + u.g10) + gam * D0.b;
    d[SYY] = -g * u.g11 - dv6 + gam * D0.c;
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
using C4_L2 = C4_L2T<false>;
using C4M_L2 = C4_L2T<true>;
// a5 = (Ma^2/dt)(p,q) + dt (grad p, grad q)                             :536-542
struct C4_A5 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 2, NP = 0;
  FX_HD static void point(const PtArgs&, double*) {}
  template <class A> FX_HDC static void terms(A& a, const double*, const Par& p) {
    a.template add<QV, QV>(p.v[P_MA] * p.v[P_MA] / p.v[P_DT]);
    a.template add<QX, QX>(p.v[P_DT]);
    a.template add<QY, QY>(p.v[P_DT]);
  }
};
// L5 = ((Ma^2/dt) p0 - Re div(rho0 us), q) + dt (grad p0, grad q)        :537-543
struct C4_L5 {
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
    d[0] = (p[P_MA] * p[P_MA] / p[P_DT]) * p0.v - p[P_RE] * (r.x * us.u0 + r.y * us.u1 + r.v * (us.g00 + us.g11));
    d[1] = p[P_DT] * p0.x;
    d[2] = p[P_DT] * p0.y;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<QV>(d[0]);
    a.template add<QX>(d[1]);
    a.template add<QY>(d[2]);
  }
};
// a6 = (1/dt)(rho, q) + ((u12 . grad rho), alpha (u12 . grad q)), alpha = h/(|u12| + eps)   :553-566
// (UFL degree 10 -> 36-point collapsed Gauss-Jacobi rule; eps = 1e-10 here, DOLFIN_EPS in compressible_dgp.py:510)
struct SU_RHO_A6 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 10, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Vel u = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
    const double al = x.g.h / (sqrt(u.u0 * u.u0 + u.u1 * u.u1) + x.p.v[P_EPS]);
    d[0] = al * u.u0 * u.u0;
    d[1] = al * u.u0 * u.u1;
    d[2] = al * u.u1 * u.u1;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    a.template add<QV, QV>(1.0 / p.v[P_DT]);
    a.template add<QX, QX>(d[0]);
    a.template add<QX, QY>(d[1]);
    a.template add<QY, QX>(d[1]);
    a.template add<QY, QY>(d[2]);
  }
};
// L6 = ((1/dt) rho0 - dot(u12, grad rho12) + rho12 div u12, q)            :558-564
struct SU_RHO_L6 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 3, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W12], x.cell, b);
    const VG r12 = eval_q(x.m, x.s.f[F_RHO12], x.cell, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v / x.p.v[P_DT] - (u.u0 * r12.x + u.u1 * r12.y)
           + r12.v * (u.g00 + u.g11);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
// a7 = ((Re/dt) rho1 u, v) + (D - grad u, R) + th 2(1-betav)(D(u), D(v))   :573-579
struct C4_A7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    w_mass(a, p.v[P_RE] * d[0] / p.v[P_DT]);
    w_d_d(a, p.v[P_TH] * 2.0 * (1.0 - p.v[P_BETAV]));
    w_devss_proj<false>(a);
  }
};
// L7 = ((Re/dt) rho0 us, v) - (1/2)(grad(p1-p0), v) + th (1-betav)(D0 + D0^T, D(v))   :574-580
struct C4_L7 {
  using Space = SpW; using Facet = NoFacet;
  static constexpr int QDEG = 5, NP = 5;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel us = eval_vel(x.m, x.s.f[F_WS], x.cell, b);
    const Sym3 D0 = eval_devss(x.m, x.s.f[F_W0], x.cell);
    const VG p1 = eval_q(x.m, x.s.f[F_P1], x.cell, b1), p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1);
    const double m = p[P_RE] / p[P_DT] * eval_q(x.m, x.s.f[F_RHO0], x.cell, b1).v;
    const double gam = p[P_TH] * 2.0 * (1.0 - p[P_BETAV]);
    d[0] = m * us.u0 - 0.5 * (p1.x - p0.x);
    d[1] = m * us.u1 - 0.5 * (p1.y - p0.y);
    d[2] = gam * D0.a; d[3] = gam * D0.b; d[4] = gam * D0.c;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<UXV>(d[0]);
    a.template add<UYV>(d[1]);
    a.template add<UXX>(d[2]);
    a.template add<UXY>(d[3]);
    a.template add<UYX>(d[3]);
    a.template add<UYY>(d[4]);
  }
};
// a4 = ((We/dt + f(tau0)) tau + We FdefG(u1, S(D1), tau), Rt) + 0.05 (kapp h grad tau, grad Rt)
//      + 0.01 (kapp h div tau, div Rt)                                    :595-608, :473
// MP: - We (phi_def(u1) - 1)/2 (tau D(u1) + D(u1) tau, S(Rt))
template <bool MP>
struct C4_A4T {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = MP ? 13 : 6, NP = MP ? 10 : 7;
  enum { M, U0, U1, GA, GB, GC, KH, CD00, CD01, CD11 };
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Sym3 D1 = eval_devss(x.m, x.s.f[F_W1], x.cell);
    const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    d[M] = p[P_WE] / p[P_DT] + fene_f(t0.t0.v, t0.t2.v, p[P_B_FENE]);
    d[U0] = u.u0; d[U1] = u.u1; d[GA] = D1.a; d[GB] = D1.b; d[GC] = D1.c;
    d[KH] = eval_dg0(x.s.f[F_KAPP], x.m.dm[S_QT], x.m.nc, x.cell, 0) * x.g.h;
    if constexpr (MP) {
      const double c = -p[P_WE] * 0.5 * (phi_def_of(u, p[P_LAMBDA_D]) - 1.0);
      d[CD00] = c * u.g00; d[CD01] = c * 0.5 * (u.g01 + u.g10); d[CD11] = c * u.g11;
    }
  }
  template <class A> FX_HDC static void terms(A& a, const double* d, const Par& p) {
    const double We = p.v[P_WE], m = d[M];
    if constexpr (MP) {  // (tau D + D tau) : S(Rt), off-diagonal counted twice
      a.template add<T0V, T0V>(2.0 * d[CD00]);
      a.template add<T0V, T1V>(2.0 * d[CD01]);
      a.template add<T1V, T0V>(2.0 * d[CD01]);
      a.template add<T1V, T1V>(2.0 * (d[CD00] + d[CD11]));
      a.template add<T1V, T2V>(2.0 * d[CD01]);
      a.template add<T2V, T1V>(2.0 * d[CD01]);
      a.template add<T2V, T2V>(2.0 * d[CD11]);
    }
    a.template add<T0V, T0V>(m);
    a.template add<T1V, T1V>(2.0 * m);
    a.template add<T2V, T2V>(m);
    // We (M - G tau - tau G^T, S(Rt)), G = S(D1) = [[ga,gb],[gb,gc]], M_jk = u_i d_k tau_ij
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
    const double c1 = 0.05 * d[KH], c2 = 0.01 * d[KH];
    a.template add<T0X, T0X>(c1);
    a.template add<T0Y, T0Y>(c1);
    a.template add<T1X, T1X>(2.0 * c1);
    a.template add<T1Y, T1Y>(2.0 * c1);
    a.template add<T2X, T2X>(c1);
    a.template add<T2Y, T2Y>(c1);
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
using C4_A4 = C4_A4T<false>;
using C4M_A4 = C4_A4T<true>;
// L4 = ((We/dt) S(tau0) + I, S(Rt))                                       :597-601
struct C4_L4 {
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double wd = x.p.v[P_WE] / x.p.v[P_DT];
    d[0] = wd * t.t0.v + 1.0;
    d[1] = 2.0 * wd * t.t1.v;
    d[2] = wd * t.t2.v + 1.0;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};
// a9 = ((1/dt) rho1 thetal + rho1 dot(u1, grad thetal), r) + (Di grad thetal + (Di/We) S(tau1) grad thetal, grad r)
//      thetal = T/(T_h - T_0)                                              :617-621
struct C4_A9 {
  using Space = SpQ; using Facet = NoFacet;
  static constexpr int QDEG = 4, NP = 6;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const double rho = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    const double idT = 1.0 / (p[P_TH_HOT] - p[P_T0]), dw = p[P_DI] / p[P_WE];
    d[0] = rho / p[P_DT] * idT;
    d[1] = rho * u.u0 * idT;
    d[2] = rho * u.u1 * idT;
    d[3] = (p[P_DI] + dw * t.t0.v) * idT;
    d[4] = dw * t.t1.v * idT;
    d[5] = (p[P_DI] + dw * t.t2.v) * idT;
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
// L9 = ((1/dt) rho1 (thetar + theta0) + rho1 dot(u1, grad thetar) + Vh inner(sigma_f(u0,p0,tau0), grad u0), r)
//      + (Di grad thetar + (Di/We) S(tau1) grad thetar, grad r) - Di Bi <theta0, r>_{ds(1)}      :615-624
// thetar = project(T_0/(T_h-T_0), Q) is the constant T_0/(T_h-T_0): its gradient terms vanish.
struct C4_L9_Facet {
  using Space = SpQ;
  static constexpr bool present = true;
  static constexpr int QDEG = 2, NP = 1, MARKERS = 1 << 1;  // ds(1): bearing
  FX_HD static void point(const PtArgs& x, double* d) {
    P1Phys b1;
    p1_phys(x.xi, x.et, x.g, b1);
    d[0] = -x.p.v[P_DI] * x.p.v[P_BI] * theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, x.p.v);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};
template <bool MP>
struct C4_L9T {
  using Space = SpQ; using Facet = C4_L9_Facet;
  static constexpr int QDEG = MP ? 14 : 6, NP = 1;
  FX_HD static void point(const PtArgs& x, double* d) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const double* p = x.p.v;
    const Vel u0 = eval_vel(x.m, x.s.f[F_W0], x.cell, b);
    const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
    const double p0 = eval_q(x.m, x.s.f[F_P0], x.cell, b1).v;
    const double rho1 = eval_q(x.m, x.s.f[F_RHO1], x.cell, b1).v;
    const double th0 = theta_of(eval_q(x.m, x.s.f[F_T0], x.cell, b1).v, p);
    const double thr = p[P_T0] / (p[P_TH_HOT] - p[P_T0]);
    const SigmaF s = sigma_f(u0, p0, t0, p, MP ? phi_def_of(u0, p[P_LAMBDA_D]) : 1.0);
    const double gamdot = s.s00 * u0.g00 + s.s01 * (u0.g01 + u0.g10) + s.s11 * u0.g11;
    d[0] = rho1 / p[P_DT] * (thr + th0) + p[P_VH] * gamdot;
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) { a.template add<QV>(d[0]); }
};

using C4_L9 = C4_L9T<false>;
using C4M_L9 = C4_L9T<true>;

// restau0 = (We/dt)(tau1v - tau0v) + We [F00,F10,F11] + f(tau0) tau1v - I_vec,  F = Fdef(u1,tau1) with
// + div(u1) tau1 (fenics_base.py:189-190)                                   :462-468
// MP: - We (phi_def(u1) - 1)/2 [tau1 D(u1) + D(u1) tau1]_{00,10,11}        :464-466
template <bool MP>
FX_HD void jbp_restau(const PtArgs& x, double* r) {
  P2Phys b;
  p2_phys(x.xi, x.et, x.g, b);
  const double* p = x.p.v;
  const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
  const Tau t1 = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
  const Tau t0 = eval_tau(x.m, x.s.f[F_TAU0], x.cell, b);
  const double We = p[P_WE], wd = We / p[P_DT];
  const double f0 = fene_f(t0.t0.v, t0.t2.v, p[P_B_FENE]);
  const double dv = u.g00 + u.g11;
  const double a0 = t1.t0.v, a1 = t1.t1.v, a2 = t1.t2.v;
  const double F00 = u.u0 * t1.t0.x + u.u1 * t1.t1.x - 2.0 * (u.g00 * a0 + u.g01 * a1) + dv * a0;
  const double F10 = u.u0 * t1.t1.x + u.u1 * t1.t2.x - (u.g10 * a0 + u.g11 * a1) - (a1 * u.g00 + a2 * u.g01)
                     + dv * a1;
  const double F11 = u.u0 * t1.t1.y + u.u1 * t1.t2.y - 2.0 * (u.g10 * a1 + u.g11 * a2) + dv * a2;
  r[0] = wd * (a0 - t0.t0.v) + We * F00 + f0 * a0 - 1.0;
  r[1] = wd * (a1 - t0.t1.v) + We * F10 + f0 * a1;
  r[2] = wd * (a2 - t0.t2.v) + We * F11 + f0 * a2 - 1.0;
  if constexpr (MP) {
    const double c = We * 0.5 * (phi_def_of(u, p[P_LAMBDA_D]) - 1.0);
    const double d00 = u.g00, d01 = 0.5 * (u.g01 + u.g10), d11 = u.g11;
    r[0] -= c * 2.0 * (a0 * d00 + a1 * d01);
    r[1] -= c * (a0 * d01 + a1 * (d00 + d11) + a2 * d01);
    r[2] -= c * 2.0 * (a1 * d01 + a2 * d11);
  }
}
template <bool MP>
struct C4_L_RES_ORTHT {  // RHS of project(restau0 - res_test, Zc)  :468
  using Space = SpZc; using Facet = NoFacet;
  static constexpr int QDEG = MP ? QD_MP_RES_ORTH : 6, NP = 3;
  FX_HD static void point(const PtArgs& x, double* d) {
    jbp_restau<MP>(x, d);
    for (int j = 0; j < 3; ++j) d[j] -= eval_dg0(x.s.f[F_RES_TEST], x.m.dm[S_ZD], x.m.nc, x.cell, j);
  }
  template <class A> FX_HDC static void vterms(A& a, const double* d, const Par&) {
    a.template add<T0V>(d[0]);
    a.template add<T1V>(d[1]);
    a.template add<T2V>(d[2]);
  }
};
using C4_L_RES_ORTH = C4_L_RES_ORTHT<false>;
using C4M_L_RES_ORTH = C4_L_RES_ORTHT<true>;
template <bool MP>
struct E_C4_RESTAUT {  // project(restau0, Zd)  :467
  static constexpr int QDEG = MP ? QD_MP_RESTAU : 4, NOUT = 3;
  FX_HD static void value(const PtArgs& x, double* v) { jbp_restau<MP>(x, v); }
};
using E_C4_RESTAU = E_C4_RESTAUT<false>;
using E_C4M_RESTAU = E_C4_RESTAUT<true>;

// ---- post-loop adaptive-refinement indicator -------------------------------------
// incompressible_benchmarkEWM.py:903-917, fenepmp_jbp.py:691-705, comp_ewm_jpb.py:735-747 (identical):
//   restau0 = We/dt*(tau1_vec - tau0_vec) + We*F1R_vec + tau1_vec - I_vec,   F1R = Fdef(u1, tau1)
//   kapp = project(inner(restau0, restau0), Qt);  tau_average = project(sum(tau1_vec)/3, Qt);
//   error_rat = |project(kapp/(tau_average + eps), Qt)|
struct E_ADAPT_RES_SQ {  // degree: restau0 3 (u1 . grad tau1: 2 + 1), squared -> 6
  static constexpr int QDEG = 6, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
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
    const double r0 = wd * (a0 - t0.t0.v) + We * F00 + a0 - 1.0;
    const double r1 = wd * (a1 - t0.t1.v) + We * F10 + a1;
    const double r2 = wd * (a2 - t0.t2.v) + We * F11 + a2 - 1.0;
    v[0] = r0 * r0 + r1 * r1 + r2 * r2;
  }
};
struct E_TAU_AVG {  // project((tau1_vec[0]+tau1_vec[1]+tau1_vec[2])/3.0, Qt)
  static constexpr int QDEG = 2, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    v[0] = (t.t0.v + t.t1.v + t.t2.v) / 3.0;
  }
};
struct E_ABS_KAPP_RATIO {  // absolute(project(kapp/(tau_average + eps), Qt)); tau_average in FX_F_QT_A, eps = FX_P_EPS
  static constexpr int QDEG = 0, NOUT = 1;
  FX_HD static void value(const PtArgs& x, double* v) {
    const double k = eval_dg0(x.s.f[F_KAPP], x.m.dm[S_QT], x.m.nc, x.cell, 0);
    const double ta = eval_dg0(x.s.f[F_QT_A], x.m.dm[S_QT], x.m.nc, x.cell, 0);
    v[0] = fabs(k / (ta + x.p.v[P_EPS]));
  }
};

// ---- functionals ---------------------------------------------------------------------------------
struct C4_EE {  // (fene_func(tau1,b)*(tau1[0,0]+tau1[1,1]) - 2.)*dx   :636
  static constexpr int QDEG = 4;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b;
    p2_phys(x.xi, x.et, x.g, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    return fene_f(t.t0.v, t.t2.v, x.p.v[P_B_FENE]) * (t.t0.v + t.t2.v) - 2.0;
  }
};
// journal load and torque on ds(0)  (:639-649): sigma = sigma_f(u1,p1,tau1), tang = (n1,-n0)
//   WHICH 0: innertorque = -inner(n, sigma tang); 1: innerforcex = (1,0).(sigma n); 2: innerforcey = (0,-1).(sigma n)
template <int WHICH, bool MP = false>
struct C4_JOURNAL {
  static constexpr int QDEG = MP ? 12 : 4, MARKERS = 1 << 0;
  FX_HD static double integrand(const PtArgs& x) {
    P2Phys b; P1Phys b1;
    p2_phys(x.xi, x.et, x.g, b);
    p1_phys(x.xi, x.et, x.g, b1);
    const Vel u = eval_vel(x.m, x.s.f[F_W1], x.cell, b);
    const Tau t = eval_tau(x.m, x.s.f[F_TAU1], x.cell, b);
    const SigmaF s = sigma_f(u, eval_q(x.m, x.s.f[F_P1], x.cell, b1).v, t, x.p.v,
                             MP ? phi_def_of(u, x.p.v[P_LAMBDA_D]) : 1.0);
    const double n0 = x.nx, n1 = x.ny;
    if (WHICH == 0) return -(n0 * (s.s00 * n1 - s.s01 * n0) + n1 * (s.s01 * n1 - s.s11 * n0));
    if (WHICH == 1) return s.s00 * n0 + s.s01 * n1;
    return -(s.s01 * n0 + s.s11 * n1);
  }
};

using C4M_TORQUE = C4_JOURNAL<0, true>;
using C4M_FORCE_X = C4_JOURNAL<1, true>;
using C4M_FORCE_Y = C4_JOURNAL<2, true>;

}  // namespace jbp
}  // namespace fx
