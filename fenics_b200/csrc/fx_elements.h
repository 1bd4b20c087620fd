// Reference elements, spaces, quadrature -- shared by the CUDA kernels and the host emulation
// used in tests (tests/hostemu).  Conventions follow SURVEY.md 8a-notes (UFC/FIAT ordering).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__ __forceinline__
#define FX_HDC __host__ __device__ constexpr
#else
#define FX_HD inline
#define FX_HDC constexpr
#endif

namespace fx {

enum ElKind : int { P2 = 0, P1 = 1, DG0 = 2 };

FX_HDC int el_nb(int k) { return k == P2 ? 6 : (k == P1 ? 3 : 1); }
FX_HDC int el_nslot(int k) { return k == DG0 ? 1 : 3; }

// ---- compile-time space descriptors ---------------------------------------------------------
template <int... Ks>
struct SpaceT {
  static constexpr int NCOMP = sizeof...(Ks);
  FX_HDC static int kind(int c) {
    constexpr int k[NCOMP] = {Ks...};
    return k[c];
  }
  FX_HDC static int nb(int c) { return el_nb(kind(c)); }
  FX_HDC static int off(int c) {  // local dof offset of component c
    int o = 0;
    for (int i = 0; i < c; ++i) o += el_nb(kind(i));
    return o;
  }
  FX_HDC static int slot0(int c) {  // first slot (value) of component c
    int o = 0;
    for (int i = 0; i < c; ++i) o += el_nslot(kind(i));
    return o;
  }
  static constexpr int LD = off(NCOMP);
  static constexpr int NS = slot0(NCOMP);
  FX_HDC static int slot_comp(int s) {
    int c = 0;
    while (c + 1 < NCOMP && slot0(c + 1) <= s) ++c;
    return c;
  }
  FX_HDC static int dof_comp(int a) {
    int c = 0;
    while (c + 1 < NCOMP && off(c + 1) <= a) ++c;
    return c;
  }
  FX_HDC static int maxnb() {
    int m = 0;
    for (int i = 0; i < NCOMP; ++i) m = nb(i) > m ? nb(i) : m;
    return m;
  }
};

using SpW = SpaceT<P2, P2, DG0, DG0, DG0>;
using SpZc = SpaceT<P2, P2, P2>;
using SpQ = SpaceT<P1>;
using SpZd = SpaceT<DG0, DG0, DG0>;
using SpQt = SpaceT<DG0>;
using SpV = SpaceT<P2, P2>;
using SpVs = SpaceT<P2>;

// runtime ids must match include/fenics_b200.h fx_space
enum SpaceId : int { S_W = 0, S_ZC = 1, S_Q = 2, S_ZD = 3, S_QT = 4, S_V = 5, S_VS = 6, S_N = 7 };
template <class Sp> struct space_id;
template <> struct space_id<SpW> { static constexpr int v = S_W; };
template <> struct space_id<SpZc> { static constexpr int v = S_ZC; };
template <> struct space_id<SpQ> { static constexpr int v = S_Q; };
template <> struct space_id<SpZd> { static constexpr int v = S_ZD; };
template <> struct space_id<SpQt> { static constexpr int v = S_QT; };
template <> struct space_id<SpV> { static constexpr int v = S_V; };
template <> struct space_id<SpVs> { static constexpr int v = S_VS; };

FX_HDC int space_ld(int s) {
  return s == S_W ? 15 : s == S_ZC ? 18 : s == S_Q ? 3 : s == S_ZD ? 3 : s == S_QT ? 1 : s == S_V ? 12 : 6;
}

// ---- compile-time loop ----------------------------------------------------------------------
template <int V> struct IC { static constexpr int value = V; constexpr operator int() const { return V; } };
template <int B, int E, class F>
FX_HD constexpr void static_for(F&& f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}

// ---- geometry of one cell: K = J^-1 (K[a][i] = d xi_a / d x_i), |detJ|, h ---------------------
struct Geo {
  double k00, k01, k10, k11, adet, h;
};

// ---- reference bases (UFC order) at a reference point ------------------------------------------
struct P2Ref {
  double v[6], dx[6], de[6];  // value, d/dxi, d/deta
};
FX_HD void p2_ref(double xi, double et, P2Ref& r) {
  const double l0 = 1.0 - xi - et, l1 = xi, l2 = et;
  r.v[0] = l0 * (2 * l0 - 1); r.v[1] = l1 * (2 * l1 - 1); r.v[2] = l2 * (2 * l2 - 1);
  r.v[3] = 4 * l1 * l2;       r.v[4] = 4 * l0 * l2;       r.v[5] = 4 * l0 * l1;
  r.dx[0] = -(4 * l0 - 1); r.de[0] = -(4 * l0 - 1);
  r.dx[1] = 4 * l1 - 1;    r.de[1] = 0.0;
  r.dx[2] = 0.0;           r.de[2] = 4 * l2 - 1;
  r.dx[3] = 4 * l2;        r.de[3] = 4 * l1;
  r.dx[4] = -4 * l2;       r.de[4] = 4 * (l0 - l2);
  r.dx[5] = 4 * (l0 - l1); r.de[5] = -4 * l1;
}
struct P1Ref {
  double v[3], dx[3], de[3];
};
FX_HD void p1_ref(double xi, double et, P1Ref& r) {
  r.v[0] = 1.0 - xi - et; r.v[1] = xi; r.v[2] = et;
  r.dx[0] = -1.0; r.de[0] = -1.0;
  r.dx[1] = 1.0;  r.de[1] = 0.0;
  r.dx[2] = 0.0;  r.de[2] = 1.0;
}

// one basis function (runtime index k, warp-uniform in the kernels)
FX_HD void ref_one(int kind, int k, double xi, double et, double& v, double& dx, double& de) {
  const double l0 = 1.0 - xi - et, l1 = xi, l2 = et;
  if (kind == DG0) { v = 1.0; dx = 0.0; de = 0.0; return; }
  if (kind == P1) {
    switch (k) {
      case 0: v = l0; dx = -1.0; de = -1.0; break;
      case 1: v = l1; dx = 1.0; de = 0.0; break;
      default: v = l2; dx = 0.0; de = 1.0; break;
    }
    return;
  }
  switch (k) {
    case 0: v = l0 * (2 * l0 - 1); dx = -(4 * l0 - 1); de = -(4 * l0 - 1); break;
    case 1: v = l1 * (2 * l1 - 1); dx = 4 * l1 - 1; de = 0.0; break;
    case 2: v = l2 * (2 * l2 - 1); dx = 0.0; de = 4 * l2 - 1; break;
    case 3: v = 4 * l1 * l2; dx = 4 * l2; de = 4 * l1; break;
    case 4: v = 4 * l0 * l2; dx = -4 * l2; de = 4 * (l0 - l2); break;
    default: v = 4 * l0 * l1; dx = 4 * (l0 - l1); de = -4 * l1; break;
  }
}

// ---- quadrature: FFC "default" scheme ---------------------------------
FX_HDC int tri_npts(int q) {
  return q <= 1 ? 1 : q == 2 ? 3 : q == 3 ? 6 : q == 4 ? 6 : q == 5 ? 7 : q == 6 ? 12
       : ((q + 2) / 2) * ((q + 2) / 2);
}
FX_HDC int seg_npts(int q) { return (q + 2) / 2; }

// large (collapsed Gauss-Jacobi) rules live in generated tables
#include "fx_quad_tables.h"

template <int Q>
FX_HD void tri_point(int i, double& xi, double& et, double& w) {
  if constexpr (Q <= 1) {
    xi = 1.0 / 3; et = 1.0 / 3; w = 0.5;
  } else if constexpr (Q == 2) {
    static constexpr double T[3][2] = {{1.0 / 6, 1.0 / 6}, {1.0 / 6, 2.0 / 3}, {2.0 / 3, 1.0 / 6}};
    xi = T[i][0]; et = T[i][1]; w = 1.0 / 6;
  } else if constexpr (Q == 3) {
    constexpr double a = 0.659027622374092, b = 0.231933368553031, c = 0.109039009072877;
    static constexpr double T[6][2] = {{a, b}, {a, c}, {b, a}, {b, c}, {c, a}, {c, b}};
    xi = T[i][0]; et = T[i][1]; w = 1.0 / 12;
  } else if constexpr (Q == 4) {
    constexpr double a = 0.816847572980459, b = 0.091576213509771;
    constexpr double c = 0.108103018168070, d = 0.445948490915965;
    static constexpr double T[6][2] = {{a, b}, {b, a}, {b, b}, {c, d}, {d, c}, {d, d}};
    xi = T[i][0]; et = T[i][1];
    w = i < 3 ? 0.109951743655322 / 2 : 0.223381589678011 / 2;
  } else if constexpr (Q == 5) {
    constexpr double a = 0.79742698535308720, b = 0.10128650732345633;
    constexpr double c = 0.05971587178976981, d = 0.47014206410511505;
    static constexpr double T[7][2] = {{1.0 / 3, 1.0 / 3}, {a, b}, {b, a}, {b, b}, {c, d}, {d, c}, {d, d}};
    xi = T[i][0]; et = T[i][1];
    w = i == 0 ? 0.225 / 2 : (i < 4 ? 0.12593918054482717 / 2 : 0.13239415278850616 / 2);
  } else if constexpr (Q == 6) {
    constexpr double a1 = 0.873821971016996, b1 = 0.063089014491502;
    constexpr double a2 = 0.501426509658179, b2 = 0.249286745170910;
    constexpr double a = 0.636502499121399, b = 0.310352451033785, c = 0.053145049844816;
    static constexpr double T[12][2] = {{a1, b1}, {b1, a1}, {b1, b1}, {a2, b2}, {b2, a2}, {b2, b2},
                                 {a, b}, {b, a}, {a, c}, {c, a}, {b, c}, {c, b}};
    xi = T[i][0]; et = T[i][1];
    w = i < 3 ? 0.050844906370207 / 2 : (i < 6 ? 0.116786275726379 / 2 : 0.082851075618374 / 2);
  } else {
    gj_tri_point<(Q + 2) / 2>(i, xi, et, w);
  }
}

// Gauss-Legendre on [0,1], m = (Q+2)/2 points
template <int Q>
FX_HD void seg_point(int i, double& s, double& w) {
  constexpr int M = (Q + 2) / 2;
  gl_seg_point<M>(i, s, w);
}

// ---- field evaluation ------------------------------------------------------------------------
// dofmaps are stored transposed on the device: dm[a * nc + cell] (coalesced across cells)
struct P2Phys {
  double v[6], x[6], y[6];
};
FX_HD void p2_phys(double xi, double et, const Geo& g, P2Phys& b) {
  P2Ref r;
  p2_ref(xi, et, r);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    b.v[i] = r.v[i];
    b.x[i] = r.dx[i] * g.k00 + r.de[i] * g.k10;
    b.y[i] = r.dx[i] * g.k01 + r.de[i] * g.k11;
  }
}
struct P1Phys {
  double v[3], x[3], y[3];
};
FX_HD void p1_phys(double xi, double et, const Geo& g, P1Phys& b) {
  P1Ref r;
  p1_ref(xi, et, r);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    b.v[i] = r.v[i];
    b.x[i] = r.dx[i] * g.k00 + r.de[i] * g.k10;
    b.y[i] = r.dx[i] * g.k01 + r.de[i] * g.k11;
  }
}

struct VG {  // value + gradient of a scalar field at a point
  double v, x, y;
};
FX_HD VG eval_p2(const double* __restrict__ vec, const int* __restrict__ dm, int nc, int cell,
                 int loff, const P2Phys& b) {
  VG r{0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double d = vec[dm[(loff + i) * (long)nc + cell]];
    r.v += d * b.v[i]; r.x += d * b.x[i]; r.y += d * b.y[i];
  }
  return r;
}
FX_HD VG eval_p1(const double* __restrict__ vec, const int* __restrict__ dm, int nc, int cell,
                 int loff, const P1Phys& b) {
  VG r{0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double d = vec[dm[(loff + i) * (long)nc + cell]];
    r.v += d * b.v[i]; r.x += d * b.x[i]; r.y += d * b.y[i];
  }
  return r;
}
FX_HD double eval_dg0(const double* __restrict__ vec, const int* __restrict__ dm, int nc, int cell,
                      int loff) {
  return vec[dm[loff * (long)nc + cell]];
}

}  // namespace fx
