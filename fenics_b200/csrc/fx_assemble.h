// Element-kernel engine (K2/K3/K4/K5 of SURVEY.md 2.3), shared between the CUDA kernels
// (fx_kernels.cu) and the host emulation used by the CPU tests (tests/hostemu).
//
// A form is a small hand-written functor:
//   struct Form {
//     using Space = SpW | SpZc | SpQ;           // test == trial space
//     static constexpr int QDEG, NP;            // quadrature degree (UFL estimate), #pointwise doubles
//     FX_HD static void point(PtArgs, double* d);              // fields at one quadrature point
//     template<class A> FX_HDC static void terms(A&, const double* d, const Par&);   // bilinear
//        -> A.template add<S,T>(c): integrand += c * (test slot S) * (trial slot T)
//     (linear forms: vterms, A.template add<S>(c))
//     struct Facet { QDEG, NP, point, terms|vterms }  or  using Facet = NoFacet;
//   };
// A "slot" is a (component, value | d/dx | d/dy) pair of the mixed element (SpaceT::slot0).
//
// Thread mapping of the cell kernels: blockIdx.y = test component I, block = 32 cells x nb(I)
// basis functions; every warp has one basis function index (uniform branches, broadcast
// constants), consecutive lanes = consecutive cells (coalesced SoA geometry/dofmap loads).
// Phase 1 stages the pointwise data of 32 cells x NQ points in shared memory; phase 2 computes
// one row of the element tensor per thread in registers and scatter-adds it.
#pragma once
#include "../../include/fenics_b200.h"  // FX_NFIELDS / FX_NPARAMS: one source of truth for the table sizes
#include "fx_elements.h"

namespace fx {

constexpr int NFIELDS = FX_NFIELDS;
constexpr int NPARAMS = FX_NPARAMS;
constexpr int CB = 32;  // cells per block

struct MeshView {
  int nc = 0;
  const double* geom = nullptr;  // [6][nc]
  const int* dm[S_N] = {};       // transposed dofmaps [ld][nc]
  int nbf = 0;
  const int* bf_cell = nullptr;
  const int* bf_local = nullptr;
  const int* bf_marker = nullptr;
  const double* bf_geo = nullptr;  // [3][nbf]: nx, ny, length
  // boundary facets chained per cell (deterministic assembly): bf_first[f] != 0 for the lowest facet of its cell,
  // bf_next[f] = the cell's next boundary facet or -1
  const unsigned char* bf_first = nullptr;
  const int* bf_next = nullptr;
};
struct State {
  const double* f[NFIELDS];
};
struct Par {
  double v[NPARAMS];
};
FX_HDC Par par_ones() {
  Par p{};
  for (int i = 0; i < NPARAMS; ++i) p.v[i] = 1.0;
  return p;
}

struct AsmArgs {
  MeshView m;
  State s;
  Par p;
  // matrix output
  double* val;
  const int* rowptr;
  const uint8_t* pos;  // [(a*LD+b)*nc + cell] offset of column b inside row a's CSR row
  // vector / scalar output
  double* vec;
  double* part;  // per-block partial sums (functionals)
  // deterministic (atomic-free) assembly: non-null => the kernels store every element-tensor row / entry here
  // ([cell][row][col] resp. [cell][row]) instead of scatter-adding it; k_gather_* then sum the contributions of each
  // CSR entry / dof in a fixed order (ascending cell)
  double* elt;
  int marker_mask;  // unused (facet functors carry their own MARKERS mask)
  int nc_owned;     // functionals: integrate over cells [0, nc_owned) only (partitioned mode)
};

struct PtArgs {
  const MeshView& m;
  const State& s;
  const Par& p;
  int cell;
  const Geo& g;
  double xi, et;   // reference coordinates of the point
  double nx, ny;   // outward normal (facet integrals only)
};

struct NoFacet {
  static constexpr bool present = false;
};

FX_HD Geo load_geo(const MeshView& m, int cell) {
  const long nc = m.nc;
  Geo g;
  g.k00 = m.geom[0 * nc + cell]; g.k01 = m.geom[1 * nc + cell];
  g.k10 = m.geom[2 * nc + cell]; g.k11 = m.geom[3 * nc + cell];
  g.adet = m.geom[4 * nc + cell]; g.h = m.geom[5 * nc + cell];
  return g;
}

FX_HD void atomic_add(double* p, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}

// ---- field helpers for the point() functors --------------------------------------------------
// velocity part (components 0,1) of a W-space vector
struct Vel {
  double u0, u1, g00, g01, g10, g11;  // gij = d u_i / d x_j  (UFL grad(u)[i,j])
};
FX_HD Vel eval_vel(const MeshView& m, const double* w, int cell, const P2Phys& b) {
  const VG a = eval_p2(w, m.dm[S_W], m.nc, cell, 0, b);
  const VG c = eval_p2(w, m.dm[S_W], m.nc, cell, 6, b);
  return Vel{a.v, c.v, a.x, a.y, c.x, c.y};
}
// second derivatives of the velocity components (P2: constant on a cell).  Needed where a form differentiates a
// function of grad(u): div(phi_def(u0) (...)) of the FENE-P-MP predictor (fenepmp_jbp.py:519, fenics_base.py:287-291)
struct VelH {
  double a_xx, a_xy, a_yy;  // u_0
  double c_xx, c_xy, c_yy;  // u_1
};
FX_HD VelH eval_vel_hess(const MeshView& m, const double* w, int cell, const Geo& g) {
  // reference second derivatives of the UFC P2 basis (xi xi, xi eta, eta eta)
  constexpr double fxx[6] = {4.0, 4.0, 0.0, 0.0, 0.0, -8.0};
  constexpr double fxe[6] = {4.0, 0.0, 0.0, 4.0, -4.0, -4.0};
  constexpr double fee[6] = {4.0, 0.0, 4.0, 0.0, -8.0, 0.0};
  VelH h{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    // d/dx = k00 d/dxi + k10 d/deta,  d/dy = k01 d/dxi + k11 d/deta  (p2_phys)
    const double hxx = g.k00 * g.k00 * fxx[i] + 2.0 * g.k00 * g.k10 * fxe[i] + g.k10 * g.k10 * fee[i];
    const double hxy = g.k00 * g.k01 * fxx[i] + (g.k00 * g.k11 + g.k10 * g.k01) * fxe[i] + g.k10 * g.k11 * fee[i];
    const double hyy = g.k01 * g.k01 * fxx[i] + 2.0 * g.k01 * g.k11 * fxe[i] + g.k11 * g.k11 * fee[i];
    const double a = w[m.dm[S_W][(long)i * m.nc + cell]], c = w[m.dm[S_W][(long)(6 + i) * m.nc + cell]];
    h.a_xx += a * hxx; h.a_xy += a * hxy; h.a_yy += a * hyy;
    h.c_xx += c * hxx; h.c_xy += c * hxy; h.c_yy += c * hyy;
  }
  return h;
}
// DEVSS part (components 2,3,4) of a W-space vector
struct Sym3 {
  double a, b, c;  // [[a,b],[b,c]]
};
FX_HD Sym3 eval_devss(const MeshView& m, const double* w, int cell) {
  return Sym3{eval_dg0(w, m.dm[S_W], m.nc, cell, 12), eval_dg0(w, m.dm[S_W], m.nc, cell, 13),
              eval_dg0(w, m.dm[S_W], m.nc, cell, 14)};
}
struct Tau {
  VG t0, t1, t2;
};
FX_HD Tau eval_tau(const MeshView& m, const double* t, int cell, const P2Phys& b) {
  return Tau{eval_p2(t, m.dm[S_ZC], m.nc, cell, 0, b), eval_p2(t, m.dm[S_ZC], m.nc, cell, 6, b),
             eval_p2(t, m.dm[S_ZC], m.nc, cell, 12, b)};
}
FX_HD VG eval_q(const MeshView& m, const double* q, int cell, const P1Phys& b) {
  return eval_p1(q, m.dm[S_Q], m.nc, cell, 0, b);
}

// ---- accumulators -----------------------------------------------------------------------------
template <class Sp>
struct MaskAcc {
  bool m[Sp::NS][Sp::NS] = {};
  template <int S, int T>
  constexpr void add(double) { m[S][T] = true; }
};
template <class TP>  // TP: provider of Space, NP, terms()
constexpr MaskAcc<typename TP::Space> form_mask() {
  MaskAcc<typename TP::Space> a;
  double d[TP::NP > 0 ? TP::NP : 1] = {};
  for (int i = 0; i < (TP::NP > 0 ? TP::NP : 1); ++i) d[i] = 1.0;
  const Par p = par_ones();
  TP::terms(a, d, p);
  return a;
}

template <class Sp, int I>
struct RowAcc {
  double g[Sp::NS];
  double bv, bx, by;
  template <int S, int T>
  FX_HD constexpr void add(double c) {
    if constexpr (Sp::slot_comp(S) == I) {
      constexpr int kk = S - Sp::slot0(I);
      g[T] += (kk == 0 ? bv : (kk == 1 ? bx : by)) * c;
    }
  }
};

template <class Sp, int I>
struct VecAcc {
  double sum;
  double bv, bx, by;
  template <int S>
  FX_HD constexpr void add(double c) {
    if constexpr (Sp::slot_comp(S) == I) {
      constexpr int kk = S - Sp::slot0(I);
      sum += (kk == 0 ? bv : (kk == 1 ? bx : by)) * c;
    }
  }
};

template <class TP, int I>
constexpr bool row_uses(int t) {  // does any test slot of component I couple to trial slot t?
  using Sp = typename TP::Space;
  constexpr auto M = form_mask<TP>();
  bool r = false;
  for (int k = 0; k < el_nslot(Sp::kind(I)); ++k) r = r || M.m[Sp::slot0(I) + k][t];
  return r;
}

// one quadrature point of one element-tensor row: test basis k of component I against all
// trial dofs.  wd = weight * measure scale.  acc[LD] accumulates the row.
template <class TP, int I>
FX_HD void row_qp(const Geo& g, int k, double xi, double et, double wd, const double* d,
                  const Par& p, double* acc) {
  using Sp = typename TP::Space;
  double tv, tdx, tde;
  ref_one(Sp::kind(I), k, xi, et, tv, tdx, tde);
  RowAcc<Sp, I> R;
  R.bv = wd * tv;
  R.bx = wd * (tdx * g.k00 + tde * g.k10);
  R.by = wd * (tdx * g.k01 + tde * g.k11);
#pragma unroll
  for (int t = 0; t < Sp::NS; ++t) R.g[t] = 0.0;
  TP::terms(R, d, p);
  P2Ref r2;
  P1Ref r1;
  p2_ref(xi, et, r2);
  p1_ref(xi, et, r1);
  static_for<0, Sp::NCOMP>([&](auto jc) {
    constexpr int J = decltype(jc)::value;
    constexpr int s0 = Sp::slot0(J), o = Sp::off(J), kind = Sp::kind(J);
    if constexpr (kind == DG0) {
      if constexpr (row_uses<TP, I>(s0)) acc[o] += R.g[s0];
    } else {
      constexpr bool mv = row_uses<TP, I>(s0);
      constexpr bool md = row_uses<TP, I>(s0 + 1) || row_uses<TP, I>(s0 + 2);
      if constexpr (mv || md) {
        double gxi = 0.0, get = 0.0;
        if constexpr (md) {  // pull the physical-gradient weights back to the reference frame
          gxi = g.k00 * R.g[s0 + 1] + g.k01 * R.g[s0 + 2];
          get = g.k10 * R.g[s0 + 1] + g.k11 * R.g[s0 + 2];
        }
        constexpr int nb = el_nb(kind);
#pragma unroll
        for (int l = 0; l < nb; ++l) {
          double c = 0.0;
          if constexpr (mv) c = R.g[s0] * (kind == P2 ? r2.v[l] : r1.v[l]);
          if constexpr (md)
            c += gxi * (kind == P2 ? r2.dx[l] : r1.dx[l]) + get * (kind == P2 ? r2.de[l] : r1.de[l]);
          acc[o + l] += c;
        }
      }
    }
  });
}

template <class TP, int I>
FX_HD double vec_qp(const Geo& g, int k, double xi, double et, double wd, const double* d,
                    const Par& p) {
  using Sp = typename TP::Space;
  double tv, tdx, tde;
  ref_one(Sp::kind(I), k, xi, et, tv, tdx, tde);
  VecAcc<Sp, I> R;
  R.sum = 0.0;
  R.bv = wd * tv;
  R.bx = wd * (tdx * g.k00 + tde * g.k10);
  R.by = wd * (tdx * g.k01 + tde * g.k11);
  TP::vterms(R, d, p);
  return R.sum;
}

// ---- cell kernels ------------------------------------------------------------------------------
template <class F>
FX_HD void cell_phase1(const AsmArgs& a, int cell0, int tid, int nthreads, double* smem) {
  constexpr int NQ = tri_npts(F::QDEG);
  if constexpr (F::NP > 0) {
    for (int task = tid; task < CB * NQ; task += nthreads) {
      const int lane = task % CB, q = task / CB, cell = cell0 + lane;
      if (cell < a.m.nc) {
        const Geo g = load_geo(a.m, cell);
        double xi, et, w;
        tri_point<F::QDEG>(q, xi, et, w);
        double d[F::NP];
        F::point(PtArgs{a.m, a.s, a.p, cell, g, xi, et, 0.0, 0.0}, d);
#pragma unroll
        for (int j = 0; j < F::NP; ++j) smem[(j * NQ + q) * CB + lane] = d[j];
      }
    }
  }
}

template <class Sp>
FX_HD void scatter_row(const AsmArgs& a, int cell, int arow, const double* acc) {
  const long nc = a.m.nc;
  const int r = a.m.dm[space_id<Sp>::v][arow * nc + cell];
  double* row = a.val + a.rowptr[r];
  const uint8_t* pp = a.pos + (long)arow * Sp::LD * nc + cell;
#pragma unroll
  for (int b = 0; b < Sp::LD; ++b) atomic_add(row + pp[b * nc], acc[b]);
}

template <class F, int I>
FX_HD void mat_rows(const AsmArgs& a, int cell0, int lane, int k, const double* smem) {
  using Sp = typename F::Space;
  constexpr int NQ = tri_npts(F::QDEG);
  const int cell = cell0 + lane;
  if (cell >= a.m.nc || k >= Sp::nb(I)) return;
  const Geo g = load_geo(a.m, cell);
  double acc[Sp::LD];
#pragma unroll
  for (int b = 0; b < Sp::LD; ++b) acc[b] = 0.0;
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    double xi, et, w;
    tri_point<F::QDEG>(q, xi, et, w);
    double d[F::NP > 0 ? F::NP : 1];
#pragma unroll
    for (int j = 0; j < F::NP; ++j) d[j] = smem[(j * NQ + q) * CB + lane];
    row_qp<F, I>(g, k, xi, et, w * g.adet, d, a.p, acc);
  }
  // scatter only the entries this form can make non-zero (the values buffer was zeroed): the rest of
  // DOLFIN's cell-block pattern stays an explicit 0.0, exactly as after PETSc ADD_VALUES of zeros
  if (a.elt != nullptr) {  // deterministic mode: the whole row (structural zeros included), gathered later
    double* dst = a.elt + ((long)cell * Sp::LD + Sp::off(I) + k) * Sp::LD;
#pragma unroll
    for (int b = 0; b < Sp::LD; ++b) dst[b] = acc[b];
  } else {
    const long nc = a.m.nc;
    const int arow = Sp::off(I) + k;
    const int r = a.m.dm[space_id<Sp>::v][arow * nc + cell];
    double* row = a.val + a.rowptr[r];
    const uint8_t* pp = a.pos + (long)arow * Sp::LD * nc + cell;
    static_for<0, Sp::NCOMP>([&](auto jc) {
      constexpr int J = decltype(jc)::value;
      constexpr int s0 = Sp::slot0(J), o = Sp::off(J);
      constexpr bool used = row_uses<F, I>(s0) ||
                            (Sp::kind(J) != DG0 && (row_uses<F, I>(s0 + 1) || row_uses<F, I>(s0 + 2)));
      if constexpr (used) {
#pragma unroll
        for (int l = 0; l < Sp::nb(J); ++l) atomic_add(row + pp[(o + l) * nc], acc[o + l]);
      }
    });
  }
}

template <class F, int I>
FX_HD void vec_rows(const AsmArgs& a, int cell0, int lane, int k, const double* smem) {
  using Sp = typename F::Space;
  constexpr int NQ = tri_npts(F::QDEG);
  const int cell = cell0 + lane;
  if (cell >= a.m.nc || k >= Sp::nb(I)) return;
  const Geo g = load_geo(a.m, cell);
  double sum = 0.0;
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    double xi, et, w;
    tri_point<F::QDEG>(q, xi, et, w);
    double d[F::NP > 0 ? F::NP : 1];
#pragma unroll
    for (int j = 0; j < F::NP; ++j) d[j] = smem[(j * NQ + q) * CB + lane];
    sum += vec_qp<F, I>(g, k, xi, et, w * g.adet, d, a.p);
  }
  if (a.elt != nullptr) {
    a.elt[(long)cell * Sp::LD + Sp::off(I) + k] = sum;
    return;
  }
  const int r = a.m.dm[space_id<Sp>::v][(long)(Sp::off(I) + k) * a.m.nc + cell];
  atomic_add(a.vec + r, sum);
}

// ---- exterior-facet kernels: one thread per (boundary facet, local test dof) -------------------
FX_HD void facet_ref_point(int lf, double s, double& xi, double& et) {
  // local facet i is opposite local vertex i; facet vertices in ascending local order
  // facet 0: v1->v2, facet 1: v0->v2, facet 2: v0->v1
  const double ax = (lf == 0) ? 1.0 : 0.0, ay = 0.0;
  const double bx = (lf == 2) ? 1.0 : 0.0, by = (lf == 2) ? 0.0 : 1.0;
  xi = ax * (1.0 - s) + bx * s;
  et = ay * (1.0 - s) + by * s;
}

// adds row `arow` of facet f's element tensor to acc; false when the facet's marker is not integrated over
template <class F>
FX_HD bool facet_mat_row_acc(const AsmArgs& a, int f, int arow, double* acc) {
  using Sp = typename F::Space;
  using FF = typename F::Facet;
  constexpr int NQ = seg_npts(FF::QDEG);
  const int marker = a.m.bf_marker[f];
  if (FF::MARKERS != -1 && !((FF::MARKERS >> marker) & 1)) return false;  // ds(i) vs ds
  const int cell = a.m.bf_cell[f], lf = a.m.bf_local[f];
  const Geo g = load_geo(a.m, cell);
  const double nx = a.m.bf_geo[f], ny = a.m.bf_geo[a.m.nbf + f], len = a.m.bf_geo[2 * a.m.nbf + f];
  const int I = Sp::dof_comp(arow);
  for (int q = 0; q < NQ; ++q) {
    double s, w, xi, et;
    seg_point<FF::QDEG>(q, s, w);
    facet_ref_point(lf, s, xi, et);
    double d[FF::NP > 0 ? FF::NP : 1];
    FF::point(PtArgs{a.m, a.s, a.p, cell, g, xi, et, nx, ny}, d);
    static_for<0, Sp::NCOMP>([&](auto ic) {
      constexpr int II = decltype(ic)::value;
      if (I == II) row_qp<FF, II>(g, arow - Sp::off(II), xi, et, w * len, d, a.p, acc);
    });
  }
  return true;
}
template <class F>
FX_HD void facet_mat_row(const AsmArgs& a, int f, int arow) {
  using Sp = typename F::Space;
  double acc[Sp::LD];
  for (int b = 0; b < Sp::LD; ++b) acc[b] = 0.0;
  if (a.elt != nullptr) {  // deterministic mode: the cell's lowest facet adds ALL of the cell's facets, in order
    if (!a.m.bf_first[f]) return;
    bool any = false;
    for (int g = f; g >= 0; g = a.m.bf_next[g]) any |= facet_mat_row_acc<F>(a, g, arow, acc);
    if (!any) return;
    double* dst = a.elt + ((long)a.m.bf_cell[f] * Sp::LD + arow) * Sp::LD;
    for (int b = 0; b < Sp::LD; ++b) dst[b] += acc[b];
    return;
  }
  if (facet_mat_row_acc<F>(a, f, arow, acc)) scatter_row<Sp>(a, a.m.bf_cell[f], arow, acc);
}

// value of facet f's contribution to entry `arow` of the element vector (0 when the marker is not integrated over)
template <class F>
FX_HD double facet_vec_row_val(const AsmArgs& a, int f, int arow) {
  using Sp = typename F::Space;
  using FF = typename F::Facet;
  constexpr int NQ = seg_npts(FF::QDEG);
  const int marker = a.m.bf_marker[f];
  if (FF::MARKERS != -1 && !((FF::MARKERS >> marker) & 1)) return 0.0;  // ds(i) vs ds
  const int cell = a.m.bf_cell[f], lf = a.m.bf_local[f];
  const Geo g = load_geo(a.m, cell);
  const double nx = a.m.bf_geo[f], ny = a.m.bf_geo[a.m.nbf + f], len = a.m.bf_geo[2 * a.m.nbf + f];
  const int I = Sp::dof_comp(arow);
  double sum = 0.0;
  for (int q = 0; q < NQ; ++q) {
    double s, w, xi, et;
    seg_point<FF::QDEG>(q, s, w);
    facet_ref_point(lf, s, xi, et);
    double d[FF::NP > 0 ? FF::NP : 1];
    FF::point(PtArgs{a.m, a.s, a.p, cell, g, xi, et, nx, ny}, d);
    static_for<0, Sp::NCOMP>([&](auto ic) {
      constexpr int II = decltype(ic)::value;
      if (I == II) sum += vec_qp<FF, II>(g, arow - Sp::off(II), xi, et, w * len, d, a.p);
    });
  }
  return sum;
}
template <class F>
FX_HD void facet_vec_row(const AsmArgs& a, int f, int arow) {
  using Sp = typename F::Space;
  using FF = typename F::Facet;
  const int cell = a.m.bf_cell[f];
  if (a.elt != nullptr) {
    if (!a.m.bf_first[f]) return;
    double sum = 0.0;
    for (int g = f; g >= 0; g = a.m.bf_next[g]) sum += facet_vec_row_val<F>(a, g, arow);
    a.elt[(long)cell * Sp::LD + arow] += sum;
    return;
  }
  const int marker = a.m.bf_marker[f];
  if (FF::MARKERS != -1 && !((FF::MARKERS >> marker) & 1)) return;
  const int r = a.m.dm[space_id<Sp>::v][(long)arow * a.m.nc + cell];
  atomic_add(a.vec + r, facet_vec_row_val<F>(a, f, arow));
}

// ---- functionals and DG0 projections: one thread per cell ---------------------------------------
template <class F>
FX_HD double cell_functional(const AsmArgs& a, int cell) {
  constexpr int NQ = tri_npts(F::QDEG);
  const Geo g = load_geo(a.m, cell);
  double sum = 0.0;
  for (int q = 0; q < NQ; ++q) {
    double xi, et, w;
    tri_point<F::QDEG>(q, xi, et, w);
    sum += w * F::integrand(PtArgs{a.m, a.s, a.p, cell, g, xi, et, 0.0, 0.0});
  }
  return sum * g.adet;
}

// exterior-facet functional (assemble(expr*ds(i))): one thread per boundary facet
template <class F>
FX_HD double facet_functional(const AsmArgs& a, int f) {
  constexpr int NQ = seg_npts(F::QDEG);
  const int marker = a.m.bf_marker[f];
  if (F::MARKERS != -1 && !((F::MARKERS >> marker) & 1)) return 0.0;
  const int cell = a.m.bf_cell[f], lf = a.m.bf_local[f];
  if (cell >= a.nc_owned) return 0.0;
  const Geo g = load_geo(a.m, cell);
  const double nx = a.m.bf_geo[f], ny = a.m.bf_geo[a.m.nbf + f], len = a.m.bf_geo[2 * a.m.nbf + f];
  double sum = 0.0;
  for (int q = 0; q < NQ; ++q) {
    double s, w, xi, et;
    seg_point<F::QDEG>(q, s, w);
    facet_ref_point(lf, s, xi, et);
    sum += w * F::integrand(PtArgs{a.m, a.s, a.p, cell, g, xi, et, nx, ny});
  }
  return sum * len;
}

// project(expr, DG0-space): b = sum_q w_q |detJ| f_q, M = |detJ|/2 (1-point rule), x = b / M
template <class E>
FX_HD void cell_project_dg0(const AsmArgs& a, int cell) {
  constexpr int NQ = tri_npts(E::QDEG);
  const Geo g = load_geo(a.m, cell);
  double sum[E::NOUT];
  for (int j = 0; j < E::NOUT; ++j) sum[j] = 0.0;
  for (int q = 0; q < NQ; ++q) {
    double xi, et, w;
    tri_point<E::QDEG>(q, xi, et, w);
    double v[E::NOUT];
    E::value(PtArgs{a.m, a.s, a.p, cell, g, xi, et, 0.0, 0.0}, v);
    for (int j = 0; j < E::NOUT; ++j) sum[j] += (w * g.adet) * v[j];
  }
  const double mass = 0.5 * g.adet;
  const int* dm = a.m.dm[E::NOUT == 3 ? S_ZD : S_QT];
  for (int j = 0; j < E::NOUT; ++j) a.vec[dm[(long)j * a.m.nc + cell]] = sum[j] / mass;
}

// ---- geometry (K1) ----------------------------------------------------------------------------
FX_HD void compute_geo(const double* xy, const int* cells, int cell, long nc, double* geom) {
  const int v0 = cells[3 * cell], v1 = cells[3 * cell + 1], v2 = cells[3 * cell + 2];
  const double x0 = xy[2 * v0], y0 = xy[2 * v0 + 1];
  const double j00 = xy[2 * v1] - x0, j10 = xy[2 * v1 + 1] - y0;  // J[:,0] = x1 - x0
  const double j01 = xy[2 * v2] - x0, j11 = xy[2 * v2 + 1] - y0;  // J[:,1] = x2 - x0
  const double det = j00 * j11 - j01 * j10;
  geom[0 * nc + cell] = j11 / det;
  geom[1 * nc + cell] = -j01 / det;
  geom[2 * nc + cell] = -j10 / det;
  geom[3 * nc + cell] = j00 / det;
  geom[4 * nc + cell] = fabs(det);
  const double e0 = hypot(j00 - j01, j10 - j11), e1 = hypot(j01, j11), e2 = hypot(j00, j10);
  geom[5 * nc + cell] = fmax(e0, fmax(e1, e2));  // CellDiameter: max vertex-vertex distance
}

FX_HD void compute_facet_geo(const double* xy, const int* cells, int cell, int lf, int f, long nbf,
                             double* out) {
  const int ia = (lf == 0) ? 1 : 0, ib = (lf == 2) ? 1 : 2;
  const int va = cells[3 * cell + ia], vb = cells[3 * cell + ib], vo = cells[3 * cell + lf];
  const double tx = xy[2 * vb] - xy[2 * va], ty = xy[2 * vb + 1] - xy[2 * va + 1];
  const double len = hypot(tx, ty);
  double nx = ty / len, ny = -tx / len;
  const double ox = xy[2 * vo] - xy[2 * va], oy = xy[2 * vo + 1] - xy[2 * va + 1];
  if (nx * ox + ny * oy > 0.0) { nx = -nx; ny = -ny; }
  out[f] = nx; out[nbf + f] = ny; out[2 * nbf + f] = len;
}

}  // namespace fx
