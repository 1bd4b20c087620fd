// TEST INFRASTRUCTURE (never loaded by the product): runs the SAME form functors and element-kernel
// phase functions as the CUDA kernels (fenics_b200/csrc/fx_assemble.h, fx_forms_*.h), serially on the
// host, emulating the block/thread mapping.  Lets the CPU test-suite check every hand-written
// functor against the oracle without a GPU.  Built by tests/hostemu/build.py with g++.
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/fenics_b200.h"
#include "../../fenics_b200/csrc/fx_plan_host.h"
#include "../../fenics_b200/csrc/fx_registry.h"

using namespace fx;

struct EmuPlan {
  int nv, nc, nbf;
  std::vector<double> xy, geom, bf_geo;
  std::vector<int> cells, bf_cell, bf_local, bf_marker;
  HostPattern pat[S_N];
  bool has[S_N] = {};
  MeshView view() const {
    MeshView m;
    m.nc = nc; m.geom = geom.data();
    for (int s = 0; s < S_N; ++s) m.dm[s] = has[s] ? pat[s].dm_t.data() : nullptr;
    m.nbf = nbf; m.bf_cell = bf_cell.data(); m.bf_local = bf_local.data();
    m.bf_marker = bf_marker.data(); m.bf_geo = bf_geo.data();
    return m;
  }
};

static AsmArgs make_args(const EmuPlan* P, const double* const* state, const double* params) {
  AsmArgs a{};
  a.m = P->view();
  for (int i = 0; i < NFIELDS; ++i) a.s.f[i] = state ? state[i] : nullptr;
  for (int i = 0; i < NPARAMS; ++i) a.p.v[i] = params ? params[i] : 0.0;
  a.marker_mask = -1;
  a.nc_owned = P->nc;
  return a;
}

template <class F>
static void run_matrix(const EmuPlan* P, AsmArgs a, double* val) {
  using Sp = typename F::Space;
  const HostPattern& pat = P->pat[space_id<Sp>::v];
  a.val = val; a.rowptr = pat.rowptr.data(); a.pos = pat.pos.data();
  std::memset(val, 0, sizeof(double) * pat.nnz);
  constexpr int NQ = tri_npts(F::QDEG);
  std::vector<double> smem((size_t)(F::NP > 0 ? F::NP : 1) * NQ * CB);
  const int nthreads = CB * Sp::maxnb();
  for (int cell0 = 0; cell0 < P->nc; cell0 += CB) {
    for (int tid = 0; tid < nthreads; ++tid) cell_phase1<F>(a, cell0, tid, nthreads, smem.data());
    static_for<0, Sp::NCOMP>([&](auto ic) {
      constexpr int I = decltype(ic)::value;
      for (int k = 0; k < Sp::nb(I); ++k)
        for (int lane = 0; lane < CB; ++lane) mat_rows<F, I>(a, cell0, lane, k, smem.data());
    });
  }
  if constexpr (F::Facet::present) {
    for (int f = 0; f < P->nbf; ++f)
      for (int arow = 0; arow < Sp::LD; ++arow) facet_mat_row<F>(a, f, arow);
  }
}

template <class F>
static void run_vector(const EmuPlan* P, AsmArgs a, double* b) {
  using Sp = typename F::Space;
  a.vec = b;
  std::memset(b, 0, sizeof(double) * P->pat[space_id<Sp>::v].n);
  constexpr int NQ = tri_npts(F::QDEG);
  std::vector<double> smem((size_t)(F::NP > 0 ? F::NP : 1) * NQ * CB);
  const int nthreads = CB * Sp::maxnb();
  for (int cell0 = 0; cell0 < P->nc; cell0 += CB) {
    for (int tid = 0; tid < nthreads; ++tid) cell_phase1<F>(a, cell0, tid, nthreads, smem.data());
    static_for<0, Sp::NCOMP>([&](auto ic) {
      constexpr int I = decltype(ic)::value;
      for (int k = 0; k < Sp::nb(I); ++k)
        for (int lane = 0; lane < CB; ++lane) vec_rows<F, I>(a, cell0, lane, k, smem.data());
    });
  }
  if constexpr (F::Facet::present) {
    for (int f = 0; f < P->nbf; ++f)
      for (int arow = 0; arow < Sp::LD; ++arow) facet_vec_row<F>(a, f, arow);
  }
}

extern "C" {

void* emu_plan_create(const double* xy, int nv, const int* cells, int nc, const int* const* cell_dofs,
                      const int* bf_cell, const int* bf_local, const int* bf_marker, int nbf) {
  EmuPlan* P = new EmuPlan();
  P->nv = nv; P->nc = nc; P->nbf = nbf;
  P->xy.assign(xy, xy + 2 * (size_t)nv);
  P->cells.assign(cells, cells + 3 * (size_t)nc);
  P->bf_cell.assign(bf_cell, bf_cell + nbf);
  P->bf_local.assign(bf_local, bf_local + nbf);
  P->bf_marker.assign(bf_marker, bf_marker + nbf);
  P->geom.resize(6 * (size_t)nc);
  for (int c = 0; c < nc; ++c) compute_geo(P->xy.data(), P->cells.data(), c, nc, P->geom.data());
  P->bf_geo.resize(3 * (size_t)nbf + 1);
  for (int f = 0; f < nbf; ++f)
    compute_facet_geo(P->xy.data(), P->cells.data(), bf_cell[f], bf_local[f], f, nbf, P->bf_geo.data());
  std::string err;
  for (int s = 0; s < S_N; ++s) {
    if (!cell_dofs[s]) continue;
    const bool want = (s == S_W || s == S_ZC || s == S_Q || s == S_V || s == S_VS);
    if (!build_pattern(cell_dofs[s], nc, space_ld(s), want, P->pat[s], err)) {
      std::fprintf(stderr, "emu_plan_create: %s\n", err.c_str());
      delete P;
      return nullptr;
    }
    P->has[s] = true;
  }
  return P;
}
void emu_plan_destroy(void* p) { delete (EmuPlan*)p; }

int emu_pattern(void* p, int space, int* n, int* nnz, const int** rowptr, const int** col) {
  EmuPlan* P = (EmuPlan*)p;
  *n = P->pat[space].n; *nnz = P->pat[space].nnz;
  *rowptr = P->pat[space].rowptr.data(); *col = P->pat[space].col.data();
  return 0;
}
const double* emu_geometry(void* p) { return ((EmuPlan*)p)->geom.data(); }
const double* emu_facet_geometry(void* p) { return ((EmuPlan*)p)->bf_geo.data(); }

int emu_assemble_matrix(void* p, int form, const double* const* state, const double* params, double* val) {
  EmuPlan* P = (EmuPlan*)p;
  AsmArgs a = make_args(P, state, params);
  switch (form) {
#define X(id, F) case id: run_matrix<F>(P, a, val); return 0;
    FX_MATRIX_FORMS(X)
#undef X
  }
  return FX_ERR_ARG;
}

int emu_assemble_vector(void* p, int form, const double* const* state, const double* params, double* b) {
  EmuPlan* P = (EmuPlan*)p;
  AsmArgs a = make_args(P, state, params);
  const bool conv0 = params[ldc::P_CONV] == 0.0;
  switch (form) {
#define X(id, F) case id: run_vector<F>(P, a, b); return 0;
#define XC(id, F0, F1) case id: if (conv0) run_vector<F0>(P, a, b); else run_vector<F1>(P, a, b); return 0;
    FX_VECTOR_FORMS(X, XC)
#undef X
#undef XC
  }
  return FX_ERR_ARG;
}

int emu_assemble_scalar(void* p, int form, const double* const* state, const double* params, double* out) {
  EmuPlan* P = (EmuPlan*)p;
  AsmArgs a = make_args(P, state, params);
  double sum = 0.0;
  switch (form) {
#define X(id, F) case id: for (int c = 0; c < P->nc; ++c) sum += cell_functional<F>(a, c); *out = sum; return 0;
    FX_SCALAR_FORMS(X)
#undef X
#define X(id, F) case id: for (int f = 0; f < P->nbf; ++f) sum += facet_functional<F>(a, f); *out = sum; return 0;
    FX_FACET_SCALAR_FORMS(X)
#undef X
  }
  return FX_ERR_ARG;
}

int emu_project_dg0(void* p, int expr, const double* const* state, const double* params, double* out) {
  EmuPlan* P = (EmuPlan*)p;
  AsmArgs a = make_args(P, state, params);
  a.vec = out;
  switch (expr) {
#define X(id, E) case id: for (int c = 0; c < P->nc; ++c) cell_project_dg0<E>(a, c); return 0;
    FX_DG0_EXPRS(X)
#undef X
  }
  return FX_ERR_ARG;
}

int emu_form_qdeg(int form, int* qdx, int* qds, double conv) {
  switch (form) {
#define X(id, F) case id: *qdx = F::QDEG; *qds = facet_qdeg<F>(); return 0;
#define XC(id, F0, F1) case id: *qdx = (conv == 0.0) ? F0::QDEG : F1::QDEG; *qds = -1; return 0;
    FX_MATRIX_FORMS(X)
    FX_VECTOR_FORMS(X, XC)
#undef X
#undef XC
#define X(id, F) case id: *qdx = F::QDEG; *qds = -1; return 0;
    FX_SCALAR_FORMS(X)
#undef X
#define X(id, F) case id: *qdx = -1; *qds = F::QDEG; return 0;
    FX_FACET_SCALAR_FORMS(X)
#undef X
  }
  return FX_ERR_ARG;
}

}  // extern "C"
