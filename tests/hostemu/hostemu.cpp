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
#include "../../fenics_b200/csrc/fx_nb_host.h"

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

// Node-block solver layout (fx_nb_host.h): builds it for a space and replays the device SpMV (fx_nb.cuh: chunk ->
// stream phase over lanes/groups -> per-(row, component) reduce) on the host, with the same bounds the kernel relies
// on asserted.  y gets (A x) on the rows inside node blocks; other entries are left alone.  Returns 0, or -1 when
// the space has no usable node-block structure (why is printed), or > 0 = number of violated kernel invariants.
int emu_nb_spmv(void* p, int space, const unsigned char* mask, int num_sms, const double* val, const double* x,
                double* y, int* info /* [8]: bs, nn, N, ne, nchunks, grid, nzchk, bad zero-checks */,
                const unsigned char* ghost_owner /* null, or 1 + owner rank of ghost dofs */) {
  EmuPlan* P = (EmuPlan*)p;
  const HostPattern& pat = P->pat[space];
  int bs = 0, nloc = 0;
  if (!nb_space_layout(space, bs, nloc)) return -1;
  NbHost H;
  std::string why;
  if (!nb_build(pat.dm_t.data(), P->nc, pat.ld, bs, nloc, pat.n, pat.rowptr.data(), pat.col.data(), mask, num_sms, H, why, ghost_owner)) {
    std::fprintf(stderr, "emu_nb_spmv: %s\n", why.c_str());
    return -1;
  }
  const int bp = H.bp, b2 = bs * bs, CH = nb_chunk(bs), RM = nb_rows(bs);
  int viol = 0, bad = 0;
  std::vector<double> bval((size_t)H.nslots);
  for (long long k = 0; k < H.nslots; ++k) bval[k] = H.src[k] >= 0 ? val[H.src[k]] : 0.0;
  for (int j : H.zchk) bad += (val[j] != 0.0);
  std::vector<double> xi((size_t)H.N, 0.0), yi((size_t)H.N, 0.0), prod((size_t)CH * bs);
  for (int i = 0; i < H.N; ++i) xi[i] = H.ext_of_int[i] >= 0 ? x[H.ext_of_int[i]] : 0.0;
  const int nw = H.grid * NB_SLOTS;
  std::vector<char> row_done(H.nn, 0);
  for (int w = 0; w < nw; ++w)
    for (int k = H.wrange[w]; k < H.wrange[w + 1]; ++k) {
      const int r0 = H.desc[4 * k], nr = H.desc[4 * k + 1], e0 = H.desc[4 * k + 2], e1 = H.desc[4 * k + 3];
      if (nr < 1 || nr > RM || e1 - e0 > CH || e1 < e0) ++viol;
      for (int e = e0; e < e1; ++e) {  // stream phase (lane = (e - e0) % 32)
        const int c = H.bcol[e];
        for (int ci = 0; ci < bs; ++ci) {
          double sacc = 0.0;
          for (int cj = 0; cj < bs; ++cj) sacc += bval[(size_t)nb_pos(e, ci * bs + cj, bs)] * xi[(size_t)c * bp + cj];
          prod[(size_t)(e - e0) * bs + ci] = sacc;
        }
      }
      for (int lane = 0; lane < 32; ++lane) {  // reduce phase
        const int rl = lane / bp, cl = lane % bp;
        if (rl >= nr) continue;
        const int s1 = H.bre[r0 + rl] - e0, s0 = rl == 0 ? 0 : H.bre[r0 + rl - 1] - e0;
        if (s0 < 0 || s1 < s0 || s1 > e1 - e0) ++viol;
        double sacc = 0.0;
        if (cl < bs)
          for (int j = s0; j < s1; ++j) sacc += prod[(size_t)j * bs + cl];
        yi[(size_t)(r0 + rl) * bp + cl] = sacc;
        if (cl == 0) { if (row_done[r0 + rl]) ++viol; row_done[r0 + rl] = 1; }
      }
    }
  for (int i = 0; i < H.nn_chunked; ++i) viol += !row_done[i];
  for (int i = H.nn_chunked; i < H.nn; ++i) viol += (row_done[i] || !(H.imask[(size_t)i * bp] & 2));
  if (H.wrange[nw] != H.nchunks) ++viol;
  for (int i = 0; i < H.N; ++i)
    if (H.ext_of_int[i] >= 0) y[H.ext_of_int[i]] = yi[i];
  // the diagonal block entries
  for (int i = 0; i < H.nn; ++i)
    if (H.diag[i] >= 0 && H.bcol[H.diag[i]] != i) ++viol;
  info[0] = bs; info[1] = H.nn; info[2] = H.N; info[3] = (int)H.ne; info[4] = H.nchunks; info[5] = H.grid;
  info[6] = (int)H.zchk.size(); info[7] = bad;
  return viol;
}

}  // extern "C"
