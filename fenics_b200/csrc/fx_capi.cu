// extern "C" entry points declared in include/fenics_b200.h + plan construction.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "fx_plan.h"
#include "fx_plan_host.h"

namespace fx {

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches.load(); }

std::string& last_error() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

template <class T>
static int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) return FX_OK;
  FX_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
  FX_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return FX_OK;
}

int ensure_work(fx_plan* P, size_t n) {
  if (P->work_n >= n) return FX_OK;
  if (P->work) cudaFree(P->work);
  P->work = nullptr;
  P->work_n = 0;
  FX_CUDA(cudaMalloc((void**)&P->work, sizeof(double) * n * KRYLOV_NWORK));
  P->work_n = n;
  return FX_OK;
}

static AsmArgs make_args(const fx_plan* P, const double* const* state, const double* params) {
  AsmArgs a{};
  a.m = P->view();
  for (int i = 0; i < NFIELDS; ++i) a.s.f[i] = state ? state[i] : nullptr;
  for (int i = 0; i < NPARAMS; ++i) a.p.v[i] = params ? params[i] : 0.0;
  a.marker_mask = -1;
  a.nc_owned = P->nc_owned;
  return a;
}

}  // namespace fx

using namespace fx;

extern "C" {

const char* fx_last_error(void) { return last_error().c_str(); }
int fx_version(void) { return 100; }
long long fx_launch_count(void) { return launch_count(); }
int fx_copy_to_host(void* dst, const void* src, long long nbytes) {
  if (!dst || !src || nbytes < 0) return set_error(FX_ERR_ARG, "fx_copy_to_host: bad argument");
  FX_CUDA(cudaMemcpy(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost));
  return FX_OK;
}
int fx_space_local_dim(int space) { return (space >= 0 && space < S_N) ? space_ld(space) : FX_ERR_ARG; }

int fx_plan_create(const double* xy, int nv, const int* cells, int nc, const int* const* cell_dofs,
                   const int* bfacet_cell, const int* bfacet_local, const int* bfacet_marker, int nbf, int device,
                   fx_plan** out) {
  if (!xy || !cells || !cell_dofs || !out || nv <= 0 || nc <= 0 || nbf < 0)
    return set_error(FX_ERR_ARG, "fx_plan_create: bad argument");
  for (int c = 0; c < nc; ++c) {
    const int* v = cells + 3 * (size_t)c;
    if (!(v[0] >= 0 && v[0] < v[1] && v[1] < v[2] && v[2] < nv))
      return set_error(FX_ERR_ARG, "fx_plan_create: cell %d vertices not ascending / out of range (UFC order)", c);
  }
  for (int f = 0; f < nbf; ++f)
    if (bfacet_cell[f] < 0 || bfacet_cell[f] >= nc || bfacet_local[f] < 0 || bfacet_local[f] > 2 ||
        bfacet_marker[f] < 0 || bfacet_marker[f] > 30)
      return set_error(FX_ERR_ARG, "fx_plan_create: bad boundary facet %d", f);
  FX_CUDA(cudaSetDevice(device));
  fx_plan* P = new fx_plan();
  P->device = device;
  P->nv = nv; P->nc = nc; P->nbf = nbf;
  P->nc_owned = nc;
  int rc = FX_OK;
  auto fail = [&](int code) { fx_plan_destroy(P); return code; };
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(set_error(FX_ERR_CUDA, "no device %d", device));
  P->num_sms = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) return fail(set_error(FX_ERR_UNSUPPORTED, "device lacks cooperative launch"));
  if ((rc = upload(&P->xy, xy, 2 * (size_t)nv))) return fail(rc);
  if ((rc = upload(&P->cells, cells, 3 * (size_t)nc))) return fail(rc);
  if ((rc = upload(&P->bf_cell, bfacet_cell, (size_t)nbf))) return fail(rc);
  if ((rc = upload(&P->bf_local, bfacet_local, (size_t)nbf))) return fail(rc);
  if ((rc = upload(&P->bf_marker, bfacet_marker, (size_t)nbf))) return fail(rc);
  {  // boundary facets of one cell chained in ascending facet order (deterministic facet integrals)
    std::vector<unsigned char> first((size_t)nbf, 1);
    std::vector<int> next((size_t)nbf, -1), last((size_t)nc, -1);
    for (int f = 0; f < nbf; ++f) {
      const int c = bfacet_cell[f];
      if (last[c] >= 0) { next[last[c]] = f; first[f] = 0; }
      last[c] = f;
    }
    if ((rc = upload(&P->bf_first, first.data(), (size_t)nbf))) return fail(rc);
    if ((rc = upload(&P->bf_next, next.data(), (size_t)nbf))) return fail(rc);
  }
  if (const char* e = getenv("FX_ASM_MODE")) P->asm_mode = atoi(e) ? 1 : 0;
  if (cudaMalloc((void**)&P->geom, sizeof(double) * 6 * (size_t)nc) != cudaSuccess ||
      cudaMalloc((void**)&P->bf_geo, sizeof(double) * (3 * (size_t)nbf + 1)) != cudaSuccess)
    return fail(set_error(FX_ERR_CUDA, "cudaMalloc geometry failed"));
  size_t maxn = 0;
  for (int s = 0; s < S_N; ++s) {
    if (!cell_dofs[s]) continue;
    HostPattern hp;
    std::string err;
    const bool want = (s == S_W || s == S_ZC || s == S_Q || s == S_V || s == S_VS);
    if (!build_pattern(cell_dofs[s], nc, space_ld(s), want, hp, err))
      return fail(set_error(FX_ERR_ARG, "fx_plan_create: space %d: %s", s, err.c_str()));
    DevPattern& d = P->sp[s];
    d.n = hp.n; d.nnz = hp.nnz; d.ld = hp.ld; d.present = true; d.has_csr = want;
    if ((rc = upload(&d.dm_t, hp.dm_t.data(), hp.dm_t.size()))) return fail(rc);
    if (want) {
      if ((rc = upload(&d.rowptr, hp.rowptr.data(), hp.rowptr.size()))) return fail(rc);
      if ((rc = upload(&d.col, hp.col.data(), hp.col.size()))) return fail(rc);
      if ((rc = upload(&d.pos, hp.pos.data(), hp.pos.size()))) return fail(rc);
      if ((rc = upload(&d.rowblk, hp.rowblk.data(), hp.rowblk.size()))) return fail(rc);
      d.nrowblk = (int)hp.rowblk.size() - 1;
      d.rowblk_h = hp.rowblk;
      for (int b = 0; b < d.nrowblk; ++b)
        d.maxblknnz = std::max(d.maxblknnz, hp.rowptr[hp.rowblk[b + 1]] - hp.rowptr[hp.rowblk[b]]);
      for (int r = 0; r < hp.n; ++r) d.maxrow = std::max(d.maxrow, hp.rowptr[r + 1] - hp.rowptr[r]);
    }
    maxn = std::max(maxn, (size_t)hp.n);
  }
  size_t maxnnz = 0, maxblk = 0;
  for (int s = 0; s < S_N; ++s) {
    maxnnz = std::max(maxnnz, (size_t)P->sp[s].nnz);
    maxblk = std::max(maxblk, (size_t)P->sp[s].nrowblk);
    if (P->sp[s].has_csr && P->sp[s].maxrow < SPMV_NNZB)  // blocks compact_reblock can build: (nnz + 16 n) / T + 2
      maxblk = std::max(maxblk, ((size_t)P->sp[s].nnz + 16 * (size_t)P->sp[s].n) / (size_t)(SPMV_NNZB - P->sp[s].maxrow) + 2);
  }
  if (maxnnz > 0 &&
      (cudaMalloc((void**)&P->cval, sizeof(double) * maxnnz) != cudaSuccess ||
       cudaMalloc((void**)&P->ccol, sizeof(int) * maxnnz) != cudaSuccess ||
       cudaMalloc((void**)&P->cre, sizeof(int) * maxn) != cudaSuccess ||
       cudaMalloc((void**)&P->cdesc, sizeof(int4) * (maxblk + 1)) != cudaSuccess ||
       cudaMalloc((void**)&P->cprefix, sizeof(int) * (maxblk + 2)) != cudaSuccess ||
       cudaMalloc((void**)&P->scanpart, sizeof(int) * 2 * MAX_GRID) != cudaSuccess ||
       cudaMalloc((void**)&P->rowpre, sizeof(int2) * (maxn + 1)) != cudaSuccess ||
       cudaMalloc((void**)&P->nblk, 64) != cudaSuccess ||
       cudaMalloc((void**)&P->bar, 64) != cudaSuccess))
    return fail(set_error(FX_ERR_CUDA, "cudaMalloc compacted-matrix buffers failed"));
  P->part_cap = std::max((nc + 255) / 256, 148 * 8);
  if (cudaMalloc((void**)&P->part, sizeof(double) * P->part_cap) != cudaSuccess ||
      cudaMalloc((void**)&P->scal, sizeof(double) * 64) != cudaSuccess ||
      cudaMalloc((void**)&P->red, sizeof(double) * (2 * RED_SLOTS * MAX_GRID + 3 * RED_SLOTS + 64)) != cudaSuccess ||
      cudaMalloc((void**)&P->info, sizeof(fx_solve_info) * FX_NTICKETS) != cudaSuccess)
    return fail(set_error(FX_ERR_CUDA, "cudaMalloc scratch failed"));
  cudaMemset(P->info, 0, sizeof(fx_solve_info) * FX_NTICKETS);
  if ((rc = ensure_work(P, maxn))) return fail(rc);
  if ((rc = launch_geometry(P, 0))) return fail(rc);
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(set_error(FX_ERR_CUDA, "geometry kernel failed"));
  *out = P;
  return FX_OK;
}

void fx_plan_destroy(fx_plan* P) {
  if (!P) return;
  cudaFree(P->xy); cudaFree(P->cells); cudaFree(P->geom);
  cudaFree(P->bf_cell); cudaFree(P->bf_local); cudaFree(P->bf_marker); cudaFree(P->bf_geo);
  for (int s = 0; s < S_N; ++s) {
    cudaFree(P->sp[s].rowptr); cudaFree(P->sp[s].col); cudaFree(P->sp[s].pos); cudaFree(P->sp[s].dm_t); cudaFree(P->sp[s].rowblk);
    cudaFree(P->sp[s].elim); cudaFree(P->sp[s].elim_rows); cudaFree(P->sp[s].agg);
    cudaFree(P->sp[s].s_first); cudaFree(P->sp[s].s_peer); cudaFree(P->sp[s].s_local); cudaFree(P->sp[s].s_remote);
    P->sp[s].nb.release();
    cudaFree(P->sp[s].det_cptr); cudaFree(P->sp[s].det_contrib); cudaFree(P->sp[s].det_aptr); cudaFree(P->sp[s].det_adj);
  }
  cudaFree(P->nbval); cudaFree(P->elt); cudaFree(P->bf_first); cudaFree(P->bf_next);
  cudaFree(P->cval); cudaFree(P->ccol); cudaFree(P->cre); cudaFree(P->cdesc); cudaFree(P->cprefix); cudaFree(P->scanpart); cudaFree(P->rowpre); cudaFree(P->nblk); cudaFree(P->bar); cudaFree(P->coarse);
  cudaFree(P->gmres_basis); cudaFree(P->red_big);
  ilu_release(P);
  cudaFree(P->part); cudaFree(P->scal); cudaFree(P->work); cudaFree(P->red); cudaFree(P->info);
  delete P;
}

int fx_plan_set_markers(fx_plan* P, const int* m, int nbf) {
  if (!P || !m || nbf != P->nbf) return set_error(FX_ERR_ARG, "fx_plan_set_markers: bad argument");
  for (int f = 0; f < nbf; ++f)
    if (m[f] < 0 || m[f] > 30) return set_error(FX_ERR_ARG, "marker out of range [0,30]");
  FX_CUDA(cudaMemcpy(P->bf_marker, m, sizeof(int) * nbf, cudaMemcpyHostToDevice));
  return FX_OK;
}

int fx_space_dim(const fx_plan* P, int space) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].present) return FX_ERR_ARG;
  return P->sp[space].n;
}

int fx_pattern(const fx_plan* P, int space, int* n, int* nnz, const int** rowptr, const int** col) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_csr)
    return set_error(FX_ERR_ARG, "fx_pattern: space %d has no pattern", space);
  if (n) *n = P->sp[space].n;
  if (nnz) *nnz = P->sp[space].nnz;
  if (rowptr) *rowptr = P->sp[space].rowptr;
  if (col) *col = P->sp[space].col;
  return FX_OK;
}

int fx_geometry(const fx_plan* P, const double** geom) {
  if (!P || !geom) return set_error(FX_ERR_ARG, "fx_geometry: bad argument");
  *geom = P->geom;
  return FX_OK;
}

int fx_assemble_matrix(fx_plan* P, int form, const double* const* state, const double* params, double* val,
                       void* stream) {
  if (!P || !val) return set_error(FX_ERR_ARG, "fx_assemble_matrix: bad argument");
  AsmArgs a = make_args(P, state, params);
  a.val = val;
  return launch_matrix(P, form, a, (cudaStream_t)stream);
}

int fx_assemble_vector(fx_plan* P, int form, const double* const* state, const double* params, double* b,
                       void* stream) {
  if (!P || !b) return set_error(FX_ERR_ARG, "fx_assemble_vector: bad argument");
  AsmArgs a = make_args(P, state, params);
  a.vec = b;
  return launch_vector(P, form, a, (cudaStream_t)stream);
}

int fx_assemble_scalar_dev(fx_plan* P, int form, const double* const* state, const double* params,
                           double* out_dev, void* stream) {
  if (!P || !out_dev) return set_error(FX_ERR_ARG, "fx_assemble_scalar_dev: bad argument");
  AsmArgs a = make_args(P, state, params);
  return launch_scalar(P, form, a, out_dev, (cudaStream_t)stream);
}

int fx_assemble_scalar(fx_plan* P, int form, const double* const* state, const double* params, double* out,
                       void* stream) {
  if (!P || !out) return set_error(FX_ERR_ARG, "fx_assemble_scalar: bad argument");
  int rc = fx_assemble_scalar_dev(P, form, state, params, P->scal, stream);
  if (rc) return rc;
  FX_CUDA(cudaMemcpyAsync(out, P->scal, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return FX_OK;
}

int fx_project_dg0(fx_plan* P, int expr, const double* const* state, const double* params, double* out_dev,
                   void* stream) {
  if (!P || !out_dev) return set_error(FX_ERR_ARG, "fx_project_dg0: bad argument");
  AsmArgs a = make_args(P, state, params);
  a.vec = out_dev;
  return launch_dg0(P, expr, a, (cudaStream_t)stream);
}

int fx_form_info(int form, int* rank, int* space, int* qdeg) {
  if (!rank || !space || !qdeg) return set_error(FX_ERR_ARG, "fx_form_info: bad argument");
  return form_info(form, rank, space, qdeg);
}

int fx_apply_bc(fx_plan* P, int space, double* val, double* b, const int* rows, const double* g, int n,
                void* stream) {
  if (!P || space < 0 || space >= S_N || (n > 0 && (!rows || !g)))
    return set_error(FX_ERR_ARG, "fx_apply_bc: bad argument");
  return la_apply_bc(P, space, val, b, rows, g, n, (cudaStream_t)stream);
}

int fx_spmv(fx_plan* P, int space, const double* val, const double* x, double* y, void* stream) {
  if (!P || space < 0 || space >= S_N || !val || !x || !y) return set_error(FX_ERR_ARG, "fx_spmv: bad argument");
  return la_spmv(P, space, val, x, y, (cudaStream_t)stream);
}

int fx_spmv_variant(fx_plan* P, int space, const double* val, const double* x, double* y, int variant, void* stream) {
  if (!P || space < 0 || space >= S_N || !val || !x || !y) return set_error(FX_ERR_ARG, "fx_spmv_variant: bad argument");
  return la_spmv_variant(P, space, val, x, y, variant, (cudaStream_t)stream);
}

int fx_axpy(int n, double alpha, const double* x, double* y, void* stream) {
  if (n < 0 || !x || !y) return set_error(FX_ERR_ARG, "fx_axpy: bad argument");
  return la_axpy(n, alpha, x, y, (cudaStream_t)stream);
}

int fx_dot(fx_plan* P, int n, const double* x, const double* y, double* out, void* stream) {
  if (!P || n < 0 || !x || !y || !out) return set_error(FX_ERR_ARG, "fx_dot: bad argument");
  int rc = la_dot(P, n, x, y, P->scal + 1, (cudaStream_t)stream);
  if (rc) return rc;
  FX_CUDA(cudaMemcpyAsync(out, P->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return FX_OK;
}

int fx_norm(fx_plan* P, int n, const double* x, int kind, double* out, void* stream) {
  if (!P || n < 0 || !x || !out) return set_error(FX_ERR_ARG, "fx_norm: bad argument");
  int rc = la_norm(P, n, x, kind, P->scal + 2, (cudaStream_t)stream);
  if (rc) return rc;
  FX_CUDA(cudaMemcpyAsync(out, P->scal + 2, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return FX_OK;
}

int fx_dot_dev(fx_plan* P, int n, const double* x, const double* y, double* out_dev, void* stream) {
  if (!P || n < 0 || !x || !y || !out_dev) return set_error(FX_ERR_ARG, "fx_dot_dev: bad argument");
  return la_dot(P, n, x, y, out_dev, (cudaStream_t)stream);
}

int fx_norm_dev(fx_plan* P, int n, const double* x, int kind, double* out_dev, void* stream) {
  if (!P || n < 0 || !x || !out_dev) return set_error(FX_ERR_ARG, "fx_norm_dev: bad argument");
  return la_norm(P, n, x, kind, out_dev, (cudaStream_t)stream);
}

int fx_lincomb3_dev(int n, double* out, const double* coef, const double* x, const double* y, const double* z,
                    void* stream) {
  if (n < 0 || !out || !coef || !x || !y || !z) return set_error(FX_ERR_ARG, "fx_lincomb3_dev: bad argument");
  return la_lincomb3(n, out, coef, x, y, z, (cudaStream_t)stream);
}

int fx_dotw_dev(fx_plan* P, int n, const double* x, const double* y, const double* w, double* out_dev,
                void* stream) {
  if (!P || n < 0 || !x || !y || !w || !out_dev) return set_error(FX_ERR_ARG, "fx_dotw_dev: bad argument");
  return la_dotw(P, n, x, y, w, out_dev, (cudaStream_t)stream);
}

int fx_plan_set_assembly_mode(fx_plan* P, int mode) {
  if (!P || (mode != FX_ASM_ATOMIC && mode != FX_ASM_DETERMINISTIC))
    return set_error(FX_ERR_ARG, "fx_plan_set_assembly_mode: bad argument");
  P->asm_mode = mode;
  return FX_OK;
}

int fx_plan_set_owned_cells(fx_plan* P, int nc_owned) {
  if (!P || nc_owned < 0 || nc_owned > P->nc) return set_error(FX_ERR_ARG, "fx_plan_set_owned_cells: bad argument");
  P->nc_owned = nc_owned;
  return FX_OK;
}

// device row mask of a space = eliminated rows (bit 0) | ghost rows (bit 1) | interface rows (bit 2)
static int rebuild_mask(DevPattern& d) {
  d.nb.release();  // the node-block layout depends on the masks: rebuilt at the next solve
  cudaFree(d.elim);
  d.elim = nullptr;
  if (d.elim_h.empty() && d.halo_h.empty()) return FX_OK;
  std::vector<uint8_t> m((size_t)d.n, 0);
  for (int i = 0; i < d.n; ++i)
    m[i] = (d.elim_h.empty() ? 0 : d.elim_h[i]) | (d.halo_h.empty() ? 0 : d.halo_h[i]);
  FX_CUDA(cudaMalloc((void**)&d.elim, (size_t)d.n));
  FX_CUDA(cudaMemcpy(d.elim, m.data(), (size_t)d.n, cudaMemcpyHostToDevice));
  return FX_OK;
}

int fx_plan_set_eliminated_dofs(fx_plan* P, int space, const int* rows, int n) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_csr || n < 0 || (n > 0 && !rows))
    return set_error(FX_ERR_ARG, "fx_plan_set_eliminated_dofs: bad argument");
  DevPattern& d = P->sp[space];
  cudaFree(d.elim_rows);
  d.elim_rows = nullptr; d.nelim = 0;
  d.elim_h.clear();
  if (n > 0) {
    std::vector<uint8_t> mask((size_t)d.n, 0);
    for (int i = 0; i < n; ++i) {
      if (rows[i] < 0 || rows[i] >= d.n || mask[rows[i]])
        return set_error(FX_ERR_ARG, "fx_plan_set_eliminated_dofs: row %d out of range or repeated", rows[i]);
      mask[rows[i]] = 1;
    }
    d.elim_h.swap(mask);
    FX_CUDA(cudaMalloc((void**)&d.elim_rows, sizeof(int) * (size_t)n));
    FX_CUDA(cudaMemcpy(d.elim_rows, rows, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    d.nelim = n;
  }
  return rebuild_mask(d);
}

long long fx_peer_region_bytes(long long nmax) {
  return nmax < 0 ? 0 : (long long)sizeof(double) * dist_region_doubles(nmax);
}

int fx_plan_set_peer_regions(fx_plan* P, int rank, int world, void* const* regions, long long nmax) {
  if (!P || world < 0 || world > DIST_MAXW || (world > 1 && (rank < 0 || rank >= world || !regions || nmax < 1)))
    return set_error(FX_ERR_ARG, "fx_plan_set_peer_regions: bad argument (world <= %d)", DIST_MAXW);
  P->dist = DistDev();
  if (world <= 1) return FX_OK;
  for (int q = 0; q < world; ++q) {
    if (!regions[q]) return set_error(FX_ERR_ARG, "fx_plan_set_peer_regions: region %d is null", q);
    P->dist.region[q] = static_cast<double*>(regions[q]);
  }
  P->dist.rank = rank; P->dist.world = world; P->dist.nmax = nmax;
  return FX_OK;
}

int fx_plan_set_halo(fx_plan* P, int space, int nsend, const int* peer, const int* local, const int* remote,
                     const unsigned char* ghost) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_csr || nsend < 0 || (nsend > 0 && (!peer || !local || !remote)))
    return set_error(FX_ERR_ARG, "fx_plan_set_halo: bad argument");
  DevPattern& d = P->sp[space];
  cudaFree(d.s_first); cudaFree(d.s_peer); cudaFree(d.s_local); cudaFree(d.s_remote);
  d.s_first = d.s_peer = d.s_local = d.s_remote = nullptr;
  d.nsend = 0; d.has_halo = false;
  d.halo_h.clear();
  d.halo_order.clear(); d.halo_first_h.clear(); d.halo_remote_int.clear(); d.ghost_owner_h.clear();
  if (!ghost) return rebuild_mask(d);  // halo removed
  std::vector<uint8_t> m((size_t)d.n, 0);
  for (int i = 0; i < d.n; ++i) m[i] = ghost[i] ? 2 : 0;
  // sort the (dof, destination) pairs by dof; the last pair of a dof carries DIST_LAST
  std::vector<int> order(nsend);
  for (int k = 0; k < nsend; ++k) {
    order[k] = k;
    if (local[k] < 0 || local[k] >= d.n || ghost[local[k]] || peer[k] < 0 || peer[k] >= DIST_MAXW || remote[k] < 0)
      return set_error(FX_ERR_ARG, "fx_plan_set_halo: pair %d (dof %d -> rank %d slot %d) is invalid", k, local[k], peer[k], remote[k]);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return local[a] < local[b]; });
  std::vector<int> first((size_t)d.n, -1), hp(nsend), hl(nsend), hr(nsend);
  for (int k = 0; k < nsend; ++k) {
    const int o = order[k];
    hl[k] = local[o]; hr[k] = remote[o]; hp[k] = peer[o];
    if (first[hl[k]] < 0) first[hl[k]] = k;
    m[hl[k]] |= 4;
    if (k + 1 == nsend || local[order[k + 1]] != hl[k]) hp[k] |= DIST_LAST;
  }
  FX_CUDA(cudaMalloc((void**)&d.s_first, sizeof(int) * (size_t)d.n));
  FX_CUDA(cudaMemcpy(d.s_first, first.data(), sizeof(int) * (size_t)d.n, cudaMemcpyHostToDevice));
  if (nsend > 0) {
    FX_CUDA(cudaMalloc((void**)&d.s_peer, sizeof(int) * (size_t)nsend));
    FX_CUDA(cudaMalloc((void**)&d.s_local, sizeof(int) * (size_t)nsend));
    FX_CUDA(cudaMalloc((void**)&d.s_remote, sizeof(int) * (size_t)nsend));
    FX_CUDA(cudaMemcpy(d.s_peer, hp.data(), sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
    FX_CUDA(cudaMemcpy(d.s_local, hl.data(), sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
    FX_CUDA(cudaMemcpy(d.s_remote, hr.data(), sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
  }
  d.nsend = nsend; d.has_halo = true;
  d.halo_h.swap(m);
  d.halo_order = order;
  d.halo_first_h = first;
  d.ghost_owner_h.assign(ghost, ghost + d.n);
  return rebuild_mask(d);
}

int fx_space_solver_index(fx_plan* P, int space, int* int_of_ext, int* n_internal) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_csr || !int_of_ext)
    return set_error(FX_ERR_ARG, "fx_space_solver_index: bad argument");
  int rc = nb_ensure(P, space);
  if (rc) return rc;
  DevPattern& d = P->sp[space];
  if (d.nb.state != 1)
    return set_error(FX_ERR_UNSUPPORTED, "space %d has no node-block solver layout (%s)", space, d.nb.why.c_str());
  std::memcpy(int_of_ext, d.nb.int_of_ext_h.data(), sizeof(int) * (size_t)d.n);
  if (n_internal) *n_internal = d.nb.N;
  return FX_OK;
}

int fx_plan_set_halo_solver_index(fx_plan* P, int space, int nsend, const int* remote_internal) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_halo || nsend != P->sp[space].nsend ||
      (nsend > 0 && !remote_internal))
    return set_error(FX_ERR_ARG, "fx_plan_set_halo_solver_index: bad argument (call fx_plan_set_halo first)");
  DevPattern& d = P->sp[space];
  int rc = nb_ensure(P, space);
  if (rc) return rc;
  if (d.nb.state != 1) return set_error(FX_ERR_UNSUPPORTED, "space %d has no node-block solver layout (%s)", space, d.nb.why.c_str());
  // pairs of dofs outside the node blocks (the eliminated DG0 rows of W) are never pushed by the Krylov kernels and
  // carry -1 on both sides; every other pair needs the destination's internal index
  std::vector<int> local_sorted((size_t)nsend);
  FX_CUDA(cudaMemcpy(local_sorted.data(), d.s_local, sizeof(int) * (size_t)nsend, cudaMemcpyDeviceToHost));
  for (int k = 0; k < nsend; ++k) {
    const int li = d.nb.int_of_ext_h[local_sorted[k]], ri = remote_internal[d.halo_order[k]];
    if ((li >= 0) != (ri >= 0))
      return set_error(FX_ERR_ARG, "fx_plan_set_halo_solver_index: pair %d: local internal index %d, remote %d", d.halo_order[k], li, ri);
  }
  d.halo_remote_int.assign(remote_internal, remote_internal + nsend);
  return nb_attach_halo(P, space);
}

int fx_solve_results(fx_plan* P, int first, int count, fx_solve_info* info, void* stream) {
  if (!P || first < 0 || count < 1 || first + count > FX_NTICKETS || !info)
    return set_error(FX_ERR_ARG, "fx_solve_results: bad argument");
  FX_CUDA(cudaMemcpyAsync(info, P->info + first, sizeof(fx_solve_info) * count, cudaMemcpyDeviceToHost,
                          (cudaStream_t)stream));
  FX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return FX_OK;
}

int fx_plan_set_aggregates(fx_plan* P, int space, const int* agg, int k) {
  if (!P || space < 0 || space >= S_N || !P->sp[space].has_csr || k < 0 || (k > 0 && !agg))
    return set_error(FX_ERR_ARG, "fx_plan_set_aggregates: bad argument");
  DevPattern& d = P->sp[space];
  cudaFree(d.agg);
  d.agg = nullptr; d.nagg = 0;
  if (k == 0) return FX_OK;
  if (k % 32 != 0 || k > COARSE_KMAX_CG)
    return set_error(FX_ERR_UNSUPPORTED, "fx_plan_set_aggregates: k must be a multiple of 32 and <= %d", COARSE_KMAX_CG);
  std::vector<int> count(k, 0);
  for (int i = 0; i < d.n; ++i) {
    if (agg[i] < 0 || agg[i] >= k) return set_error(FX_ERR_ARG, "fx_plan_set_aggregates: agg[%d] = %d out of range", i, agg[i]);
    ++count[agg[i]];
  }
  for (int g = 0; g < k; ++g)
    if (count[g] == 0) return set_error(FX_ERR_UNSUPPORTED, "fx_plan_set_aggregates: aggregate %d is empty", g);
  // the resident CG kernel gives CTA c the row blocks c*8R .. c*8R+8R-1 (R per warp); each CTA keeps the rows
  // of (P^T A P)^-1 of the aggregates its rows touch in shared memory
  // (systems too large for the resident kernel use the coarse correction inside k_cg: no per-CTA limit)
  const PcgConfig cfg = pcg_resident_config(P, d.nrowblk, d.maxblknnz);
  const int per_cta = 8 * std::max(cfg.R, 1);
  int nl_max = 0;
  for (int c = 0; c < cfg.grid && cfg.R > 0 && k <= COARSE_KMAX; ++c) {
    std::vector<char> seen(k, 0);
    int nl = 0;
    for (int b = c * per_cta; b < std::min(d.nrowblk, (c + 1) * per_cta); ++b)
      for (int r = d.rowblk_h[b]; r < d.rowblk_h[b + 1]; ++r)
        if (!seen[agg[r]]) { seen[agg[r]] = 1; ++nl; }
    nl_max = std::max(nl_max, nl);
  }
  // nl_max > cfg.nlmax (aggregates not local in the dof numbering, e.g. the O-grid annulus at n = 384): the resident
  // kernel cannot keep the needed rows of E^-1 in shared memory; la_krylov then runs the coarse correction inside k_cg
  d.agg_nl = nl_max;
  if (!P->coarse) {
    FX_CUDA(cudaMalloc((void**)&P->coarse, sizeof(double) * (2 * (size_t)COARSE_KMAX_CG * COARSE_KMAX_CG + 5 * COARSE_KMAX_CG + 16)));
  }
  FX_CUDA(cudaMalloc((void**)&d.agg, sizeof(int) * (size_t)d.n));
  FX_CUDA(cudaMemcpy(d.agg, agg, sizeof(int) * (size_t)d.n, cudaMemcpyHostToDevice));
  d.nagg = k;
  return FX_OK;
}

int fx_solve_results_async(fx_plan* P, int first, int count, fx_solve_info* info, void* stream) {
  if (!P || first < 0 || count < 1 || first + count > FX_NTICKETS || !info)
    return set_error(FX_ERR_ARG, "fx_solve_results_async: bad argument");
  FX_CUDA(cudaMemcpyAsync(info, P->info + first, sizeof(fx_solve_info) * count, cudaMemcpyDeviceToHost,
                          (cudaStream_t)stream));
  return FX_OK;
}

int fx_krylov_solve_async(fx_plan* P, int space, const double* val, double* x, const double* b, int method,
                          int pc, double rtol, double atol, int maxit, int ticket, void* stream) {
  if (!P || space < 0 || space >= S_N || !val || !x || !b || maxit < 0 || ticket < 0 || ticket >= FX_NTICKETS)
    return set_error(FX_ERR_ARG, "fx_krylov_solve_async: bad argument");
  return la_krylov(P, space, val, x, b, method, pc, rtol, atol, maxit, ticket, (cudaStream_t)stream);
}

int fx_krylov_solve(fx_plan* P, int space, const double* val, double* x, const double* b, int method, int pc,
                    double rtol, double atol, int maxit, fx_solve_info* info, void* stream) {
  const int ticket = FX_NTICKETS - 1;
  int rc = fx_krylov_solve_async(P, space, val, x, b, method, pc, rtol, atol, maxit, ticket, stream);
  if (rc) return rc;
  fx_solve_info local;
  rc = fx_solve_results(P, ticket, 1, &local, stream);
  if (rc) return rc;
  if (info) *info = local;
  if (local.converged != 1)
    return set_error(FX_ERR_NOCONV, "Krylov solver did not converge: %d iterations, residual %.3e (ref %.3e), status %d",
                     local.iterations, local.residual, local.residual0, local.converged);
  return FX_OK;
}

}  // extern "C"
