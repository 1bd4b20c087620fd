// CUDA kernels for finite-element assembly (K1-K5 of SURVEY.md 2.3) and their launchers.
// The element arithmetic lives in fx_assemble.h / fx_forms_*.h; this file is the thread mapping.
#include <vector>

#include "fx_plan.h"
#include "fx_registry.h"

namespace fx {

// ---- K1 geometry ---------------------------------------------------------------------------------
__global__ void k_geometry(const double* __restrict__ xy, const int* __restrict__ cells, int nc,
                           double* __restrict__ geom) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) compute_geo(xy, cells, c, nc, geom);
}
__global__ void k_facet_geometry(const double* __restrict__ xy, const int* __restrict__ cells,
                                 const int* __restrict__ bf_cell, const int* __restrict__ bf_local, int nbf,
                                 double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nbf) compute_facet_geo(xy, cells, bf_cell[f], bf_local[f], f, nbf, out);
}

int launch_geometry(fx_plan* P, cudaStream_t st) {
  k_geometry<<<(P->nc + 255) / 256, 256, 0, st>>>(P->xy, P->cells, P->nc, P->geom);
  if (P->nbf > 0)
    k_facet_geometry<<<(P->nbf + 127) / 128, 128, 0, st>>>(P->xy, P->cells, P->bf_cell, P->bf_local, P->nbf,
                                                           P->bf_geo);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// ---- K2/K3 cell kernels: grid = cell blocks of 32, block = (32 cells, 6 | 3 row warps) --------------
// Phase 1 (pointwise data of 32 cells x NQ points -> shared memory) runs ONCE per cell block; the LD rows
// of the element tensor are then dealt to the warps as "row tasks": warp y takes rows y, y+ny, ... so
// every warp works on one (test component, basis function) at a time (warp-uniform dispatch).
template <class F, bool MATRIX>
__device__ __forceinline__ void cell_rows_dispatch(const AsmArgs& a, int cell0, const double* smem) {
  using Sp = typename F::Space;
  for (int t0 = 0; t0 < Sp::LD; t0 += blockDim.y) {
    const int task = t0 + threadIdx.y;  // local row index, uniform within the warp
    static_for<0, Sp::NCOMP>([&](auto ic) {
      constexpr int I = decltype(ic)::value;
      if (task >= Sp::off(I) && task < Sp::off(I) + Sp::nb(I)) {
        if constexpr (MATRIX) mat_rows<F, I>(a, cell0, threadIdx.x, task - Sp::off(I), smem);
        else vec_rows<F, I>(a, cell0, threadIdx.x, task - Sp::off(I), smem);
      }
    });
  }
}

template <class F>
__global__ void __launch_bounds__(CB * 6, 4) k_cell_matrix(const __grid_constant__ AsmArgs a) {
  extern __shared__ double smem[];
  const int cell0 = blockIdx.x * CB;
  if constexpr (F::NP > 0) {
    cell_phase1<F>(a, cell0, threadIdx.y * CB + threadIdx.x, blockDim.x * blockDim.y, smem);
    __syncthreads();
  }
  cell_rows_dispatch<F, true>(a, cell0, smem);
}

template <class F>
__global__ void __launch_bounds__(CB * 6) k_cell_vector(const __grid_constant__ AsmArgs a) {
  extern __shared__ double smem[];
  const int cell0 = blockIdx.x * CB;
  if constexpr (F::NP > 0) {
    cell_phase1<F>(a, cell0, threadIdx.y * CB + threadIdx.x, blockDim.x * blockDim.y, smem);
    __syncthreads();
  }
  cell_rows_dispatch<F, false>(a, cell0, smem);
}

// ---- K4 exterior-facet kernels: one thread per (boundary facet, local test dof) -------------------
template <class F>
__global__ void k_facet_matrix(const __grid_constant__ AsmArgs a) {
  constexpr int LD = F::Space::LD;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.m.nbf * LD) facet_mat_row<F>(a, idx / LD, idx % LD);
}
template <class F>
__global__ void k_facet_vector(const __grid_constant__ AsmArgs a) {
  constexpr int LD = F::Space::LD;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.m.nbf * LD) facet_vec_row<F>(a, idx / LD, idx % LD);
}

// ---- K5 functionals: per-cell value, deterministic two-stage reduction ----------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

template <class F>
__global__ void __launch_bounds__(256) k_cell_scalar(const __grid_constant__ AsmArgs a) {
  __shared__ double sh[32];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double v = (c < a.nc_owned) ? cell_functional<F>(a, c) : 0.0;
  v = block_sum(v, sh);
  if (threadIdx.x == 0) a.part[blockIdx.x] = v;
}
__global__ void __launch_bounds__(1024) k_sum_partials(const double* __restrict__ part, int n,
                                                       double* __restrict__ out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = block_sum(v, sh);
  if (threadIdx.x == 0) out[0] = v;
}

template <class F>
__global__ void __launch_bounds__(256) k_facet_scalar(const __grid_constant__ AsmArgs a) {
  __shared__ double sh[32];
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  double v = (f < a.m.nbf) ? facet_functional<F>(a, f) : 0.0;
  v = block_sum(v, sh);
  if (threadIdx.x == 0) a.part[blockIdx.x] = v;
}

template <class E>
__global__ void __launch_bounds__(256) k_project_dg0(const __grid_constant__ AsmArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.m.nc) cell_project_dg0<E>(a, c);
}

// ---- deterministic (atomic-free) assembly: gather of the element-tensor buffer ---------------------------------
// one thread per CSR entry / dof; contributions listed in ascending cell order => bitwise reproducible sums
__global__ void __launch_bounds__(256) k_gather_matrix(const int* __restrict__ cptr, const int* __restrict__ contrib,
                                                       const double* __restrict__ elt, double* __restrict__ val, int nnz) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  double s = 0.0;
  for (int k = cptr[j]; k < cptr[j + 1]; ++k) s += elt[contrib[k]];
  val[j] = s;
}
__global__ void __launch_bounds__(256) k_gather_vector(const int* __restrict__ aptr, const int* __restrict__ adj,
                                                       const double* __restrict__ elt, double* __restrict__ vec, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = aptr[i]; k < aptr[i + 1]; ++k) s += elt[adj[k]];
  vec[i] = s;
}

// contribution lists of a space (once per plan) + the element-tensor buffer
static int det_ensure(fx_plan* P, int space, bool matrix) {
  DevPattern& d = P->sp[space];
  const size_t need = (size_t)P->nc * d.ld * (matrix ? d.ld : 1);
  if (P->elt_cap < need) {
    cudaFree(P->elt);
    P->elt = nullptr; P->elt_cap = 0;
    FX_CUDA(cudaMalloc((void**)&P->elt, sizeof(double) * need));
    P->elt_cap = need;
  }
  if (d.det_ready) return FX_OK;
  const int nc = P->nc, ld = d.ld;
  std::vector<int> dm((size_t)ld * nc);
  FX_CUDA(cudaMemcpy(dm.data(), d.dm_t, sizeof(int) * dm.size(), cudaMemcpyDeviceToHost));
  {  // dof -> (cell, local row) pairs, ascending cell
    std::vector<int> aptr((size_t)d.n + 1, 0), adj((size_t)nc * ld);
    for (size_t i = 0; i < dm.size(); ++i) ++aptr[dm[i] + 1];
    for (int i = 0; i < d.n; ++i) aptr[i + 1] += aptr[i];
    std::vector<int> fill(aptr.begin(), aptr.end() - 1);
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < ld; ++a) adj[fill[dm[(size_t)a * nc + c]]++] = c * ld + a;
    FX_CUDA(cudaMalloc((void**)&d.det_aptr, sizeof(int) * aptr.size()));
    FX_CUDA(cudaMalloc((void**)&d.det_adj, sizeof(int) * adj.size()));
    FX_CUDA(cudaMemcpy(d.det_aptr, aptr.data(), sizeof(int) * aptr.size(), cudaMemcpyHostToDevice));
    FX_CUDA(cudaMemcpy(d.det_adj, adj.data(), sizeof(int) * adj.size(), cudaMemcpyHostToDevice));
  }
  if (d.has_csr) {  // CSR entry -> positions in the [cell][row][col] buffer, ascending cell
    if ((size_t)nc * ld * ld > 2147483647u) return set_error(FX_ERR_UNSUPPORTED, "deterministic assembly: mesh too large for int32 indices");
    std::vector<int> rowptr((size_t)d.n + 1);
    std::vector<uint8_t> pos((size_t)nc * ld * ld);
    FX_CUDA(cudaMemcpy(rowptr.data(), d.rowptr, sizeof(int) * rowptr.size(), cudaMemcpyDeviceToHost));
    FX_CUDA(cudaMemcpy(pos.data(), d.pos, pos.size(), cudaMemcpyDeviceToHost));
    std::vector<int> cptr((size_t)d.nnz + 1, 0), contrib((size_t)nc * ld * ld);
    auto entry = [&](int c, int a, int b) { return rowptr[dm[(size_t)a * nc + c]] + pos[((size_t)a * ld + b) * nc + c]; };
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < ld; ++a)
        for (int b = 0; b < ld; ++b) ++cptr[entry(c, a, b) + 1];
    for (int j = 0; j < d.nnz; ++j) cptr[j + 1] += cptr[j];
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < ld; ++a)
        for (int b = 0; b < ld; ++b) contrib[fill[entry(c, a, b)]++] = (c * ld + a) * ld + b;
    FX_CUDA(cudaMalloc((void**)&d.det_cptr, sizeof(int) * cptr.size()));
    FX_CUDA(cudaMalloc((void**)&d.det_contrib, sizeof(int) * contrib.size()));
    FX_CUDA(cudaMemcpy(d.det_cptr, cptr.data(), sizeof(int) * cptr.size(), cudaMemcpyHostToDevice));
    FX_CUDA(cudaMemcpy(d.det_contrib, contrib.data(), sizeof(int) * contrib.size(), cudaMemcpyHostToDevice));
  }
  d.det_ready = true;
  return FX_OK;
}

// ---- launchers ------------------------------------------------------------------------------------
template <class F>
static int run_matrix(fx_plan* P, AsmArgs a, cudaStream_t st) {
  using Sp = typename F::Space;
  DevPattern& pat = P->sp[space_id<Sp>::v];
  if (!pat.has_csr) return set_error(FX_ERR_ARG, "space %d has no dofmap/pattern in this plan", space_id<Sp>::v);
  a.rowptr = pat.rowptr;
  a.pos = pat.pos;
  a.elt = nullptr;
  const bool det = P->asm_mode == FX_ASM_DETERMINISTIC;
  if (det) {
    if (int rc = det_ensure(P, space_id<Sp>::v, true)) return rc;
    a.elt = P->elt;
  } else {
    FX_CUDA(cudaMemsetAsync(a.val, 0, sizeof(double) * (size_t)pat.nnz, st));
  }
  constexpr int NQ = tri_npts(F::QDEG);
  constexpr size_t smem = sizeof(double) * (size_t)F::NP * NQ * CB;
  static bool attr_done = false;
  if (smem > 48 * 1024 && !attr_done) {
    FX_CUDA(cudaFuncSetAttribute(k_cell_matrix<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((P->nc + CB - 1) / CB), block(CB, Sp::maxnb());
  k_cell_matrix<F><<<grid, block, smem, st>>>(a);
  count_launch(1);
  if constexpr (F::Facet::present) {
    if (P->nbf > 0) {
      const int n = P->nbf * Sp::LD;
      k_facet_matrix<F><<<(n + 127) / 128, 128, 0, st>>>(a);
      count_launch(1);
    }
  }
  if (det) {
    k_gather_matrix<<<(pat.nnz + 255) / 256, 256, 0, st>>>(pat.det_cptr, pat.det_contrib, P->elt, a.val, pat.nnz);
    count_launch(1);
  }
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

template <class F>
static int run_vector(fx_plan* P, AsmArgs a, cudaStream_t st) {
  using Sp = typename F::Space;
  DevPattern& pat = P->sp[space_id<Sp>::v];
  if (!pat.present) return set_error(FX_ERR_ARG, "space %d has no dofmap in this plan", space_id<Sp>::v);
  a.elt = nullptr;
  const bool det = P->asm_mode == FX_ASM_DETERMINISTIC;
  double* out = a.vec;
  if (det) {
    if (int rc = det_ensure(P, space_id<Sp>::v, false)) return rc;
    a.elt = P->elt;
  } else {
    FX_CUDA(cudaMemsetAsync(a.vec, 0, sizeof(double) * (size_t)pat.n, st));
  }
  constexpr int NQ = tri_npts(F::QDEG);
  constexpr size_t smem = sizeof(double) * (size_t)F::NP * NQ * CB;
  static bool attr_done = false;
  if (smem > 48 * 1024 && !attr_done) {
    FX_CUDA(cudaFuncSetAttribute(k_cell_vector<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((P->nc + CB - 1) / CB), block(CB, Sp::maxnb());
  k_cell_vector<F><<<grid, block, smem, st>>>(a);
  count_launch(1);
  if constexpr (F::Facet::present) {
    if (P->nbf > 0) {
      const int n = P->nbf * Sp::LD;
      k_facet_vector<F><<<(n + 127) / 128, 128, 0, st>>>(a);
      count_launch(1);
    }
  }
  if (det) {
    k_gather_vector<<<(pat.n + 255) / 256, 256, 0, st>>>(pat.det_aptr, pat.det_adj, P->elt, out, pat.n);
    count_launch(1);
  }
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

template <class F>
static int run_scalar(fx_plan* P, AsmArgs a, double* out_dev, cudaStream_t st) {
  const int nb = (P->nc + 255) / 256;
  if (nb > P->part_cap) return set_error(FX_ERR_ARG, "partial buffer too small");
  a.part = P->part;
  k_cell_scalar<F><<<nb, 256, 0, st>>>(a);
  k_sum_partials<<<1, 1024, 0, st>>>(P->part, nb, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

template <class F>
static int run_facet_scalar(fx_plan* P, AsmArgs a, double* out_dev, cudaStream_t st) {
  const int nb = max(1, (P->nbf + 255) / 256);
  if (nb > P->part_cap) return set_error(FX_ERR_ARG, "partial buffer too small");
  a.part = P->part;
  k_facet_scalar<F><<<nb, 256, 0, st>>>(a);
  k_sum_partials<<<1, 1024, 0, st>>>(P->part, nb, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// the config-4 functors that carry phi_def(u, lambda_D) exist in two instantiations: FX_FORM_C4_* fold phi_def to 1
// (lambda_D = 0, as UFL does) and reject anything else loudly; FX_FORM_C4M_* keep it
static int check_form_params(int form, const AsmArgs& a) {
  const bool phi_form = form == FX_FORM_C4_A4 || form == FX_FORM_C4_L2 || form == FX_FORM_C4_L9 ||
                        form == FX_FORM_C4_L_RES_ORTH || form == FX_FORM_C4_TORQUE || form == FX_FORM_C4_FORCE_X ||
                        form == FX_FORM_C4_FORCE_Y;
  if (phi_form && a.p.v[jbp::P_LAMBDA_D] != 0.0)
    return set_error(FX_ERR_UNSUPPORTED, "form %d folds phi_def to 1 (lambda_D = 0); lambda_D = %g needs its "
                     "FX_FORM_C4M_* counterpart", form, a.p.v[jbp::P_LAMBDA_D]);
  if (form >= FX_FORM_C3_A0 && form <= FX_FORM_C3_FORCE_Y && a.p.v[ldc::P_CONV] == 0.0)
    return set_error(FX_ERR_UNSUPPORTED, "config-3 forms are implemented for conv != 0 only");
  // the FX_FORM_C3_* functors below fold phi_s == phi_ewm == 1 (the script's A_0 = k_ewm = B = 0); the true EWM
  // model goes through their FX_FORM_C3E_* counterparts
  const bool folded = form == FX_FORM_C3_A1 || form == FX_FORM_C3_L1 || form == FX_FORM_C3_L2 || form == FX_FORM_C3_A4 ||
                      form == FX_FORM_C3_L4 || form == FX_FORM_C3_L9;
  if (folded && (a.p.v[jbp::P_A0] != 0.0 || a.p.v[jbp::P_K_EWM] != 0.0 || a.p.v[jbp::P_B_EWM] != 0.0))
    return set_error(FX_ERR_UNSUPPORTED, "form %d assumes A_0 = k_ewm = B = 0 (got %g, %g, %g): use the FX_FORM_C3E_* forms",
                     form, a.p.v[jbp::P_A0], a.p.v[jbp::P_K_EWM], a.p.v[jbp::P_B_EWM]);
  return FX_OK;
}

template <class E>
static int run_dg0(fx_plan* P, AsmArgs a, cudaStream_t st) {
  k_project_dg0<E><<<(P->nc + 255) / 256, 256, 0, st>>>(a);
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int launch_matrix(fx_plan* P, int form, const AsmArgs& a, cudaStream_t st) {
  if (int rc = check_form_params(form, a)) return rc;
  switch (form) {
#define X(id, F) case id: return run_matrix<F>(P, a, st);
    FX_MATRIX_FORMS(X)
#undef X
  }
  return set_error(FX_ERR_ARG, "unknown bilinear form id %d", form);
}

int launch_vector(fx_plan* P, int form, const AsmArgs& a, cudaStream_t st) {
  const bool conv0 = a.p.v[ldc::P_CONV] == 0.0;
  if (int rc = check_form_params(form, a)) return rc;
  switch (form) {
#define X(id, F) case id: return run_vector<F>(P, a, st);
#define XC(id, F0, F1) case id: return conv0 ? run_vector<F0>(P, a, st) : run_vector<F1>(P, a, st);
    FX_VECTOR_FORMS(X, XC)
#undef X
#undef XC
  }
  return set_error(FX_ERR_ARG, "unknown linear form id %d", form);
}

int launch_scalar(fx_plan* P, int form, const AsmArgs& a, double* out_dev, cudaStream_t st) {
  if (int rc = check_form_params(form, a)) return rc;
  switch (form) {
#define X(id, F) case id: return run_scalar<F>(P, a, out_dev, st);
    FX_SCALAR_FORMS(X)
#undef X
#define X(id, F) case id: return run_facet_scalar<F>(P, a, out_dev, st);
    FX_FACET_SCALAR_FORMS(X)
#undef X
  }
  return set_error(FX_ERR_ARG, "unknown functional id %d", form);
}

int launch_dg0(fx_plan* P, int expr, const AsmArgs& a, cudaStream_t st) {
  if (expr == FX_EXPR_C4_RESTAU && a.p.v[jbp::P_LAMBDA_D] != 0.0)
    return set_error(FX_ERR_UNSUPPORTED, "FX_EXPR_C4_RESTAU folds phi_def to 1 (lambda_D = 0); lambda_D = %g needs "
                     "FX_EXPR_C4M_RESTAU", a.p.v[jbp::P_LAMBDA_D]);
  switch (expr) {
#define X(id, E) case id: return run_dg0<E>(P, a, st);
    FX_DG0_EXPRS(X)
#undef X
  }
  return set_error(FX_ERR_ARG, "unknown DG0 expression id %d", expr);
}

int form_info(int form, int* rank, int* space, int* qdeg) {
  switch (form) {
#define X(id, F) case id: *rank = 2; *space = space_id<typename F::Space>::v; qdeg[0] = F::QDEG; qdeg[1] = facet_qdeg<F>(); return FX_OK;
    FX_MATRIX_FORMS(X)
#undef X
#define X(id, F) case id: *rank = 1; *space = space_id<typename F::Space>::v; qdeg[0] = F::QDEG; qdeg[1] = facet_qdeg<F>(); return FX_OK;
#define XC(id, F0, F1) case id: *rank = 1; *space = space_id<typename F0::Space>::v; qdeg[0] = F1::QDEG; qdeg[1] = -1; return FX_OK;
    FX_VECTOR_FORMS(X, XC)
#undef X
#undef XC
#define X(id, F) case id: *rank = 0; *space = -1; qdeg[0] = F::QDEG; qdeg[1] = -1; return FX_OK;
    FX_SCALAR_FORMS(X)
#undef X
#define X(id, F) case id: *rank = 0; *space = -1; qdeg[0] = -1; qdeg[1] = F::QDEG; return FX_OK;
    FX_FACET_SCALAR_FORMS(X)
#undef X
  }
  return set_error(FX_ERR_ARG, "unknown form id %d", form);
}

}  // namespace fx
