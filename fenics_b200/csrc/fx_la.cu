// Sparse/dense vector kernels on the DOLFIN-pattern CSR matrices (K6-K8 of SURVEY.md 2.3):
// stand-alone CSR SpMV (A*x), DirichletBC row replacement, axpy / dot / norms.
// The Krylov solvers (fx_krylov.cu) fuse the same building blocks into persistent kernels.
#include "fx_la_common.cuh"

namespace fx {

// ---- SpMV -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LA_THREADS, 4) k_spmv(const __grid_constant__ CsrView A,
                                                        const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double prod[SPMV_SMEM_DOUBLES];
  spmv_stream(A, round_robin_range(A), x, prod, [&](int row, double s) { y[row] = s; });
}

// CSR-vector variant kept for A/B measurements (tools/spmv_bench.py): TPR lanes per row, RU rows in
// flight per sub-warp, no shared memory.
template <int TPR, int RU>
__global__ void __launch_bounds__(LA_THREADS, 4) k_spmv_vec(const __grid_constant__ CsrView A,
                                                            const double* __restrict__ x, double* __restrict__ y) {
  constexpr int SPW = 32 / TPR;  // sub-warps per warp
  const int lane = threadIdx.x % TPR, subw = (threadIdx.x & 31) / TPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int base = warp * SPW * RU; base < A.n; base += nwarp * SPW * RU) {  // warp-uniform trip count
    const int r0 = base + subw * RU;
    int b[RU], e[RU];
    double acc[RU];
#pragma unroll
    for (int i = 0; i < RU; ++i) {
      const int row = r0 + i;
      b[i] = row < A.n ? A.rowptr[row] + lane : 0;
      e[i] = row < A.n ? A.rowptr[row + 1] : 0;
      acc[i] = 0.0;
    }
    bool any = true;
    while (any) {
      any = false;
      int c[RU];
      double v[RU];
#pragma unroll
      for (int i = 0; i < RU; ++i) {
        c[i] = -1;
        if (b[i] < e[i]) { c[i] = ld_stream(A.col + b[i]); v[i] = ld_stream(A.val + b[i]); }
      }
#pragma unroll
      for (int i = 0; i < RU; ++i)
        if (c[i] >= 0) {
          acc[i] += v[i] * x[c[i]];
          b[i] += TPR;
          any = any || (b[i] < e[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < RU; ++i) {
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o, TPR);
      if (lane == 0 && r0 + i < A.n) y[r0 + i] = acc[i];
    }
  }
}

int la_spmv_variant(fx_plan* P, int space, const double* val, const double* x, double* y, int variant,
                    cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  CsrView A{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  const int nb = P->num_sms * 8;
  switch (variant) {
    case 0: return la_spmv(P, space, val, x, y, st);
    case 1: k_spmv_vec<4, 4><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 2: k_spmv_vec<8, 4><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 3: k_spmv_vec<8, 2><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 4: k_spmv_vec<16, 2><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 5: k_spmv_vec<4, 8><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    default: return set_error(FX_ERR_ARG, "unknown SpMV variant %d", variant);
  }
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int la_spmv(fx_plan* P, int space, const double* val, const double* x, double* y, cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  CsrView A{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  const int nb = max(1, min((p.nrowblk + LA_WARPS - 1) / LA_WARPS, P->num_sms * 8));
  k_spmv<<<nb, LA_THREADS, 0, st>>>(A, x, y);
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// ---- DirichletBC.apply ---------------------------------------------------------------------------------
// one warp per constrained row: zero the row, put 1.0 on the diagonal (pattern kept), b[row] = g.
__global__ void k_apply_bc(const int* __restrict__ rowptr, const int* __restrict__ col, double* __restrict__ val,
                           double* __restrict__ b, const int* __restrict__ rows, const double* __restrict__ g,
                           int n) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int r = rows[w];
  if (val != nullptr)
    for (int j = rowptr[r] + lane; j < rowptr[r + 1]; j += 32) val[j] = (col[j] == r) ? 1.0 : 0.0;
  if (b != nullptr && lane == 0) b[r] = g[w];
}

int la_apply_bc(fx_plan* P, int space, double* val, double* b, const int* rows, const double* g, int n,
                cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  if (n <= 0) return FX_OK;
  k_apply_bc<<<(n * 32 + 255) / 256, 256, 0, st>>>(p.rowptr, p.col, val, b, rows, g, n);
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// ---- BLAS-1 --------------------------------------------------------------------------------------------
__global__ void k_axpy(int n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += alpha * x[i];
}
int la_axpy(int n, double alpha, const double* x, double* y, cudaStream_t st) {
  if (n <= 0) return FX_OK;
  const int nb = min((n + 255) / 256, 148 * 8);
  k_axpy<<<nb, 256, 0, st>>>(n, alpha, x, y);
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// out = c0 x + c1 y + c2 z, coefficients in device memory (partitioned Krylov: no host round trip)
__global__ void k_lincomb3(int n, double* out, const double* __restrict__ coef, const double* x, const double* y,
                           const double* z) {
  const double c0 = coef[0], c1 = coef[1], c2 = coef[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = c0 * x[i] + c1 * y[i] + c2 * z[i];
}
int la_lincomb3(int n, double* out, const double* coef, const double* x, const double* y, const double* z,
                cudaStream_t st) {
  if (n <= 0) return FX_OK;
  k_lincomb3<<<min((n + 255) / 256, 148 * 8), 256, 0, st>>>(n, out, coef, x, y, z);
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}
__global__ void __launch_bounds__(256) k_dotw_partial(int n, const double* __restrict__ x,
                                                      const double* __restrict__ y, const double* __restrict__ w,
                                                      double* __restrict__ part) {
  __shared__ double sh[33];
  double v = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += w[i] * x[i] * y[i];
  v = block_sum_bcast(v, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// mode 0: dot(x,y) partials, 1: max|x| partials
__global__ void __launch_bounds__(256) k_dot_partial(int n, const double* __restrict__ x,
                                                     const double* __restrict__ y, int mode,
                                                     double* __restrict__ part) {
  __shared__ double sh[33];
  double v = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    v = mode == 0 ? v + x[i] * y[i] : fmax(v, fabs(x[i]));
  if (mode == 0) {
    v = block_sum_bcast(v, sh);
  } else {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
      double t = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
      t = warp_max(t);
      if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    v = sh[32];
  }
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}
// mode 0: sum, 1: max, 2: sqrt(sum)
__global__ void __launch_bounds__(256) k_finish(const double* __restrict__ part, int n, int mode,
                                                double* __restrict__ out) {
  __shared__ double sh[33];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v = (mode == 1) ? fmax(v, part[i]) : v + part[i];
  if (mode == 1) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, sh[i]);
      out[0] = t;
    }
  } else {
    v = block_sum_bcast(v, sh);
    if (threadIdx.x == 0) out[0] = (mode == 2) ? sqrt(v) : v;
  }
}

int la_dot(fx_plan* P, int n, const double* x, const double* y, double* out_dev, cudaStream_t st) {
  const int nb = max(1, min((n + 255) / 256, min(P->part_cap, 148 * 4)));
  k_dot_partial<<<nb, 256, 0, st>>>(n, x, y, 0, P->part);
  k_finish<<<1, 256, 0, st>>>(P->part, nb, 0, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}
int la_dotw(fx_plan* P, int n, const double* x, const double* y, const double* w, double* out_dev, cudaStream_t st) {
  const int nb = max(1, min((n + 255) / 256, min(P->part_cap, 148 * 4)));
  k_dotw_partial<<<nb, 256, 0, st>>>(n, x, y, w, P->part);
  k_finish<<<1, 256, 0, st>>>(P->part, nb, 0, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}
int la_norm(fx_plan* P, int n, const double* x, int kind, double* out_dev, cudaStream_t st) {
  const int nb = max(1, min((n + 255) / 256, min(P->part_cap, 148 * 4)));
  k_dot_partial<<<nb, 256, 0, st>>>(n, x, x, kind == 1 ? 1 : 0, P->part);
  k_finish<<<1, 256, 0, st>>>(P->part, nb, kind == 1 ? 1 : 2, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

}  // namespace fx
