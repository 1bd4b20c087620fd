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

// Software-pipelined CSR-stream variant (experiment): the value / column stream of row block k+1 is copied
// global -> shared with cp.async (LDGSTS: no registers, no L1 allocation) while the warp gathers and reduces block k,
// so a block's dependent chain shrinks from [stream -> gather -> reduce] to [gather -> reduce] and the bytes in flight
// are no longer bounded by registers.  Two staging buffers per warp; the products overwrite the values in place.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
constexpr int PIPE_NB = SPMV_NNZB + 8;  // staging capacity: a block plus the 16-byte alignment slack on both ends
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_spmv_pipe(const __grid_constant__ CsrView A,
                                                             const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* vb[2];
  int* cb[2];
  {
    unsigned char* base = dyn + (size_t)wib * 2 * PIPE_NB * 12;
    vb[0] = reinterpret_cast<double*>(base);
    vb[1] = vb[0] + PIPE_NB;
    cb[0] = reinterpret_cast<int*>(vb[1] + PIPE_NB);
    cb[1] = cb[0] + PIPE_NB;
  }
  const int nw = gridDim.x * WARPS, gw = blockIdx.x * WARPS + wib;
  const int per = (A.nrowblk + nw - 1) / nw;
  const int k0 = min(gw * per, A.nrowblk), k1 = min(k0 + per, A.nrowblk);
  if (k0 >= k1) return;
  auto bounds = [&](int k, int& r0, int& r1, int& j0, int& j1) {
    r0 = A.rowblk[k]; r1 = A.rowblk[k + 1];
    j0 = A.rowptr[r0]; j1 = A.rowptr[r1];
  };
  const int nnz_total = A.rowptr[A.n];
  auto issue = [&](int j0, int j1, int b) {
    // 16-byte chunks from the aligned start; the last chunk of the whole matrix must not read past the arrays, so
    // its (< 4) tail entries are copied with plain loads
    const int s0 = j0 & ~3, m4 = min((j1 - s0 + 3) >> 2, (nnz_total - s0) >> 2);
    for (int c = lane; c < m4; c += 32) cp_async16(cb[b] + 4 * c, A.col + s0 + 4 * c);
    for (int c = lane; c < 2 * m4; c += 32) cp_async16(vb[b] + 2 * c, A.val + s0 + 2 * c);
    for (int j = s0 + 4 * m4 + lane; j < j1; j += 32) { cb[b][j - s0] = A.col[j]; vb[b][j - s0] = A.val[j]; }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int r0, r1, j0, j1, nr0 = 0, nr1 = 0, nj0 = 0, nj1 = 0;
  bounds(k0, r0, r1, j0, j1);
  issue(j0, j1, 0);
  if (k0 + 1 < k1) bounds(k0 + 1, nr0, nr1, nj0, nj1);
  for (int k = k0; k < k1; ++k) {
    const int b = (k - k0) & 1;
    int fr0 = 0, fr1 = 0, fj0 = 0, fj1 = 0;
    if (k + 1 < k1) {
      issue(nj0, nj1, b ^ 1);
      if (k + 2 < k1) bounds(k + 2, fr0, fr1, fj0, fj1);  // two blocks ahead: in flight while this block is processed
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");  // empty group keeps the wait count uniform
    }
    const int row = r0 + lane;
    const bool act = row < r1;
    const int e1 = act ? A.rowptr[row + 1] : j0;
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const int s0 = j0 & ~3, lo = j0 - s0, hi = j1 - s0;
    double* pv = vb[b];
    const int* pc = cb[b];
    for (int jb = lo; jb < hi; jb += 256) {
      double xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = jb + u * 32 + lane;
        xv[u] = (j < hi) ? x[pc[j]] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = jb + u * 32 + lane;
        if (j < hi) pv[j] *= xv[u];
      }
    }
    __syncwarp();
    const int t1 = e1 - s0;
    int t0 = __shfl_up_sync(0xffffffffu, t1, 1);
    if (lane == 0) t0 = lo;
    if (act) {
      double sum = 0.0;
      for (int j = t0; j < t1; ++j) sum += pv[j];
      y[row] = sum;
    }
    __syncwarp();
    r0 = nr0; r1 = nr1; j0 = nj0; j1 = nj1;
    nr0 = fr0; nr1 = fr1; nj0 = fj0; nj1 = fj1;
  }
}

int la_spmv_variant(fx_plan* P, int space, const double* val, const double* x, double* y, int variant,
                    cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  if ((variant & 0xff) == 8)  // node-block layout; bits 8.. = number of products (default 1)
    return la_spmv_nb(P, space, val, x, y, std::max(1, variant >> 8), st);
  CsrView A{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr};
  const int nb = P->num_sms * 8;
  switch (variant) {
    case 0: return la_spmv(P, space, val, x, y, st);
    case 1: k_spmv_vec<4, 4><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 2: k_spmv_vec<8, 4><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 3: k_spmv_vec<8, 2><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 4: k_spmv_vec<16, 2><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 5: k_spmv_vec<4, 8><<<nb, LA_THREADS, 0, st>>>(A, x, y); break;
    case 6: case 7: {  // software-pipelined stream, 16 warps (6) or 8 warps x 2 CTAs (7) per SM
      constexpr size_t per_warp = (size_t)2 * PIPE_NB * 12;
      static bool attr = false;
      if (!attr) {
        FX_CUDA(cudaFuncSetAttribute(k_spmv_pipe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * per_warp)));
        FX_CUDA(cudaFuncSetAttribute(k_spmv_pipe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * per_warp)));
        attr = true;
      }
      if (variant == 6) k_spmv_pipe<16><<<P->num_sms, 512, 16 * per_warp, st>>>(A, x, y);
      else k_spmv_pipe<8><<<2 * P->num_sms, 256, 8 * per_warp, st>>>(A, x, y);
      break;
    }
    default: return set_error(FX_ERR_ARG, "unknown SpMV variant %d", variant);
  }
  count_launch(1);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

int la_spmv(fx_plan* P, int space, const double* val, const double* x, double* y, cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  CsrView A{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr};
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
