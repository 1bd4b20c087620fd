// Sparse linear algebra on the DOLFIN-pattern CSR matrices (K6-K10 of SURVEY.md 2.3):
//   * CSR SpMV, sub-warp (TPR lanes) per row, coalesced val/col loads, warp-shuffle row reduction
//   * DirichletBC row replacement
//   * BiCGStab / CG as ONE persistent cooperative kernel per solve: SpMV, fused axpy/dot passes and
//     the convergence test all run on the device, separated by grid.sync(); reductions are
//     deterministic (fixed-order two-stage sums, no atomics) so every CTA takes the same branch.
// Replaces PETSc KSPBCGS/KSPCG + PC behind dolfin solve(A,x,b,"bicgstab","default")
// (lid-driven-cavity/incomp_LDC.py:331 and every other solve site).
#include <cooperative_groups.h>

#include "fx_plan.h"

namespace cg = cooperative_groups;

namespace fx {

// ---- SpMV -------------------------------------------------------------------------------------------
template <int TPR>
__device__ __forceinline__ double row_dot(const int* __restrict__ rowptr, const int* __restrict__ col,
                                          const double* __restrict__ val, const double* x,
                                          int row, int n, int lane) {  // x may be written by this kernel: no restrict
  double s = 0.0;
  if (row < n) {
    const int e = rowptr[row + 1];
    for (int j = rowptr[row] + lane; j < e; j += TPR) s += val[j] * x[col[j]];
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, TPR);
  return s;
}

template <int TPR>
__global__ void __launch_bounds__(256) k_spmv(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                              const double* __restrict__ val, const double* __restrict__ x,
                                              double* __restrict__ y) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gt / TPR, lane = gt % TPR;
  const double s = row_dot<TPR>(rowptr, col, val, x, row, n, lane);
  if (lane == 0 && row < n) y[row] = s;
}

static int pick_tpr(const DevPattern& p) {
  const double avg = p.n > 0 ? (double)p.nnz / p.n : 0.0;
  return avg <= 8 ? 4 : (avg <= 20 ? 8 : (avg <= 48 ? 16 : 32));
}

int la_spmv(fx_plan* P, int space, const double* val, const double* x, double* y, cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  const int tpr = pick_tpr(p);
  const long nt = (long)p.n * tpr;
  const int nb = (int)((nt + 255) / 256);
  switch (tpr) {
    case 4: k_spmv<4><<<nb, 256, 0, st>>>(p.n, p.rowptr, p.col, val, x, y); break;
    case 8: k_spmv<8><<<nb, 256, 0, st>>>(p.n, p.rowptr, p.col, val, x, y); break;
    case 16: k_spmv<16><<<nb, 256, 0, st>>>(p.n, p.rowptr, p.col, val, x, y); break;
    default: k_spmv<32><<<nb, 256, 0, st>>>(p.n, p.rowptr, p.col, val, x, y); break;
  }
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result broadcast to every thread; sh must hold 33 doubles
__device__ __forceinline__ double block_sum_bcast(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < nw) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
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
__global__ void __launch_bounds__(256) k_finish(const double* __restrict__ part, int n, int mode,
                                                double* __restrict__ out) {
  __shared__ double sh[33];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v = (mode == 1) ? fmax(v, part[i]) : v + part[i];
  if (mode == 1) {
    // max: reuse the sum machinery on a single warp pass
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
int la_norm(fx_plan* P, int n, const double* x, int kind, double* out_dev, cudaStream_t st) {
  const int nb = max(1, min((n + 255) / 256, min(P->part_cap, 148 * 4)));
  k_dot_partial<<<nb, 256, 0, st>>>(n, x, x, kind == 1 ? 1 : 0, P->part);
  k_finish<<<1, 256, 0, st>>>(P->part, nb, kind == 1 ? 1 : 2, out_dev);
  count_launch(2);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

// ---- persistent Krylov kernels ------------------------------------------------------------------------
struct KArgs {
  int n;
  const int* rowptr;
  const int* col;
  const double* val;
  double* x;
  const double* b;
  double* w;        // KRYLOV_NWORK x n work vectors
  double* red;      // 2 x RED_SLOTS x MAX_GRID
  double rtol, atol, dtol;
  int maxit, pc;
  fx_solve_info* info;
};

// Deterministic grid-wide sum of NV values: block partials -> scratch -> grid.sync -> every CTA sums
// all partials in the same order (bitwise identical result in every CTA).  `phase` alternates the
// scratch half so a fast CTA cannot overwrite partials a slow CTA is still reading.
template <int NV>
__device__ __forceinline__ void grid_sum(cg::grid_group& grid, double (&v)[NV], double* red, int& phase,
                                         double* sh) {
  double* buf = red + (size_t)(phase & 1) * RED_SLOTS * MAX_GRID;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = block_sum_bcast(v[i], sh);
    if (threadIdx.x == 0) buf[i * MAX_GRID + blockIdx.x] = s;
  }
  grid.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) s += buf[i * MAX_GRID + j];
    v[i] = block_sum_bcast(s, sh);
  }
  ++phase;
}

__device__ __forceinline__ void finish(const KArgs& a, int it, int conv, double rn, double bn) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.info->iterations = it;
    a.info->converged = conv;
    a.info->residual = rn;
    a.info->residual0 = bn;
  }
}

// Left-preconditioned BiCGStab (M = diag(A) or I), PETSc KSPBCGS semantics: the recurrence runs on
// M^-1 A, the monitored norm is ||M^-1 r||_2 and the test is ||r|| <= max(rtol ||M^-1 b||, atol).
template <int TPR>
__global__ void __launch_bounds__(256) k_bicgstab(const __grid_constant__ KArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[33];
  const int n = a.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  const int lane = gt % TPR, sub = gt / TPR, nsub = gs / TPR;  // global sub-warp id / count
  const int subw = (threadIdx.x & 31) / TPR;  // sub-warp index inside the warp
  double* r = a.w;
  double* rh = a.w + (size_t)n;
  double* p = a.w + 2 * (size_t)n;
  double* v = a.w + 3 * (size_t)n;
  double* s = a.w + 4 * (size_t)n;
  double* t = a.w + 5 * (size_t)n;
  double* dinv = a.w + 6 * (size_t)n;
  int phase = 0;

  // setup: dinv, r = M^-1 (b - A x), rhat = r, p = v = 0
  double acc[2] = {0.0, 0.0};
  for (int r0 = sub - subw; r0 < n; r0 += nsub) {
    const int row = r0 + subw;
    const double ax = row_dot<TPR>(a.rowptr, a.col, a.val, a.x, row, n, lane);
    double dg = 0.0;
    if (row < n && a.pc == FX_PC_JACOBI)
      for (int j = a.rowptr[row] + lane; j < a.rowptr[row + 1]; j += TPR)
        if (a.col[j] == row) dg = a.val[j];
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o, TPR);
    if (lane == 0 && row < n) {
      const double di = (a.pc == FX_PC_JACOBI && dg != 0.0) ? 1.0 / dg : 1.0;
      const double ri = di * (a.b[row] - ax), bi = di * a.b[row];
      dinv[row] = di; r[row] = ri; rh[row] = ri; p[row] = 0.0; v[row] = 0.0;
      acc[0] += ri * ri; acc[1] += bi * bi;
    }
  }
  grid_sum<2>(grid, acc, a.red, phase, sh);
  double rn = sqrt(acc[0]);
  const double bn = sqrt(acc[1]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(rn == rn)) { finish(a, 0, -1, rn, bn); return; }
  if (rn <= tol) { finish(a, 0, 1, rn, bn); return; }
  double rho = acc[0], rho_old = 1.0, alpha = 1.0, omega = 1.0;
  double rhn = rn;  // ||rhat||
  int restarts = 0;
  for (int it = 1; it <= a.maxit; ++it) {
    // (rhat, r) ~ 0 or omega ~ 0: the recurrence has broken down (e.g. a residual supported on
    // Dirichlet rows only is annihilated in one step, leaving rhat orthogonal to r).  Restart with
    // rhat = r, the standard cure; give up after a few consecutive restarts.
    if (fabs(rho) <= 1e-14 * rhn * rn || omega == 0.0) {
      if (++restarts > 50) { finish(a, it - 1, -1, rn, bn); return; }
      for (int i = gt; i < n; i += gs) { rh[i] = r[i]; p[i] = 0.0; v[i] = 0.0; }
      rho = rn * rn; rhn = rn; rho_old = 1.0; alpha = 1.0; omega = 1.0;
    }
    const double beta = (rho / rho_old) * (alpha / omega);
    for (int i = gt; i < n; i += gs) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    grid.sync();
    double d1[1] = {0.0};
    for (int r0 = sub - subw; r0 < n; r0 += nsub) {
      const int row = r0 + subw;
      const double ap = row_dot<TPR>(a.rowptr, a.col, a.val, p, row, n, lane);
      if (lane == 0 && row < n) {
        const double vi = dinv[row] * ap;
        v[row] = vi;
        d1[0] += rh[row] * vi;
      }
    }
    grid_sum<1>(grid, d1, a.red, phase, sh);
    if (d1[0] == 0.0) { finish(a, it - 1, -1, rn, bn); return; }
    alpha = rho / d1[0];
    double d2[1] = {0.0};
    for (int i = gt; i < n; i += gs) {
      const double si = r[i] - alpha * v[i];
      s[i] = si;
      d2[0] += si * si;
    }
    grid_sum<1>(grid, d2, a.red, phase, sh);
    if (sqrt(d2[0]) <= tol) {  // early exit on the half step (KSPBCGS does the same)
      for (int i = gt; i < n; i += gs) a.x[i] += alpha * p[i];
      finish(a, it, 1, sqrt(d2[0]), bn);
      return;
    }
    double d3[2] = {0.0, 0.0};
    for (int r0 = sub - subw; r0 < n; r0 += nsub) {
      const int row = r0 + subw;
      const double as = row_dot<TPR>(a.rowptr, a.col, a.val, s, row, n, lane);
      if (lane == 0 && row < n) {
        const double ti = dinv[row] * as;
        t[row] = ti;
        d3[0] += ti * s[row];
        d3[1] += ti * ti;
      }
    }
    grid_sum<2>(grid, d3, a.red, phase, sh);
    if (d3[1] == 0.0) { finish(a, it, -1, sqrt(d2[0]), bn); return; }
    omega = d3[0] / d3[1];
    double d4[2] = {0.0, 0.0};
    for (int i = gt; i < n; i += gs) {
      a.x[i] += alpha * p[i] + omega * s[i];
      const double ri = s[i] - omega * t[i];
      r[i] = ri;
      d4[0] += rh[i] * ri;
      d4[1] += ri * ri;
    }
    grid_sum<2>(grid, d4, a.red, phase, sh);
    rho_old = rho;
    rho = d4[0];
    rn = sqrt(d4[1]);
    if (!(rn == rn) || rn > a.dtol * bn) { finish(a, it, -1, rn, bn); return; }
    if (rn <= tol) { finish(a, it, 1, rn, bn); return; }
  }
  finish(a, a.maxit, 0, rn, bn);
}

// Preconditioned CG (M = diag(A) or I); monitored norm ||M^-1 r||_2 (PETSc default for KSPCG).
template <int TPR>
__global__ void __launch_bounds__(256) k_cg(const __grid_constant__ KArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[33];
  const int n = a.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  const int lane = gt % TPR, sub = gt / TPR, nsub = gs / TPR;
  const int subw = (threadIdx.x & 31) / TPR;
  double* r = a.w;
  double* p = a.w + (size_t)n;
  double* q = a.w + 2 * (size_t)n;
  double* dinv = a.w + 3 * (size_t)n;
  int phase = 0;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int r0 = sub - subw; r0 < n; r0 += nsub) {
    const int row = r0 + subw;
    const double ax = row_dot<TPR>(a.rowptr, a.col, a.val, a.x, row, n, lane);
    double dg = 0.0;
    if (row < n && a.pc == FX_PC_JACOBI)
      for (int j = a.rowptr[row] + lane; j < a.rowptr[row + 1]; j += TPR)
        if (a.col[j] == row) dg = a.val[j];
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o, TPR);
    if (lane == 0 && row < n) {
      const double di = (a.pc == FX_PC_JACOBI && dg != 0.0) ? 1.0 / dg : 1.0;
      const double ri = a.b[row] - ax, zi = di * ri, bi = di * a.b[row];
      dinv[row] = di; r[row] = ri; p[row] = zi;
      acc[0] += ri * zi; acc[1] += zi * zi; acc[2] += bi * bi;
    }
  }
  grid_sum<3>(grid, acc, a.red, phase, sh);
  double rz = acc[0], zn = sqrt(acc[1]);
  const double bn = sqrt(acc[2]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(zn == zn)) { finish(a, 0, -1, zn, bn); return; }
  if (zn <= tol) { finish(a, 0, 1, zn, bn); return; }
  for (int it = 1; it <= a.maxit; ++it) {
    double d1[1] = {0.0};
    for (int r0 = sub - subw; r0 < n; r0 += nsub) {
      const int row = r0 + subw;
      const double ap = row_dot<TPR>(a.rowptr, a.col, a.val, p, row, n, lane);
      if (lane == 0 && row < n) {
        q[row] = ap;
        d1[0] += p[row] * ap;
      }
    }
    grid_sum<1>(grid, d1, a.red, phase, sh);
    if (d1[0] == 0.0) { finish(a, it - 1, -1, zn, bn); return; }
    const double alpha = rz / d1[0];
    double d2[2] = {0.0, 0.0};
    for (int i = gt; i < n; i += gs) {
      a.x[i] += alpha * p[i];
      const double ri = r[i] - alpha * q[i];
      r[i] = ri;
      const double zi = dinv[i] * ri;
      d2[0] += ri * zi;
      d2[1] += zi * zi;
    }
    grid_sum<2>(grid, d2, a.red, phase, sh);
    zn = sqrt(d2[1]);
    if (!(zn == zn) || zn > a.dtol * bn) { finish(a, it, -1, zn, bn); return; }
    if (zn <= tol) { finish(a, it, 1, zn, bn); return; }
    const double beta = d2[0] / rz;
    rz = d2[0];
    for (int i = gt; i < n; i += gs) p[i] = dinv[i] * r[i] + beta * p[i];
    grid.sync();
  }
  finish(a, a.maxit, 0, zn, bn);
}

template <int TPR>
static int launch_krylov(fx_plan* P, int method, KArgs& ka, cudaStream_t st) {
  void* fn = method == FX_KSP_CG ? (void*)k_cg<TPR> : (void*)k_bicgstab<TPR>;
  int per_sm = 0;
  FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
  if (per_sm < 1) return set_error(FX_ERR_CUDA, "Krylov kernel does not fit on an SM");
  per_sm = min(per_sm, 4);
  long want = ((long)ka.n * TPR + 255) / 256;
  int grid = (int)min((long)per_sm * P->num_sms, max(1L, want));
  grid = min(grid, MAX_GRID);
  void* args[] = {(void*)&ka};
  FX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, st));
  count_launch(1);
  return FX_OK;
}

int la_krylov(fx_plan* P, int space, const double* val, double* x, const double* b, int method, int pc,
              double rtol, double atol, int maxit, int ticket, cudaStream_t st) {
  const DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  if (method != FX_KSP_CG && method != FX_KSP_BICGSTAB)
    return set_error(FX_ERR_UNSUPPORTED, "Krylov method %d not implemented (bicgstab, cg available)", method);
  if (pc != FX_PC_NONE && pc != FX_PC_JACOBI)
    return set_error(FX_ERR_UNSUPPORTED, "preconditioner %d not implemented (none, jacobi available)", pc);
  int rc = ensure_work(P, (size_t)p.n);
  if (rc) return rc;
  KArgs ka;
  ka.n = p.n; ka.rowptr = p.rowptr; ka.col = p.col; ka.val = val; ka.x = x; ka.b = b;
  ka.w = P->work; ka.red = P->red;
  ka.rtol = rtol; ka.atol = atol; ka.dtol = 1e5; ka.maxit = maxit; ka.pc = pc;
  ka.info = P->info + ticket;
  switch (pick_tpr(p)) {
    case 4: return launch_krylov<4>(P, method, ka, st);
    case 8: return launch_krylov<8>(P, method, ka, st);
    case 16: return launch_krylov<16>(P, method, ka, st);
    default: return launch_krylov<32>(P, method, ka, st);
  }
}

}  // namespace fx
