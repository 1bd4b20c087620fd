// Krylov solvers (K9/K10 of SURVEY.md 2.3) as ONE persistent cooperative kernel per solve.
//   * k_bicgstab:  left-preconditioned BiCGStab, PETSc KSPBCGS semantics
//   * k_cg:        preconditioned CG in the Chronopoulos-Gear form (one fused reduction per iteration)
//   * k_pcg_res<R>: pipelined CG for small systems, matrix slice resident in shared memory
// SpMV (CSR-stream through shared memory), fused vector updates, dot products and the convergence test
// all run on the device, separated by grid.sync(); reductions are deterministic two-stage sums so every
// CTA takes the same branch.  Preconditioner: Jacobi (M = diag A) or none.
// Replaces PETSc KSP behind dolfin solve(A, x, b, "bicgstab", "default")
// (lid-driven-cavity/incomp_LDC.py:331 and every other solve site).
#include <cooperative_groups.h>

#include "fx_la_common.cuh"

namespace cg = cooperative_groups;

namespace fx {

struct KArgs {
  CsrView A;
  double* x;
  const double* b;
  double* w;    // KRYLOV_NWORK x n work vectors
  double* red;  // 2 x RED_SLOTS x MAX_GRID
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
  block_sum_bcast_n<NV>(v, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[i * MAX_GRID + blockIdx.x] = v[i];
  }
  grid.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) s += buf[i * MAX_GRID + j];
    v[i] = s;
  }
  block_sum_bcast_n<NV>(v, sh);
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

// Leaving a solve: record the outcome, then recover the eliminated rows from the final x.  Every CTA
// takes the same exit (deterministic reductions), so the barrier is safe.
__device__ __forceinline__ void leave(cg::grid_group& grid, const KArgs& a, int it, int conv, double rn, double bn) {
  finish(a, it, conv, rn, bn);
  if (a.A.elim != nullptr) {
    grid.sync();
    if (back_substitute(a.A, a.x, a.b) != 0) a.info->converged = -3;
  }
}

__device__ __forceinline__ double row_diag_inv(const CsrView& A, int row, int pc) {
  if (A.elim != nullptr && A.elim[row]) return 0.0;  // eliminated rows stay identically zero in the iteration
  if (pc != FX_PC_JACOBI) return 1.0;
  double dg = 0.0;
  for (int j = A.rowptr[row]; j < A.rowptr[row + 1]; ++j)
    if (A.col[j] == row) dg = A.val[j];
  return dg != 0.0 ? 1.0 / dg : 1.0;
}

// ---- BiCGStab -------------------------------------------------------------------------------------------
// The recurrence runs on M^-1 A; the monitored norm is ||M^-1 r||_2 and the test is
// ||r|| <= max(rtol ||M^-1 b||, atol) (PETSc default for left preconditioning).
__global__ void __launch_bounds__(LA_THREADS, 4) k_bicgstab(const __grid_constant__ KArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[4 * 33];
  __shared__ double prod[SPMV_SMEM_DOUBLES];
  const CsrView& A = a.A;
  const int n = A.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  double* r = a.w;
  double* rh = a.w + (size_t)n;
  double* p = a.w + 2 * (size_t)n;
  double* v = a.w + 3 * (size_t)n;
  double* s = a.w + 4 * (size_t)n;
  double* t = a.w + 5 * (size_t)n;
  double* dinv = a.w + 6 * (size_t)n;
  int phase = 0;

  // setup: compacted matrix, dinv, r = M^-1 (b - A x), rhat = r, p = v = 0
  double acc[3] = {0.0, 0.0, 0.0};
  if (A.rs != nullptr) acc[2] = (double)compact_zeros(A);
  spmv_stream(A, a.x, prod, [&](int row, double ax) {
    const double di = row_diag_inv(A, row, a.pc);
    const double ri = di * (a.b[row] - ax), bi = di * a.b[row];
    dinv[row] = di; r[row] = ri; rh[row] = ri; p[row] = 0.0; v[row] = 0.0;
    acc[0] += ri * ri; acc[1] += bi * bi;
  });
  grid_sum<3>(grid, acc, a.red, phase, sh);
  if (acc[2] != 0.0) { finish(a, 0, -3, 0.0, 0.0); return; }  // A[active, eliminated] != 0: not block triangular
  double rn = sqrt(acc[0]);
  const double bn = sqrt(acc[1]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(rn == rn)) { leave(grid, a, 0, -1, rn, bn); return; }
  if (rn <= tol) { leave(grid, a, 0, 1, rn, bn); return; }
  double rho = acc[0], rho_old = 1.0, alpha = 1.0, omega = 1.0;
  double rhn = rn;  // ||rhat||
  int restarts = 0;
  for (int it = 1; it <= a.maxit; ++it) {
    // (rhat, r) ~ 0 or omega ~ 0: the recurrence has broken down (e.g. a residual supported on
    // Dirichlet rows only is annihilated in one step, leaving rhat orthogonal to r).  Restart with
    // rhat = r, the standard cure.
    if (fabs(rho) <= 1e-14 * rhn * rn || omega == 0.0) {
      if (++restarts > 50) { leave(grid, a, it - 1, -1, rn, bn); return; }
      for (int i = gt; i < n; i += gs) { rh[i] = r[i]; p[i] = 0.0; v[i] = 0.0; }
      rho = rn * rn; rhn = rn; rho_old = 1.0; alpha = 1.0; omega = 1.0;
    }
    const double beta = (rho / rho_old) * (alpha / omega);
    for (int i = gt; i < n; i += gs) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    grid.sync();
    double d1[1] = {0.0};
    spmv_stream(A, p, prod, [&](int row, double ap) {
      const double vi = dinv[row] * ap;
      v[row] = vi;
      d1[0] += rh[row] * vi;
    });
    grid_sum<1>(grid, d1, a.red, phase, sh);
    if (d1[0] == 0.0) { leave(grid, a, it - 1, -1, rn, bn); return; }
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
      leave(grid, a, it, 1, sqrt(d2[0]), bn);
      return;
    }
    double d3[2] = {0.0, 0.0};
    spmv_stream(A, s, prod, [&](int row, double as) {
      const double ti = dinv[row] * as;
      t[row] = ti;
      d3[0] += ti * s[row];
      d3[1] += ti * ti;
    });
    grid_sum<2>(grid, d3, a.red, phase, sh);
    if (d3[1] == 0.0) { leave(grid, a, it, -1, sqrt(d2[0]), bn); return; }
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
    if (!(rn == rn) || rn > a.dtol * bn) { leave(grid, a, it, -1, rn, bn); return; }
    if (rn <= tol) { leave(grid, a, it, 1, rn, bn); return; }
  }
  leave(grid, a, a.maxit, 0, rn, bn);
}

// ---- CG (Chronopoulos-Gear) ---------------------------------------------------------------------------------
// u = M^-1 r, w = A u, p = u + beta p, s = w + beta s (= A p).  One SpMV and ONE fused reduction
// (gamma = (r,u), delta = (w,u), ||u||^2) per iteration: two grid barriers instead of three.
// Monitored norm ||M^-1 r||_2 (PETSc default for KSPCG).
__global__ void __launch_bounds__(LA_THREADS, 4) k_cg(const __grid_constant__ KArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[4 * 33];
  __shared__ double prod[SPMV_SMEM_DOUBLES];
  const CsrView& A = a.A;
  const int n = A.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  double* r = a.w;
  double* u = a.w + (size_t)n;
  double* w = a.w + 2 * (size_t)n;
  double* p = a.w + 3 * (size_t)n;
  double* s = a.w + 4 * (size_t)n;
  double* dinv = a.w + 5 * (size_t)n;
  int phase = 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (A.rs != nullptr) acc[3] = (double)compact_zeros(A);
  spmv_stream(A, a.x, prod, [&](int row, double ax) {
    const double di = row_diag_inv(A, row, a.pc);
    const double ri = a.b[row] - ax, ui = di * ri, bi = di * a.b[row];
    dinv[row] = di; r[row] = ri; u[row] = ui; p[row] = 0.0; s[row] = 0.0;
    acc[0] += ri * ui; acc[1] += ui * ui; acc[2] += bi * bi;
  });
  grid_sum<4>(grid, acc, a.red, phase, sh);
  if (acc[3] != 0.0) { finish(a, 0, -3, 0.0, 0.0); return; }
  double gamma = acc[0], un = sqrt(acc[1]);
  const double bn = sqrt(acc[2]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(un == un)) { leave(grid, a, 0, -1, un, bn); return; }
  if (un <= tol) { leave(grid, a, 0, 1, un, bn); return; }
  double d0[1] = {0.0};
  spmv_stream(A, u, prod, [&](int row, double au) {
    w[row] = au;
    d0[0] += au * u[row];
  });
  grid_sum<1>(grid, d0, a.red, phase, sh);
  if (!(d0[0] > 0.0)) { leave(grid, a, 0, -1, un, bn); return; }
  double alpha = gamma / d0[0], beta = 0.0;
  for (int it = 1; it <= a.maxit; ++it) {
    double d[3] = {0.0, 0.0, 0.0};
    for (int i = gt; i < n; i += gs) {
      const double pi = u[i] + beta * p[i];
      const double si = w[i] + beta * s[i];
      p[i] = pi; s[i] = si;
      a.x[i] += alpha * pi;
      const double ri = r[i] - alpha * si;
      r[i] = ri;
      const double ui = dinv[i] * ri;
      u[i] = ui;
      d[0] += ri * ui;
      d[1] += ui * ui;
    }
    grid.sync();
    spmv_stream(A, u, prod, [&](int row, double au) {
      w[row] = au;
      d[2] += au * u[row];
    });
    grid_sum<3>(grid, d, a.red, phase, sh);
    un = sqrt(d[1]);
    if (!(un == un) || un > a.dtol * bn) { leave(grid, a, it, -1, un, bn); return; }
    if (un <= tol) { leave(grid, a, it, 1, un, bn); return; }
    beta = d[0] / gamma;
    const double denom = d[2] - beta * d[0] / alpha;
    if (!(denom > 0.0)) { leave(grid, a, it, -1, un, bn); return; }  // not SPD / breakdown
    alpha = d[0] / denom;
    gamma = d[0];
  }
  leave(grid, a, a.maxit, 0, un, bn);
}

// ---- pipelined CG (Ghysels-Vanroose), matrix resident in shared memory -----------------------------------
// Small systems (the P1 pressure/density matrices: 37k rows, 3 MB at n=192) need ~10^3 iterations of
// microseconds each: the cost is latency, not bandwidth.  Two measures:
//  (1) the pipelined recurrence (u = M^-1 r, w = A u, m = M^-1 w, n = A m; z, q, s, p auxiliary) makes
//      the SpMV independent of the iteration's dot products, so ONE grid barrier per iteration serves
//      both the reduction and the SpMV's data dependency, and every vector update lives in the SpMV
//      epilogue (the lane that sums row i owns entry i of every vector): an iteration is one pass;
//  (2) each warp keeps its <= R row blocks (zero-compacted values, columns, row offsets) in shared
//      memory for the whole solve, so an iteration issues only the m-gather and its own-row updates.
// The recurrences lose a few digits of attainable accuracy: la_krylov selects this kernel only for
// rtol >= 1e-9 and when the matrix fits (R <= PCG_MAXR row blocks per warp with one CTA per SM).
constexpr int PCG_MAXR = 3;
__host__ __device__ constexpr size_t pcg_smem_bytes(int R) {
  return (size_t)LA_WARPS * ((size_t)R * SPMV_NNZB * (sizeof(double) + sizeof(int)) + SPMV_NNZB * sizeof(double));
}

template <int R>
__global__ void __launch_bounds__(LA_THREADS, 1) k_pcg_res(const __grid_constant__ KArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[3 * 33];
  extern __shared__ __align__(16) unsigned char dyn[];
  const CsrView& A = a.A;
  const size_t n = (size_t)A.n;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nw = gridDim.x * LA_WARPS, gw = blockIdx.x * LA_WARPS + wib;
  // per-warp shared slices
  double* sval = reinterpret_cast<double*>(dyn) + (size_t)wib * R * SPMV_NNZB;
  double* prod = reinterpret_cast<double*>(dyn) + (size_t)LA_WARPS * R * SPMV_NNZB + (size_t)wib * SPMV_NNZB;
  int* scol = reinterpret_cast<int*>(dyn + (size_t)LA_WARPS * (R + 1) * SPMV_NNZB * sizeof(double)) +
              (size_t)wib * R * SPMV_NNZB;
  double* r = a.w;
  double* u = a.w + n;
  double* w = a.w + 2 * n;
  double* z = a.w + 3 * n;
  double* q = a.w + 4 * n;
  double* s = a.w + 5 * n;
  double* p = a.w + 6 * n;
  double* dinv = a.w + 7 * n;
  double* mbuf[2] = {a.w + 8 * n, a.w + 9 * n};
  // row-block registers: my row, slice of `prod` holding my row, entries in the block
  int my_row[R], my_s0[R], my_s1[R], blk_nnz[R];
  double my_dinv[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int k = gw + rr * nw;
    my_row[rr] = -1; my_s0[rr] = 0; my_s1[rr] = 0; blk_nnz[rr] = 0; my_dinv[rr] = 1.0;
    if (k < A.nrowblk) {
      const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
      const int row = r0 + lane;
      const bool act = row < r1;
      int b = 0, e = 0, cnt = 0;
      double dg = 0.0;
      if (act) {
        b = A.rowptr[row]; e = A.rowptr[row + 1];
        for (int j = b; j < e; ++j) {
          const double v = A.val[j];
          cnt += (v != 0.0);
          if (A.col[j] == row) dg = v;
        }
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      blk_nnz[rr] = __shfl_sync(0xffffffffu, incl, 31);
      if (act) {
        int dst = incl - cnt;
        my_row[rr] = row; my_s0[rr] = dst; my_s1[rr] = dst + cnt;
        my_dinv[rr] = (a.pc == FX_PC_JACOBI && dg != 0.0) ? 1.0 / dg : 1.0;
        for (int j = b; j < e; ++j) {
          const double v = A.val[j];
          if (v != 0.0) { sval[rr * SPMV_NNZB + dst] = v; scol[rr * SPMV_NNZB + dst] = A.col[j]; ++dst; }
        }
      }
    }
  }
  __syncwarp();
  // y_row = sum_j val_j x[col_j] for my rows, from the resident slices
  auto spmv_res = [&](const double* x, double (&y)[R]) {
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int nnz = blk_nnz[rr];
      const double* sv = sval + rr * SPMV_NNZB;
      const int* sc = scol + rr * SPMV_NNZB;
      for (int jb = 0; jb < nnz; jb += 128) {
        double t[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int j = jb + uu * 32 + lane;
          t[uu] = (j < nnz) ? x[sc[j]] : 0.0;
        }
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int j = jb + uu * 32 + lane;
          if (j < nnz) prod[j] = sv[j] * t[uu];
        }
      }
      __syncwarp();
      double acc = 0.0;
      for (int j = my_s0[rr]; j < my_s1[rr]; ++j) acc += prod[j];
      y[rr] = acc;
      __syncwarp();
    }
  };
  int phase = 0;
  double y[R];
  // r = b - A x, u = M^-1 r
  double acc1[1] = {0.0};
  spmv_res(a.x, y);
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = my_row[rr];
    if (row >= 0) {
      const double ri = a.b[row] - y[rr], bi = my_dinv[rr] * a.b[row];
      dinv[row] = my_dinv[rr]; r[row] = ri; u[row] = my_dinv[rr] * ri;
      z[row] = 0.0; q[row] = 0.0; s[row] = 0.0; p[row] = 0.0;
      acc1[0] += bi * bi;
    }
  }
  grid_sum<1>(grid, acc1, a.red, phase, sh);
  const double bn = sqrt(acc1[0]);
  const double tol = fmax(a.rtol * bn, a.atol);
  // w = A u, m = M^-1 w
  double d[3] = {0.0, 0.0, 0.0};  // gamma = (r,u), delta = (w,u), ||u||^2
  spmv_res(u, y);
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = my_row[rr];
    if (row >= 0) {
      const double ui = u[row];
      w[row] = y[rr];
      mbuf[0][row] = my_dinv[rr] * y[rr];
      d[0] += r[row] * ui; d[1] += y[rr] * ui; d[2] += ui * ui;
    }
  }
  grid_sum<3>(grid, d, a.red, phase, sh);
  double gamma_old = 1.0, alpha = 1.0, un = sqrt(d[2]);
  for (int it = 0; it <= a.maxit; ++it) {
    un = sqrt(d[2]);
    if (!(un == un) || un > a.dtol * bn) { finish(a, it, -1, un, bn); return; }
    if (un <= tol) { finish(a, it, 1, un, bn); return; }
    if (it == a.maxit) break;
    const double gamma = d[0], delta = d[1];
    double beta = 0.0;
    if (it > 0) {
      beta = gamma / gamma_old;
      const double den = delta - beta * gamma / alpha;
      if (!(den > 0.0)) { finish(a, it, -1, un, bn); return; }
      alpha = gamma / den;
    } else {
      if (!(delta > 0.0)) { finish(a, it, -1, un, bn); return; }
      alpha = gamma / delta;
    }
    gamma_old = gamma;
    const double* m = mbuf[it & 1];
    double* mn = mbuf[(it + 1) & 1];
    // own-row operands first (independent of the gather): their latency overlaps the SpMV
    double zo[R], qo[R], so[R], po[R], uo[R], wo[R], ro[R], mo[R], xo[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = my_row[rr];
      if (row >= 0) {
        zo[rr] = z[row]; qo[rr] = q[row]; so[rr] = s[row]; po[rr] = p[row]; uo[rr] = u[row];
        wo[rr] = w[row]; ro[rr] = r[row]; mo[rr] = m[row]; xo[rr] = a.x[row];
      }
    }
    spmv_res(m, y);
    d[0] = d[1] = d[2] = 0.0;
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = my_row[rr];
      if (row >= 0) {
        const double zi = y[rr] + beta * zo[rr];
        const double qi = mo[rr] + beta * qo[rr];
        const double si = wo[rr] + beta * so[rr];
        const double pi = uo[rr] + beta * po[rr];
        z[row] = zi; q[row] = qi; s[row] = si; p[row] = pi;
        a.x[row] = xo[rr] + alpha * pi;
        const double ri = ro[rr] - alpha * si;
        const double ui = uo[rr] - alpha * qi;
        const double wi = wo[rr] - alpha * zi;
        r[row] = ri; u[row] = ui; w[row] = wi;
        mn[row] = my_dinv[rr] * wi;
        d[0] += ri * ui; d[1] += wi * ui; d[2] += ui * ui;
      }
    }
    grid_sum<3>(grid, d, a.red, phase, sh);
  }
  finish(a, a.maxit, 0, un, bn);
}

template <int R>
static int launch_pcg_res(fx_plan* P, KArgs& ka, int grid, cudaStream_t st) {
  const size_t smem = pcg_smem_bytes(R);
  static bool attr_done = false;
  if (!attr_done) {
    FX_CUDA(cudaFuncSetAttribute(k_pcg_res<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  void* args[] = {(void*)&ka};
  FX_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg_res<R>, dim3(grid), dim3(LA_THREADS), args, smem, st));
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
  ka.A = CsrView{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, P->crs, P->cre, P->cblkend, P->ccol, P->cval,
                 p.nelim > 0 ? p.elim : nullptr, p.elim_rows, p.nelim};
  ka.x = x; ka.b = b; ka.w = P->work; ka.red = P->red;
  ka.rtol = rtol; ka.atol = atol; ka.dtol = 1e5; ka.maxit = maxit; ka.pc = pc;
  ka.info = P->info + ticket;
  // small SPD systems at loose tolerance: shared-memory-resident pipelined CG, one CTA per SM
  if (method == FX_KSP_CG && rtol >= 1e-9 && p.nelim == 0) {
    const int warps = P->num_sms * LA_WARPS;
    const int R = (p.nrowblk + warps - 1) / warps;
    if (R <= PCG_MAXR) {
      const int grid = min(P->num_sms, (p.nrowblk + LA_WARPS - 1) / LA_WARPS);
      if (R <= 1) return launch_pcg_res<1>(P, ka, grid, st);
      if (R == 2) return launch_pcg_res<2>(P, ka, grid, st);
      return launch_pcg_res<3>(P, ka, grid, st);
    }
  }
  void* fn = method == FX_KSP_CG ? (void*)k_cg : (void*)k_bicgstab;
  int per_sm = 0;
  FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, LA_THREADS, 0));
  if (per_sm < 1) return set_error(FX_ERR_CUDA, "Krylov kernel does not fit on an SM");
  // one CTA per 8 row blocks up to the co-resident limit; small systems get a small grid (cheap barriers)
  int grid = max(1, min(min(per_sm, 8) * P->num_sms, (p.nrowblk + LA_WARPS - 1) / LA_WARPS));
  grid = min(grid, MAX_GRID);
  void* args[] = {(void*)&ka};
  FX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LA_THREADS), args, 0, st));
  count_launch(1);
  return FX_OK;
}

}  // namespace fx
