// FX_PC_ILU0: what `solve(A, x, b, "bicgstab", "default")` runs inside a serial PETSc process -- KSPBCGS / KSPCG with
// PCILU, levels 0, natural ordering, left preconditioning, convergence on ||M^-1 r||_2 (reference call sites:
// lid-driven-cavity/incomp_LDC.py:331,363,375,394,416; PETSc 3.7.5 pinned by fenics-install.sh:28-40).
//
// ILU(0) in the natural ordering is one dependency chain per row of the pattern.  Both the factorisation and the two
// triangular solves run "synchronisation-free": one warp per row, rows claimed in dependency order from an atomic
// ticket (so every row a warp waits for belongs to a warp that is already running), a per-row epoch flag published
// with st.release / polled with ld.acquire.  No level analysis on the host, one launch per factorisation and one per
// application of M^-1.  Sums run in the order of the CPU algorithm (ascending column, rounded product then subtract),
// so the factors and M^-1 r are bit-identical to the oracle's (oracle/krylov_c.c) -- only the dot products of the
// Krylov recurrences are reduced in a different order.
//
// The Krylov recurrences around it are plain stream-ordered kernels with their scalars in device memory; the host
// looks at the `done` flag every ILU_CHUNK iterations.  The depth of the dependency graph (tools/ilu_levels.py: ~14 n
// levels on an n x n mesh, thousands on the bench mesh) bounds this path, not HBM, and the measured cost (0.15 us per
// row and pass, profiles/r02_ilu0_vs_jacobi.jsonl) is still 2.5-5x above that bound: it exists for fidelity with the
// reference's default (same iteration counts as PETSc's); the Jacobi / two-level kernels of fx_krylov.cu are the fast
// path.
#include <algorithm>
#include <cmath>

#include "fx_plan.h"

namespace fx {

namespace {

constexpr int ILU_THREADS = 256;
constexpr int ILU_WPB = ILU_THREADS / 32;
constexpr int ILU_CHUNK = 4;  // iterations enqueued between two looks at the done flag

struct IluScal {
  double bn, tol, rho, alpha, omega, beta, rn;
  double acc[2];
  int done;  // 0 running, 1 finished (rc holds the outcome), 2 finished by the half-step exit (x += alpha p pending)
  int rc;    // fx_solve_info::converged
  int it;
  int err;   // factorisation: 1 + first row without a usable pivot
};

struct IluWork {
  double* lu = nullptr;  // factors in A's pattern
  size_t lu_cap = 0;
  int* dptr = nullptr;   // position of the diagonal of every row
  int* flagF = nullptr;  // epoch flags: factorisation / lower / upper solve
  int* flagL = nullptr;
  int* flagU = nullptr;
  double* vec = nullptr;  // 9 work vectors
  size_t n_cap = 0;
  IluScal* scal = nullptr;
  unsigned* ticket = nullptr;
  IluScal* scal_h = nullptr;  // page-locked
  int dptr_space = -1;        // the space dptr was built for
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int* f, int epoch) {
  while (ld_acquire(f) < epoch) __nanosleep(20);
}

// every warp takes the next row of the dependency order; warp-uniform
__device__ __forceinline__ unsigned claim(unsigned* ticket) {
  unsigned t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(ticket, 1u);
  return __shfl_sync(0xffffffffu, t, 0);
}

__global__ void k_ilu_diag(int n, const int* __restrict__ rowptr, const int* __restrict__ col, int* __restrict__ dptr,
                           IluScal* S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = -1;
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
    if (col[j] == i) { d = j; break; }
  dptr[i] = d;
  if (d < 0) atomicMax(&S->err, i + 1);
}

// IKJ elimination restricted to the pattern (oracle/krylov_c.c:fxo_ilu0).  lu holds a copy of A on entry.
__global__ void __launch_bounds__(ILU_THREADS) k_ilu_factor(int n, const int* __restrict__ rowptr,
                                                            const int* __restrict__ col, const int* __restrict__ dptr,
                                                            double* lu, int* flag, unsigned* ticket, IluScal* S) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    const unsigned i = claim(ticket);
    if (i >= (unsigned)n) return;
    const int r0 = rowptr[i], r1 = rowptr[i + 1], di = dptr[i];
    if (di >= 0) {
      for (int j = r0; j < di; ++j) {
        const int k = col[j];
        wait_flag(flag + k, 1);
        const int dk = dptr[k];
        const double piv = dk >= 0 ? __ldcg(lu + dk) : 1.0;
        const double f = __ddiv_rn(__ldcg(lu + j), piv);
        __syncwarp();
        if (lane == 0) lu[j] = f;
        if (f != 0.0 && dk >= 0) {
          const int k1 = rowptr[k + 1];
          for (int q = dk + 1 + lane; q < k1; q += 32) {
            const int c = col[q];
            int lo = j + 1, hi = r1 - 1, p = -1;  // columns of row i ascend: c > k = col[j]
            while (lo <= hi) {
              const int mid = (lo + hi) >> 1;
              const int cm = col[mid];
              if (cm == c) { p = mid; break; }
              if (cm < c) lo = mid + 1; else hi = mid - 1;
            }
            if (p >= 0) lu[p] = __dsub_rn(__ldcg(lu + p), __dmul_rn(f, __ldcg(lu + q)));
          }
        }
        __syncwarp();
      }
      if (lane == 0 && __ldcg(lu + di) == 0.0) atomicMax(&S->err, (int)i + 1);
    }
    __syncwarp();
    __threadfence();
    if (lane == 0) st_release(flag + i, 1);
  }
}

// z = U^-1 L^-1 r.  Tickets [0, n): forward rows ascending (y = L^-1 r, unit diagonal); [n, 2n): backward rows
// descending.  Ordered sums: the 32 products of a chunk are formed by the lanes, then subtracted in column order.
__global__ void __launch_bounds__(ILU_THREADS) k_ilu_apply(int n, const int* __restrict__ rowptr,
                                                           const int* __restrict__ col, const int* __restrict__ dptr,
                                                           const double* __restrict__ lu, const double* __restrict__ r,
                                                           double* y, double* z, int* flagL, int* flagU, int epoch,
                                                           unsigned* ticket, const IluScal* S) {
  if (S->done) return;
  const int lane = threadIdx.x & 31;
  for (;;) {
    const unsigned t = claim(ticket);
    if (t >= 2u * (unsigned)n) return;
    const bool fwd = t < (unsigned)n;
    const int i = fwd ? (int)t : 2 * n - 1 - (int)t;
    const int di = dptr[i];
    const int j0 = fwd ? rowptr[i] : di + 1, j1 = fwd ? di : rowptr[i + 1];
    const double* src = fwd ? y : z;
    const int* flag = fwd ? flagL : flagU;
    double s;
    if (fwd) s = r[i];
    else { wait_flag(flagL + i, epoch); s = __ldcg(y + i); }
    for (int jb = j0; jb < j1; jb += 32) {
      const int j = jb + lane;
      double term = 0.0;
      if (j < j1) {
        const int c = col[j];
        wait_flag(flag + c, epoch);
        term = __dmul_rn(lu[j], __ldcg(src + c));
      }
      const int cnt = min(32, j1 - jb);
      for (int l = 0; l < cnt; ++l) s = __dsub_rn(s, __shfl_sync(0xffffffffu, term, l));
    }
    if (lane == 0) {  // the value and its flag leave from the same thread: st.release orders them
      if (fwd) y[i] = s; else z[i] = __ddiv_rn(s, lu[di]);
      st_release((fwd ? flagL : flagU) + i, epoch);
    }
    __syncwarp();
  }
}

// y = A x (sub == 0) or y = b - A x (sub == 1); one warp per row of the plain CSR
__global__ void __launch_bounds__(ILU_THREADS) k_ilu_spmv(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                          const double* __restrict__ val, const double* __restrict__ x,
                                                          const double* __restrict__ b, double* __restrict__ y, int sub,
                                                          const IluScal* S) {
  if (S->done) return;
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = w; i < n; i += nw) {
    double s = 0.0;
    for (int j = rowptr[i] + lane; j < rowptr[i + 1]; j += 32) s += val[j] * x[col[j]];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = sub ? b[i] - s : s;
  }
}

__device__ __forceinline__ void block_add(double a, double b, double* acc) {
  __shared__ double sh[2][ILU_WPB];
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = a; sh[1][w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < ILU_WPB; ++k) { sa += sh[0][k]; sb += sh[1][k]; }
    atomicAdd(acc, sa);
    atomicAdd(acc + 1, sb);
  }
}

// vector passes of the recurrences; every one adds up to two dot products into S->acc
enum { V_DOT2 = 0, V_S, V_XR, V_P, V_HALF, V_CG_XR, V_CG_P, V_COPY2 };
__global__ void __launch_bounds__(ILU_THREADS) k_ilu_vec(int mode, int n, int it, double* a, double* b, double* c,
                                                         double* d, double* e, double* f, IluScal* S) {
  const int done = S->done;
  if (mode == V_HALF) { if (!(done == 2 && S->it == it)) return; }
  else if (done) return;
  const double alpha = S->alpha, omega = S->omega, beta = S->beta;
  double s0 = 0.0, s1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    switch (mode) {
      case V_DOT2:  // acc0 += (a, b), acc1 += (c, d)
        s0 += a[i] * b[i];
        if (c) s1 += c[i] * d[i];
        break;
      case V_S: {  // a = s, b = r, c = v:  s = r - alpha v; acc0 += (s, s)
        const double v = b[i] - alpha * c[i];
        a[i] = v; s0 += v * v;
      } break;
      case V_XR: {  // a = x, b = p, c = s, d = r, e = t, f = rh: x += alpha p + omega s; r = s - omega t
        a[i] += alpha * b[i] + omega * c[i];
        const double v = c[i] - omega * e[i];
        d[i] = v; s0 += v * v; s1 += f[i] * v;
      } break;
      case V_P:  // a = p, b = r, c = v: p = r + beta (p - omega v)
        a[i] = b[i] + beta * (a[i] - omega * c[i]);
        break;
      case V_HALF:  // a = x, b = p
        a[i] += alpha * b[i];
        break;
      case V_CG_XR:  // a = x, b = p, c = r, d = q: x += alpha p; r -= alpha q
        a[i] += alpha * b[i];
        c[i] -= alpha * d[i];
        break;
      case V_CG_P:  // a = p, b = z: p = z + beta p
        a[i] = b[i] + beta * a[i];
        break;
      case V_COPY2:  // a = b, c = b (c may be null)
        a[i] = b[i];
        if (c) c[i] = b[i];
        break;
    }
  }
  if (mode == V_DOT2 || mode == V_S || mode == V_XR) block_add(s0, s1, S->acc);
}

// scalar steps between the vector passes (one thread)
enum { SC_BN = 0, SC_R0, SC_ALPHA, SC_SN, SC_OMEGA, SC_END, SC_CG_R0, SC_CG_GAMMA0, SC_CG_ALPHA, SC_CG_RN, SC_CG_BETA, SC_INFO };
__global__ void k_ilu_scal(int step, double rtol, double atol, int maxit, IluScal* S, fx_solve_info* info) {
  if (step == SC_INFO) {
    info->iterations = S->it;
    info->converged = S->err ? -2 : S->rc;
    info->residual = S->rn;
    info->residual0 = S->bn;
    return;
  }
  if (S->done) return;
  const double a0 = S->acc[0], a1 = S->acc[1];
  S->acc[0] = 0.0; S->acc[1] = 0.0;
  switch (step) {
    case SC_BN:  // acc0 = ||M^-1 b||^2
      S->bn = sqrt(a0);
      S->tol = fmax(rtol * S->bn, atol);
      break;
    case SC_R0:  // acc0 = ||r0||^2 = (rh, r0)
    case SC_CG_R0:
      S->rn = sqrt(a0);
      if (S->rn <= S->tol) { S->done = 1; S->rc = 1; }
      else if (!(S->rn == S->rn)) { S->done = 1; S->rc = -1; }
      else if (maxit == 0) { S->done = 1; S->rc = 0; }
      S->rho = a0;
      S->it = S->done ? 0 : 1;
      break;
    case SC_CG_GAMMA0:  // acc0 = (r, z)
      S->rho = a0;
      break;
    case SC_ALPHA:  // acc0 = (rh, v)
      if (a0 == 0.0) { S->done = 1; S->rc = -1; break; }
      S->alpha = S->rho / a0;
      break;
    case SC_SN: {  // acc0 = (s, s)
      const double sn = sqrt(a0);
      if (sn <= S->tol) { S->rn = sn; S->done = 2; S->rc = 1; }
    } break;
    case SC_OMEGA:  // acc0 = (t, t), acc1 = (t, s)
      if (a0 == 0.0) { S->done = 1; S->rc = -1; break; }
      S->omega = a1 / a0;
      break;
    case SC_END: {  // acc0 = (r, r), acc1 = (rh, r)
      S->rn = sqrt(a0);
      if (S->rn <= S->tol) { S->done = 1; S->rc = 1; break; }
      if (!(S->rn == S->rn)) { S->done = 1; S->rc = -1; break; }
      S->beta = (a1 / S->rho) * (S->alpha / S->omega);
      S->rho = a1;
      if (S->it >= maxit) { S->done = 1; S->rc = 0; break; }
      S->it += 1;
    } break;
    case SC_CG_ALPHA:  // acc0 = (p, q)
      if (!(a0 > 0.0)) { S->done = 1; S->rc = -1; break; }
      S->alpha = S->rho / a0;
      break;
    case SC_CG_RN:  // acc0 = (z, z), acc1 = (r, z)
      S->rn = sqrt(a0);
      if (S->rn <= S->tol) { S->done = 1; S->rc = 1; break; }
      if (!(S->rn == S->rn)) { S->done = 1; S->rc = -1; break; }
      S->beta = a1 / S->rho;
      S->rho = a1;
      if (S->it >= maxit) { S->done = 1; S->rc = 0; break; }
      S->it += 1;
      break;
  }
}

}  // namespace

void ilu_release(fx_plan* P) {
  IluWork* W = static_cast<IluWork*>(P->ilu);
  if (!W) return;
  cudaFree(W->lu); cudaFree(W->dptr); cudaFree(W->flagF); cudaFree(W->flagL); cudaFree(W->flagU); cudaFree(W->vec);
  cudaFree(W->scal); cudaFree(W->ticket);
  cudaFreeHost(W->scal_h);
  delete W;
  P->ilu = nullptr;
}

static int ilu_ensure(fx_plan* P, size_t n, size_t nnz) {
  if (!P->ilu) P->ilu = new IluWork();
  IluWork* W = static_cast<IluWork*>(P->ilu);
  if (!W->scal) {
    FX_CUDA(cudaMalloc((void**)&W->scal, sizeof(IluScal)));
    FX_CUDA(cudaMalloc((void**)&W->ticket, sizeof(unsigned)));
    FX_CUDA(cudaMallocHost((void**)&W->scal_h, sizeof(IluScal)));
  }
  if (W->lu_cap < nnz) {
    cudaFree(W->lu);
    W->lu = nullptr; W->lu_cap = 0;
    FX_CUDA(cudaMalloc((void**)&W->lu, sizeof(double) * nnz));
    W->lu_cap = nnz;
  }
  if (W->n_cap < n) {
    cudaFree(W->dptr); cudaFree(W->flagF); cudaFree(W->flagL); cudaFree(W->flagU); cudaFree(W->vec);
    W->dptr = W->flagF = W->flagL = W->flagU = nullptr; W->vec = nullptr; W->n_cap = 0; W->dptr_space = -1;
    FX_CUDA(cudaMalloc((void**)&W->dptr, sizeof(int) * n));
    FX_CUDA(cudaMalloc((void**)&W->flagF, sizeof(int) * n));
    FX_CUDA(cudaMalloc((void**)&W->flagL, sizeof(int) * n));
    FX_CUDA(cudaMalloc((void**)&W->flagU, sizeof(int) * n));
    FX_CUDA(cudaMalloc((void**)&W->vec, sizeof(double) * 9 * n));
    W->n_cap = n;
  }
  return FX_OK;
}

// BiCGStab / CG with ILU(0), the algorithm of oracle/krylov_c.c:fxo_krylov statement by statement.  Waits for the
// solve (the host polls the done flag); the outcome lands in the plan's result slot `ticket` like every other solve.
int la_krylov_ilu0(fx_plan* P, int space, const double* val, double* x, const double* b, int method, double rtol,
                   double atol, int maxit, int ticket, cudaStream_t st) {
  DevPattern& p = P->sp[space];
  if (method != FX_KSP_BICGSTAB && method != FX_KSP_CG)
    return set_error(FX_ERR_UNSUPPORTED, "preconditioner ilu is not implemented for this Krylov method (bicgstab, cg are)");
  if (P->dist.world > 1)
    return set_error(FX_ERR_UNSUPPORTED, "ilu is not available on a partitioned plan (PETSc's parallel default is "
                     "block Jacobi + ILU(0), a different operator on every partition)");
  const int n = p.n;
  int rc = ilu_ensure(P, (size_t)n, (size_t)p.nnz);
  if (rc) return rc;
  IluWork* W = static_cast<IluWork*>(P->ilu);
  IluScal* S = W->scal;
  const int blocks_n = (n + ILU_THREADS - 1) / ILU_THREADS;
  const int grid_v = std::max(1, std::min(blocks_n, 8 * P->num_sms));
  const int grid_w = std::max(1, std::min((n + ILU_WPB - 1) / ILU_WPB, 8 * P->num_sms));
  int launches = 0;
  FX_CUDA(cudaMemsetAsync(S, 0, sizeof(IluScal), st));
  // ---- factorisation --------------------------------------------------------------------------------
  if (W->dptr_space != space) {
    k_ilu_diag<<<blocks_n, ILU_THREADS, 0, st>>>(n, p.rowptr, p.col, W->dptr, S);
    ++launches;
    W->dptr_space = space;
  }
  FX_CUDA(cudaMemcpyAsync(W->lu, val, sizeof(double) * (size_t)p.nnz, cudaMemcpyDeviceToDevice, st));
  FX_CUDA(cudaMemsetAsync(W->flagF, 0, sizeof(int) * (size_t)n, st));
  FX_CUDA(cudaMemsetAsync(W->flagL, 0, sizeof(int) * (size_t)n, st));
  FX_CUDA(cudaMemsetAsync(W->flagU, 0, sizeof(int) * (size_t)n, st));
  FX_CUDA(cudaMemsetAsync(W->ticket, 0, sizeof(unsigned), st));
  k_ilu_factor<<<grid_w, ILU_THREADS, 0, st>>>(n, p.rowptr, p.col, W->dptr, W->lu, W->flagF, W->ticket, S);
  ++launches;
  FX_CUDA(cudaMemcpyAsync(W->scal_h, S, sizeof(IluScal), cudaMemcpyDeviceToHost, st));
  FX_CUDA(cudaStreamSynchronize(st));
  if (W->scal_h->err) {
    W->dptr_space = -1;
    k_ilu_scal<<<1, 1, 0, st>>>(SC_INFO, rtol, atol, maxit, S, P->info + ticket);
    count_launch(launches + 1);
    FX_CUDA(cudaGetLastError());
    return FX_OK;  // converged = -2 in the result slot: fx_krylov_solve reports FX_ERR_NOCONV with that status
  }
  double* v0 = W->vec;
  auto V = [&](int k) { return v0 + (size_t)k * n; };
  double *r = V(0), *rh = V(1), *pp = V(2), *v = V(3), *s = V(4), *t = V(5), *tmp = V(6), *yy = V(7), *z = V(8);
  int epoch = 0;
  auto prec = [&](const double* src, double* dst) -> int {
    ++epoch;
    FX_CUDA(cudaMemsetAsync(W->ticket, 0, sizeof(unsigned), st));
    k_ilu_apply<<<grid_w, ILU_THREADS, 0, st>>>(n, p.rowptr, p.col, W->dptr, W->lu, src, yy, dst, W->flagL, W->flagU, epoch,
                                               W->ticket, S);
    ++launches;
    return FX_OK;
  };
  auto spmv = [&](const double* xin, double* yout, int sub) {
    k_ilu_spmv<<<grid_w, ILU_THREADS, 0, st>>>(n, p.rowptr, p.col, val, xin, b, yout, sub, S);
    ++launches;
  };
  auto vec = [&](int mode, int it, double* a, double* b2, double* c, double* d, double* e, double* f) {
    k_ilu_vec<<<grid_v, ILU_THREADS, 0, st>>>(mode, n, it, a, b2, c, d, e, f, S);
    ++launches;
  };
  auto scal = [&](int step) {
    k_ilu_scal<<<1, 1, 0, st>>>(step, rtol, atol, maxit, S, P->info + ticket);
    ++launches;
  };
  // ---- set-up: ||M^-1 b||, r0 -----------------------------------------------------------------------
  if ((rc = prec(b, tmp))) return rc;
  vec(V_DOT2, 0, tmp, tmp, nullptr, nullptr, nullptr, nullptr);
  scal(SC_BN);
  spmv(x, tmp, 1);
  if (method == FX_KSP_BICGSTAB) {
    if ((rc = prec(tmp, r))) return rc;
    vec(V_DOT2, 0, r, r, nullptr, nullptr, nullptr, nullptr);
    scal(SC_R0);
    vec(V_COPY2, 0, rh, r, pp, nullptr, nullptr, nullptr);
  } else {  // CG: r = b - A x, z = M^-1 r, monitored norm ||z||
    vec(V_COPY2, 0, r, tmp, nullptr, nullptr, nullptr, nullptr);
    if ((rc = prec(r, z))) return rc;
    vec(V_DOT2, 0, z, z, nullptr, nullptr, nullptr, nullptr);
    scal(SC_CG_R0);
    vec(V_DOT2, 0, r, z, nullptr, nullptr, nullptr, nullptr);
    scal(SC_CG_GAMMA0);
    vec(V_COPY2, 0, pp, z, nullptr, nullptr, nullptr, nullptr);
  }
  // ---- iterations ---------------------------------------------------------------------------------------
  int it = 1;
  for (;;) {
    FX_CUDA(cudaMemcpyAsync(W->scal_h, S, sizeof(IluScal), cudaMemcpyDeviceToHost, st));
    FX_CUDA(cudaStreamSynchronize(st));
    if (W->scal_h->done || it > maxit) break;
    for (int c = 0; c < ILU_CHUNK && it <= maxit; ++c, ++it) {
      if (method == FX_KSP_BICGSTAB) {
        spmv(pp, tmp, 0);
        if ((rc = prec(tmp, v))) return rc;
        vec(V_DOT2, it, rh, v, nullptr, nullptr, nullptr, nullptr);
        scal(SC_ALPHA);
        vec(V_S, it, s, r, v, nullptr, nullptr, nullptr);
        scal(SC_SN);
        vec(V_HALF, it, x, pp, nullptr, nullptr, nullptr, nullptr);
        spmv(s, tmp, 0);
        if ((rc = prec(tmp, t))) return rc;
        vec(V_DOT2, it, t, t, t, s, nullptr, nullptr);
        scal(SC_OMEGA);
        vec(V_XR, it, x, pp, s, r, t, rh);
        scal(SC_END);
        vec(V_P, it, pp, r, v, nullptr, nullptr, nullptr);
      } else {
        spmv(pp, tmp, 0);
        vec(V_DOT2, it, pp, tmp, nullptr, nullptr, nullptr, nullptr);
        scal(SC_CG_ALPHA);
        vec(V_CG_XR, it, x, pp, r, tmp, nullptr, nullptr);
        if ((rc = prec(r, z))) return rc;
        vec(V_DOT2, it, z, z, r, z, nullptr, nullptr);
        scal(SC_CG_RN);
        vec(V_CG_P, it, pp, z, nullptr, nullptr, nullptr, nullptr);
      }
    }
  }
  scal(SC_INFO);
  count_launch(launches);
  FX_CUDA(cudaGetLastError());
  return FX_OK;
}

}  // namespace fx
