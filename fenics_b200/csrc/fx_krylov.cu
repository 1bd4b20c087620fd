// Krylov solvers (K9/K10 of SURVEY.md 2.3) as ONE persistent cooperative kernel per solve.
//   * k_bicgstab:  left-preconditioned BiCGStab, PETSc KSPBCGS semantics
//   * k_cg:        preconditioned CG in the Chronopoulos-Gear form (one fused reduction per iteration)
//   * k_pcg_res<R>: pipelined CG for small systems, matrix slice resident in shared memory
//   * k_gmres:     left-preconditioned restarted GMRES(30), PETSc KSPGMRES defaults
// SpMV (CSR-stream through shared memory), fused vector updates, dot products and the convergence test
// all run on the device, separated by grid.sync(); reductions are deterministic two-stage sums so every
// CTA takes the same branch.  Preconditioner: Jacobi (M = diag A) or none.
// Replaces PETSc KSP behind dolfin solve(A, x, b, "bicgstab", "default")
// (lid-driven-cavity/incomp_LDC.py:331 and every other solve site).
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "fx_la_common.cuh"
#include "fx_nb.cuh"

namespace cg = cooperative_groups;

#ifndef FX_NB_PU
#define FX_NB_PU 2
#endif

namespace fx {

// optional phase timestamps (FX_KRYLOV_PROF=1): block 0 / thread 0 stamps %globaltimer at phase boundaries
__device__ unsigned long long g_prof[1024];
__device__ unsigned long long g_prof_warp[8192];  // per-warp end of the first SpMV of iteration 2
__device__ __forceinline__ void stamp(int prof, int& np) {
  if (prof && blockIdx.x == 0 && threadIdx.x == 0 && np < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_prof[np++] = t;
  }
}

// two-level preconditioner of k_pcg_res (fx_plan_set_aggregates); agg == null: plain Jacobi
struct Coarse {
  int slice;       // k_pcg_res: entries of shared memory per resident row block (>= the largest block, multiple of 32)
  int nlmax;       // k_pcg_res: rows of E^-1 a CTA can keep in shared memory (<= COARSE_NLMAX)
  const int* agg;  // [n]
  int k;           // aggregates (multiple of 32, <= COARSE_KMAX)
  double* E0;      // k x k
  double* E1;      // k x k (ping-pong partner of the blocked inversion)
  double* cvec;    // [3][COARSE_KMAX] rotating coarse right-hand sides
};

struct KArgs {
  DistDev dd;  // mesh-partitioned run (world > 1): peer-mapped work regions, halo push lists
  Coarse cz;
  int prof;
  CsrView A;
  double* x;
  const double* b;
  double* w;    // KRYLOV_NWORK x n work vectors
  double* red;  // 2 x RED_SLOTS x MAX_GRID
  unsigned* bar;  // grid-barrier counter (zeroed before the launch)
  int red_atomic;  // grid sums through fp64 atomics (3 rotating accumulators) instead of per-CTA partials
  double rtol, atol, dtol;
  int maxit, pc;
  fx_solve_info* info;
  NbView nb;        // node-block solver layout of the space (k_bicgstab_nb)
  double* basis;    // GMRES: (GMRES_M + 1) x n Krylov basis
  double* red_big;  // GMRES: 2 x GMRES_NRED x MAX_GRID block partials of the Gram-Schmidt coefficients
};

// Deterministic grid-wide sum of NV values: block partials -> scratch -> grid barrier -> every CTA sums
// all partials in the same order (bitwise identical result in every CTA).  `phase` alternates the
// scratch half so a fast CTA cannot overwrite partials a slow CTA is still reading.
template <int NV, class Bar>
__device__ __forceinline__ void grid_sum(Bar& bar, double (&v)[NV], double* red, int& phase, double* sh) {
  double* buf = red + (size_t)(phase & 1) * RED_SLOTS * MAX_GRID;
  block_sum_bcast_n<NV>(v, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[i * MAX_GRID + blockIdx.x] = v[i];
  }
  bar.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) s += __ldcg(buf + i * MAX_GRID + j);
    v[i] = s;
  }
  block_sum_bcast_n<NV>(v, sh);
  ++phase;
}

// Variant for latency-bound solves: the CTA sums are added into one of three rotating accumulators with
// fp64 atomics (fire-and-forget REDs ahead of the barrier's fence), so after the barrier every thread
// reads the NV totals directly -- no partials array to re-read, no second block reduction.  The sum order
// varies from run to run (last-bit differences), but every CTA reads the same totals, so control flow
// stays uniform.  Accumulator (phase+1)%3 is cleared by CTA 0 before it arrives: everybody finished
// reading it before arriving at the previous barrier, and nobody adds to it until this barrier opens.
template <int NV, bool PUSHED, class Bar>
__device__ __forceinline__ void grid_sum_atomic(Bar& bar, double (&v)[NV], double* red, int& phase, double* sh) {
  double* acc = red + (size_t)2 * RED_SLOTS * MAX_GRID;  // [3][RED_SLOTS]
  double* cur = acc + (phase % 3) * RED_SLOTS;
  block_sum_bcast_n<NV>(v, sh);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (threadIdx.x == i) atomicAdd(cur + i, v[i]);
  if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + RED_SLOTS)
    acc[((phase + 1) % 3) * RED_SLOTS + threadIdx.x - 32] = 0.0;
  bar.template sync_x<NV, PUSHED>(cur);
  if (!bar.dist()) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcg(cur + i);
  } else {
    // partitioned run: the barrier handed this GPU's totals to every peer and left theirs in bar.xs.
    // Everybody adds the per-rank totals in rank order, so all GPUs hold bitwise identical sums and take
    // the same branches.
    const DistDev& dd = *bar.dd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = 0.0;
      for (int q = 0; q < dd.world; ++q) t += (q == dd.rank) ? __ldcg(cur + i) : bar.xs[q * RED_SLOTS + i];
      v[i] = t;
    }
  }
  ++phase;
}

// Single-GPU variant with the block reduction folded into the barrier: warp sums -> shared memory -> (the barrier's
// first __syncthreads) -> warp 0 finishes the NV sums, adds them to the rotating accumulator and arrives.  Two CTA
// barriers per grid sum instead of five.  sh: NV x 32 doubles.
constexpr int FAST_SLOTS = 8;
template <int NV>
__device__ __forceinline__ void grid_sum_fast(GridBar& bar, double (&v)[NV], double* red, int& phase, double* sh) {
  static_assert(NV <= FAST_SLOTS, "too many values");
  double* acc = red + (size_t)2 * RED_SLOTS * MAX_GRID + 3 * RED_SLOTS;  // [3][FAST_SLOTS]
  double* cur = acc + (phase % 3) * FAST_SLOTS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 32 + w] = v[i];
  }
  bar.target += gridDim.x;
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = lane < nw ? sh[i * 32 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) atomicAdd(cur + i, t);
    }
    if (blockIdx.x == 0 && lane < FAST_SLOTS) acc[((phase + 1) % 3) * FAST_SLOTS + lane] = 0.0;
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicAdd(bar.ctr, 1u);
      bar.wait_local();
      unsigned t;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(t) : "l"(bar.ctr) : "memory");
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldcg(cur + i);
  ++phase;
}

// PUSHED = false: no thread stored halo values to a peer since the previous barrier
template <int NV, bool PUSHED = true, class Bar>
__device__ __forceinline__ void gsum(const KArgs& a, Bar& bar, double (&v)[NV], int& phase, double* sh) {
  if (a.red_atomic || bar.dist()) grid_sum_atomic<NV, PUSHED>(bar, v, a.red, phase, sh);
  else grid_sum<NV>(bar, v, a.red, phase, sh);
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
template <class Bar>
__device__ __forceinline__ void leave(Bar& bar, const KArgs& a, int it, int conv, double rn, double bn) {
  finish(a, it, bar.dead ? -4 : conv, rn, bn);
  if (bar.dist()) {
    // partitioned run: the solution's ghost entries are coefficients of the next forms.  Owners store their
    // interface values into the neighbours' exchange slot (P2P), a barrier, ghosts copy them out.
    // Eliminated rows are skipped: every rank back-substitutes the DG0 rows of its ghost cells itself
    // (those rows are complete on the sub-mesh).
    const DistDev& dd = a.dd;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    bar.sync();  // x is final on every rank
    for (int k = gt; k < dd.nsend; k += gs) {
      const int i = dd.s_local[k];
      if (!(a.A.elim[i] & MASK_ELIM))
        dd.region[dd.s_peer[k] & (DIST_MAXW - 1)][(long long)DIST_XH * dd.nmax + dd.s_remote[k]] = a.x[i];
    }
    bar.sync();
    const double* xh = a.w + (size_t)DIST_XH * dd.nmax;
    for (int i = gt; i < a.A.n; i += gs)
      if ((a.A.elim[i] & MASK_INACTIVE) == MASK_GHOST) a.x[i] = __ldcg(xh + i);
  }
  if (a.A.elim != nullptr && a.A.nelim > 0) {
    bar.sync();
    if (back_substitute(a.A, a.x, a.b) != 0) a.info->converged = -3;
  }
  bar.save();
}
__device__ __forceinline__ double row_diag_inv(const CsrView& A, int row, int pc) {
  if (A.elim != nullptr && (A.elim[row] & MASK_INACTIVE)) return 0.0;  // eliminated / ghost rows stay identically zero in the iteration
  if (pc != FX_PC_JACOBI && pc != FX_PC_JACOBI_COARSE) return 1.0;
  double dg = 0.0;
  for (int j = A.rowptr[row]; j < A.rowptr[row + 1]; ++j)
    if (A.col[j] == row) dg = A.val[j];
  return dg != 0.0 ? 1.0 / dg : 1.0;
}

struct Ops2 { double a, b; };
struct Ops3 { double a, b, c; };
struct Ops4 { double a, b, c, d; };

// ---- BiCGStab -------------------------------------------------------------------------------------------
// The recurrence runs on M^-1 A; the monitored norm is ||M^-1 r||_2 and the test is
// ||r|| <= max(rtol ||M^-1 b||, atol) (PETSc default for left preconditioning).
// Four grid barriers per iteration: the scalars of the second half step ((t,s), (t,t), (rhat,s), (rhat,t))
// are reduced together, which gives omega, rho_new = (rhat, s - omega t) and, from (s,s), the residual
// norm ||s - omega t|| without a reduction of their own, so x, r and the next search direction p are
// updated in ONE pass over the vectors.  A residual that the recurrence reports as converged is
// confirmed with a true norm before the solve returns.
template <int THREADS, bool DIST>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_bicgstab(const __grid_constant__ KArgs a) {
  __shared__ double sh[4 * 33];
  __shared__ int sh_i[64];
  extern __shared__ __align__(16) double prod[];  // (THREADS/32) x SPMV_NNZB
  const CsrView& A = a.A;
  const int n = A.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  // work vectors (stride nmax in a partitioned run: the same vector sits at the same offset on every GPU);
  // p (index 2) and s (index 4) are SpMV inputs: their ghost entries are written by the owners' halo pushes
  // and never locally
  const size_t ld = DIST ? (size_t)a.dd.nmax : (size_t)n;
  const unsigned char* mask = DIST ? A.elim : nullptr;
  double* r = a.w;
  double* rh = a.w + ld;
  double* p = a.w + 2 * ld;
  double* v = a.w + 3 * ld;
  double* s = a.w + 4 * ld;
  double* t = a.w + 5 * ld;
  double* dinv = a.w + 6 * ld;
  int phase = 0;
  __shared__ double sh_x[DIST ? DIST_MAXW * RED_SLOTS : 1];
  GridBarT<DIST> bar;
  bar.init(a.bar, &a.dd, sh_x);

  // setup: compacted matrix, balanced partition, dinv, r = M^-1 (b - A x), rhat = p = r, v = 0
  double acc[3] = {0.0, 0.0, 0.0};
  int nb = A.nrowblk;
  if (A.blk_target > 0) {
    acc[2] = (double)compact_reblock(A, bar, sh_i, nb);
  } else {
    acc[2] = (double)compact_zeros(A);
    bar.sync();
  }
  const BlockRange rg = balance_partition(A, bar, sh_i, nb);
  spmv_stream_c(A, rg, a.x, prod,
                [&](int row) { return Ops2{row_diag_inv(A, row, a.pc), a.b[row]}; },
                [&](int row, double ax, const Ops2& o) {
                  const double ri = o.a * (o.b - ax), bi = o.a * o.b;
                  dinv[row] = o.a; r[row] = ri; rh[row] = ri; v[row] = 0.0;
                  const unsigned char mk = mask != nullptr ? mask[row] : 0;
                  if (!(mk & MASK_GHOST)) p[row] = ri;
                  if (mk & MASK_IFACE) push_iface(a.dd, 2, row, ri);
                  acc[0] += ri * ri; acc[1] += bi * bi;
                });
  gsum<3>(a, bar, acc, phase, sh);
  if (acc[2] != 0.0) { finish(a, 0, -3, 0.0, 0.0); bar.save(); return; }  // A[active, eliminated] != 0: not block triangular
  double rn = sqrt(acc[0]);
  const double bn = sqrt(acc[1]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(rn == rn)) { leave(bar, a, 0, -1, rn, bn); return; }
  if (rn <= tol) { leave(bar, a, 0, 1, rn, bn); return; }
  double rho = acc[0];
  double rhn = rn;  // ||rhat||
  int restarts = 0;
  int np = 0;
  stamp(a.prof, np);
  for (int it = 1; it <= a.maxit; ++it) {
    // v = M^-1 A p, (rhat, v)
    double d1[1] = {0.0};
    spmv_stream_c(A, rg, p, prod,
                  [&](int row) { return Ops2{dinv[row], rh[row]}; },
                  [&](int row, double ap, const Ops2& o) {
                    const double vi = o.a * ap;
                    v[row] = vi;
                    d1[0] += o.b * vi;
                  });
    if (a.prof && it == 2 && (threadIdx.x & 31) == 0) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      const int gw_ = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
      if (gw_ < 8192) g_prof_warp[gw_] = tt;
    }
    stamp(a.prof, np);
    gsum<1, false>(a, bar, d1, phase, sh);
    stamp(a.prof, np);
    if (d1[0] == 0.0) { leave(bar, a, it - 1, -1, rn, bn); return; }
    const double alpha = rho / d1[0];
    double d2[1] = {0.0};
    // partitioned run: interface entries first, so their P2P stores are acknowledged while the rest of the
    // pass runs (the barrier's system-scope fence then finds nothing in flight)
    auto upd_s = [&](int i, bool iface) {
      const double si = r[i] - alpha * v[i];
      s[i] = si;
      if (iface) push_iface(a.dd, 4, i, si);
      d2[0] += si * si;
    };
    if (DIST)
      for (int k = gt; k < a.dd.nsend; k += gs) {
        const int i = a.dd.s_local[k];
        if (k == 0 || a.dd.s_local[k - 1] != i) upd_s(i, true);
      }
    for (int i = gt; i < n; i += gs) {
      if (DIST && (mask[i] & (MASK_GHOST | MASK_IFACE))) continue;
      upd_s(i, false);
    }
    stamp(a.prof, np);
    gsum<1>(a, bar, d2, phase, sh);
    stamp(a.prof, np);
    const double ss = d2[0];
    if (sqrt(ss) <= tol) {  // early exit on the half step (KSPBCGS does the same)
      for (int i = gt; i < n; i += gs)
        if (mask == nullptr || !(mask[i] & MASK_GHOST)) a.x[i] += alpha * p[i];
      leave(bar, a, it, 1, sqrt(ss), bn);
      return;
    }
    // t = M^-1 A s and the four scalars of the second half step
    double d3[4] = {0.0, 0.0, 0.0, 0.0};
    spmv_stream_c(A, rg, s, prod,
                  [&](int row) { return Ops3{dinv[row], s[row], rh[row]}; },
                  [&](int row, double as, const Ops3& o) {
                    const double ti = o.a * as;
                    t[row] = ti;
                    d3[0] += ti * o.b; d3[1] += ti * ti; d3[2] += o.c * o.b; d3[3] += o.c * ti;
                  });
    stamp(a.prof, np);
    gsum<4, false>(a, bar, d3, phase, sh);
    stamp(a.prof, np);
    if (d3[1] == 0.0) { leave(bar, a, it, -1, sqrt(ss), bn); return; }
    const double omega = d3[0] / d3[1];
    double rho_new = d3[2] - omega * d3[3];
    double rn2 = ss - 2.0 * omega * d3[0] + omega * omega * d3[1];  // ||s - omega t||^2
    rn = sqrt(fmax(rn2, 0.0));
    if (!(rn2 == rn2) || rn > a.dtol * bn) { leave(bar, a, it, -1, rn, bn); return; }
    // (rhat, r) ~ 0 or omega ~ 0: the recurrence has broken down (e.g. a residual supported on
    // Dirichlet rows only is annihilated in one step, leaving rhat orthogonal to r).  Restart with
    // rhat = r, the standard cure.
    const bool restart = fabs(rho_new) <= 1e-14 * rhn * rn || omega == 0.0;
    if (restart && ++restarts > 50) { leave(bar, a, it, -1, rn, bn); return; }
    const double beta = restart ? 0.0 : (rho_new / rho) * (alpha / omega);
    // x += alpha p + omega s;  r = s - omega t;  p = r + beta (p - omega v)   [restart: rhat = p = r]
    const bool check = rn <= 2.0 * tol;  // confirm with a true norm
    double d4[1] = {0.0};
    auto upd_p = [&](int i, bool iface) {
      const double pi = p[i];
      a.x[i] += alpha * pi + omega * s[i];
      const double ri = s[i] - omega * t[i];
      r[i] = ri;
      const double pn = ri + beta * (pi - omega * v[i]);
      p[i] = pn;
      if (iface) push_iface(a.dd, 2, i, pn);
      if (restart) rh[i] = ri;
      d4[0] += ri * ri;
    };
    if (DIST)
      for (int k = gt; k < a.dd.nsend; k += gs) {
        const int i = a.dd.s_local[k];
        if (k == 0 || a.dd.s_local[k - 1] != i) upd_p(i, true);
      }
    for (int i = gt; i < n; i += gs) {
      if (DIST && (mask[i] & (MASK_GHOST | MASK_IFACE))) continue;
      upd_p(i, false);
    }
    stamp(a.prof, np);
    if (check || restart) {
      gsum<1>(a, bar, d4, phase, sh);
      rn = sqrt(d4[0]);
      if (rn <= tol) { leave(bar, a, it, 1, rn, bn); return; }
      if (restart) { rho_new = d4[0]; rhn = rn; }
    } else {
      bar.sync();
    }
    stamp(a.prof, np);
    rho = rho_new;
  }
  leave(bar, a, a.maxit, 0, rn, bn);
}

// ---- BiCGStab on the node-block layout (fx_nb_host.h / fx_nb.cuh) ------------------------------------------------
// Same recurrence, reductions and exits as k_bicgstab; what differs is the data layout: the work vectors live in the
// internal node-major numbering (length N = nodes x bp), the matrix is the gathered block copy (one column index per
// bs x bs node block, one aligned x-gather per block), and the set-up is a plain value gather -- no compaction scans.
// x is permuted in at the start and out at every exit; eliminated rows (W's DG0 DEVSS rows) never enter the internal
// vectors and are recovered by leave()'s back-substitution on the external arrays.
// DIST: mesh-partitioned run -- ghost block rows are empty, the owner of an interface entry of p / s stores it straight
// into the ghost slot of every neighbour's copy (P2P, internal indices exchanged at set-up: fx_plan_set_halo_solver_index),
// and every grid barrier closes across the GPUs (GridBarT<true>), exactly as in k_bicgstab<.., DIST>.
template <int BS, bool DUAL, bool DIST>
__global__ void __launch_bounds__(NB_WARPS * 32, 1) k_bicgstab_nb(const __grid_constant__ KArgs a) {
  constexpr int BP = nb_bp(BS);
  __shared__ double sh[5 * 32];
  extern __shared__ __align__(16) double prod_all[];
  const NbView& B = a.nb;
  const int N = B.N;
  static_assert(!(DIST && DUAL), "the partitioned kernel keeps the s vector (its interface entries are pushed)");
  const size_t ld = DIST ? (size_t)a.dd.nmax : (((size_t)N + 3) & ~(size_t)3);
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const unsigned char* mk = DIST ? B.imask : nullptr;
  auto ghost = [&](int i) { return DIST && (mk[i] & MASK_GHOST); };
  auto push = [&](int j, int i, double val) {  // interface entry i of work vector j -> the neighbours' ghost slots
    if (DIST && (mk[i] & MASK_IFACE)) {
      int q = B.s_first[i];
      for (;;) {
        const int pe = a.dd.s_peer[q];
        a.dd.region[pe & (DIST_MAXW - 1)][(long long)j * a.dd.nmax + B.s_remote[q]] = val;
        if (pe & DIST_LAST) break;
        ++q;
      }
    }
  };
  // one (dof, destination) pair: consecutive pairs address consecutive ghost slots of the destination (nb_build numbers
  // ghost nodes per owner in the owner's order), so a warp's stores coalesce into a few NVLink packets
  auto push_pair = [&](int j, int q, double val) {
    a.dd.region[a.dd.s_peer[q] & (DIST_MAXW - 1)][(long long)j * a.dd.nmax + B.s_remote[q]] = val;
  };
  double* r = a.w;
  double* rh = a.w + ld;
  double* p = a.w + 2 * ld;
  double* v = a.w + 3 * ld;
  double* s = a.w + 4 * ld;
  double* t = a.w + 5 * ld;
  double* dinv = a.w + 6 * ld;
  double* xi = a.w + 7 * ld;
  double* prod = prod_all + (size_t)(threadIdx.x >> 5) * (nb_chunk(BS) * BS);
  const int gw = blockIdx.x * NB_WARPS + (threadIdx.x >> 5);
  const int k0 = B.wrange[gw * (NB_SLOTS / NB_WARPS)], k1 = B.wrange[(gw + 1) * (NB_SLOTS / NB_WARPS)];
  // the warp OWNS the internal rows of its chunks (a contiguous range): it computes them in the SpMV epilogues and
  // updates them in the vector passes, so a pass is one batch of coalesced loads per lane -- no grid-stride loop
  int i0 = 0, i1 = 0;
  if (k0 < k1) {
    const int4 da = B.desc[k0], db = B.desc[k1 - 1];
    i0 = da.x * BP; i1 = (db.x + db.y) * BP;
  }
  const unsigned long long pol = nb_make_policy(B.l2policy);
  int phase = 0;
  __shared__ double sh_x[DIST ? DIST_MAXW * RED_SLOTS : 1];
  GridBarT<DIST> bar;
  bar.init(a.bar, &a.dd, sh_x);
  auto out = [&](int it, int conv, double rn, double bn) {  // every exit: x back to the external numbering
    for (int ib = i0 + lane; ib < i1; ib += 32 * 4) {       // (4 independent index -> store chains per lane)
      int e[4];
      double xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = ib + 32 * u;
        e[u] = i < i1 ? B.ext_of_int[i] : -1;
        xv[u] = i < i1 ? xi[i] : 0.0;
        if (i < i1 && ghost(i)) e[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e[u] >= 0) a.x[e[u]] = xv[u];
    }
    leave(bar, a, it, conv, rn, bn);
  };
#define NB_GSUM(NV, PUSHED, v)                                             \
  do {                                                                     \
    if constexpr (DIST) gsum<NV, PUSHED>(a, bar, v, phase, sh);            \
    else grid_sum_fast<NV>(bar, v, a.red, phase, sh);                      \
  } while (0)
  double acc[3] = {0.0, 0.0, 0.0};
  acc[2] = (double)nb_gather_values<BS>(B, a.A.val);
  for (int ib = gt; ib < N; ib += 4 * gs) {  // x into the internal numbering (4 independent gathers per thread)
    int e[4];
    double xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = ib + u * gs;
      e[u] = i < N ? B.ext_of_int[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = e[u] >= 0 ? a.x[e[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = ib + u * gs;
      if (i < N) {
        xi[i] = xv[u];
        if (ghost(i)) { dinv[i] = 0.0; r[i] = 0.0; rh[i] = 0.0; v[i] = 0.0; t[i] = 0.0; }  // ghost rows have no chunk
      }
    }
  }
  bar.sync();
  const unsigned pm = B.planemask ? __ldcg(B.pmask) : 0xffffffffu;
  spmv_nb<BS>(B, k0, k1, xi, prod, pol, pm,
              [&](int i) {
                const int e = B.ext_of_int[i];
                return Ops2{nb_diag_inv<BS>(B, i, a.pc), e >= 0 ? a.b[e] : 0.0};
              },
              [&](int i, double ax, const Ops2& o) {
                const double ri = o.a * (o.b - ax), bi = o.a * o.b;
                dinv[i] = o.a; r[i] = ri; rh[i] = ri; v[i] = 0.0;
                if (!ghost(i)) { p[i] = ri; push(2, i, ri); }  // ghost entries of p arrive from their owners
                acc[0] += ri * ri; acc[1] += bi * bi;
              });
  NB_GSUM(3, true, acc);
  if (acc[2] != 0.0) { finish(a, 0, -3, 0.0, 0.0); bar.save(); return; }  // A[active, eliminated] != 0: not block triangular
  double rn = sqrt(acc[0]);
  const double bn = sqrt(acc[1]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(rn == rn)) { out(0, -1, rn, bn); return; }
  if (rn <= tol) { out(0, 1, rn, bn); return; }
  double rho = acc[0];
  double rhn = rn;
  int restarts = 0;
  int np = 0;
  stamp(a.prof, np);
  constexpr int PU = FX_NB_PU;  // pass unroll: 32 * PU owned entries per batch of loads
  for (int it = 1; it <= a.maxit; ++it) {
    double d1[1] = {0.0};
    spmv_nb<BS>(B, k0, k1, p, prod, pol, pm, [&](int i) { return Ops2{dinv[i], rh[i]}; },
                [&](int i, double ap, const Ops2& o) {
                  const double vi = o.a * ap;
                  v[i] = vi;
                  d1[0] += o.b * vi;
                });
    stamp(a.prof, np);
    NB_GSUM(1, false, d1);
    stamp(a.prof, np);
    if (d1[0] == 0.0) { out(it - 1, -1, rn, bn); return; }
    const double alpha = rho / d1[0];
    double ss;
    double d3[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (DUAL) {
      // s = r - alpha v is never stored: the product gathers r and v and forms it on the fly, the owner of row i
      // forms s_i for the dot products.  One barrier and one vector pass fewer; the half-step test moves behind
      // the product (same outcome, one product later).
      double d5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      spmv_nb<BS, true>(B, k0, k1, r, prod, pol, pm, [&](int i) { return Ops4{dinv[i], r[i] - alpha * v[i], rh[i], 0.0}; },
                        [&](int i, double as, const Ops4& o) {
                          const double ti = o.a * as;
                          t[i] = ti;
                          d5[0] += ti * o.b; d5[1] += ti * ti; d5[2] += o.c * o.b; d5[3] += o.c * ti; d5[4] += o.b * o.b;
                        }, v, -alpha);
      stamp(a.prof, np); stamp(a.prof, np); stamp(a.prof, np);
      NB_GSUM(5, false, d5);
      stamp(a.prof, np);
      ss = d5[4];
      d3[0] = d5[0]; d3[1] = d5[1]; d3[2] = d5[2]; d3[3] = d5[3];
      if (sqrt(ss) <= tol) {
        for (int i = i0 + lane; i < i1; i += 32) xi[i] += alpha * p[i];
        out(it, 1, sqrt(ss), bn);
        return;
      }
    } else {
    double d2[1] = {0.0};
    if constexpr (DIST) {
      // interface entries: spread over the WHOLE grid (one (dof, destination) pair per thread) and first, so the P2P
      // stores are in flight while the owned ranges are updated -- a warp's owned range would put the ~10^3 stores
      // of a cut line on a handful of warps (measured: 17 us per pass instead of 3)
      for (int q = gt; q < a.dd.nsend; q += gs) {
        const int i = B.s_local[q];
        if (i >= 0 && (q == 0 || B.s_local[q - 1] != i)) {  // first pair of the dof: update + all its destinations
          const double si = r[i] - alpha * v[i];
          for (int qq = q; qq < a.dd.nsend && B.s_local[qq] == i; ++qq) push_pair(4, qq, si);
          s[i] = si;
          d2[0] += si * si;
        }
      }
    }
    for (int base = i0 + lane; base < i1; base += 32 * PU) {
      double rr[PU], vv[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int i = base + 32 * u;
        rr[u] = i < i1 ? r[i] : 0.0;
        vv[u] = i < i1 ? v[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int i = base + 32 * u;
        const bool mine = i < i1 && !(DIST && (mk[i] & (MASK_GHOST | MASK_IFACE)));
        const double si = mine ? rr[u] - alpha * vv[u] : 0.0;
        if (mine) s[i] = si;
        d2[0] += si * si;
      }
    }
    stamp(a.prof, np);
    NB_GSUM(1, true, d2);
    stamp(a.prof, np);
    ss = d2[0];
    if (sqrt(ss) <= tol) {  // early exit on the half step (KSPBCGS does the same)
      for (int i = i0 + lane; i < i1; i += 32)
        if (!ghost(i)) xi[i] += alpha * p[i];
      out(it, 1, sqrt(ss), bn);
      return;
    }
    spmv_nb<BS>(B, k0, k1, s, prod, pol, pm, [&](int i) { return Ops3{dinv[i], s[i], rh[i]}; },
                [&](int i, double as, const Ops3& o) {
                  const double ti = o.a * as;
                  t[i] = ti;
                  d3[0] += ti * o.b; d3[1] += ti * ti; d3[2] += o.c * o.b; d3[3] += o.c * ti;
                });
    stamp(a.prof, np);
    NB_GSUM(4, false, d3);
    stamp(a.prof, np);
    }
    if (d3[1] == 0.0) { out(it, -1, sqrt(ss), bn); return; }
    const double omega = d3[0] / d3[1];
    double rho_new = d3[2] - omega * d3[3];
    const double rn2 = ss - 2.0 * omega * d3[0] + omega * omega * d3[1];  // ||s - omega t||^2
    rn = sqrt(fmax(rn2, 0.0));
    if (!(rn2 == rn2) || rn > a.dtol * bn) { out(it, -1, rn, bn); return; }
    const bool restart = fabs(rho_new) <= 1e-14 * rhn * rn || omega == 0.0;
    if (restart && ++restarts > 50) { out(it, -1, rn, bn); return; }
    const double beta = restart ? 0.0 : (rho_new / rho) * (alpha / omega);
    const bool check = rn <= 2.0 * tol;  // confirm with a true norm
    double d4[1] = {0.0};
    if constexpr (DIST) {
      for (int q = gt; q < a.dd.nsend; q += gs) {
        const int i = B.s_local[q];
        if (i >= 0 && (q == 0 || B.s_local[q - 1] != i)) {
          const double pi = p[i], si = s[i];
          const double ri = si - omega * t[i];
          const double pn = ri + beta * (pi - omega * v[i]);
          for (int qq = q; qq < a.dd.nsend && B.s_local[qq] == i; ++qq) push_pair(2, qq, pn);
          xi[i] += alpha * pi + omega * si;
          r[i] = ri;
          p[i] = pn;
          if (restart) rh[i] = ri;
          d4[0] += ri * ri;
        }
      }
    }
    for (int base = i0 + lane; base < i1; base += 32 * PU) {
      double pp[PU], sv[PU], tt[PU], vv[PU], xx[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int i = base + 32 * u;
        const bool ok = i < i1;
        pp[u] = ok ? p[i] : 0.0; sv[u] = ok ? (DUAL ? r[i] : s[i]) : 0.0; tt[u] = ok ? t[i] : 0.0;
        vv[u] = ok ? v[i] : 0.0; xx[u] = ok ? xi[i] : 0.0;
        if (DUAL) sv[u] -= alpha * vv[u];
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const int i = base + 32 * u;
        const bool own = i < i1 && !(DIST && (mk[i] & (MASK_GHOST | MASK_IFACE)));
        const double ri = own ? sv[u] - omega * tt[u] : 0.0;
        if (own) {
          xi[i] = xx[u] + alpha * pp[u] + omega * sv[u];
          r[i] = ri;
          p[i] = ri + beta * (pp[u] - omega * vv[u]);
          if (restart) rh[i] = ri;
        }
        d4[0] += ri * ri;
      }
    }
    stamp(a.prof, np);
    if (check || restart) {
      NB_GSUM(1, true, d4);
      rn = sqrt(d4[0]);
      if (rn <= tol) { out(it, 1, rn, bn); return; }
      if (restart) { rho_new = d4[0]; rhn = rn; }
    } else {
      bar.sync();
    }
    stamp(a.prof, np);
    rho = rho_new;
  }
  out(a.maxit, 0, rn, bn);
#undef NB_GSUM
}

// ---- CG (Chronopoulos-Gear) on the node-block layout ------------------------------------------------------------
// Same recurrence and stopping rule as k_cg (u = M^-1 r, w = A u, one fused reduction per iteration, two grid
// barriers); data layout and set-up as k_bicgstab_nb.  Jacobi / none only: the two-level preconditioner stays with k_cg
// and k_pcg_res.  Serves the consistent-mass projections of the loops (project(., Zc): 29 iterations per step on a 3 x 3
// node-block matrix) and SPD systems too large or too tightly toleranced for the resident kernel.
template <int BS>
__global__ void __launch_bounds__(NB_WARPS_CG * 32, 1) k_cg_nb(const __grid_constant__ KArgs a) {
  constexpr int BP = nb_bp(BS);
  __shared__ double sh[4 * 32];
  extern __shared__ __align__(16) double prod_all[];
  const NbView& B = a.nb;
  const int N = B.N;
  const size_t ld = ((size_t)N + 3) & ~(size_t)3;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  double* r = a.w;
  double* u = a.w + ld;
  double* w = a.w + 2 * ld;
  double* p = a.w + 3 * ld;
  double* s = a.w + 4 * ld;
  double* dinv = a.w + 5 * ld;
  double* xi = a.w + 6 * ld;
  double* prod = prod_all + (size_t)(threadIdx.x >> 5) * (nb_chunk(BS) * BS);
  const int gw = blockIdx.x * NB_WARPS_CG + (threadIdx.x >> 5);
  const int k0 = B.wrange[gw * (NB_SLOTS / NB_WARPS_CG)], k1 = B.wrange[(gw + 1) * (NB_SLOTS / NB_WARPS_CG)];
  int i0 = 0, i1 = 0;
  if (k0 < k1) {
    const int4 da = B.desc[k0], db = B.desc[k1 - 1];
    i0 = da.x * BP; i1 = (db.x + db.y) * BP;
  }
  const unsigned long long pol = nb_make_policy(B.l2policy);
  int phase = 0;
  GridBar bar;
  bar.init(a.bar);
  auto out = [&](int it, int conv, double rn, double bn) {
    for (int i = i0 + lane; i < i1; i += 32) {
      const int e = B.ext_of_int[i];
      if (e >= 0) a.x[e] = xi[i];
    }
    leave(bar, a, it, conv, rn, bn);
  };
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  acc[3] = (double)nb_gather_values<BS>(B, a.A.val);
  for (int i = gt; i < N; i += gs) {
    const int e = B.ext_of_int[i];
    xi[i] = e >= 0 ? a.x[e] : 0.0;
  }
  bar.sync();
  const unsigned pm = B.planemask ? __ldcg(B.pmask) : 0xffffffffu;
  spmv_nb<BS>(B, k0, k1, xi, prod, pol, pm,
              [&](int i) {
                const int e = B.ext_of_int[i];
                return Ops2{nb_diag_inv<BS>(B, i, a.pc), e >= 0 ? a.b[e] : 0.0};
              },
              [&](int i, double ax, const Ops2& o) {
                const double ri = o.b - ax, ui = o.a * ri, bi = o.a * o.b;
                dinv[i] = o.a; r[i] = ri; u[i] = ui; p[i] = 0.0; s[i] = 0.0;
                acc[0] += ri * ui; acc[1] += ui * ui; acc[2] += bi * bi;
              });
  grid_sum_fast<4>(bar, acc, a.red, phase, sh);
  if (acc[3] != 0.0) { finish(a, 0, -3, 0.0, 0.0); return; }
  double gamma = acc[0], un = sqrt(acc[1]);
  const double bn = sqrt(acc[2]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(un == un)) { out(0, -1, un, bn); return; }
  if (un <= tol) { out(0, 1, un, bn); return; }
  double d0[1] = {0.0};
  spmv_nb<BS>(B, k0, k1, u, prod, pol, pm, [&](int i) { return Ops2{u[i], 0.0}; },
              [&](int i, double au, const Ops2& o) {
                w[i] = au;
                d0[0] += au * o.a;
              });
  grid_sum_fast<1>(bar, d0, a.red, phase, sh);
  if (!(d0[0] > 0.0)) { out(0, -1, un, bn); return; }
  double alpha = gamma / d0[0], beta = 0.0;
  constexpr int PU = FX_NB_PU;
  for (int it = 1; it <= a.maxit; ++it) {
    double d[3] = {0.0, 0.0, 0.0};
    for (int base = i0 + lane; base < i1; base += 32 * PU) {
      double uu[PU], pp[PU], ww[PU], ss[PU], rr[PU], xx[PU], dd[PU];
#pragma unroll
      for (int q = 0; q < PU; ++q) {
        const int i = base + 32 * q;
        const bool ok = i < i1;
        uu[q] = ok ? u[i] : 0.0; pp[q] = ok ? p[i] : 0.0; ww[q] = ok ? w[i] : 0.0; ss[q] = ok ? s[i] : 0.0;
        rr[q] = ok ? r[i] : 0.0; xx[q] = ok ? xi[i] : 0.0; dd[q] = ok ? dinv[i] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < PU; ++q) {
        const int i = base + 32 * q;
        const double pi = uu[q] + beta * pp[q], si = ww[q] + beta * ss[q];
        const double ri = rr[q] - alpha * si, ui = dd[q] * ri;
        if (i < i1) { p[i] = pi; s[i] = si; xi[i] = xx[q] + alpha * pi; r[i] = ri; u[i] = ui; }
        d[0] += ri * ui; d[1] += ui * ui;
      }
    }
    bar.sync();
    spmv_nb<BS>(B, k0, k1, u, prod, pol, pm, [&](int i) { return Ops2{u[i], 0.0}; },
                [&](int i, double au, const Ops2& o) {
                  w[i] = au;
                  d[2] += au * o.a;
                });
    grid_sum_fast<3>(bar, d, a.red, phase, sh);
    un = sqrt(d[1]);
    if (!(un == un) || un > a.dtol * bn) { out(it, -1, un, bn); return; }
    if (un <= tol) { out(it, 1, un, bn); return; }
    beta = d[0] / gamma;
    const double denom = d[2] - beta * d[0] / alpha;
    if (!(denom > 0.0)) { out(it, -1, un, bn); return; }  // not SPD / breakdown
    alpha = d[0] / denom;
    gamma = d[0];
  }
  out(a.maxit, 0, un, bn);
}

// y = A x through the node-block layout, stand-alone (tests and tools/spmv_bench.py): gathers the values, permutes x in,
// runs `reps` products (a grid barrier between them, as inside a solve) and writes the rows of the node blocks to y
template <int BS>
__global__ void __launch_bounds__(NB_WARPS_CG * 32, 1) k_nb_spmv(const __grid_constant__ KArgs a, int reps) {
  extern __shared__ __align__(16) double prod_all[];
  const NbView& B = a.nb;
  const int N = B.N;
  const size_t ld = ((size_t)N + 3) & ~(size_t)3;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  double* xi = a.w;
  double* yi = a.w + ld;
  double* prod = prod_all + (size_t)(threadIdx.x >> 5) * (nb_chunk(BS) * BS);
  const int gw = blockIdx.x * NB_WARPS_CG + (threadIdx.x >> 5);
  const int k0 = B.wrange[gw * (NB_SLOTS / NB_WARPS_CG)], k1 = B.wrange[(gw + 1) * (NB_SLOTS / NB_WARPS_CG)];
  GridBar bar;
  bar.init(a.bar);
  const unsigned long long pol = nb_make_policy(B.l2policy);
  const int bad = nb_gather_values<BS>(B, a.A.val);
  if (bad) a.info->converged = -3;
  for (int i = gt; i < N; i += gs) {
    const int e = B.ext_of_int[i];
    xi[i] = e >= 0 ? a.b[e] : 0.0;
  }
  bar.sync();
  const unsigned pm = B.planemask ? __ldcg(B.pmask) : 0xffffffffu;
  for (int rep = 0; rep < reps; ++rep) {
    spmv_nb<BS>(B, k0, k1, xi, prod, pol, pm, [&](int i) { return Ops2{0.0, 0.0}; },
                [&](int i, double ax, const Ops2&) { yi[i] = ax; });
    bar.sync();
  }
  for (int i = gt; i < N; i += gs) {
    const int e = B.ext_of_int[i];
    if (e >= 0) a.x[e] = yi[i];
  }
}

__device__ __forceinline__ double* invert_blocked(double* E0, double* E1, int k, GridBar& bar, double* scratch);

// Singular and nearly singular coarse matrices.  Pure-Neumann pressure systems (config 1, incompressible_dgp.py) have
// A 1 = 0, so E = P^T A P is singular along the constant coarse vector; the compressible loops at Ma = 1e-6 (config 3's
// script parameters: mass term 1e-12 / dt) make it singular to ~1e-10, which the unpivoted Gauss-Jordan cannot resolve
// (CG then diverges at n = 384).  Both are detected on the device (largest |row sum| of E against its largest
// diagonal entry; the same verdict in every CTA) and handled by a rank-one shift:
//   1. invert  E^ = E + gamma 1 1^T  (well conditioned);
//   2. undo the shift exactly with Sherman-Morrison:  E^-1 = E^^-1 + mu z z^T,  z = E^^-1 1,  mu = gamma / den,
//      den = 1 - gamma 1^T z, evaluated WITHOUT the cancellation as den = (E 1)^T z / k from the row sums of the
//      unshifted E (E^ z = 1  =>  E z = den 1).
// mu z z^T is symmetric positive semi-definite however inaccurately the tiny denominator is known, so the
// preconditioner stays SPD.  For an exactly singular E the denominator is rounding noise (<= 1e-14): step 2 is skipped
// and E^^-1 acts as the pseudo-inverse on the range of E (all that a consistent right-hand side ever excites).
// flag: 16 doubles (2 used, zeroed before E was assembled) followed by k row sums.  Returns gamma (0: E is regular).
__device__ __forceinline__ double coarse_shift_nullspace(double* E0, int k, GridBar& bar, double* flag) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
  double mrs = 0.0, mdg = 0.0;
  for (int g = gt >> 5; g < k; g += gs >> 5) {
    double rs = 0.0;
    for (int h = lane; h < k; h += 32) rs += __ldcg(E0 + (size_t)g * k + h);
    rs = warp_sum(rs);
    if (lane == 0) { flag[16 + g] = rs; mrs = fmax(mrs, fabs(rs)); mdg = fmax(mdg, fabs(__ldcg(E0 + (size_t)g * k + g))); }
  }
  if (lane == 0 && mdg > 0.0) {  // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(flag), (unsigned long long)__double_as_longlong(mrs));
    atomicMax(reinterpret_cast<unsigned long long*>(flag) + 1, (unsigned long long)__double_as_longlong(mdg));
  }
  bar.sync();
  const double rs = __ldcg(flag), dg = __ldcg(flag + 1);
  if (!(rs <= 1e-6 * dg)) return 0.0;
  const double gamma = dg / k;
  for (int i = gt; i < k * k; i += gs) E0[i] += gamma;
  bar.sync();
  return gamma;
}
__device__ __forceinline__ void coarse_unshift(double* Einv, int k, double gamma, GridBar& bar, double* z,
                                               const double* flag, double* sh) {
  if (gamma == 0.0) return;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
  for (int g = gt >> 5; g < k; g += gs >> 5) {
    double rs = 0.0;
    for (int h = lane; h < k; h += 32) rs += __ldcg(Einv + (size_t)g * k + h);
    rs = warp_sum(rs);
    if (lane == 0) z[g] = rs;
  }
  bar.sync();
  double part = 0.0;
  for (int h = threadIdx.x; h < k; h += blockDim.x) part += __ldcg(flag + 16 + h) * __ldcg(z + h);
  const double den = block_sum_bcast(part, sh) / k;  // fixed tree: the same value in every CTA
  if (den > 1e-14) {
    const double mu = gamma / den;
    for (int i = gt; i < k * k; i += gs) Einv[i] += mu * __ldcg(z + i / k) * __ldcg(z + i % k);
  }
  bar.sync();
}

// ---- CG (Chronopoulos-Gear) ---------------------------------------------------------------------------------
// u = M^-1 r, w = A u, p = u + beta p, s = w + beta s (= A p).  One SpMV and ONE fused reduction
// (gamma = (r,u), delta = (w,u), ||u||^2) per iteration: two grid barriers instead of three.
// Monitored norm ||M^-1 r||_2 (PETSc default for KSPCG).
template <int THREADS, bool DIST>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_cg(const __grid_constant__ KArgs a) {
  __shared__ double sh[4 * 33];
  __shared__ int sh_i[64];
  extern __shared__ __align__(16) double prod[];
  const CsrView& A = a.A;
  const int n = A.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  // u (index 1) is the SpMV input: its ghost entries come from the owners' halo pushes (partitioned run)
  const size_t ld = DIST ? (size_t)a.dd.nmax : (size_t)n;
  const unsigned char* mask = DIST ? A.elim : nullptr;
  double* r = a.w;
  double* u = a.w + ld;
  double* w = a.w + 2 * ld;
  double* p = a.w + 3 * ld;
  double* s = a.w + 4 * ld;
  double* dinv = a.w + 5 * ld;
  int phase = 0;
  __shared__ double sh_x[DIST ? DIST_MAXW * RED_SLOTS : 1];
  GridBarT<DIST> bar;
  bar.init(a.bar, &a.dd, sh_x);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int nb = A.nrowblk;
  if (A.blk_target > 0) {
    acc[3] = (double)compact_reblock(A, bar, sh_i, nb);
  } else {
    acc[3] = (double)compact_zeros(A);
    bar.sync();
  }
  const BlockRange rg = balance_partition(A, bar, sh_i, nb);
  // ---- two-level preconditioner M^-1 = D^-1 + P E^-1 P^T, E = P^T A P, for systems too large for the resident kernel
  // (k_pcg_res): the coarse matrix is assembled from the values and inverted across the grid once per solve; every
  // application costs two grid barriers: restriction c = P^T v (shared-memory atomics per CTA, one global reduction
  // per aggregate and CTA) | e = E^-1 c, one warp per aggregate | prolongation.
  const bool coarse = !DIST && a.cz.agg != nullptr;
  const int ck = a.cz.k;
  __shared__ double c_sh[DIST ? 1 : COARSE_KMAX_CG];
  const double* Einv = nullptr;
  int cph = 0;
  double* evec = a.cz.cvec + 3 * COARSE_KMAX_CG;
  // vout += P E^-1 P^T vin; with dsum != null also dsum[0] += (vin, vout), dsum[1] += (vout, vout) on the final values
  auto coarse_apply = [&](const double* vin, double* vout, double* dsum) {
    double* cur = a.cz.cvec + (cph % 3) * COARSE_KMAX_CG;
    for (int g = threadIdx.x; g < ck; g += blockDim.x) c_sh[g] = 0.0;
    __syncthreads();
    for (int i = gt; i < n; i += gs) atomicAdd(&c_sh[a.cz.agg[i]], vin[i]);
    __syncthreads();
    for (int g = threadIdx.x; g < ck; g += blockDim.x) {
      const double v = c_sh[g];
      if (v != 0.0) atomicAdd(cur + g, v);
    }
    if (blockIdx.x == 0)  // the buffer of the next application: nobody reads or adds to it now
      for (int g = threadIdx.x; g < ck; g += blockDim.x) a.cz.cvec[((cph + 1) % 3) * COARSE_KMAX_CG + g] = 0.0;
    bar.sync();
    const int lane = threadIdx.x & 31;
    for (int g = gt >> 5; g < ck; g += gs >> 5) {  // one warp per aggregate (small systems run on a small grid)
      double e = 0.0;
      for (int h = lane; h < ck; h += 32) e += __ldcg(Einv + (size_t)g * ck + h) * __ldcg(cur + h);
      e = warp_sum(e);
      if (lane == 0) evec[g] = e;
    }
    bar.sync();
    for (int g = threadIdx.x; g < ck; g += blockDim.x) c_sh[g] = __ldcg(evec + g);
    __syncthreads();
    for (int i = gt; i < n; i += gs) {
      const double vo = vout[i] + c_sh[a.cz.agg[i]];
      vout[i] = vo;
      if (dsum != nullptr) { dsum[0] += vin[i] * vo; dsum[1] += vo * vo; }
    }
    __syncthreads();
    ++cph;
  };
  if constexpr (!DIST) {
    if (coarse) {
      for (int i = gt; i < ck * ck; i += gs) a.cz.E0[i] = 0.0;
      for (int i = gt; i < 4 * COARSE_KMAX_CG + 2; i += gs) a.cz.cvec[i] = 0.0;  // + the two null-space flags behind them
      bar.sync();
      for (int row = gt; row < n; row += gs) {
        const int ga = a.cz.agg[row];
        for (int j = A.rowptr[row]; j < A.rowptr[row + 1]; ++j) {
          const double v = A.val[j];
          if (v != 0.0) atomicAdd(a.cz.E0 + (size_t)ga * ck + a.cz.agg[A.col[j]], v);
        }
      }
      bar.sync();
      const double shift = coarse_shift_nullspace(a.cz.E0, ck, bar, a.cz.cvec + 4 * COARSE_KMAX_CG);
      double* Ei = invert_blocked(a.cz.E0, a.cz.E1, ck, bar, prod);  // scratch: the SpMV staging area (idle here)
      coarse_unshift(Ei, ck, shift, bar, evec, a.cz.cvec + 4 * COARSE_KMAX_CG, sh);
      Einv = Ei;
    }
  }
  spmv_stream_c(A, rg, a.x, prod,
                [&](int row) { return Ops2{row_diag_inv(A, row, a.pc), a.b[row]}; },
                [&](int row, double ax, const Ops2& o) {
                  const double ri = o.b - ax, ui = o.a * ri, bi = o.a * o.b;
                  dinv[row] = o.a; r[row] = ri; p[row] = 0.0; s[row] = coarse ? bi : 0.0;
                  const unsigned char mk = mask != nullptr ? mask[row] : 0;
                  if (!(mk & MASK_GHOST)) u[row] = ui;
                  if (mk & MASK_IFACE) push_iface(a.dd, 1, row, ui);
                  if (!coarse) { acc[0] += ri * ui; acc[1] += ui * ui; acc[2] += bi * bi; }
                });
  if (coarse) {  // u = M^-1 r and ||M^-1 b|| with the coarse part; s holds D^-1 b meanwhile
    bar.sync();
    double du[2] = {0.0, 0.0}, db[2] = {0.0, 0.0};
    coarse_apply(r, u, du);
    coarse_apply(a.b, s, db);
    for (int i = gt; i < n; i += gs) s[i] = 0.0;
    acc[0] = du[0]; acc[1] = du[1]; acc[2] = db[1];
  }
  gsum<4>(a, bar, acc, phase, sh);
  if (acc[3] != 0.0) { finish(a, 0, -3, 0.0, 0.0); bar.save(); return; }
  double gamma = acc[0], un = sqrt(acc[1]);
  const double bn = sqrt(acc[2]);
  const double tol = fmax(a.rtol * bn, a.atol);
  if (!(un == un)) { leave(bar, a, 0, -1, un, bn); return; }
  if (un <= tol) { leave(bar, a, 0, 1, un, bn); return; }
  double d0[1] = {0.0};
  spmv_stream_c(A, rg, u, prod, [&](int row) { return Ops2{u[row], 0.0}; },
                [&](int row, double au, const Ops2& o) {
                  w[row] = au;
                  d0[0] += au * o.a;
                });
  gsum<1, false>(a, bar, d0, phase, sh);
  if (!(d0[0] > 0.0)) { leave(bar, a, 0, -1, un, bn); return; }
  double alpha = gamma / d0[0], beta = 0.0;
  for (int it = 1; it <= a.maxit; ++it) {
    double d[3] = {0.0, 0.0, 0.0};
    auto upd = [&](int i, bool iface) {
      const double pi = u[i] + beta * p[i];
      const double si = w[i] + beta * s[i];
      p[i] = pi; s[i] = si;
      a.x[i] += alpha * pi;
      const double ri = r[i] - alpha * si;
      r[i] = ri;
      const double ui = dinv[i] * ri;
      u[i] = ui;
      if (iface) push_iface(a.dd, 1, i, ui);
      d[0] += ri * ui;
      d[1] += ui * ui;
    };
    if (DIST)  // interface entries first: their P2P stores are acknowledged while the rest of the pass runs
      for (int k = gt; k < a.dd.nsend; k += gs) {
        const int i = a.dd.s_local[k];
        if (k == 0 || a.dd.s_local[k - 1] != i) upd(i, true);
      }
    if (coarse) {  // u = D^-1 r here, the coarse part and the two sums follow in coarse_apply
      for (int i = gt; i < n; i += gs) {
        const double pi = u[i] + beta * p[i];
        const double si = w[i] + beta * s[i];
        p[i] = pi; s[i] = si;
        a.x[i] += alpha * pi;
        const double ri = r[i] - alpha * si;
        r[i] = ri;
        u[i] = dinv[i] * ri;
      }
      coarse_apply(r, u, d);
    } else {
      for (int i = gt; i < n; i += gs) {
        if (DIST && (mask[i] & (MASK_GHOST | MASK_IFACE))) continue;
        upd(i, false);
      }
    }
    bar.sync();
    spmv_stream_c(A, rg, u, prod, [&](int row) { return Ops2{u[row], 0.0}; },
                  [&](int row, double au, const Ops2& o) {
                    w[row] = au;
                    d[2] += au * o.a;
                  });
    gsum<3, false>(a, bar, d, phase, sh);
    un = sqrt(d[1]);
    if (!(un == un) || un > a.dtol * bn) { leave(bar, a, it, -1, un, bn); return; }
    if (un <= tol) { leave(bar, a, it, 1, un, bn); return; }
    beta = d[0] / gamma;
    const double denom = d[2] - beta * d[0] / alpha;
    if (!(denom > 0.0)) { leave(bar, a, it, -1, un, bn); return; }  // not SPD / breakdown
    alpha = d[0] / denom;
    gamma = d[0];
  }
  leave(bar, a, a.maxit, 0, un, bn);
}

// ---- restarted GMRES -----------------------------------------------------------------------------------------
// Left-preconditioned GMRES(30) with classical Gram-Schmidt, PETSc KSPGMRES's defaults (restart 30,
// monitored norm ||M^-1 r||, which the Givens recurrence delivers for free).  Two grid barriers per
// iteration:
//   * w = M^-1 A v_j and ALL Gram-Schmidt coefficients (w, v_k), k <= j, in the SpMV epilogue (the lane
//     that sums row i owns entry i of every basis vector), ONE fused deterministic reduction of up to 30
//     values;
//   * w -= sum_k h_k v_k and ||w||^2 in one pass, stored UNnormalised as v_{j+1}: the factor 1 / h_{j+1,j}
//     is a scalar every thread knows after the second reduction, so it is folded into the next SpMV
//     epilogue, the coefficients and the final update instead of costing a pass and a barrier of its own.
// The small least-squares problem (Givens rotations, back substitution) is solved redundantly by every CTA
// in shared memory from bitwise identical sums, so all CTAs take the same branches.  A cycle that the
// recurrence reports as converged is confirmed by the true residual at the start of the next one.
constexpr int GMRES_M = 30;
constexpr int GMRES_NRED = 32;

template <int NV>
__device__ __forceinline__ void grid_sum_big(GridBar& bar, double (&v)[NV], double* red, int& phase, double* sh) {
  double* buf = red + (size_t)(phase & 1) * GMRES_NRED * MAX_GRID;
  block_sum_bcast_n<NV>(v, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[i * MAX_GRID + blockIdx.x] = v[i];
  }
  bar.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) t += __ldcg(buf + i * MAX_GRID + j);
    v[i] = t;
  }
  block_sum_bcast_n<NV>(v, sh);
  ++phase;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_gmres(const __grid_constant__ KArgs a) {
  __shared__ double sh[GMRES_NRED * 33];
  __shared__ int sh_i[64];
  __shared__ double H[(GMRES_M + 1) * GMRES_M];  // column j at H + j * (GMRES_M + 1)
  __shared__ double cs[GMRES_M], sn[GMRES_M], g[GMRES_M + 1], scale[GMRES_M + 1], coef[GMRES_M + 1];
  extern __shared__ __align__(16) double prod[];
  const CsrView& A = a.A;
  const int n = A.n;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
  double* dinv = a.w;
  double* w = a.w + (size_t)n;
  double* V = a.basis;
  int phase = 0, phase_big = 0;
  GridBar bar;
  bar.init(a.bar);
  int nb = A.nrowblk;
  double bad[1] = {0.0};
  if (A.blk_target > 0) {
    bad[0] = (double)compact_reblock(A, bar, sh_i, nb);
  } else {
    bad[0] = (double)compact_zeros(A);
    bar.sync();
  }
  const BlockRange rg = balance_partition(A, bar, sh_i, nb);
  double bn = 0.0, tol = 0.0, rn = 0.0;
  int it = 0;
  for (int cycle = 0;; ++cycle) {
    // v_0 = r = M^-1 (b - A x), unnormalised
    double acc[3] = {0.0, 0.0, cycle == 0 ? bad[0] : 0.0};
    spmv_stream_c(A, rg, a.x, prod,
                  [&](int row) { return Ops2{cycle == 0 ? row_diag_inv(A, row, a.pc) : dinv[row], a.b[row]}; },
                  [&](int row, double ax, const Ops2& o) {
                    const double ri = o.a * (o.b - ax), bi = o.a * o.b;
                    if (cycle == 0) dinv[row] = o.a;
                    V[row] = ri;
                    acc[0] += ri * ri; acc[1] += bi * bi;
                  });
    gsum<3>(a, bar, acc, phase, sh);
    if (cycle == 0) {
      if (acc[2] != 0.0) { finish(a, 0, -3, 0.0, 0.0); return; }  // A[active, eliminated] != 0
      bn = sqrt(acc[1]);
      tol = fmax(a.rtol * bn, a.atol);
    }
    rn = sqrt(acc[0]);
    if (!(rn == rn) || rn > a.dtol * bn) { leave(bar, a, it, -1, rn, bn); return; }
    if (rn <= tol) { leave(bar, a, it, 1, rn, bn); return; }
    if (it >= a.maxit) { leave(bar, a, it, 0, rn, bn); return; }
    __syncthreads();
    if (threadIdx.x == 0) { g[0] = rn; scale[0] = 1.0 / rn; }
    __syncthreads();
    int m = 0;  // basis vectors used by this cycle's update
    for (int j = 0; j < GMRES_M && it < a.maxit; ++j) {
      ++it;
      const double* vj = V + (size_t)j * n;
      const double sj = scale[j];
      double hacc[GMRES_NRED];
#pragma unroll
      for (int k = 0; k < GMRES_NRED; ++k) hacc[k] = 0.0;
      spmv_stream_c(A, rg, vj, prod, [&](int row) { return Ops2{dinv[row], 0.0}; },
                    [&](int row, double av, const Ops2& o) {
                      const double wi = o.a * sj * av;
                      w[row] = wi;
#pragma unroll
                      for (int k = 0; k < GMRES_M; ++k)
                        if (k <= j) hacc[k] += wi * V[(size_t)k * n + row];
                    });
      grid_sum_big<GMRES_NRED>(bar, hacc, a.red_big, phase_big, sh);
      // h_k = (w, v_k) = scale_k * (w, v_k raw);  w -= sum_k h_k v_k = sum_k (h_k scale_k) v_k raw
      __syncthreads();
      double* Hj = H + j * (GMRES_M + 1);
#pragma unroll
      for (int k = 0; k < GMRES_M; ++k)
        if (k <= j && threadIdx.x == 0) { Hj[k] = hacc[k] * scale[k]; coef[k] = Hj[k] * scale[k]; }
      __syncthreads();
      double nacc[1] = {0.0};
      double* vn = V + (size_t)(j + 1) * n;
      for (int i = gt; i < n; i += gs) {
        double wi = w[i];
        for (int k = 0; k <= j; ++k) wi -= coef[k] * V[(size_t)k * n + i];
        vn[i] = wi;
        nacc[0] += wi * wi;
      }
      gsum<1>(a, bar, nacc, phase, sh);
      const double hn = sqrt(nacc[0]);
      __syncthreads();
      if (threadIdx.x == 0) {
        // previous rotations on the new column, then the rotation that annihilates h_{j+1,j}
        for (int k = 0; k < j; ++k) {
          const double t0 = cs[k] * Hj[k] + sn[k] * Hj[k + 1];
          Hj[k + 1] = -sn[k] * Hj[k] + cs[k] * Hj[k + 1];
          Hj[k] = t0;
        }
        const double den = sqrt(Hj[j] * Hj[j] + hn * hn);
        cs[j] = den > 0.0 ? Hj[j] / den : 1.0;
        sn[j] = den > 0.0 ? hn / den : 0.0;
        Hj[j] = den;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
        scale[j + 1] = hn > 0.0 ? 1.0 / hn : 0.0;
      }
      __syncthreads();
      m = j + 1;
      rn = fabs(g[j + 1]);
      if (!(rn == rn) || rn <= tol || hn == 0.0) break;
    }
    // y = R^-1 g (back substitution), x += sum_k y_k v_k = sum_k (y_k scale_k) v_k raw
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = m - 1; k >= 0; --k) {
        double t0 = g[k];
        for (int l = k + 1; l < m; ++l) t0 -= H[l * (GMRES_M + 1) + k] * coef[l];
        const double d = H[k * (GMRES_M + 1) + k];
        coef[k] = d != 0.0 ? t0 / d : 0.0;  // y_k
      }
      for (int k = 0; k < m; ++k) coef[k] *= scale[k];
    }
    __syncthreads();
    for (int i = gt; i < n; i += gs) {
      double xi = a.x[i];
      for (int k = 0; k < m; ++k) xi += coef[k] * V[(size_t)k * n + i];
      a.x[i] = xi;
    }
    if (!(rn == rn)) { leave(bar, a, it, -1, rn, bn); return; }
    bar.sync();  // x complete before the next cycle's residual
  }
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
constexpr int PCG_MAXR = 4;
// per CTA: R resident row blocks per warp (`slice` values + columns each) and one staging slice per warp.
// P1 row blocks are row-limited (32 rows x <= 9 entries), so `slice` is ~288 rather than SPMV_NNZB = 512:
// that is what lets R = 4 blocks per warp (148 k pressure rows, n = 384) stay resident.
__host__ __device__ constexpr size_t pcg_smem_bytes(int R, int slice) {
  return (size_t)LA_WARPS * ((size_t)R * slice * (sizeof(double) + sizeof(int)) + slice * sizeof(double));
}

// shared memory of the coarse correction: rows of E^-1 for the CTA's aggregates, the coarse vector, per-CTA
// partial sums, the corrections, the aggregate list and the aggregate -> slot map
constexpr int COARSE_LD = COARSE_KMAX + 8;  // row stride of the resident E^-1 rows (bank-conflict-free dots)
__host__ __device__ constexpr size_t pcg_coarse_smem_bytes(int nlmax) {
  return sizeof(double) * ((size_t)nlmax * COARSE_LD + COARSE_KMAX + 2 * COARSE_NLMAX) +
         sizeof(int) * (COARSE_NLMAX + COARSE_KMAX + 4);
}

// In-place blocked Gauss-Jordan inversion of the SPD k x k matrix E0 (row major) across the grid:
// block size 32, K = k/32 steps, step p rewrites every 32x32 tile from the OLD matrix
//   (p,p) -> P^-1   (p,j) -> P^-1 A_pj   (i,p) -> -A_ip P^-1   (i,j) -> A_ij - A_ip P^-1 A_pj,   P = A_pp
// into the ping-pong partner, so one grid barrier per step suffices.  Every CTA that owns tiles inverts
// the pivot block itself (one warp, shared memory).  scratch: 4 x 32 x 33 doubles.  Returns the buffer
// holding the inverse.
__device__ __forceinline__ double* invert_blocked(double* E0, double* E1, int k, GridBar& bar, double* scratch) {
  const int K = k >> 5, tiles = K * K, tid = threadIdx.x, lane = threadIdx.x & 31;
  double* sP = scratch;
  double* sA = scratch + 32 * 33;
  double* sB = scratch + 2 * 32 * 33;
  double* sT = scratch + 3 * 32 * 33;
  double* Ein = E0;
  double* Eout = E1;
  for (int p = 0; p < K; ++p) {
    if ((int)blockIdx.x < tiles) {
      for (int e = tid; e < 1024; e += blockDim.x) {
        const int r = e >> 5, c = e & 31;
        sP[r * 33 + c] = __ldcg(Ein + (size_t)(p * 32 + r) * k + p * 32 + c);
      }
      __syncthreads();
      if (tid < 32) {  // Gauss-Jordan without pivoting (SPD), lane = row
        for (int c = 0; c < 32; ++c) {
          const double pinv = 1.0 / sP[c * 33 + c];
          __syncwarp();
          double pv = sP[c * 33 + lane];
          pv = (lane == c) ? pinv : pv * pinv;
          __syncwarp();
          sP[c * 33 + lane] = pv;
          __syncwarp();
          if (lane != c) {
            const double f = sP[lane * 33 + c];
            for (int j = 0; j < 32; ++j) {
              const double base = (j == c) ? 0.0 : sP[lane * 33 + j];
              sP[lane * 33 + j] = base - f * sP[c * 33 + j];
            }
          }
          __syncwarp();
        }
      }
      __syncthreads();
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int i = t / K, j = t % K;
        for (int e = tid; e < 1024; e += blockDim.x) {
          const int r = e >> 5, c = e & 31;
          if (i != p) sA[r * 33 + c] = __ldcg(Ein + (size_t)(i * 32 + r) * k + p * 32 + c);
          if (j != p) sB[r * 33 + c] = __ldcg(Ein + (size_t)(p * 32 + r) * k + j * 32 + c);
        }
        __syncthreads();
        if (i != p && j != p) {
          for (int e = tid; e < 1024; e += blockDim.x) {
            const int r = e >> 5, c = e & 31;
            double acc = 0.0;
            for (int q = 0; q < 32; ++q) acc += sP[r * 33 + q] * sB[q * 33 + c];
            sT[r * 33 + c] = acc;
          }
          __syncthreads();
        }
        for (int e = tid; e < 1024; e += blockDim.x) {
          const int r = e >> 5, c = e & 31;
          double out;
          if (i == p && j == p) {
            out = sP[r * 33 + c];
          } else if (i == p) {
            double acc = 0.0;
            for (int q = 0; q < 32; ++q) acc += sP[r * 33 + q] * sB[q * 33 + c];
            out = acc;
          } else if (j == p) {
            double acc = 0.0;
            for (int q = 0; q < 32; ++q) acc += sA[r * 33 + q] * sP[q * 33 + c];
            out = -acc;
          } else {
            double acc = 0.0;
            for (int q = 0; q < 32; ++q) acc += sA[r * 33 + q] * sT[q * 33 + c];
            out = __ldcg(Ein + (size_t)(i * 32 + r) * k + j * 32 + c) - acc;
          }
          Eout[(size_t)(i * 32 + r) * k + j * 32 + c] = out;
        }
        __syncthreads();
      }
    }
    bar.sync();
    double* tmp = Ein; Ein = Eout; Eout = tmp;
  }
  return Ein;
}

template <int R>
__global__ void __launch_bounds__(LA_THREADS, 1) k_pcg_res(const __grid_constant__ KArgs a) {
  GridBar bar;
  bar.init(a.bar, &a.dd);
  __shared__ double sh[3 * 33];
  extern __shared__ __align__(16) unsigned char dyn[];
  const CsrView& A = a.A;
  const size_t n = (size_t)A.n;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * LA_WARPS + wib;
  // per-warp shared slices
  const int SL = a.cz.slice;
  double* sval = reinterpret_cast<double*>(dyn) + (size_t)wib * R * SL;
  double* prod = reinterpret_cast<double*>(dyn) + (size_t)LA_WARPS * R * SL + (size_t)wib * SL;
  int* scol = reinterpret_cast<int*>(dyn + (size_t)LA_WARPS * (R + 1) * SL * sizeof(double)) + (size_t)wib * R * SL;
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
    const int k = gw * R + rr;  // contiguous blocks per warp and CTA: few aggregates per CTA
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
        my_dinv[rr] = (a.pc != FX_PC_NONE && dg != 0.0) ? 1.0 / dg : 1.0;
        for (int j = b; j < e; ++j) {
          const double v = A.val[j];
          if (v != 0.0) { sval[rr * SL + dst] = v; scol[rr * SL + dst] = A.col[j]; ++dst; }
        }
      }
    }
  }
  __syncwarp();
  // ---- two-level preconditioner M^-1 = D^-1 + P E^-1 P^T, E = P^T A P (aggregates a.cz.agg) --------------
  const bool coarse = a.cz.agg != nullptr;
  const int ck = a.cz.k;
  const int NLMAX = a.cz.nlmax;
  double* einv_s = reinterpret_cast<double*>(dyn + pcg_smem_bytes(R, SL));  // [nl][ck]
  double* c_s = einv_s + (size_t)NLMAX * COARSE_LD;                          // [ck]
  double* cpart = c_s + COARSE_KMAX;                                     // [nl]
  double* e_s = cpart + COARSE_NLMAX;                                    // [nl]
  int* list = reinterpret_cast<int*>(e_s + COARSE_NLMAX);                // [nl]
  int* slot_of = list + COARSE_NLMAX;                                    // [ck]
  int* nl_s = slot_of + COARSE_KMAX;
  int my_slot[R], my_agg[R], nl = 0, cph = 0;
#pragma unroll
  for (int rr = 0; rr < R; ++rr) { my_slot[rr] = 0; my_agg[rr] = 0; }
  if (coarse) {
    for (int i = threadIdx.x; i < ck; i += blockDim.x) slot_of[i] = -1;
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (my_row[rr] >= 0) { my_agg[rr] = a.cz.agg[my_row[rr]]; slot_of[my_agg[rr]] = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cnt = 0;
      for (int g = 0; g < ck; ++g)
        if (slot_of[g] == 0) {
          slot_of[g] = min(cnt, NLMAX - 1);  // the host checked cnt <= NLMAX
          if (cnt < NLMAX) list[cnt] = g;
          ++cnt;
        }
      nl_s[0] = min(cnt, NLMAX);
    }
    __syncthreads();
    nl = nl_s[0];
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (my_row[rr] >= 0) my_slot[rr] = slot_of[my_agg[rr]];
    // E = P^T A P from the resident slices
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ck * ck; i += gridDim.x * blockDim.x) a.cz.E0[i] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * COARSE_KMAX; i += gridDim.x * blockDim.x) a.cz.cvec[i] = 0.0;
    if (blockIdx.x == 0 && threadIdx.x < 2) a.cz.cvec[4 * COARSE_KMAX_CG + threadIdx.x] = 0.0;  // null-space flags
    bar.sync();
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (my_row[rr] >= 0)
        for (int j = my_s0[rr]; j < my_s1[rr]; ++j)
          atomicAdd(a.cz.E0 + (size_t)my_agg[rr] * ck + a.cz.agg[scol[rr * SL + j]], sval[rr * SL + j]);
    bar.sync();
    const double shift = coarse_shift_nullspace(a.cz.E0, ck, bar, a.cz.cvec + 4 * COARSE_KMAX_CG);
    double* Einv = invert_blocked(a.cz.E0, a.cz.E1, ck, bar, einv_s);
    coarse_unshift(Einv, ck, shift, bar, a.cz.cvec + 3 * COARSE_KMAX_CG, a.cz.cvec + 4 * COARSE_KMAX_CG, sh);
    for (int sidx = 0; sidx < nl; ++sidx)
      for (int h = threadIdx.x; h < ck; h += blockDim.x)
        einv_s[(size_t)sidx * COARSE_LD + h] = __ldcg(Einv + (size_t)list[sidx] * ck + h);
    __syncthreads();
  }
  // restriction: add this CTA's part of P^T v to the current coarse right-hand side (before a grid barrier).
  // A warp's 32 consecutive rows fall into a few aggregates: lanes of one aggregate are summed with
  // shuffles and their first lane issues ONE global reduction.
  auto coarse_restrict = [&](const double (&v)[R]) {
    double* cur = a.cz.cvec + (cph % 3) * COARSE_KMAX;
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int g = (my_row[rr] >= 0) ? my_agg[rr] : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, g);
      double sum = 0.0;
#pragma unroll 8
      for (int src = 0; src < 32; ++src) {
        const double o = __shfl_sync(0xffffffffu, v[rr], src);
        if ((peers >> src) & 1u) sum += o;
      }
      if (g >= 0 && lane == __ffs(peers) - 1) atomicAdd(cur + g, sum);
    }
    if (blockIdx.x == 0)  // clear the buffer of the next application (nobody reads or adds to it now)
      for (int i = threadIdx.x; i < ck; i += blockDim.x) a.cz.cvec[((cph + 1) % 3) * COARSE_KMAX + i] = 0.0;
  };
  // after that barrier: e[rr] = (E^-1 P^T v)[aggregate of my row]
  auto coarse_correct = [&](double (&e)[R]) {
    const double* cur = a.cz.cvec + (cph % 3) * COARSE_KMAX;
    for (int i = threadIdx.x; i < ck; i += blockDim.x) c_s[i] = __ldcg(cur + i);
    __syncthreads();
    const int seg = threadIdx.x & 7, len = ck >> 3;
    for (int s0 = 0; s0 < nl; s0 += (int)(blockDim.x >> 3)) {
      const int sl = s0 + (threadIdx.x >> 3);
      double acc = 0.0;
      if (sl < nl) {
        const double* row = einv_s + (size_t)sl * COARSE_LD + seg;
        const double* cc = c_s + seg;
        for (int i = 0; i < len; ++i) acc += row[i * 8] * cc[i * 8];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (sl < nl && seg == 0) e_s[sl] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < R; ++rr) e[rr] = (my_row[rr] >= 0) ? e_s[my_slot[rr]] : 0.0;
    ++cph;
  };
  // y_row = sum_j val_j x[col_j] for my rows, from the resident slices
  auto spmv_res = [&](const double* x, double (&y)[R]) {
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int nnz = blk_nnz[rr];
      const double* sv = sval + rr * SL;
      const int* sc = scol + rr * SL;
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
  double y[R], ec[R], tmp[R];
  // The pipelined recurrences drift: the recursively updated ||u|| can pass the test while b - A x has not.  A
  // claimed convergence is therefore CONFIRMED by restarting from the true residual (the set-up block below, ~one
  // extra product and three barriers per solve): the restart's own first test sees ||M^-1 (b - A x)|| itself; if it
  // fails the iteration simply carries on from the replaced residual.
  int it = 0;
  double bn = 0.0, tol = 0.0, un = 0.0;
  for (int cycle = 0;; ++cycle) {
  // r = b - A x, u = M^-1 r
  double acc1[1] = {0.0};
  spmv_res(a.x, y);
  if (coarse) {  // ||M^-1 b||: restrict b, barrier, correct
#pragma unroll
    for (int rr = 0; rr < R; ++rr) tmp[rr] = (my_row[rr] >= 0) ? a.b[my_row[rr]] : 0.0;
    coarse_restrict(tmp);
    bar.sync();
    coarse_correct(ec);
#pragma unroll
    for (int rr = 0; rr < R; ++rr) tmp[rr] = (my_row[rr] >= 0) ? a.b[my_row[rr]] - y[rr] : 0.0;
    coarse_restrict(tmp);  // for u = M^-1 r below
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = my_row[rr];
    if (row >= 0) {
      const double ri = a.b[row] - y[rr], bi = my_dinv[rr] * a.b[row] + (coarse ? ec[rr] : 0.0);
      dinv[row] = my_dinv[rr]; r[row] = ri; u[row] = my_dinv[rr] * ri;
      z[row] = 0.0; q[row] = 0.0; s[row] = 0.0; p[row] = 0.0;
      acc1[0] += bi * bi;
    }
  }
  gsum<1>(a, bar, acc1, phase, sh);
  bn = sqrt(acc1[0]);
  tol = fmax(a.rtol * bn, a.atol);
  if (coarse) {  // finish u = D^-1 r + P E^-1 P^T r and publish it for the product below
    coarse_correct(ec);
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (my_row[rr] >= 0) u[my_row[rr]] += ec[rr];
    bar.sync();
  }
  // w = A u, m = M^-1 w
  double d[3] = {0.0, 0.0, 0.0};  // gamma = (r,u), delta = (w,u), ||u||^2
  spmv_res(u, y);
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = my_row[rr];
    tmp[rr] = y[rr];
    if (row >= 0) {
      const double ui = u[row];
      w[row] = y[rr];
      mbuf[0][row] = my_dinv[rr] * y[rr];
      d[0] += r[row] * ui; d[1] += y[rr] * ui; d[2] += ui * ui;
    }
  }
  if (coarse) coarse_restrict(tmp);
  gsum<3>(a, bar, d, phase, sh);
  if (coarse) {
    coarse_correct(ec);
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (my_row[rr] >= 0) mbuf[0][my_row[rr]] += ec[rr];
    bar.sync();
  }
  double gamma_old = 1.0, alpha = 1.0;
  un = sqrt(d[2]);
  bool again = false;
  int np = 0;
  for (int lit = 0; it <= a.maxit; ++it, ++lit) {  // lit: iterations since the last (re)start
    un = sqrt(d[2]);
    if (!(un == un) || un > a.dtol * bn) { finish(a, it, -1, un, bn); return; }
    if (un <= tol) {
      if (lit == 0 || cycle >= 3) { finish(a, it, 1, un, bn); return; }  // lit == 0: un is the true residual norm
      again = true;
      break;
    }
    if (it == a.maxit) { finish(a, a.maxit, 0, un, bn); return; }
    const double gamma = d[0], delta = d[1];
    double beta = 0.0;
    if (lit > 0) {
      beta = gamma / gamma_old;
      const double den = delta - beta * gamma / alpha;
      if (!(den > 0.0)) { finish(a, it, -1, un, bn); return; }
      alpha = gamma / den;
    } else {
      if (!(delta > 0.0)) { finish(a, it, -1, un, bn); return; }
      alpha = gamma / delta;
    }
    gamma_old = gamma;
    const double* m = mbuf[lit & 1];
    double* mn = mbuf[(lit + 1) & 1];
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
    stamp(a.prof, np);
    spmv_res(m, y);
    stamp(a.prof, np);
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
        tmp[rr] = wi;
        d[0] += ri * ui; d[1] += wi * ui; d[2] += ui * ui;
      } else {
        tmp[rr] = 0.0;
      }
    }
    stamp(a.prof, np);
    if (coarse) coarse_restrict(tmp);
    gsum<3>(a, bar, d, phase, sh);
    if (coarse) {  // m_new = D^-1 w + P E^-1 P^T w, visible to every CTA before the next product
      coarse_correct(ec);
#pragma unroll
      for (int rr = 0; rr < R; ++rr)
        if (my_row[rr] >= 0) mn[my_row[rr]] += ec[rr];
      bar.sync();
    }
    stamp(a.prof, np);
  }
  if (!again) break;
  }
  finish(a, a.maxit, 0, un, bn);
}

template <int R>
static int launch_pcg_res(fx_plan* P, KArgs& ka, int grid, cudaStream_t st) {
  const size_t smem = pcg_smem_bytes(R, ka.cz.slice) + (ka.cz.agg != nullptr ? pcg_coarse_smem_bytes(ka.cz.nlmax) : 0);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    FX_CUDA(cudaFuncSetAttribute(k_pcg_res<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  void* args[] = {(void*)&ka};
  FX_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg_res<R>, dim3(grid), dim3(LA_THREADS), args, smem, st));
  count_launch(1);
  if (ka.prof) {  // debugging aid: microseconds of [own-row loads | SpMV | updates | reduction+barrier], iterations 10..13
    static unsigned long long h[1024];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
    fprintf(stderr, "[pcg prof] grid %d R %d:", grid, R);
    for (int it = 10; it < 14; ++it) {
      fprintf(stderr, "  |");
      for (int k = 0; k < 4; ++k) fprintf(stderr, " %.2f", (double)(h[4 * it + k + 1] - h[4 * it + k]) * 1e-3);
    }
    fprintf(stderr, "\n");
  }
  return FX_OK;
}

// Configuration of the resident CG kernel for a space: R row blocks per warp (0 = does not fit), the
// shared-memory slice per block, the grid, and how many rows of E^-1 fit beside them (two-level variant).
PcgConfig pcg_resident_config(const fx_plan* P, int nrowblk, int maxblknnz) {
  PcgConfig c{0, 0, 0, 0};
  if (nrowblk <= 0) return c;
  const int warps = P->num_sms * LA_WARPS;
  const int R = (nrowblk + warps - 1) / warps;
  const int slice = std::max(32, (maxblknnz + 31) / 32 * 32);
  const size_t budget = 227 * 1024 - 4096;  // opt-in shared memory per CTA minus the static arrays
  if (R > PCG_MAXR || pcg_smem_bytes(R, slice) > budget) return c;
  c.R = R; c.slice = slice;
  c.grid = std::min(P->num_sms, (nrowblk + LA_WARPS * R - 1) / (LA_WARPS * R));
  int nl = COARSE_NLMAX;
  while (nl > 0 && pcg_smem_bytes(R, slice) + pcg_coarse_smem_bytes(nl) > budget) --nl;
  c.nlmax = nl;
  return c;
}

// Build (once) and upload the node-block layout of a space.  Spaces without a usable structure keep state -1 and the
// reason; CUDA failures are the only errors.
int nb_ensure(fx_plan* P, int space) {
  DevPattern& d = P->sp[space];
  DevPattern::NbDev& nb = d.nb;
  if (nb.state != 0) return FX_OK;
  int bs = 0, nloc = 0;
  if (!d.has_csr || !nb_space_layout(space, bs, nloc)) { nb.state = -1; nb.why = "space has no node-block layout"; return FX_OK; }
  const int nc = P->nc;
  std::vector<int> rowptr((size_t)d.n + 1), col((size_t)d.nnz), dm((size_t)d.ld * nc);
  FX_CUDA(cudaMemcpy(rowptr.data(), d.rowptr, sizeof(int) * rowptr.size(), cudaMemcpyDeviceToHost));
  FX_CUDA(cudaMemcpy(col.data(), d.col, sizeof(int) * col.size(), cudaMemcpyDeviceToHost));
  FX_CUDA(cudaMemcpy(dm.data(), d.dm_t, sizeof(int) * dm.size(), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> mask;
  if (!d.elim_h.empty() || !d.halo_h.empty()) {
    mask.assign((size_t)d.n, 0);
    for (int i = 0; i < d.n; ++i) mask[i] = (d.elim_h.empty() ? 0 : d.elim_h[i]) | (d.halo_h.empty() ? 0 : d.halo_h[i]);
  }
  NbHost H;
  std::string why;
  if (!nb_build(dm.data(), nc, d.ld, bs, nloc, d.n, rowptr.data(), col.data(), mask.empty() ? nullptr : mask.data(),
                P->num_sms, H, why, d.ghost_owner_h.empty() ? nullptr : d.ghost_owner_h.data())) {
    nb.state = -1; nb.why = why;
    return FX_OK;
  }
  auto up = [](auto** dst, const auto& v) -> cudaError_t {
    *dst = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(v[0]));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice);
  };
  int* desc_i = nullptr;
  FX_CUDA(up(&desc_i, H.desc));
  nb.desc = reinterpret_cast<int4*>(desc_i);
  FX_CUDA(up(&nb.wrange, H.wrange));
  FX_CUDA(up(&nb.bre, H.bre));
  FX_CUDA(up(&nb.bcol, H.bcol));
  FX_CUDA(up(&nb.src, H.src));
  FX_CUDA(up(&nb.diag, H.diag));
  FX_CUDA(up(&nb.ext_of_int, H.ext_of_int));
  FX_CUDA(up(&nb.int_of_ext, H.int_of_ext));
  FX_CUDA(up(&nb.zchk, H.zchk));
  FX_CUDA(up(&nb.imask, H.imask));
  nb.bs = H.bs; nb.bp = H.bp; nb.nn = H.nn; nb.N = H.N; nb.nchunks = H.nchunks; nb.grid = H.grid;
  nb.nzchk = (int)H.zchk.size(); nb.ne = H.ne; nb.nslots = H.nslots;
  if (P->nbval_cap < (size_t)H.nslots) {
    cudaFree(P->nbval);
    P->nbval = nullptr; P->nbval_cap = 0;
    FX_CUDA(cudaMalloc((void**)&P->nbval, sizeof(double) * (size_t)H.nslots));
    P->nbval_cap = (size_t)H.nslots;
  }
  nb.int_of_ext_h = H.int_of_ext;
  nb.state = 1;
  return nb_attach_halo(P, space);
}

// Partitioned run: the halo push lists of the space in INTERNAL numbering (needs the destinations' internal indices,
// fx_plan_set_halo_solver_index).  No-op until both the layout and the indices exist.
int nb_attach_halo(fx_plan* P, int space) {
  DevPattern& d = P->sp[space];
  DevPattern::NbDev& nb = d.nb;
  cudaFree(nb.s_first); cudaFree(nb.s_remote); cudaFree(nb.s_local);
  nb.s_first = nb.s_remote = nb.s_local = nullptr;
  if (nb.state != 1 || !d.has_halo || (int)d.halo_remote_int.size() != d.nsend || d.halo_first_h.empty()) return FX_OK;
  std::vector<int> ext_of_int((size_t)nb.N), first((size_t)nb.N, -1), remote((size_t)std::max(d.nsend, 1), 0);
  FX_CUDA(cudaMemcpy(ext_of_int.data(), nb.ext_of_int, sizeof(int) * ext_of_int.size(), cudaMemcpyDeviceToHost));
  for (int i = 0; i < nb.N; ++i)
    if (ext_of_int[i] >= 0) first[i] = d.halo_first_h[ext_of_int[i]];
  std::vector<int> local_ext((size_t)std::max(d.nsend, 1), 0), local_int((size_t)std::max(d.nsend, 1), -1);
  if (d.nsend > 0) FX_CUDA(cudaMemcpy(local_ext.data(), d.s_local, sizeof(int) * (size_t)d.nsend, cudaMemcpyDeviceToHost));
  for (int k = 0; k < d.nsend; ++k) {
    remote[k] = d.halo_remote_int[d.halo_order[k]];
    local_int[k] = nb.int_of_ext_h[local_ext[k]];
  }
  FX_CUDA(cudaMalloc((void**)&nb.s_local, sizeof(int) * local_int.size()));
  FX_CUDA(cudaMemcpy(nb.s_local, local_int.data(), sizeof(int) * local_int.size(), cudaMemcpyHostToDevice));
  FX_CUDA(cudaMalloc((void**)&nb.s_first, sizeof(int) * first.size()));
  FX_CUDA(cudaMalloc((void**)&nb.s_remote, sizeof(int) * remote.size()));
  FX_CUDA(cudaMemcpy(nb.s_first, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice));
  FX_CUDA(cudaMemcpy(nb.s_remote, remote.data(), sizeof(int) * remote.size(), cudaMemcpyHostToDevice));
  return FX_OK;
}

static NbView nb_view(const fx_plan* P, const DevPattern& d) {
  const DevPattern::NbDev& nb = d.nb;
  static const int pol_env = getenv("FX_NB_L2POLICY") ? atoi(getenv("FX_NB_L2POLICY")) : 1;
  static const int planemask_env = getenv("FX_NB_PLANEMASK") ? atoi(getenv("FX_NB_PLANEMASK")) : 1;
  static const int pol_small_env = getenv("FX_NB_L2POLICY_SMALL") ? atoi(getenv("FX_NB_L2POLICY_SMALL")) : 2;
  static const int pre_env = getenv("FX_NB_DUAL") ? atoi(getenv("FX_NB_DUAL")) : 1;
  return NbView{nb.nn, nb.N, nb.nchunks, nb.nzchk, nb.ne, nb.nslots, nb.desc, nb.wrange, nb.bre, nb.bcol, nb.src,
                nb.diag, nb.ext_of_int, nb.zchk, nb.imask, P->nbval,
                // matrices that fit the L2 beside the work vectors (<= 72 MB of block values) may ask for another policy
                (pol_small_env >= 0 && nb.nslots * 8 + nb.ne * 4 <= (72ll << 20)) ? pol_small_env : pol_env, pre_env,
                nb.s_first, nb.s_remote, nb.s_local, P->bar + 4, planemask_env};
}

template <int BS>
static int launch_nb(void* fn, const DevPattern& d, KArgs& ka, int reps, cudaStream_t st, int warps = NB_WARPS_CG) {
  const size_t smem = sizeof(double) * warps * nb_chunk(BS) * BS;
  static std::vector<void*> done;
  if (std::find(done.begin(), done.end(), fn) == done.end()) {
    FX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done.push_back(fn);
  }
  void* args1[] = {(void*)&ka};
  void* args2[] = {(void*)&ka, (void*)&reps};
  FX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(d.nb.grid), dim3(warps * 32), reps >= 0 ? args2 : args1, smem, st));
  count_launch(1);
  return FX_OK;
}

// y = A x on the rows of the node blocks, through the node-block layout (variant 8 of fx_spmv_variant; reps products)
int la_spmv_nb(fx_plan* P, int space, const double* val, const double* x, double* y, int reps, cudaStream_t st) {
  DevPattern& p = P->sp[space];
  int rc = nb_ensure(P, space);
  if (rc) return rc;
  if (p.nb.state != 1) return set_error(FX_ERR_UNSUPPORTED, "space %d: no node-block layout (%s)", space, p.nb.why.c_str());
  if ((rc = ensure_work(P, ((size_t)std::max(p.n, p.nb.N) + 3) & ~(size_t)3))) return rc;
  KArgs ka{};
  ka.A = CsrView{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, nullptr, nullptr, nullptr, nullptr, p.elim, p.elim_rows,
                 p.nelim, nullptr, nullptr, 0, nullptr, nullptr};
  ka.nb = nb_view(P, p);
  ka.bar = P->bar; ka.w = P->work; ka.x = y; ka.b = x; ka.info = P->info + (FX_NTICKETS - 1);
  FX_CUDA(cudaMemsetAsync(P->bar, 0, 8 * sizeof(unsigned), st));  // barrier counter + the node-block plane mask
  switch (p.nb.bs) {
    case 1: return launch_nb<1>((void*)k_nb_spmv<1>, p, ka, reps, st);
    case 2: return launch_nb<2>((void*)k_nb_spmv<2>, p, ka, reps, st);
    default: return launch_nb<3>((void*)k_nb_spmv<3>, p, ka, reps, st);
  }
}

int la_krylov(fx_plan* P, int space, const double* val, double* x, const double* b, int method, int pc,
              double rtol, double atol, int maxit, int ticket, cudaStream_t st) {
  DevPattern& p = P->sp[space];
  if (!p.has_csr) return set_error(FX_ERR_ARG, "space %d has no CSR pattern", space);
  if (method != FX_KSP_CG && method != FX_KSP_BICGSTAB && method != FX_KSP_GMRES)
    return set_error(FX_ERR_UNSUPPORTED, "Krylov method %d not implemented (bicgstab, cg, gmres available)", method);
  if (pc == FX_PC_ILU0)  // PETSc's serial default: natural-ordering ILU(0), dependency-driven kernels (fx_ilu.cu)
    return la_krylov_ilu0(P, space, val, x, b, method, rtol, atol, maxit, ticket, st);
  if (pc != FX_PC_NONE && pc != FX_PC_JACOBI && pc != FX_PC_JACOBI_COARSE)
    return set_error(FX_ERR_UNSUPPORTED, "preconditioner %d not implemented (none, jacobi, ilu, jacobi_coarse available)", pc);
  PcgConfig pcfg = pcg_resident_config(P, p.nrowblk, p.maxblknnz);
  static const int resident_env = getenv("FX_PCG_RESIDENT") ? atoi(getenv("FX_PCG_RESIDENT")) : 1;
  if (!resident_env) pcfg.R = 0;  // testing aid: send CG solves through k_cg even when the system would fit k_pcg_res
  // jacobi_coarse: the resident kernel when the system fits it (and rtol >= 1e-9: pipelined recurrences), else k_cg
  const bool coarse_resident = pcfg.R > 0 && p.nagg <= COARSE_KMAX && p.agg_nl <= pcfg.nlmax && rtol >= 1e-9;
  if (pc == FX_PC_JACOBI_COARSE && (method != FX_KSP_CG || p.nagg == 0 || p.nelim != 0))
    return set_error(FX_ERR_UNSUPPORTED, "jacobi_coarse needs CG, aggregates on the space (fx_plan_set_aggregates) and no "
                     "eliminated rows");
  const bool part = P->dist.world > 1;  // mesh-partitioned run: fused halo pushes + cross-GPU barriers
  if (part && !p.has_halo)
    return set_error(FX_ERR_ARG, "partitioned plan: space %d has no halo (fx_plan_set_halo)", space);
  if (part && (size_t)p.n > (size_t)P->dist.nmax)
    return set_error(FX_ERR_ARG, "partitioned plan: space %d has %d dofs, peer regions hold %lld", space, p.n, P->dist.nmax);
  if (part && pc == FX_PC_JACOBI_COARSE)
    return set_error(FX_ERR_UNSUPPORTED, "jacobi_coarse is not available on a partitioned plan");
  int rc = ensure_work(P, (size_t)p.n);
  if (rc) return rc;
  KArgs ka;
  static const int reblock_env = getenv("FX_REBLOCK") ? atoi(getenv("FX_REBLOCK")) : 1;
  ka.A = CsrView{p.n, p.rowptr, p.col, val, p.rowblk, p.nrowblk, P->cdesc, P->cre, P->ccol, P->cval,
                 (p.nelim > 0 || part) ? p.elim : nullptr, p.elim_rows, p.nelim, P->cprefix, P->scanpart,
                 (reblock_env && p.maxrow <= SPMV_NNZB / 2) ? SPMV_NNZB - p.maxrow : 0, P->rowpre, P->nblk};
  // partitioned run: every rank must take the SAME set-up branch (compact_reblock crosses two cross-GPU barriers,
  // compact_zeros one).  It does: build_pattern rejects rows longer than 256 = SPMV_NNZB / 2 entries, so the
  // condition above is true on every rank whatever its sub-mesh looks like.
  static_assert(SPMV_NNZB / 2 >= 256, "rank-local maxrow would decide the number of cross-GPU barriers");
  ka.bar = P->bar;
  ka.dd = P->dist;
  if (part) {
    ka.dd.nsend = p.nsend; ka.dd.s_first = p.s_first; ka.dd.s_peer = p.s_peer;
    ka.dd.s_local = p.s_local; ka.dd.s_remote = p.s_remote;
  }
  ka.cz = Coarse{pcfg.slice, pcfg.nlmax, nullptr, 0, nullptr, nullptr, nullptr};
  if (pc == FX_PC_JACOBI_COARSE)
    ka.cz = Coarse{pcfg.slice, pcfg.nlmax, p.agg, p.nagg, P->coarse, P->coarse + (size_t)COARSE_KMAX_CG * COARSE_KMAX_CG,
                   P->coarse + (size_t)2 * COARSE_KMAX_CG * COARSE_KMAX_CG};
  static const int prof_env = getenv("FX_KRYLOV_PROF") ? atoi(getenv("FX_KRYLOV_PROF")) : 0;
  ka.prof = prof_env;
  FX_CUDA(cudaMemsetAsync(P->bar, 0, 8 * sizeof(unsigned), st));  // barrier counter + the node-block plane mask
  static const int red_env = getenv("FX_RED_ATOMIC") ? atoi(getenv("FX_RED_ATOMIC")) : 1;
  ka.red_atomic = red_env;
  if (red_env) FX_CUDA(cudaMemsetAsync(P->red + (size_t)2 * RED_SLOTS * MAX_GRID, 0, sizeof(double) * (3 * RED_SLOTS + 3 * 8), st));
  ka.x = x; ka.b = b; ka.w = part ? P->dist.region[P->dist.rank] : P->work; ka.red = P->red;
  ka.rtol = rtol; ka.atol = atol; ka.dtol = 1e5; ka.maxit = maxit; ka.pc = pc;
  ka.info = P->info + ticket;
  ka.basis = nullptr; ka.red_big = nullptr;
  if (method == FX_KSP_GMRES) {
    if (part) return set_error(FX_ERR_UNSUPPORTED, "gmres is not available on a partitioned plan (bicgstab, cg are)");
    if (pc == FX_PC_JACOBI_COARSE) return set_error(FX_ERR_UNSUPPORTED, "gmres: preconditioners none, jacobi");
    const size_t need = (size_t)(GMRES_M + 1) * (size_t)p.n;
    if (P->gmres_cap < need) {
      cudaFree(P->gmres_basis);
      P->gmres_basis = nullptr; P->gmres_cap = 0;
      FX_CUDA(cudaMalloc((void**)&P->gmres_basis, sizeof(double) * need));
      P->gmres_cap = need;
    }
    if (!P->red_big) FX_CUDA(cudaMalloc((void**)&P->red_big, sizeof(double) * 2 * GMRES_NRED * MAX_GRID));
    ka.basis = P->gmres_basis; ka.red_big = P->red_big;
    constexpr int T = 256;
    const size_t smem = sizeof(double) * (T / 32) * SPMV_NNZB;
    static bool attr = false;
    if (!attr) {
      FX_CUDA(cudaFuncSetAttribute(k_gmres<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    int per_sm = 0;
    FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gmres<T>, T, smem));
    if (per_sm < 1) return set_error(FX_ERR_CUDA, "GMRES kernel does not fit on an SM");
    const int grid = max(1, min(min(per_sm * P->num_sms, MAX_GRID), (p.nrowblk + T / 32 - 1) / (T / 32)));
    void* gargs[] = {(void*)&ka};
    FX_CUDA(cudaLaunchCooperativeKernel((void*)k_gmres<T>, dim3(grid), dim3(T), gargs, smem, st));
    count_launch(1);
    return FX_OK;
  }
  // CG on the node-block layout: everything the resident kernel below does not take (tight tolerances, systems that
  // do not fit it, the 3 x 3-block Zc mass projections) with Jacobi / no preconditioning
  static const int nb_cg_env = getenv("FX_NB") ? atoi(getenv("FX_NB")) : 1;
  const bool resident = method == FX_KSP_CG && rtol >= 1e-9 && p.nelim == 0 && !part && pcfg.R > 0 &&
                        (pc != FX_PC_JACOBI_COARSE || coarse_resident);
  if (nb_cg_env && method == FX_KSP_CG && !part && !resident && (pc == FX_PC_NONE || pc == FX_PC_JACOBI)) {
    if ((rc = nb_ensure(P, space))) return rc;
    if (p.nb.state == 1) {
      if ((rc = ensure_work(P, ((size_t)std::max(p.n, p.nb.N) + 3) & ~(size_t)3))) return rc;
      ka.w = P->work;
      ka.nb = nb_view(P, p);
      switch (p.nb.bs) {
        case 1: return launch_nb<1>((void*)k_cg_nb<1>, p, ka, -1, st);
        case 2: return launch_nb<2>((void*)k_cg_nb<2>, p, ka, -1, st);
        default: return launch_nb<3>((void*)k_cg_nb<3>, p, ka, -1, st);
      }
    }
  }
  // small SPD systems at loose tolerance: shared-memory-resident pipelined CG, one CTA per SM
  if (method == FX_KSP_CG && rtol >= 1e-9 && p.nelim == 0 && !part && pcfg.R > 0 &&
      (pc != FX_PC_JACOBI_COARSE || coarse_resident)) {
    switch (pcfg.R) {
      case 1: return launch_pcg_res<1>(P, ka, pcfg.grid, st);
      case 2: return launch_pcg_res<2>(P, ka, pcfg.grid, st);
      case 3: return launch_pcg_res<3>(P, ka, pcfg.grid, st);
      default: return launch_pcg_res<4>(P, ka, pcfg.grid, st);
    }
  }
  // BiCGStab on the node-block layout when the space has one (FX_NB=0: the scalar compacted-CSR kernel below)
  static const int nb_env = getenv("FX_NB") ? atoi(getenv("FX_NB")) : 1;
  if (nb_env && method == FX_KSP_BICGSTAB) {
    if ((rc = nb_ensure(P, space))) return rc;
    // partitioned run: the halo pushes need the neighbours' INTERNAL indices (fx_plan_set_halo_solver_index) and the
    // peer regions must hold the internal vectors; otherwise the scalar kernel below serves the solve
    // (... and a vector stride that keeps every work vector 32-byte aligned: node blocks are gathered with 256-bit loads)
    const bool nb_ok = p.nb.state == 1 &&
                       (!part || (p.nb.s_first != nullptr && (long long)p.nb.N <= P->dist.nmax && P->dist.nmax % 4 == 0));
    if (nb_ok) {
      if (!part) {
        if ((rc = ensure_work(P, ((size_t)std::max(p.n, p.nb.N) + 3) & ~(size_t)3))) return rc;
        ka.w = P->work;
      }
      ka.nb = nb_view(P, p);
      const bool dual = ka.nb.dual != 0;
      if (part) {
        switch (p.nb.bs) {
          case 1: rc = launch_nb<1>((void*)k_bicgstab_nb<1, false, true>, p, ka, -1, st, NB_WARPS); break;
          case 2: rc = launch_nb<2>((void*)k_bicgstab_nb<2, false, true>, p, ka, -1, st, NB_WARPS); break;
          default: rc = launch_nb<3>((void*)k_bicgstab_nb<3, false, true>, p, ka, -1, st, NB_WARPS); break;
        }
      } else {
        switch (p.nb.bs) {
          case 1: rc = launch_nb<1>(dual ? (void*)k_bicgstab_nb<1, true, false> : (void*)k_bicgstab_nb<1, false, false>, p, ka, -1, st, NB_WARPS); break;
          case 2: rc = launch_nb<2>(dual ? (void*)k_bicgstab_nb<2, true, false> : (void*)k_bicgstab_nb<2, false, false>, p, ka, -1, st, NB_WARPS); break;
          default: rc = launch_nb<3>(dual ? (void*)k_bicgstab_nb<3, true, false> : (void*)k_bicgstab_nb<3, false, false>, p, ka, -1, st, NB_WARPS); break;
        }
      }
      if (rc == FX_OK && prof_env && p.n > 100000) {  // debugging aid: per-phase microseconds of iterations 2..4
        static unsigned long long h[1024];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
        fprintf(stderr, "[krylov prof nb] space %d grid %d [spmv1 red+bar pass_s red+bar spmv2 red+bar pass_p bar]:", space, p.nb.grid);
        for (int it = 1; it < 4; ++it) {
          fprintf(stderr, "  |");
          for (int k = 0; k < 8; ++k) fprintf(stderr, " %.1f", (double)(h[1 + 8 * it + k] - h[8 * it + k]) * 1e-3);
        }
        fprintf(stderr, "\n");
      }
      return rc;
    }
  }
  // big CTAs (1024 threads, one per SM) keep the grid barrier cheap: its cost grows with the number of
  // arriving CTAs, and every iteration crosses four of them
  static const int threads_env = getenv("FX_KRYLOV_THREADS") ? atoi(getenv("FX_KRYLOV_THREADS")) : 1024;
  const int threads = (threads_env == 256 && !part) ? 256 : 1024;
  void* fn = method == FX_KSP_CG ? (threads == 256 ? (void*)k_cg<256, false> : (void*)k_cg<1024, false>)
                                 : (threads == 256 ? (void*)k_bicgstab<256, false> : (void*)k_bicgstab<1024, false>);
  if (part) fn = method == FX_KSP_CG ? (void*)k_cg<1024, true> : (void*)k_bicgstab<1024, true>;
  const size_t smem = sizeof(double) * (threads / 32) * SPMV_NNZB;
  static bool attr_done[6] = {false, false, false, false, false, false};
  const int ai = (method == FX_KSP_CG ? 0 : 1) + (part ? 4 : (threads == 256 ? 0 : 2));
  if (!attr_done[ai]) {
    FX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[ai] = true;
  }
  int per_sm = 0;
  FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
  if (per_sm < 1) return set_error(FX_ERR_CUDA, "Krylov kernel does not fit on an SM");
  // one warp per 1-2 row blocks up to the co-resident limit; small systems get a small grid (cheap barriers)
  const int wpb = threads / 32;
  int grid = max(1, min(min(per_sm, 1024 / threads) * P->num_sms, (p.nrowblk + wpb - 1) / wpb));
  grid = min(grid, MAX_GRID);
  void* args[] = {(void*)&ka};
  FX_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, st));
  count_launch(1);
  if (prof_env && method == FX_KSP_BICGSTAB && p.n > 100000 && (!part || P->dist.rank == 0)) {  // debugging aid: per-phase microseconds of iterations 2..4
    static unsigned long long h[1024];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
    fprintf(stderr, "[krylov prof] space %d grid %d:", space, grid);
    for (int it = 1; it < 4; ++it) {
      fprintf(stderr, "  |");
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %.1f", (double)(h[1 + 8 * it + k] - h[8 * it + k]) * 1e-3);
    }
    fprintf(stderr, "\n");
    static unsigned long long hw[8192];
    cudaMemcpyFromSymbol(hw, g_prof_warp, sizeof(hw));
    const int nwarp = grid * wpb;
    // SpMV1 of iteration 2 started at h[8] (block 0's view); per-warp finish offsets
    std::vector<double> dt;
    for (int i = 0; i < nwarp && i < 8192; ++i) dt.push_back((double)(long long)(hw[i] - h[8]) * 1e-3);
    std::vector<double> srt(dt);
    std::sort(srt.begin(), srt.end());
    fprintf(stderr, "[krylov prof] warp finish (us after start) min %.1f p10 %.1f med %.1f p90 %.1f max %.1f\n", srt[0],
            srt[nwarp / 10], srt[nwarp / 2], srt[nwarp * 9 / 10], srt[nwarp - 1]);
    // per-CTA max and spread
    double cmax_min = 1e9, cmax_max = 0, spread = 0;
    for (int c = 0; c < grid; ++c) {
      double mx = 0, mn = 1e9;
      for (int w2 = 0; w2 < wpb; ++w2) { mx = std::max(mx, dt[c * wpb + w2]); mn = std::min(mn, dt[c * wpb + w2]); }
      cmax_min = std::min(cmax_min, mx); cmax_max = std::max(cmax_max, mx); spread += (mx - mn) / grid;
    }
    fprintf(stderr, "[krylov prof] per-CTA slowest warp: fastest CTA %.1f slowest CTA %.1f; mean in-CTA spread %.1f\n", cmax_min,
            cmax_max, spread);
  }
  return FX_OK;
}

}  // namespace fx
