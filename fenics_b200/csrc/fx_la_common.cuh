// Device helpers shared by fx_la.cu and fx_krylov.cu: warp/block reductions and the CSR "stream"
// SpMV building block.
#pragma once
#include "fx_plan.h"
#include "fx_plan_host.h"  // SPMV_NNZB, SPMV_ROWB

namespace fx {

constexpr int LA_THREADS = 256;
constexpr int LA_WARPS = LA_THREADS / 32;
constexpr int SPMV_SMEM_DOUBLES = LA_WARPS * SPMV_NNZB;  // shared staging per CTA (32 KB)

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
// block-wide sum broadcast to every thread (fixed reduction tree => deterministic); sh: 33 doubles
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
// NV sums at once (same barriers as one): sh must hold NV*33 doubles
template <int NV>
__device__ __forceinline__ void block_sum_bcast_n(double (&v)[NV], double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 33 + w] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = (lane < nw) ? sh[i * 33 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) sh[i * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[i * 33 + 32];
}

// streaming (read-once) loads: do not allocate in L1 so the x-gather lines stay resident
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// CSR matrix on the plan's pattern, optionally with a zero-compacted copy (see compact_zeros).
struct CsrView {
  int n;
  const int* rowptr;
  const int* col;
  const double* val;
  const int* rowblk;
  int nrowblk;
  // compacted copy (structural zeros of the DOLFIN cell-block pattern dropped); null = not compacted
  int* rs;      // [n] first entry of row
  int* re;      // [n] one past the last entry of row
  int* blkend;  // [nrowblk] one past the last entry of the row block
  int* col2;
  double* val2;
  // block-triangular elimination (null = off): rows with elim[row] != 0 are left out of the Krylov
  // iteration and recovered by back_substitute(); requires A[active, eliminated] == 0
  const unsigned char* elim;
  const int* elim_rows;
  int nelim;
};

// Drop exactly-zero entries, row block by row block, into (col2, val2) at the block's original offset.
// DOLFIN's pattern stores every (test dof, trial dof) pair of every cell whether or not the form
// couples them: 20-60 % of the stored values of the W and Zc matrices are exact
// zeros, which a Krylov solve would otherwise stream from HBM twice per iteration.  The warp that
// compacts a block is the warp that later streams it, so no grid barrier is needed in between.
// With elimination on, eliminated rows become empty in the copy (never streamed); the return value
// counts non-zero couplings A[active, eliminated] seen by this thread (must total 0).
__device__ __forceinline__ int compact_zeros(const CsrView& A) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nw = gridDim.x * LA_WARPS;
  int bad = 0;
  for (int k = blockIdx.x * LA_WARPS + wib; k < A.nrowblk; k += nw) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0];
    const int row = r0 + lane;
    const bool act = row < r1;
    int b = 0, e = 0, cnt = 0;
    if (act) {
      b = A.rowptr[row]; e = A.rowptr[row + 1];
      if (A.elim != nullptr && A.elim[row]) {
        e = b;  // eliminated row: empty
      } else if (A.elim != nullptr) {
        for (int j = b; j < e; ++j) {
          const bool nz = A.val[j] != 0.0;
          cnt += nz;
          bad += (nz && A.elim[A.col[j]]);
        }
      } else {
        for (int j = b; j < e; ++j) cnt += (A.val[j] != 0.0);
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (act) {
      int dst = j0 + incl - cnt;
      A.rs[row] = dst;
      A.re[row] = dst + cnt;
      for (int j = b; j < e; ++j) {
        const double v = A.val[j];
        if (v != 0.0) { A.val2[dst] = v; A.col2[dst] = A.col[j]; ++dst; }
      }
    }
    if (lane == 0) A.blkend[k] = j0 + total;
    __syncwarp();
  }
  return bad;
}

// x[row] = (b[row] - sum_{j != row} A[row,j] x[j]) / A[row,row] for the eliminated rows, 16 lanes per row
// (exact for a trailing block with diagonal self-coupling, e.g. the DG0 DEVSS rows of the W systems).
// Returns the number of non-zero couplings between DIFFERENT eliminated rows seen by this thread.
__device__ __forceinline__ int back_substitute(const CsrView& A, double* x, const double* b) {
  int bad = 0;
  const int sub = threadIdx.x & 15;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, ng = (gridDim.x * blockDim.x) >> 4;
  const int rounds = (A.nelim + ng - 1) / ng;  // uniform trip count: the shuffles need whole warps
  for (int it = 0; it < rounds; ++it) {
    const int k = it * ng + g;
    const int row = k < A.nelim ? A.elim_rows[k] : -1;
    double s = 0.0, dg = 0.0;
    if (row >= 0)
      for (int j = A.rowptr[row] + sub; j < A.rowptr[row + 1]; j += 16) {
        const int c = A.col[j];
        const double v = A.val[j];
        if (c == row) dg = v;
        else if (v != 0.0) { s += v * x[c]; bad += A.elim[c]; }
      }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o, 16);
      dg += __shfl_xor_sync(0xffffffffu, dg, o, 16);
    }
    if (row >= 0 && sub == 0) x[row] = (b[row] - s) / dg;
  }
  return bad;
}

// CSR-stream SpMV at warp granularity (rows of 7-57 nonzeros are too short for a warp per row and a
// thread per row cannot coalesce): a warp takes a block of consecutive rows holding <= SPMV_NNZB
// nonzeros (<= 32 rows), streams val[j]*x[col[j]] into its private shared-memory slice with fully
// coalesced val/col loads, 8 independent loads per lane in flight, then one lane per row sums its
// slice and runs the fused epilogue epi(row, (A x)[row]).  Only __syncwarp() -- warps of a CTA drift
// apart, so streaming and reducing overlap on the SM.
// `x` may have been written earlier in the same (cooperative) kernel: plain coherent loads.
template <class Epi>
__device__ __forceinline__ void spmv_stream(const CsrView& A, const double* x, double* smem, Epi&& epi) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* prod = smem + wib * SPMV_NNZB;
  const int nw = gridDim.x * LA_WARPS;
  const bool compact = A.rs != nullptr;
  const int* col = compact ? A.col2 : A.col;
  const double* val = compact ? A.val2 : A.val;
  for (int k = blockIdx.x * LA_WARPS + wib; k < A.nrowblk; k += nw) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0], j1 = compact ? A.blkend[k] : A.rowptr[r1];
    for (int jb = j0; jb < j1; jb += 256) {
      int c[8];
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = jb + u * 32 + lane;
        c[u] = (j < j1) ? ld_stream(col + j) : -1;
        v[u] = (j < j1) ? ld_stream(val + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c[u] >= 0) prod[jb - j0 + u * 32 + lane] = v[u] * x[c[u]];
    }
    __syncwarp();
    if (lane < r1 - r0) {
      const int row = r0 + lane;
      const int s0 = (compact ? A.rs[row] : A.rowptr[row]) - j0;
      const int s1 = (compact ? A.re[row] : A.rowptr[row + 1]) - j0;
      double s = 0.0;
      for (int j = s0; j < s1; ++j) s += prod[j];
      epi(row, s);
    }
    __syncwarp();
  }
}

}  // namespace fx
