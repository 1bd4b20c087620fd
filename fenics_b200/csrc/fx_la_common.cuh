// Device helpers shared by fx_la.cu and fx_krylov.cu: warp/block reductions and the CSR "stream"
// SpMV building block.
#pragma once
#include "fx_plan.h"
#include "fx_plan_host.h"  // SPMV_NNZB, SPMV_ROWB

namespace fx {

constexpr int LA_THREADS = 256;

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

struct CsrView {
  int n;
  const int* rowptr;
  const int* col;
  const double* val;
  const int* rowblk;
  int nrowblk;
};

// CSR-stream SpMV (rows of 7-57 nonzeros are too short for a warp per row): a CTA takes a block of
// consecutive rows holding <= SPMV_NNZB nonzeros, streams val[j]*x[col[j]] into shared memory with
// fully coalesced val/col loads (many independent loads in flight per thread), then one thread per
// row sums its slice and runs the fused epilogue epi(row, (A x)[row]).
// `x` may have been written earlier in the same (cooperative) kernel: no __restrict__, no ld.nc.
template <class Epi>
__device__ __forceinline__ void spmv_stream(const CsrView& A, const double* x, double* prod, Epi&& epi) {
  for (int k = blockIdx.x; k < A.nrowblk; k += gridDim.x) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0], j1 = A.rowptr[r1];
#pragma unroll 4
    for (int j = j0 + threadIdx.x; j < j1; j += LA_THREADS) prod[j - j0] = A.val[j] * x[A.col[j]];
    __syncthreads();
    if (threadIdx.x < r1 - r0) {
      const int row = r0 + threadIdx.x;
      const int s0 = A.rowptr[row] - j0, s1 = A.rowptr[row + 1] - j0;
      double s = 0.0;
      for (int j = s0; j < s1; ++j) s += prod[j];
      epi(row, s);
    }
    __syncthreads();
  }
}

}  // namespace fx
