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

// streaming (read-once) loads: non-coherent path, do not allocate in L1 so the x-gather lines stay
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

struct CsrView {
  int n;
  const int* rowptr;
  const int* col;
  const double* val;   // must not be written while a kernel using ld_stream on it runs
  const int* rowblk;
  int nrowblk;
};

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
  for (int k = blockIdx.x * LA_WARPS + wib; k < A.nrowblk; k += nw) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0], j1 = A.rowptr[r1];
    for (int jb = j0; jb < j1; jb += 256) {
      int c[8];
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = jb + u * 32 + lane;
        c[u] = (j < j1) ? ld_stream(A.col + j) : -1;
        v[u] = (j < j1) ? ld_stream(A.val + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c[u] >= 0) prod[jb - j0 + u * 32 + lane] = v[u] * x[c[u]];
    }
    __syncwarp();
    if (lane < r1 - r0) {
      const int row = r0 + lane;
      const int s0 = A.rowptr[row] - j0, s1 = A.rowptr[row + 1] - j0;
      double s = 0.0;
      for (int j = s0; j < s1; ++j) s += prod[j];
      epi(row, s);
    }
    __syncwarp();
  }
}

}  // namespace fx
