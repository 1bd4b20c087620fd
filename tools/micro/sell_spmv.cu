// Stand-alone experiment: SELL-32-sigma SpMV (one lane per row, column-major slices) vs the CSR-stream kernel,
// on a matrix exported by tools/sell_experiment.py.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -shared -Xcompiler -fPIC tools/micro/sell_spmv.cu -o tools/micro/libsell.so
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) k_sell(int nslices, const long long* __restrict__ sptr, const int* __restrict__ slen,
                                              const double* __restrict__ val, const int* __restrict__ col,
                                              const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int s = warp; s < nslices; s += nwarp) {
    const long long base = sptr[s];
    const int len = slen[s];
    double acc = 0.0;
    int j = 0;
    for (; j + U <= len; j += U) {
      double v[U];
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = __ldcs(val + base + (long long)(j + u) * 32 + lane);
        c[u] = __ldcs(col + base + (long long)(j + u) * 32 + lane);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += v[u] * x[c[u]];
    }
    for (; j < len; ++j) acc += __ldcs(val + base + (long long)j * 32 + lane) * x[__ldcs(col + base + (long long)j * 32 + lane)];
    y[s * 32 + lane] = acc;
  }
}

extern "C" int sell_spmv(int nslices, const long long* sptr, const int* slen, const double* val, const int* col,
                         const double* x, double* y, int unroll, int grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (unroll == 4) k_sell<4><<<grid, 256, 0, st>>>(nslices, sptr, slen, val, col, x, y);
  else if (unroll == 8) k_sell<8><<<grid, 256, 0, st>>>(nslices, sptr, slen, val, col, x, y);
  else k_sell<16><<<grid, 256, 0, st>>>(nslices, sptr, slen, val, col, x, y);
  return (int)cudaGetLastError();
}
