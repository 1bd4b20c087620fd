#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k(int iters, double* out) {
  cg::grid_group g = cg::this_grid();
  double s = 0;
  for (int i = 0; i < iters; ++i) { g.sync(); s += 1.0; }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = s;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  int iters = 2000;
  for (int bs : {256, 1024}) for (int grid : {16, 37, 74, 148, 296, 592, 1184}) {
    if (bs == 1024 && grid > 296) continue;
    void* args[] = {&iters, &out};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(bs), args, 0, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(bs), args, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("block %4d grid %5d: %.3f us per grid.sync (%s)\n", bs, grid, ms * 1000 / iters, cudaGetErrorString(e));
  }
  return 0;
}
