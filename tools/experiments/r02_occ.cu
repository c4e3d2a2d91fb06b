// cluster co-residency table (all sizes), 352 threads, ~212 KB dynamic smem
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(352, 1) k(int* p) { if (p) *p = 1; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  for (int cs = 1; cs <= 16; ++cs) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = 212 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cs %2d: %3d clusters = %3d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    (void)cudaGetLastError();
  }
  return 0;
}
