// Scratch probe: how many clusters of a 1-CTA-per-SM kernel (all of shared memory) fit on the device at once.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s SMs %d\n", prop.name, prop.multiProcessorCount);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16}) {
    cudaLaunchConfig_t lc{}; lc.gridDim = dim3(cs * 64); lc.blockDim = dim3(256); lc.dynamicSmemBytes = 232448;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &lc);
    printf("cluster %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}
