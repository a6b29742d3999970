// How many thread-block clusters of 8 / 16 CTAs with ~220 KB of shared memory per CTA can be resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k8(int* p) { extern __shared__ double s[]; if (p) p[0] = (int)s[0]; }
int main() {
  for (int cs : {2, 4, 8, 16}) {
    for (int smem : {100 * 1024, 200 * 1024, 227 * 1024}) {
      cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (cs > 8) cudaFuncSetAttribute(k8, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k8, &cfg);
      printf("cluster %2d smem %3d KB: max active clusters %d (%s)\n", cs, smem / 1024, n, cudaGetErrorString(e));
    }
  }
  return 0;
}
