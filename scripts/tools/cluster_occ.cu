// Prints cudaOccupancyMaxActiveClusters for a dummy kernel over cluster sizes / shared-memory sizes / block sizes.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int *p) { extern __shared__ int sm[]; if (p) p[0] = sm[0]; }
int main() {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int smems[] = {48 * 1024, 97 * 1024, 110 * 1024, 160 * 1024, 200 * 1024, 225 * 1024};
    const int blocks[] = {128, 256, 512};
    for (int cs : {1, 2, 4, 8, 16})
        for (int sm : smems)
            for (int bs : blocks) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(bs); cfg.dynamicSmemBytes = sm;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int n = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
                printf("cluster %2d smem %3d KB block %3d -> max active clusters %d (%d CTAs) %s\n", cs, sm / 1024, bs, n, n * cs,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
