// How many thread-block clusters of 2 / 4 / 8 CTAs (352 threads, ~150 KiB dynamic shared memory each, i.e. the footprint of
// k_screen_tc) can be resident on this GPU at once.   nvcc -arch=sm_100a -o cluster_occ cluster_occ.cu && ./cluster_occ
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(352, 1) k_dummy(int* out) {
    extern __shared__ unsigned char smem[];
    if (out && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
    const int smem = 150 * 1024;
    cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64);
        cfg.blockDim = dim3(352);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = cs;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
