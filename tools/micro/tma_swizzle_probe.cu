// Where does the TMA engine put the elements of a box whose inner extent (32 bytes) is smaller than the swizzle span?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_swizzle_probe tma_swizzle_probe.cu -lcuda && ./tma_swizzle_probe
// Loads a box of {8 floats, 64 rows} from a [256 rows][64 floats] tensor whose element (r, c) holds r * 64 + c with every
// swizzle mode, dumps shared memory and prints, for the first rows, the byte offset at which each 16-byte chunk landed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int words) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sm = reinterpret_cast<float*>(smem);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar = base + words * 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = -1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8 * 4 * 64) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(base),
                     "l"(&tmap), "r"(16), "r"(32), "r"(bar)
                     : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = sm[i];
    if (threadIdx.x == 0) out[words] = ok ? 1.0f : 0.0f;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int R = 256, Cn = 64, words = 16384;     // 64 KiB dump
    std::vector<float> h((size_t)R * Cn);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < Cn; ++c) h[(size_t)r * Cn + c] = (float)(r * Cn + c);
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, (words + 1) * 4);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4 + 64);
    const CUtensorMapSwizzle modes[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
    const char* names[4] = {"NONE", "32B", "64B", "128B"};
    for (int m = 0; m < 4; ++m) {
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)Cn, (cuuint64_t)R};
        const cuuint64_t strides[1] = {(cuuint64_t)Cn * 4};
        const cuuint32_t box[2] = {8, 64};
        const cuuint32_t estr[2] = {1, 1};
        CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("swizzle %s: encode rc %d\n", names[m], (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        cudaMemset(out, 0, (words + 1) * 4);
        probe<<<1, 256, words * 4 + 64>>>(tm, out, words);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> o(words + 1);
        cudaMemcpy(o.data(), out, (words + 1) * 4, cudaMemcpyDeviceToHost);
        printf("  completed %g; bytes written: ", o[words]);
        int last = -1, count = 0;
        for (int i = 0; i < words; ++i)
            if (o[i] >= 0) { last = i; ++count; }
        printf("%d words, last word index %d\n", count, last);
        // element (row 32 + rr, col 16 + cc) = (32+rr)*64 + 16 + cc: print the 16-byte chunk offsets of the first 20 rows
        for (int rr = 0; rr < 20; ++rr) {
            printf("  row %2d:", rr);
            for (int half = 0; half < 2; ++half) {
                const float want = (float)((32 + rr) * Cn + 16 + half * 4);
                int at = -1;
                for (int i = 0; i < words; i += 4)
                    if (o[i] == want) { at = i * 4; break; }
                printf(" chunk%d@%5d", half, at);
            }
            printf("   (linear would be %d, %d)\n", rr * 32, rr * 32 + 16);
        }
    }
    return 0;
}
