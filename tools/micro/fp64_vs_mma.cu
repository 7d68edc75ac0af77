// Micro-benchmark: does work on the CUDA-core pipes (FP64, FP32, LDS) slow down while tcgen05.mma instructions stream
// on the same SM?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_vs_mma fp64_vs_mma.cu && ./fp64_vs_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// mode bits: 1 = run MMAs; work: 0 = DFMA, 1 = FFMA, 2 = LDS, 3 = nothing
__global__ void __launch_bounds__(288, 1) k(int mode, int work, int n_mma, int n_work, long long* out, double* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0 && (mode & 1)) {
            const uint32_t sa = smem_u32(smem), sb = sa + 16384;
            const uint64_t ad = umma_desc(sa, 2048, 128), bd = umma_desc(sb, 4096, 128);
            for (int i = 0; i < n_mma; ++i) {
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                             "l"(ad), "l"(bd), "r"(IDESC), "r"(i) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            out[blockIdx.x * 16 + 8] = clock64() - t0;
        }
    } else {
        double a0 = 1.0 + threadIdx.x * 1e-9, a1 = 1.1, a2 = 1.2, a3 = 1.3;
        float f0 = 1.0f + threadIdx.x * 1e-6f, f1 = 1.1f, f2 = 1.2f, f3 = 1.3f;
        const double m = 1.0000001, c = 1e-9;
        if (work == 0) {
            for (int i = 0; i < n_work; ++i) {
                a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            }
        } else if (work == 1) {
            for (int i = 0; i < n_work; ++i) {
                f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
            }
        } else if (work == 2) {
            const double* sp = reinterpret_cast<const double*>(smem + 49152) + threadIdx.x;
            for (int i = 0; i < n_work; ++i) {
                a0 += sp[(i & 7) * 256]; a1 += sp[((i + 1) & 7) * 256]; a2 += sp[((i + 2) & 7) * 256]; a3 += sp[((i + 3) & 7) * 256];
            }
        }
        sink[blockIdx.x * 288 + threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3;
        if (lane == 0) out[blockIdx.x * 16 + warp] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* out;
    double* sink;
    cudaMalloc(&out, 148 * 16 * sizeof(long long));
    cudaMalloc(&sink, 148 * 288 * sizeof(double));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
    const char* names[] = {"DFMA x4 per iter", "FFMA x4 per iter", "LDS.64 x4 per iter", "idle"};
    const int n_mma = 2000;
    for (int work = 0; work < 4; ++work) {
        const int n_work = work == 0 ? 8000 : (work == 1 ? 60000 : 20000);
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(out, 0, 148 * 16 * sizeof(long long));
            k<<<148, 288, 65536 + 8192>>>(mode, work, n_mma, n_work, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[16];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-20s mma=%d : worker warps %lld cycles (warp0), mma warp %lld cycles (%d MMAs -> %.1f cyc/MMA)\n", names[work], mode,
                   h[0], h[8], n_mma, mode ? (double)h[8] / n_mma : 0.0);
        }
    }
    return 0;
}
