// Variants of the streaming outer stage of the split column pass (fft_split.cuh: k_col_outer) on one 8192^2 complex64 field:
// how far is the shipped kernel from the copy bandwidth, and would another shape of it get closer?
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -I pyatmosphere_b200/csrc -o tools/micro/outer_stage_probe tools/micro/outer_stage_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <vector>

#include "fft_split.cuh"

using namespace pa;

// the shipped radix-32 stage with other CTA shapes / residency / cache hints
template <int THREADS, int MINB, bool STREAM> __global__ void __launch_bounds__(THREADS, MINB) outer32(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    const cplx<float>* w = otw + t * 32;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) v[a] = STREAM ? __ldcs(p + a * STEP) : p[a * STEP];
    dft32_fwd<float>(v);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int j = freq32(r);
        const cplx<float> o = j == 0 ? v[r] : cmul(v[r], ldg_c<float>(w + j));
        if (STREAM) __stcs(p + j * STEP, o);
        else p[j * STEP] = o;
    }
}

template <int THREADS, int MINB, bool STREAM> __global__ void __launch_bounds__(THREADS, MINB) outer32_inv(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    const cplx<float>* w = otw + t * 32;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int j = freq32(r);
        const cplx<float> x = STREAM ? __ldcs(p + j * STEP) : p[j * STEP];
        v[r] = j == 0 ? x : cmulc(x, ldg_c<float>(w + j));
    }
    dft32_inv<float>(v);
#pragma unroll
    for (int a = 0; a < 32; ++a) {
        if (STREAM) __stcs(p + a * STEP, v[a]);
        else p[a * STEP] = v[a];
    }
}

// inverse with every field load issued before the first twiddle is touched
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) outer32_inv2(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    const cplx<float>* w = otw + t * 32;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = __ldcs(p + freq32(r) * STEP);
    asm volatile("" ::: "memory");
#pragma unroll
    for (int r = 1; r < 32; ++r) {
        const int j = freq32(r);
        v[r] = cmulc(v[r], ldg_c<float>(w + j));
    }
    dft32_inv<float>(v);
#pragma unroll
    for (int a = 0; a < 32; ++a) __stcs(p + a * STEP, v[a]);
}

// inverse through the forward butterfly: x'[a] = conj( DFT32( conj(z_j) w_j ) )[a]; natural-order loads, permuted stores
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) outer32_inv3(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    const cplx<float>* w = otw + t * 32;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        cplx<float> x = __ldcs(p + j * STEP);
        x.y = -x.y;
        v[j] = j == 0 ? x : cmul(x, ldg_c<float>(w + j));
    }
    dft32_fwd<float>(v);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        cplx<float> o = v[r];
        o.y = -o.y;
        __stcs(p + freq32(r) * STEP, o);
    }
}

// inverse with the CTA's 32 twiddles staged in shared memory
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) outer32_inv4(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    __shared__ cplx<float> sw[32];
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = __ldcs(p + freq32(r) * STEP);
    if (threadIdx.x < 32) sw[threadIdx.x] = otw[t * 32 + threadIdx.x];
    __syncthreads();
#pragma unroll
    for (int r = 1; r < 32; ++r) v[r] = cmulc(v[r], sw[freq32(r)]);
    dft32_inv<float>(v);
#pragma unroll
    for (int a = 0; a < 32; ++a) __stcs(p + a * STEP, v[a]);
}
// forward with the twiddles in shared memory too
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) outer32_fwd4(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 32;
    __shared__ cplx<float> sw[32];
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) v[a] = __ldcs(p + a * STEP);
    if (threadIdx.x < 32) sw[threadIdx.x] = otw[t * 32 + threadIdx.x];
    __syncthreads();
    dft32_fwd<float>(v);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int j = freq32(r);
        __stcs(p + j * STEP, j == 0 ? v[r] : cmul(v[r], sw[j]));
    }
}

// radix-16 outer stage (8192 = 16 * 512): 16 values per thread, half the registers
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) outer16(cplx<float>* __restrict__ field, const cplx<float>* __restrict__ otw) {
    constexpr int N = 8192, M = N / 16;
    const int t = blockIdx.y;
    cplx<float>* p = field + (size_t)t * N + blockIdx.x * THREADS + threadIdx.x;
    const cplx<float>* w = otw + (t % 256) * 32;
    constexpr size_t STEP = (size_t)M * N;
    cplx<float> v[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) v[a] = p[a * STEP];
    dftR<float, 16, false, 0, 16>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j * STEP] = j == 0 ? v[j] : cmul(v[j], ldg_c<float>(w + j));
}

__global__ void copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    constexpr int N = 8192;
    cplx<float>*field, *other;
    cplx<float>* otw;
    const size_t bytes = (size_t)N * N * sizeof(cplx<float>);
    cudaMalloc(&field, bytes);
    cudaMalloc(&other, bytes);
    cudaMemset(field, 0, bytes);
    std::vector<float> tw(2 * 256 * 32, 0.5f);
    cudaMalloc(&otw, tw.size() * sizeof(float));
    cudaMemcpy(otw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            launch();
            cudaEventRecord(e0);
            for (int i = 0; i < 10; ++i) launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("%-44s %7.1f us per sweep, %5.0f GB/s  (%s)\n", name, ms * 100, 2.0 * bytes / (ms * 1e-4) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    time("copy (float4, grid-stride)", [&] { copy_kernel<<<148 * 16, 256>>>((const float4*)field, (float4*)other, bytes / 16); });
    time("cudaMemcpyAsync device to device", [&] { cudaMemcpyAsync(other, field, bytes, cudaMemcpyDeviceToDevice); });
    time("shipped k_col_outer (256 thr, 2 CTAs/SM)", [&] { k_col_outer<float, N, false><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("radix 32, 256 thr, 2/SM, ld.cs/st.cs", [&] { outer32<256, 2, true><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("shipped k_col_outer INVERSE", [&] { k_col_outer<float, N, true><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse, 256 thr, 2/SM, plain", [&] { outer32_inv<256, 2, false><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse, 256 thr, 2/SM, ld.cs/st.cs", [&] { outer32_inv<256, 2, true><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse, 128 thr, 4/SM, ld.cs/st.cs", [&] { outer32_inv<128, 4, true><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("inverse, loads first, 256 thr, 2/SM", [&] { outer32_inv2<256, 2><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse, loads first, 128 thr, 4/SM", [&] { outer32_inv2<128, 4><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("inverse via forward butterfly, 256 thr, 2/SM", [&] { outer32_inv3<256, 2><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse via forward butterfly, 128 thr, 4/SM", [&] { outer32_inv3<128, 4><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("inverse, smem twiddles, 256 thr, 2/SM", [&] { outer32_inv4<256, 2><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("inverse, smem twiddles, 128 thr, 4/SM", [&] { outer32_inv4<128, 4><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("forward, smem twiddles, 128 thr, 4/SM", [&] { outer32_fwd4<128, 4><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("forward, 128 thr, 4/SM, ld.cs/st.cs", [&] { outer32<128, 4, true><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("radix 32, 128 thr, 4/SM", [&] { outer32<128, 4, false><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("radix 32, 128 thr, 5/SM (<= 102 regs)", [&] { outer32<128, 5, false><<<dim3(N / 128, 256, 1), 128>>>(field, otw); });
    time("radix 32, 256 thr, 3/SM (<= 85 regs)", [&] { outer32<256, 3, false><<<dim3(N / 256, 256, 1), 256>>>(field, otw); });
    time("radix 32, 512 thr, 1/SM", [&] { outer32<512, 1, false><<<dim3(N / 512, 256, 1), 512>>>(field, otw); });
    time("radix 16, 256 thr, 4/SM", [&] { outer16<256, 4><<<dim3(N / 256, 512, 1), 256>>>(field, otw); });
    time("radix 16, 256 thr, 6/SM", [&] { outer16<256, 6><<<dim3(N / 256, 512, 1), 256>>>(field, otw); });
    time("radix 16, 128 thr, 8/SM", [&] { outer16<128, 8><<<dim3(N / 128, 512, 1), 128>>>(field, otw); });
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
