// How fast does the streaming outer stage of the split column pass (k_col_outer, fft_split.cuh) run when its slab is
// L2-resident, and how much does a launch per slab cost?
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -I pyatmosphere_b200/csrc -o tools/micro/l2_slab_probe tools/micro/l2_slab_probe.cu
// Times forward + inverse outer stage (a) once over a whole 8192^2 complex64 field (512 MiB, HBM), (b) alternating on ONE slab
// of W columns (the slab is in L2 after the first launch), (c) walking over all slabs of the field, slab by slab.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <vector>

#include "fft_split.cuh"

using namespace pa;

int main() {
    constexpr int N = 8192, M = N / 32;
    cplx<float>* field;
    cplx<float>* otw;
    cudaMalloc(&field, (size_t)N * N * sizeof(cplx<float>));
    cudaMemset(field, 0, (size_t)N * N * sizeof(cplx<float>));
    std::vector<float> tw(2 * M * 32);
    for (int t = 0; t < M; ++t)
        for (int j = 0; j < 32; ++j) {
            tw[2 * (t * 32 + j)] = (float)cos(-2 * M_PI * t * j / N);
            tw[2 * (t * 32 + j) + 1] = (float)sin(-2 * M_PI * t * j / N);
        }
    cudaMalloc(&otw, tw.size() * sizeof(float));
    cudaMemcpy(otw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto pair = [&](int x0, int W) {
        const dim3 g(W / 256, M, 1);
        k_col_outer<float, N, false><<<g, 256>>>(field + x0, otw);
        k_col_outer<float, N, true><<<g, 256>>>(field + x0, otw);
    };
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        pair(0, N);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) pair(0, N);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("whole field (HBM): %.1f us per sweep, %.0f GB/s\n", ms * 1e3 / 20, 2.0 * N * N * 8 / (ms * 1e-3 / 20) / 1e9);
    for (int W : {256, 512, 1024, 2048}) {
        const double bytes = 2.0 * W * N * 8;
        pair(0, W);
        cudaEventRecord(e0);
        for (int i = 0; i < 50; ++i) pair(0, W);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double hot = ms * 1e3 / 100;
        cudaEventRecord(e0);
        for (int i = 0; i < 4; ++i)
            for (int x0 = 0; x0 < N; x0 += W) pair(x0, W);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double walk = ms * 1e3 / (8.0 * N / W);
        printf("slab %4d columns (%3.0f MiB): L2-hot %.2f us per sweep (%.0f GB/s); walking the field: %.2f us per sweep (%.0f GB/s, "
               "the inverse sweep of each pair finds the slab in L2)\n",
               W, W * (double)N * 8 / 1048576, hot, bytes / hot / 1e3, walk, bytes / walk / 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
