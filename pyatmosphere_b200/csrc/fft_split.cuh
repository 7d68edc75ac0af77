// Split column pass for grids whose columns are too long for a multi-column shared-memory tile (N = 8192).
//
// A column transform of length N = R0 * M is the first decimation-in-frequency stage
//     y_j[t] = W_N^(t j) * sum_a x[t + M a] W_R0^(a j),      t < M, j < R0,
// followed by R0 independent M-point transforms of the contiguous row blocks [M j, M j + M):
//     X[j + R0 q] = sum_t y_j[t] W_M^(t q).
// The outer stage touches rows that are M apart, so it needs no shared memory at all: `k_col_outer` streams the
// field once with one thread per (column, t), fully coalesced along x.  The inner part - M-point transform,
// transfer function, inverse M-point transform - is the ordinary TMA-fed column kernel run on [M rows] x [32
// columns] tiles (256-byte row segments), with the transfer-function factor indexed in the composite order
// storage row M j + s  <->  frequency j + R0 * perm_M[s].  The inverse outer stage streams the field once more.
// Three sweeps instead of one, each at streaming speed, against 8-byte accesses in the single-column kernel.
// complex128 uses the same factorisation (32 x 256, 16-column tiles of 256 rows; the 32 double-precision values of the
// outer butterfly take 128 registers, so that kernel runs one CTA of 256 threads per SM partition less).
#pragma once
#include "fft_core.cuh"

#ifndef PA_OUTER_THREADS_F64
#define PA_OUTER_THREADS_F64 64
#endif

namespace pa {

// exp(-2 pi i k / 32), k = 0..15; the switch folds once the calling loop is unrolled
template <typename T> __device__ __forceinline__ cplx<T> w32(int k) {
    switch (k) {
        case 1: return mkc<T>((T)0.98078528040323043058L, (T)-0.19509032201612824808L);
        case 2: return mkc<T>((T)0.92387953251128673848L, (T)-0.38268343236508978178L);
        case 3: return mkc<T>((T)0.83146961230254523567L, (T)-0.55557023301960217765L);
        case 4: return mkc<T>((T)0.70710678118654757274L, (T)-0.70710678118654746172L);
        case 5: return mkc<T>((T)0.55557023301960228867L, (T)-0.83146961230254523567L);
        case 6: return mkc<T>((T)0.38268343236508983729L, (T)-0.92387953251128673848L);
        case 7: return mkc<T>((T)0.19509032201612833135L, (T)-0.98078528040323043058L);
        case 8: return mkc<T>((T)0, (T)-1);
        case 9: return mkc<T>((T)-0.19509032201612819257L, (T)-0.98078528040323043058L);
        case 10: return mkc<T>((T)-0.38268343236508972627L, (T)-0.92387953251128673848L);
        case 11: return mkc<T>((T)-0.55557023301960195560L, (T)-0.83146961230254545772L);
        case 12: return mkc<T>((T)-0.70710678118654746172L, (T)-0.70710678118654757274L);
        case 13: return mkc<T>((T)-0.83146961230254534669L, (T)-0.55557023301960217765L);
        case 14: return mkc<T>((T)-0.92387953251128673848L, (T)-0.38268343236508989280L);
        case 15: return mkc<T>((T)-0.98078528040323043058L, (T)-0.19509032201612860891L);
        default: return mkc<T>((T)1, (T)0);
    }
}

// register r of the 32-point transforms below holds frequency freq32(r): evens in the lower half, odds in the upper
__host__ __device__ constexpr int freq32(int r) { return r < 16 ? 2 * r : 2 * (r - 16) + 1; }

// forward 32-point DFT: natural input, output X[freq32(r)] in register r
template <typename T> __device__ __forceinline__ void dft32_fwd(cplx<T> (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const cplx<T> a = v[i], b = v[i + 16];
        v[i] = cadd(a, b);
        const cplx<T> d = csub(a, b);
        v[i + 16] = i == 0 ? d : cmul(d, w32<T>(i));
    }
    dftR<T, 16, false, 0, 32>(v);
    dftR<T, 16, false, 16, 32>(v);
}
// unnormalised inverse: input X[freq32(r)] in register r, natural output (times 32)
template <typename T> __device__ __forceinline__ void dft32_inv(cplx<T> (&v)[32]) {
    dftR<T, 16, true, 0, 32>(v);
    dftR<T, 16, true, 16, 32>(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const cplx<T> u = v[i];
        const cplx<T> w = i == 0 ? v[16] : cmulc(v[i + 16], w32<T>(i));
        v[i] = cadd(u, w);
        v[i + 16] = csub(u, w);
    }
}

// Outer stage of the split column transform, in place.  grid = (N / THREADS, M, batch), one thread per column; otw[t * 32 + j] = exp(-2 pi i t j / N).  A CTA works on one t, so its 32 twiddles are staged in shared memory
// once: as per-thread global loads they miss in an L1 that the streaming field keeps flushing, and the inverse stage, which
// needs them before its butterfly, ran at 4.7 TB/s instead of 6.5 (tools/micro/outer_stage_probe.cu).  The field is read
// and written exactly once per sweep: ld.cs / st.cs.
// CTA shape: 256 threads x 2 per SM (complex64, 122-126 registers); complex128 holds its 32 values in 128 registers
// (210-226 in all) and runs 64 threads x 4 per SM -- small CTAs whose load, butterfly and store phases interleave: column
// pass at 8192^2 complex128 1251 / 1180 / 1150 us with 256 / 128 / 64 threads.
template <typename T> struct OuterGeo { static constexpr int THREADS = sizeof(T) == 4 ? 256 : PA_OUTER_THREADS_F64; };
template <typename T, int N, bool INV>
__global__ void __launch_bounds__(OuterGeo<T>::THREADS, sizeof(T) == 4 ? 2 : 256 / OuterGeo<T>::THREADS)
k_col_outer(cplx<T>* __restrict__ field, const cplx<T>* __restrict__ otw) {
    using C = cplx<T>;
    constexpr int R0 = 32, M = N / R0;
    __shared__ C sw[R0];
    const int t = blockIdx.y;
    C* p = field + ((size_t)blockIdx.z * N + t) * N + blockIdx.x * OuterGeo<T>::THREADS + threadIdx.x;
    constexpr size_t STEP = (size_t)M * N;
    C v[32];
    if constexpr (!INV) {
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = __ldcs(p + a * STEP);
        if (threadIdx.x < R0) sw[threadIdx.x] = otw[t * R0 + threadIdx.x];
        __syncthreads();
        dft32_fwd<T>(v);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int j = freq32(r);
            __stcs(p + j * STEP, j == 0 ? v[r] : cmul(v[r], sw[j]));
        }
    } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = __ldcs(p + freq32(r) * STEP);
        if (threadIdx.x < R0) sw[threadIdx.x] = otw[t * R0 + threadIdx.x];
        __syncthreads();
#pragma unroll
        for (int r = 1; r < 32; ++r) v[r] = cmulc(v[r], sw[freq32(r)]);
        dft32_inv<T>(v);
#pragma unroll
        for (int a = 0; a < 32; ++a) __stcs(p + a * STEP, v[a]);
    }
}

}  // namespace pa
