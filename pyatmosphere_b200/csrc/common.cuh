// Shared helpers for the pyatmosphere_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace pa {

// ---- error plumbing (status codes cross the C ABI, text stays here) --------------------------------
void set_error(const char* fmt, ...);
const char* last_error();

#define PA_OK 0
#define PA_ERR_ARG 1
#define PA_ERR_CUDA 2
#define PA_ERR_STATE 3

#define PA_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            pa::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return PA_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

#define PA_REQUIRE(cond, ...)                                                                       \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            pa::set_error(__VA_ARGS__);                                                             \
            return PA_ERR_ARG;                                                                      \
        }                                                                                           \
    } while (0)

// ---- complex arithmetic on float2 / double2 ---------------------------------------------------------
template <typename T> struct cx2;
template <> struct cx2<float> { using type = float2; };
template <> struct cx2<double> { using type = double2; };
template <typename T> using cplx = typename cx2<T>::type;

template <typename T> __host__ __device__ __forceinline__ cplx<T> mkc(T a, T b) {
    cplx<T> r;
    r.x = a;
    r.y = b;
    return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
#if defined(PA_F32X2) && defined(__CUDA_ARCH__)
// sm_100a packed single precision: one FADD2 / FFMA2 per complex add / subtract (halves the issue slots of the
// butterflies, which are ~40 % of the instructions of the FFT passes).
__device__ __forceinline__ unsigned long long f2_bits(float2 a) { return (unsigned long long)__float_as_uint(a.x) | ((unsigned long long)__float_as_uint(a.y) << 32); }
__device__ __forceinline__ float2 bits_f2(unsigned long long r) { return make_float2(__uint_as_float((unsigned)r), __uint_as_float((unsigned)(r >> 32))); }
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) {
    unsigned long long r;
    const unsigned long long m1 = 0xBF800000BF800000ull;     // (-1.0f, -1.0f)
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(b)), "l"(m1), "l"(f2_bits(a)));
    return bits_f2(r);
}
#endif
// a * b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.y * b.x - a.x * b.y;
    return r;
}
// a * (-i)
template <typename C> __device__ __forceinline__ C cmul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
// a * (+i)
template <typename C> __device__ __forceinline__ C cmul_pi(C a) { C r; r.x = -a.y; r.y = a.x; return r; }
template <typename C> __device__ __forceinline__ C cswap(C a) { C r; r.x = a.y; r.y = a.x; return r; }

// cudaFuncSetAttribute is a per-device setting: remember on which devices a kernel's dynamic shared-memory limit has
// been raised (one process may hold contexts on several GPUs even though the intended model is one process per GPU).
struct SmemOptIn {
    std::atomic<unsigned long long> devices{0};
    template <typename K> cudaError_t raise(K kern, int bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = 1ull << (dev & 63);
        if (devices.load(std::memory_order_relaxed) & bit) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) devices.fetch_or(bit, std::memory_order_relaxed);
        return e;
    }
};

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

}  // namespace pa
