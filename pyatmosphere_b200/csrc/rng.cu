// Counter-based generation of the random spectrum of every screen (production mode; parity runs feed
// coefficients drawn by numpy on the host instead).  One Philox4x32-10 block per (realization, screen, ring)
// gives theta and the two normals, a second block per (realization, screen) gives the single uniform number
// the reference shares between all annuli (grids.py:98-103).  Because the counter is
// (ring | flag, screen, realization) and the key is the user seed, the numbers do not depend on how
// realizations are spread over GPUs.
//
// Mirrors, in float32 like the reference: rho (grids.py:98-103), theta (grids.py:105-107),
// value = (n0 + i n1) sqrt(psd) (phase_screens.py:98-103), fx = rho cos(theta), fy = rho sin(theta)
// (grids.py:117-119).
#include "common.cuh"
#include "internal_rng.h"

namespace pa {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }   // [0,1)

// grid: (ceil(m/128), nscreens_per_real, batch)
__global__ void k_rng_spectrum(RngLaunch a) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    const int b = blockIdx.z;
    if (m >= a.m) return;
    const unsigned long long real = a.realization0 + (unsigned long long)b * a.realization_stride;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    uint32_t shared[4] = {0xFFFFFFFFu, (uint32_t)(a.screen0 + s), (uint32_t)real, (uint32_t)(real >> 32)};
    philox4x32_10(shared, k0, k1);
    uint32_t c[4] = {(uint32_t)m, (uint32_t)(a.screen0 + s), (uint32_t)real, (uint32_t)(real >> 32)};
    philox4x32_10(c, k0, k1);

    const float rnd = u01(shared[0]);
    const float f = a.base[m];
    const float fp = m > 0 ? a.base[m - 1] : 0.0f;
    const float fp2 = __fmul_rn(fp, fp);
    const float rho = __fsqrt_rn(__fadd_rn(fp2, __fmul_rn(rnd, __fsub_rn(__fmul_rn(f, f), fp2))));
    const float theta = __fmul_rn(6.2831855f, u01(c[0]));
    float st, ct;
    sincosf(theta, &st, &ct);
    const double u1 = ((double)c[1] + 1.0) * 2.3283064365386963e-10;    // (0,1]
    const double u2 = (double)c[2] * 2.3283064365386963e-10;             // [0,1)
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cn;
    sincospi(2.0 * u2, &sn, &cn);
    const float amp = __fsqrt_rn(a.psd[m]);
    const size_t o = ((size_t)s * a.batch + b) * a.m + m;     // [screen][batch][ring]
    a.fx[o] = __fmul_rn(rho, ct);
    a.fy[o] = __fmul_rn(rho, st);
    a.coef[o] = make_float2(__fmul_rn((float)(rad * cn), amp), __fmul_rn((float)(rad * sn), amp));
    if (a.rho) a.rho[o] = rho;
    if (a.theta) a.theta[o] = theta;
}

int launch_rng_spectrum(const RngLaunch& a, cudaStream_t st) {
    dim3 g((a.m + 127) / 128, a.nscreens, a.batch);
    k_rng_spectrum<<<g, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace pa
