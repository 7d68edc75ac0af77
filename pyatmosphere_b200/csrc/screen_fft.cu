// FFT phase screens (phase_screens.py:37-67 FFTPhaseScreen.generate_phase_screen):
//
//   screen = ifft2(cn, 1)  +  sum over the subharmonic terms  c_t exp(2 pi i (fx_t x + fy_t y))  -  mean
//
// `ifft2(cn, 1)` (utils.py:47-50) is ifftshift(IFFT2(ifftshift(cn))) * N^2, i.e. for even N the UNNORMALISED
// inverse DFT of  Y[p][q] = (-1)^(p+q) cn[(p + N/2) % N][(q + N/2) % N].  The transform itself is the inverse half
// of the split-step passes (k_cols<INV_ONLY> + k_rows<IN_PERM, !OUT_PERM>, fft_passes.cuh); this file holds the
// kernels around it: the gather of the centred spectrum into spectrum storage order, the subharmonic tables, the
// sum + mean reduction, and the final centring.
#include "common.cuh"
#include "internal_fftscreen.h"

namespace pa {
namespace {

constexpr int kGatherThreads = 256;
constexpr int kAddThreads = 256;

// ws[b][py][px] = (-1)^(perm[py] + perm[px]) * cn[b][(perm[py] + n/2) % n][(perm[px] + n/2) % n]
template <typename T>
__global__ void __launch_bounds__(kGatherThreads) k_fftscreen_gather(const cplx<T>* __restrict__ cn, cplx<T>* __restrict__ ws,
                                                                      const int* __restrict__ perm, int n) {
    const int px = blockIdx.x * kGatherThreads + threadIdx.x;
    const int py = blockIdx.y;
    if (px >= n) return;
    const int qy = perm[py], qx = perm[px];
    const size_t plane = (size_t)blockIdx.z * n * n;
    cplx<T> v = cn[plane + (size_t)((qy + n / 2) % n) * n + ((qx + n / 2) % n)];
    if ((qy + qx) & 1) {
        v.x = -v.x;
        v.y = -v.y;
    }
    ws[plane + (size_t)py * n + px] = v;
}

// EX[b][t][j] = exp(2 pi i fx_t x_j),  EY[b][t][i] = c_t exp(2 pi i fy_t y_i); terms[b][t] = {fx, fy, re c, im c}.
// The reference forms f*x + f*y in float32 (phase_screens.py:63-65); here the float32 axes are promoted exactly and
// the phase is evaluated in float64 (the float64 oracle's definition).
__global__ void k_fftscreen_tables(const double* __restrict__ terms, int nterms, const float* __restrict__ x,
                                   const float* __restrict__ y, int n, double2* __restrict__ ex, double2* __restrict__ ey) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const double* tm = terms + ((size_t)b * nterms + t) * 4;
    double s, c;
    sincospi(2.0 * tm[0] * (double)x[j], &s, &c);
    ex[((size_t)b * nterms + t) * n + j] = make_double2(c, s);
    sincospi(2.0 * tm[1] * (double)y[j], &s, &c);
    ey[((size_t)b * nterms + t) * n + j] = make_double2(tm[2] * c - tm[3] * s, tm[2] * s + tm[3] * c);
}

// out = ws + subharmonics; per-block sums of the result (float64) for the mean.  One block = kAddThreads columns
// of one row.
template <typename T>
__global__ void __launch_bounds__(kAddThreads) k_fftscreen_add(cplx<T>* __restrict__ ws, const double2* __restrict__ ex,
                                                                const double2* __restrict__ ey, int nterms, int n,
                                                                double2* __restrict__ partials) {
    const int j = blockIdx.x * kAddThreads + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    double re = 0.0, im = 0.0;
    if (j < n) {
        const size_t at = (size_t)b * n * n + (size_t)i * n + j;
        const cplx<T> v = ws[at];
        re = (double)v.x;
        im = (double)v.y;
        for (int t = 0; t < nterms; ++t) {
            const double2 a = ey[((size_t)b * nterms + t) * n + i];
            const double2 e = ex[((size_t)b * nterms + t) * n + j];
            re += a.x * e.x - a.y * e.y;
            im += a.x * e.y + a.y * e.x;
        }
        ws[at] = mkc<T>((T)re, (T)im);
    }
    __shared__ double sre[kAddThreads / 32], sim[kAddThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sre[threadIdx.x >> 5] = re;
        sim[threadIdx.x >> 5] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, q = 0.0;
        for (int w = 0; w < kAddThreads / 32; ++w) {
            r += sre[w];
            q += sim[w];
        }
        partials[((size_t)b * gridDim.y + i) * gridDim.x + blockIdx.x] = make_double2(r, q);
    }
}

// row sums of the block partials in a fixed order -> rowsum[b][i]
__global__ void k_fftscreen_rowsum(const double2* __restrict__ partials, int per_row, int n, double2* __restrict__ rowsum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n) return;
    double r = 0.0, q = 0.0;
    for (int k = 0; k < per_row; ++k) {
        const double2 p = partials[((size_t)b * n + i) * per_row + k];
        r += p.x;
        q += p.y;
    }
    rowsum[(size_t)b * n + i] = make_double2(r, q);
}

// subtract the mean (every block folds the n row sums in the same order -> deterministic) and write the outputs
template <typename T>
__global__ void __launch_bounds__(kAddThreads) k_fftscreen_center(const cplx<T>* __restrict__ ws, const double2* __restrict__ rowsum,
                                                                   int n, cplx<T>* __restrict__ out_c, T* __restrict__ out_re) {
    __shared__ double2 mean_s;
    const int b = blockIdx.z;
    if (threadIdx.x < 32) {
        double r = 0.0, q = 0.0;
        for (int k = threadIdx.x; k < n; k += 32) {
            const double2 p = rowsum[(size_t)b * n + k];
            r += p.x;
            q += p.y;
        }
        for (int o = 16; o > 0; o >>= 1) {
            r += __shfl_xor_sync(0xffffffffu, r, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (threadIdx.x == 0) mean_s = make_double2(r / ((double)n * n), q / ((double)n * n));
    }
    __syncthreads();
    const int j = blockIdx.x * kAddThreads + threadIdx.x;
    if (j >= n) return;
    const size_t at = (size_t)b * n * n + (size_t)blockIdx.y * n + j;
    const cplx<T> v = ws[at];
    const T re = (T)((double)v.x - mean_s.x), im = (T)((double)v.y - mean_s.y);
    if (out_c) out_c[at] = mkc<T>(re, im);
    if (out_re) out_re[at] = re;
}

template <typename T> int gather_t(const FftScreenLaunch& a, cudaStream_t st) {
    const dim3 grid((a.n + kGatherThreads - 1) / kGatherThreads, a.n, a.nscreens);
    k_fftscreen_gather<T><<<grid, kGatherThreads, 0, st>>>((const cplx<T>*)a.spectrum, (cplx<T>*)a.ws, a.perm, a.n);
    return (int)cudaGetLastError();
}

template <typename T> int finish_t(const FftScreenLaunch& a, cudaStream_t st) {
    const int per_row = (a.n + kAddThreads - 1) / kAddThreads;
    if (a.nterms > 0) {
        const dim3 tg((a.n + 127) / 128, a.nterms, a.nscreens);
        k_fftscreen_tables<<<tg, 128, 0, st>>>(a.terms, a.nterms, a.x, a.y, a.n, a.ex, a.ey);
    }
    const dim3 grid(per_row, a.n, a.nscreens);
    k_fftscreen_add<T><<<grid, kAddThreads, 0, st>>>((cplx<T>*)a.ws, a.ex, a.ey, a.nterms, a.n, a.partials);
    const dim3 rg((a.n + 127) / 128, a.nscreens);
    k_fftscreen_rowsum<<<rg, 128, 0, st>>>(a.partials, per_row, a.n, a.rowsum);
    k_fftscreen_center<T><<<grid, kAddThreads, 0, st>>>((const cplx<T>*)a.ws, a.rowsum, a.n, (cplx<T>*)a.out_complex, (T*)a.out_real);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_fftscreen_gather(int prec, const FftScreenLaunch& a, cudaStream_t st) {
    return prec == 0 ? gather_t<float>(a, st) : gather_t<double>(a, st);
}
int launch_fftscreen_finish(int prec, const FftScreenLaunch& a, cudaStream_t st) {
    return prec == 0 ? finish_t<float>(a, st) : finish_t<double>(a, st);
}
int fftscreen_finish_launches(const FftScreenLaunch& a) { return a.nterms > 0 ? 4 : 3; }

}  // namespace pa
