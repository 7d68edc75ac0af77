// FFT phase screens (phase_screens.py:37-67 FFTPhaseScreen.generate_phase_screen):
//
//   screen = ifft2(cn, 1)  +  sum over the subharmonic terms  c_t exp(2 pi i (fx_t x + fy_t y))  -  mean
//
// `ifft2(cn, 1)` (utils.py:47-50) is ifftshift(IFFT2(ifftshift(cn))) * N^2, i.e. for even N the UNNORMALISED
// inverse DFT of  Y[p][q] = (-1)^(p+q) cn[(p + N/2) % N][(q + N/2) % N].  The transform itself is the inverse half
// of the split-step passes (k_cols<INV_ONLY> + k_rows<IN_PERM, !OUT_PERM>, fft_passes.cuh); this file holds the
// kernels around it: the gather of the centred spectrum into spectrum storage order, the subharmonic tables, the
// sum + mean reduction, and the final centring (7 launches per batch of screens).
#include "common.cuh"
#include "internal_fftscreen.h"

namespace pa {
namespace {

constexpr int kGatherThreads = 256;
constexpr int kAddThreads = 256;

// ws[b][py][px] = (-1)^(perm[py] + perm[px]) * cn[b][(perm[py] + n/2) % n][(perm[px] + n/2) % n]
// CONJ: the conjugate of the input is gathered (the forward transform of pa_fft2c is conj(inverse(conj(.))))
template <typename T, bool CONJ = false>
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
    if (CONJ) v.y = -v.y;
    ws[plane + (size_t)py * n + px] = v;
}

// The terms of a screen are grouped by their x-frequency (the host sorts them: terms[b][goff[b][g] .. goff[b][g+1]) share
// fx), because  sum_t c_t e^{2 pi i (fx_t x + fy_t y)} = sum_g e^{2 pi i fx_g x} G_g(y),  G_g(y) = sum_{t in g} c_t e^{2 pi i fy_t y}:
// a 3 x 3 subharmonic patch costs 3 complex multiply-adds per pixel instead of 8.
//   EX[b][g][j] = exp(2 pi i fx_g x_j),   GY[b][g][i] = G_g(y_i);   terms[b][t] = {fx, fy, re c, im c}.
// The reference forms f*x + f*y in float32 (phase_screens.py:63-65); here the float32 axes are promoted exactly and
// the phase is evaluated in float64 (the float64 oracle's definition).
__global__ void k_fftscreen_tables(const double* __restrict__ terms, const int* __restrict__ goff, int nterms, int ngroups,
                                   const float* __restrict__ x, const float* __restrict__ y, int n, double2* __restrict__ ex,
                                   double2* __restrict__ gy) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const int t0 = goff[b * (ngroups + 1) + g], t1 = goff[b * (ngroups + 1) + g + 1];
    const double* tm = terms + (size_t)b * nterms * 4;
    double s, c;
    sincospi(2.0 * (t0 < t1 ? tm[t0 * 4] : 0.0) * (double)x[j], &s, &c);
    ex[((size_t)b * ngroups + g) * n + j] = make_double2(c, s);
    const double yv = (double)y[j];
    double re = 0.0, im = 0.0;
    for (int t = t0; t < t1; ++t) {
        sincospi(2.0 * tm[t * 4 + 1] * yv, &s, &c);
        re += tm[t * 4 + 2] * c - tm[t * 4 + 3] * s;
        im += tm[t * 4 + 2] * s + tm[t * 4 + 3] * c;
    }
    gy[((size_t)b * ngroups + g) * n + j] = make_double2(re, im);
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

// mean[b] = (sum of all block partials of screen b) / n^2, folded in a fixed order by ONE block per screen: thread k
// sums the partials of rows k, k + 1024, ... and a shared-memory tree adds the 1024 thread sums -> deterministic.
constexpr int kMeanThreads = 1024;
__global__ void __launch_bounds__(kMeanThreads) k_fftscreen_mean(const double2* __restrict__ partials, int per_row, int n,
                                                                  double2* __restrict__ mean) {
    __shared__ double sre[kMeanThreads], sim[kMeanThreads];
    const int b = blockIdx.x;
    double r = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += kMeanThreads)
        for (int k = 0; k < per_row; ++k) {
            const double2 p = partials[((size_t)b * n + i) * per_row + k];
            r += p.x;
            q += p.y;
        }
    sre[threadIdx.x] = r;
    sim[threadIdx.x] = q;
    __syncthreads();
    for (int o = kMeanThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sre[threadIdx.x] += sre[threadIdx.x + o];
            sim[threadIdx.x] += sim[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[b] = make_double2(sre[0] / ((double)n * n), sim[0] / ((double)n * n));
}

// subtract the mean and write the outputs
template <typename T>
__global__ void __launch_bounds__(kAddThreads) k_fftscreen_center(const cplx<T>* __restrict__ ws, const double2* __restrict__ mean,
                                                                   int n, cplx<T>* __restrict__ out_c, T* __restrict__ out_re) {
    const int b = blockIdx.z;
    const double2 m = mean[b];
    const int j = blockIdx.x * kAddThreads + threadIdx.x;
    if (j >= n) return;
    const size_t at = (size_t)b * n * n + (size_t)blockIdx.y * n + j;
    const cplx<T> v = ws[at];
    const T re = (T)((double)v.x - m.x), im = (T)((double)v.y - m.y);
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
    if (a.ngroups > 0) {
        const dim3 tg((a.n + 127) / 128, a.ngroups, a.nscreens);
        k_fftscreen_tables<<<tg, 128, 0, st>>>(a.terms, a.goff, a.nterms, a.ngroups, a.x, a.y, a.n, a.ex, a.ey);
    }
    const dim3 grid(per_row, a.n, a.nscreens);
    k_fftscreen_add<T><<<grid, kAddThreads, 0, st>>>((cplx<T>*)a.ws, a.ex, a.ey, a.ngroups, a.n, a.partials);
    k_fftscreen_mean<<<a.nscreens, kMeanThreads, 0, st>>>(a.partials, per_row, a.n, a.rowsum);
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
int fftscreen_finish_launches(const FftScreenLaunch& a) { return a.ngroups > 0 ? 4 : 3; }

int launch_fft2c_gather(int prec, const void* in, void* ws, const int* perm, int n, int batch, bool conj, cudaStream_t st) {
    const dim3 grid((n + kGatherThreads - 1) / kGatherThreads, n, batch);
    if (prec == 0) {
        if (conj) k_fftscreen_gather<float, true><<<grid, kGatherThreads, 0, st>>>((const cplx<float>*)in, (cplx<float>*)ws, perm, n);
        else k_fftscreen_gather<float, false><<<grid, kGatherThreads, 0, st>>>((const cplx<float>*)in, (cplx<float>*)ws, perm, n);
    } else {
        if (conj) k_fftscreen_gather<double, true><<<grid, kGatherThreads, 0, st>>>((const cplx<double>*)in, (cplx<double>*)ws, perm, n);
        else k_fftscreen_gather<double, false><<<grid, kGatherThreads, 0, st>>>((const cplx<double>*)in, (cplx<double>*)ws, perm, n);
    }
    return (int)cudaGetLastError();
}

}  // namespace pa
