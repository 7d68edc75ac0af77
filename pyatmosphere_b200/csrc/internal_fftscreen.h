// Internal launch descriptor of the FFT phase-screen kernels (screen_fft.cu).
#pragma once
#include <cuda_runtime.h>

namespace pa {

struct FftScreenLaunch {
    int n, nscreens;
    const void* spectrum;   // [nscreens][n][n] complex, centred frequency order (index n/2 = zero frequency)
    void* ws;               // [nscreens][n][n] complex workspace: spectrum in storage order, then the transform
    const int* perm;        // [n] frequency held at storage position p (pa_ctx_permutation), device copy
    const double* terms;    // [nscreens][nterms][4] {fx, fy, re c, im c}, device, sorted by group; may be null when nterms == 0
    int nterms;
    const int* goff;        // [nscreens][ngroups + 1] first term of every group of equal fx (device)
    int ngroups;
    const float* x;         // float32 axes of the context
    const float* y;
    double2* ex;            // workspace [nscreens][ngroups][n]
    double2* ey;            // workspace [nscreens][ngroups][n]
    double2* partials;      // workspace [nscreens][n][ceil(n / 256)]
    double2* rowsum;        // workspace: mean of every screen ([nscreens] used)
    void* out_complex;      // [nscreens][n][n] complex or null
    void* out_real;         // [nscreens][n][n] real or null
};

int launch_fftscreen_gather(int prec, const FftScreenLaunch& a, cudaStream_t st);
int launch_fftscreen_finish(int prec, const FftScreenLaunch& a, cudaStream_t st);
int fftscreen_finish_launches(const FftScreenLaunch& a);
// centred spectrum / field -> storage order with the (-1)^(p+q) modulation, optionally conjugated (pa_fft2c)
int launch_fft2c_gather(int prec, const void* in, void* ws, const int* perm, int n, int batch, bool conj, cudaStream_t st);

}  // namespace pa
