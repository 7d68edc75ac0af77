// Internal launch descriptor of the phase-screen kernels.
#pragma once
#include <cuda_runtime.h>

namespace pa {

constexpr int kMaxPolyDegree = 63;

struct ScreenLaunch {
    int n;              // grid size
    int m;              // harmonics per screen
    int m_split;        // harmonics [0, m_split) go to the polynomial, [m_split, m) to the contraction
    int degree;         // total degree of the polynomial, -1 = none
    int nscreens;       // screens in this launch (batch)
    const float* x;     // [n] float32 axes as the reference builds them (grids.py:63-69)
    const float* y;
    float shift_x, shift_y;
    double x0, y0, inv_x0, inv_y0;   // normalisation of the polynomial variables: xh = xs / x0
    const float* fx;    // [nscreens][m]
    const float* fy;    // [nscreens][m]
    const float2* coef; // [nscreens][m]
    double* P;          // workspace [nscreens][2 (m - m_split)][n]
    double* Q;          // workspace [nscreens][2 (m - m_split)][n]
    double* polyc;      // workspace [nscreens][(degree+1)^2]
    void* turns;        // out [nscreens][n][n] float or double, may be null
    int turns_f64;
    void* phi;          // out, optional full phase [nscreens][n][n] float or double
    int phi_f64;
    double p_scale;     // tensor-core path: power-of-two scale of the P operand (fp16 range)
    // uniform axes underlying the float32 ones (tensor-core path): xu_j = x_first + j * dxu, shift included
    double x_first, dxu, y_first, dyu;
};

int screen_init_constants();
int launch_screen_exact(const ScreenLaunch& a, cudaStream_t st);
int launch_screen_poly(const ScreenLaunch& a, cudaStream_t st);
size_t screen_tc_workspace(int n, int m, int m_split, int nscreens);
// factors_stream (optional): the operand generation of the preparation phase is launched there instead of `st`; the caller
// has made it wait for everything earlier on `st` and joins it back before the contraction (api.cu: screens()).
int launch_screen_tc(const ScreenLaunch& a, void* workspace, int* err_flag, int num_sms, int swap, cudaStream_t st, int phase = 2,
                     int first_screen = 0, int total_screens = -1, cudaStream_t factors_stream = nullptr,
                     cudaEvent_t factors_done = nullptr);

}  // namespace pa
