// Row and column passes of the split-step propagator (see DESIGN.md "FFT passes").
//
// The field lives in HBM as [batch][N rows (y)][N columns (x)] complex.  Between passes it is either in
// NATURAL order (space domain) or in ROW-SPECTRUM form U~(y, kx): space in y, permuted frequency in x.
//
//   k_rows : per row   [source | load natural | load permuted -> IFFT_x] -> [* scale * exp(-2 pi i turns)]
//                      -> [FFT_x -> store permuted | store natural]
//   k_cols : per column  load -> FFT_y -> * H(ky,kx) -> IFFT_y -> store      (always in place)
//
// One vacuum leg = FFT_x, (FFT_y, H, IFFT_y), IFFT_x; the trailing IFFT_x of a leg, the screen multiply and
// the leading FFT_x of the next leg run in ONE k_rows launch, so a steady-state split-step stage is two
// read+write sweeps of the field.
#pragma once
#include "fft_core.cuh"

namespace pa {

// ---- shared-memory address maps ----------------------------------------------------------------------------
// XOR swizzles that keep runs aligned to their own size intact and spread the strided accesses of the
// late stages over the banks (checked with tools/bank_sim.py).
template <int N, int E> __device__ __forceinline__ int swz_row(int p) {
    // bits 1-2 <- bits 4-5 (stage with SIGMA = 1, 16-byte chunks), bit 3 <- bit 7 (stage with SIGMA = 8)
    return p ^ (((p >> 4) & 3) << 1) ^ (((p >> 7) & 1) << 3);
}
template <int N, int E, int TC> __device__ __forceinline__ int swz_col(int p) {
    return p ^ ((p >> 3) & 3);
}

template <int N, int E> struct RowAddr {
    static constexpr bool kContiguous = true;
    int base;
    __device__ __forceinline__ int operator()(int p) const { return base + swz_row<N, E>(p); }
};
template <int N, int E, int TC> struct ColAddr {
    static constexpr bool kContiguous = false;
    int c;
    __device__ __forceinline__ int operator()(int p) const { return swz_col<N, E, TC>(p) * TC + c; }
};

// exp(-2 pi i t) for t in turns.  float: MUFU sin/cos (|abs err| < 4e-7 on [-pi, pi]); double: sincospi.
__device__ __forceinline__ float2 expm2pi(float turns) {
    float s, c;
    const float a = -6.283185307179586f * turns;
    __sincosf(a, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ double2 expm2pi(double turns) {
    double s, c;
    sincospi(-2.0 * turns, &s, &c);
    return make_double2(c, s);
}

template <typename T> struct RowArgs {
    cplx<T>* field;        // [rows_total][N]  in/out (in place)
    const cplx<T>* tw;     // concatenated stage twiddles
    const T* turns;        // screen phase in turns (phi / 2 pi, any integer part allowed), [rows_total][N]; may be null
    T scale;               // real amplitude factor applied together with the screen (dB losses)
    int rows_total;        // batch * N
    // source generation (SRC = true): u0 = amp * exp(-(aw + i ac) * rho2), rho2 = x^2 + y^2 in float32 as the
    // reference builds it (grids.py:74-76, theory/sources.py:16-18)
    const float* x;
    const float* y;
    double amp, aw, ac;
};

// One CTA = FPB rows, N/E threads per row.
template <typename T, int N, int E, int FPB, bool IN_PERM, bool OUT_PERM, bool SRC>
__global__ void __launch_bounds__(FPB * (N / E)) k_rows(RowArgs<T> a) {
    using C = cplx<T>;
    constexpr int TPF = N / E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);
    const int f = threadIdx.x / TPF;
    const int t = threadIdx.x % TPF;
    const int row = blockIdx.x * FPB + f;     // grid is sized so that row < rows_total
    C* ptr = a.field + (size_t)row * N;
    const RowAddr<N, E> addr{f * N};
    C v[E];

    if constexpr (SRC) {
        const float yv = a.y[row % N];
        const float y2 = __fmul_rn(yv, yv);
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int p = reg_pos<N, E, 0>(t, i);
            const float xv = a.x[p];
            const double rho2 = (double)__fadd_rn(__fmul_rn(xv, xv), y2);
            const double mag = a.amp * exp(-a.aw * rho2);
            if (a.ac != 0.0) {
                double s, c;
                sincos(-a.ac * rho2, &s, &c);
                v[i] = mkc<T>((T)(mag * c), (T)(mag * s));
            } else {
                v[i] = mkc<T>((T)mag, (T)0);
            }
        }
    } else if constexpr (IN_PERM) {
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = ptr[io_pos<N, E>(t, i)];      // spectrum in storage order
        fft_inv<T, N, E>(v, t, sm, addr, a.tw);
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = ptr[reg_pos<N, E, 0>(t, i)];
    }

    // registers now hold natural positions p = reg_pos<0>(t, i)
    if (a.turns != nullptr) {
        const T* tr = a.turns + (size_t)row * N;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            C e = expm2pi(tr[reg_pos<N, E, 0>(t, i)]);
            e.x *= a.scale;
            e.y *= a.scale;
            v[i] = cmul(v[i], e);
        }
    } else if (a.scale != (T)1) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
            v[i].x *= a.scale;
            v[i].y *= a.scale;
        }
    }

    if constexpr (OUT_PERM) {
        fft_fwd<T, N, E>(v, t, sm, addr, a.tw);
#pragma unroll
        for (int i = 0; i < E; ++i) ptr[io_pos<N, E>(t, i)] = v[i];
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) ptr[reg_pos<N, E, 0>(t, i)] = v[i];
    }
}

template <typename T> struct ColArgs {
    cplx<T>* field;        // [batch][N][N] in place, row-spectrum form
    const cplx<T>* tw;
    const cplx<T>* hp;     // transfer-function factor h[freq(q)] in spectrum storage order, N entries
    T alpha_re, alpha_im;  // e^{ikL} / N^2 (times any loss factor)
};

// One CTA = TC adjacent columns of one field, N/E threads per column; thread index = c + TC * t so that a warp
// touches TC*sizeof(C)-byte segments of 32/TC consecutive rows.
template <typename T, int N, int E, int TC>
__global__ void __launch_bounds__(TC * (N / E)) k_cols(ColArgs<T> a) {
    using C = cplx<T>;
    constexpr int L = plan_len(N, E);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);
    const int c = threadIdx.x % TC;
    const int t = threadIdx.x / TC;
    const int col = blockIdx.x * TC + c;
    C* ptr = a.field + (size_t)blockIdx.y * N * N + col;
    const ColAddr<N, E, TC> addr{c};
    C v[E];
#pragma unroll
    for (int i = 0; i < E; ++i) v[i] = ptr[(size_t)reg_pos<N, E, 0>(t, i) * N];

    fft_fwd<T, N, E>(v, t, sm, addr, a.tw);

    {   // multiply by H = alpha * h[ky] * h[kx]; registers are in the last-stage distribution
        const C hx = cmul(ldg_c<T>(a.hp + col), mkc<T>(a.alpha_re, a.alpha_im));
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const C h = cmul(ldg_c<T>(a.hp + io_pos<N, E>(t, i)), hx);     // register (t,i) holds storage index t + i*T
            v[i] = cmul(v[i], h);
        }
    }

    fft_inv<T, N, E>(v, t, sm, addr, a.tw);
#pragma unroll
    for (int i = 0; i < E; ++i) ptr[(size_t)reg_pos<N, E, 0>(t, i) * N] = v[i];
}

}  // namespace pa
