// Row and column passes of the split-step propagator (see DESIGN.md "FFT passes").
//
// The field lives in HBM as [batch][N rows (y)][N columns (x)] complex.  Between passes it is either in
// NATURAL order (space domain) or in ROW-SPECTRUM form U~(y, kx): space in y, permuted frequency in x.
//
//   k_rows : per row   [source | load natural | load permuted -> IFFT_x] -> [* scale * exp(-2 pi i turns)]
//                      -> [FFT_x -> store permuted | store natural]
//   k_cols : per column  load -> FFT_y -> * H(ky,kx) -> IFFT_y -> store      (always in place)
//            INV_ONLY: load a column spectrum (storage order) -> * real scale -> IFFT_y -> store natural; with a
//            k_rows<IN_PERM, !OUT_PERM> launch this is a plain inverse 2-D DFT (FFT phase screens, screen_fft.cu)
//
// One vacuum leg = FFT_x, (FFT_y, H, IFFT_y), IFFT_x; the trailing IFFT_x of a leg, the screen multiply and
// the leading FFT_x of the next leg run in ONE k_rows launch, so a steady-state split-step stage is two
// read+write sweeps of the field.
#pragma once
#include "fft_core.cuh"
#include "internal_measure.h"

namespace pa {

// ---- shared-memory address maps ----------------------------------------------------------------------------
// XOR swizzles that keep runs aligned to their own size intact and spread the strided accesses of the
// late stages over the banks (checked with tools/bank_sim2.py).

// Rows (complex64, E = 16; one 128-byte bank row = 16 elements).  Stage with SIGMA = 1: register pairs move as
// 16-byte chunks, a quarter warp needs 8 distinct chunks -> thread bits enter chunk bits 1-3.  Stages with
// 1 < SIGMA < 16: a half warp needs 16 distinct slots -> the thread bits above SIGMA enter the free slot bits.
// Sources are always bits above the targets, so every map is a bijection of [0, N).
template <int N, int E> __device__ __forceinline__ int swz_row(int p) {
    // (found with tools/bank_sim2.py for the last-stage thread assignment of fft_core.cuh: plan_last_butterfly)
    if constexpr (E == 16 && N == 1024) {           // 16.16.4, SIGMA = 64, 4, 1
        return p ^ (((p >> 6) & 1) << 1) ^ (((p >> 5) & 7) << 1);
    } else if constexpr (E == 16 && (N == 4096 || N == 256)) {   // last radix 16, SIGMA = .., 16, 1
        return p ^ (((p >> 4) & 7) << 1);
    } else if constexpr (E == 16 && (N == 8192 || N == 512)) {   // .., 8, 4 with SIGMA = .., 4, 1
        return p ^ (((p >> 5) & 1) << 1) ^ (((p >> 4) & 7) << 1);
    } else if constexpr (E == 16) {
        // 2048 = 16.16.8: bits 1-2 <- bits 4-5 (SIGMA = 1, 16-byte chunks), bit 3 <- bit 7 (SIGMA = 8)
        return p ^ (((p >> 4) & 3) << 1) ^ (((p >> 7) & 1) << 3);
    } else if constexpr (N == 512 || N == 4096) {   // complex128 (E = 8, 16-byte elements: 8 per bank row), radices 8.8.(8.)8
        return p ^ ((p >> 3) & 7);
    } else if constexpr (N == 1024 || N == 8192) {  // 8.8.(8.)4.4
        return p ^ ((p >> 4) & 1) ^ (((p >> 3) & 3) << 1);
    } else {                                        // 256 = 8.8.4, 2048 = 8.8.8.4
        return p ^ ((p >> 4) & 3) ^ ((p >> 3) & 7);
    }
}
// Columns: a tile keeps TC columns interleaved, so a half warp (quarter warp for complex128) covers 16/TC (8/TC) consecutive
// threads of one transform; in the last stage (SIGMA = 1) thread t owns runs of R_last positions that start at multiples of
// R_last -> those bits are folded onto the low bits.  Every other stage is conflict-free in the natural order.
template <int N, int E, int TC> __device__ __forceinline__ int swz_col(int p) {
    constexpr int RL = plan_radix(N, E, plan_len(N, E) - 1);
    constexpr int SLOTS = (E == 16 || PA_E32 == 8) ? 4 : 3;          // log2 of the elements per 128-byte bank row (PA_E32 == 8: experiment)
    constexpr int W = ilog2(TC) >= SLOTS ? 0 : SLOTS - ilog2(TC);
    if constexpr (W == 0) {
        return p;
    } else if constexpr (E == 16 && N == 8192 && TC == 1) {            // SIGMA = 4 stage as well: bit 6 -> bit 2
        return p ^ ((p >> 2) & 15) ^ (((p >> 6) & 1) << 2);
    } else {
        return p ^ ((p >> ilog2(RL)) & ((1 << W) - 1));
    }
}

// `sync()` is the barrier of one exchange.  The rows of a CTA are independent transforms in private shared-memory
// regions, so when a row is made of whole warps its N/E threads synchronise on a named barrier of their own (id 1 + row
// in CTA) and the rows of a CTA drift apart instead of all waiting for the slowest warp of the block.
template <int N, int E> struct RowAddr {
    static constexpr bool kContiguous = true;
    static constexpr int TPF = N / E;
    // the R_last adjacent threads that exchange between the last two stages always share a warp (rows start at multiples
    // of TPF, R_last divides TPF and 32)
    static constexpr bool kLocalLast = plan_len(N, E) >= 2;
    static constexpr bool kDualLayout = false;
    static constexpr int kLow = E == 16 ? 0xE : 0xF;       // bits the row swizzles write (complex64 keeps 16-byte pairs intact)
    int base;
    __device__ __forceinline__ int operator()(int p) const { return base + swz_row<N, E>(p); }
    template <int S, bool LOCAL> __device__ __forceinline__ int at(int t, int idx) const {
        const int vt = swz_row<N, E>(reg_pos<N, E, S>(t, 0));
        const int k = swz_row<N, E>(reg_pos<N, E, S>(0, idx));          // compile-time constant once unrolled
        return base + ((vt ^ (k & kLow)) + (k & ~kLow));
    }
    __device__ __forceinline__ void sync() const {
#ifndef PA_ROW_BLOCK_BARRIER
        if constexpr (TPF % 32 == 0) {
            const int row_in_cta = base / N;
            if (row_in_cta < 15) {
                asm volatile("bar.sync %0, %1;" ::"r"(row_in_cta + 1), "n"(TPF) : "memory");
                return;
            }
        }
#endif
        __syncthreads();
    }
};
template <int N, int E, int TC> struct ColAddr {
    static constexpr bool kContiguous = false;
    // thread index = c + TC * t: the R_last adjacent t of all TC columns are R_last * TC consecutive threads
    static constexpr bool kLocalLast = plan_len(N, E) >= 2 && plan_radix(N, E, plan_len(N, E) - 1) * TC <= 32;
    static constexpr bool kDualLayout = false;
    static constexpr int kLow = 0xF;       // bits the column swizzles write (position space)
    int c;
    __device__ __forceinline__ int operator()(int p) const { return swz_col<N, E, TC>(p) * TC + c; }
    template <int S, bool LOCAL> __device__ __forceinline__ int at(int t, int idx) const {
        const int vt = swz_col<N, E, TC>(reg_pos<N, E, S>(t, 0));
        const int k = swz_col<N, E, TC>(reg_pos<N, E, S>(0, idx));
        return ((vt ^ (k & kLow)) * TC + c) + (k & ~kLow) * TC;
    }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
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
    // separable start (SRC = true and sep != nullptr): the field is  sep_scale * sep[row] * sep[column]; used for
    // the Gaussian source carried analytically through the first vacuum leg (see api.cu: first_leg_table)
    const cplx<T>* sep;
    T sep_re, sep_im;
    // fused reductions (MEAS = true, natural-order output): per-row sums {I, I x, I x^2, I [inside aperture p]} are
    // written to rowsums[row][kRowSums]; k_measure_finish_rows folds them with the y weights.  store = 0 skips
    // writing the field itself (Monte-Carlo runs that only need the statistics).
    double* rowsums;
    const float* pupils;   // [npupil][3] {r^2, sx, sy} float32, npupil <= kFusedPupils
    int npupil;
    int store;
};

// One CTA = FPB rows, N/E threads per row.
// Resident warps per SM the complex64 row pass is compiled for (register cap = 64 K / (32 * warps)): the pass overlaps the
// shared-memory phases of some rows with the arithmetic of others only if enough independent rows are resident.
// Measured per size (tools/gpu/fft_variants.py, us per pass at 2048^2 x 8 / 4096^2 x 2 / 8192^2): 24 resident warps (80
// registers) + screen values requested before the last inverse stage: 166 -> 131 / 168 -> 150; 8192^2 (one 512-thread CTA per
// row) is best left alone (481 against 489); 32 warps (64 registers) spills: 171.
#ifndef PA_ROW_MINWARPS
#define PA_ROW_MINWARPS 24
#endif
template <typename T, int N> struct RowTune {
    static constexpr bool kTuned = sizeof(T) == 4 && (N == 2048 || N == 4096);
    static constexpr int kMinWarps = kTuned ? PA_ROW_MINWARPS : 0;
#ifdef PA_NO_PREFETCH_TURNS
    static constexpr bool kPrefetchTurns = false;
#else
    static constexpr bool kPrefetchTurns = kTuned;    // request the screen values before the last inverse stage
#endif
};
template <typename T, int N, int E, int FPB> __host__ __device__ constexpr int row_min_blocks() {
    constexpr int warps_per_cta = FPB * (N / E) / 32;
    if (warps_per_cta == 0 || RowTune<T, N>::kMinWarps == 0) return 1;
    return RowTune<T, N>::kMinWarps / warps_per_cta > 0 ? RowTune<T, N>::kMinWarps / warps_per_cta : 1;
}
// MEAS: 0 = plain pass; 1 + NP = final pass with the reductions fused in, NP apertures (0..kFusedPupils) compiled in
template <typename T, int N, int E, int FPB, bool IN_PERM, bool OUT_PERM, bool SRC, int MEAS = 0>
__global__ void __launch_bounds__(FPB * (N / E), row_min_blocks<T, N, E, FPB>()) k_rows(RowArgs<T> a) {
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
    constexpr bool PF = RowTune<T, N>::kPrefetchTurns && IN_PERM;
    T trv[PF ? E : 1];

    if constexpr (SRC) {
      if (a.sep != nullptr) {
        const C sy = cmul(ldg_c<T>(a.sep + row % N), mkc<T>(a.sep_re, a.sep_im));
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = cmul(ldg_c<T>(a.sep + reg_pos<N, E, 0>(t, i)), sy);
      } else {
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
      }
    } else if constexpr (IN_PERM) {
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = ptr[io_pos<N, E>(t, i)];      // spectrum in storage order
        if constexpr (PF) {
            fft_inv_head<T, N, E>(v, t, sm, addr, a.tw);
            if (a.turns != nullptr) {          // screen values requested before the last inverse stage hides their latency
                const T* tr = a.turns + (size_t)row * N;
#pragma unroll
                for (int i = 0; i < E; ++i) trv[i] = tr[reg_pos<N, E, 0>(t, i)];
            }
            fft_inv_tail<T, N, E>(v, t, a.tw);
        } else {
            fft_inv<T, N, E>(v, t, sm, addr, a.tw);
        }
    } else {
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = ptr[reg_pos<N, E, 0>(t, i)];
    }

    // registers now hold natural positions p = reg_pos<0>(t, i)
    if (a.turns != nullptr) {
        const T* tr = a.turns + (size_t)row * N;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            C e = expm2pi(PF ? trv[PF ? i : 0] : tr[reg_pos<N, E, 0>(t, i)]);
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
        if (MEAS == 0 || a.store) {
#pragma unroll
            for (int i = 0; i < E; ++i) ptr[reg_pos<N, E, 0>(t, i)] = v[i];
        }
        if constexpr (MEAS != 0) {
            // same arithmetic as k_measure_partial (measure.cu): intensity and x-weights in the field's precision,
            // aperture predicate in float32 without FMA contraction (pupils.py:10)
            constexpr int NP = MEAS - 1;                           // apertures reduced here
            constexpr int NS = 3 + NP;                             // live sums
            const float yv = a.y[row % N];
            float pr2[NP > 0 ? NP : 1], psx[NP > 0 ? NP : 1], dy2[NP > 0 ? NP : 1];
            T acc[NS];
#pragma unroll
            for (int q = 0; q < NS; ++q) acc[q] = 0;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                pr2[p] = a.pupils[3 * p];
                psx[p] = a.pupils[3 * p + 1];
                const float dy = __fadd_rn(yv, a.pupils[3 * p + 2]);
                dy2[p] = __fmul_rn(dy, dy);
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const float xv = a.x[reg_pos<N, E, 0>(t, i)];
                const T in = v[i].x * v[i].x + v[i].y * v[i].y;
                acc[0] += in;
                acc[1] += in * (T)xv;
                acc[2] += in * ((T)xv * (T)xv);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float dx = __fsub_rn(xv, psx[p]);
                    acc[3 + p] += (__fadd_rn(__fmul_rn(dx, dx), dy2[p]) <= pr2[p]) ? in : (T)0;
                }
            }
            // reduce over the N/E threads of the row: shuffles, then one slot per warp in shared memory
            double red[NS];
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                double r = (double)acc[q];
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                red[q] = r;
            }
            constexpr int WPR = TPF / 32 > 0 ? TPF / 32 : 1;      // warps per row (TPF >= 32 in the fused configuration)
            __syncthreads();                                       // exchanges are over: reuse the row's smem
            double* sred = reinterpret_cast<double*>(smem_raw) + (size_t)f * (N * sizeof(C) / sizeof(double));
            if ((t & 31) == 0) {
#pragma unroll
                for (int q = 0; q < NS; ++q) sred[(t >> 5) * kRowSums + q] = red[q];
            }
            __syncthreads();
            if (t < kRowSums) {
                double r = 0.0;
                if (t < NS)
                    for (int w = 0; w < WPR; ++w) r += sred[w * kRowSums + t];
                a.rowsums[(size_t)row * kRowSums + t] = r;
            }
        }
    }
}

template <typename T> struct ColArgs {
    cplx<T>* field;        // [batch][N][N] in place, row-spectrum form
    const cplx<T>* tw;
    const cplx<T>* hp;     // transfer-function factor h[freq(q)] in spectrum storage order, N entries (kx, and ky unless hpy is set)
    T alpha_re, alpha_im;  // e^{ikL} / N^2 (times any loss factor)
    // split column pass (fft_split.cuh): the kernel transforms blocks of N rows of a field that is `ncols` wide;
    // block b uses the ky factors hpy[(b % nsub) * N ...].  Ordinary pass: hpy = hp, ncols = N, nsub = 1.
    const cplx<T>* hpy;
    int ncols, nsub;
};

// One CTA = TC adjacent columns of one field, N/E threads per column; thread index = c + TC * t so that a warp
// touches TC*sizeof(C)-byte segments of 32/TC consecutive rows.
template <typename T, int N, int E, int TC, bool INV_ONLY = false>
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
    if constexpr (INV_ONLY) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
            v[i] = ptr[(size_t)io_pos<N, E>(t, i) * N];
            v[i].x *= a.alpha_re;
            v[i].y *= a.alpha_re;
        }
    } else {
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
    }

    fft_inv<T, N, E>(v, t, sm, addr, a.tw);
#pragma unroll
    for (int i = 0; i < E; ++i) ptr[(size_t)reg_pos<N, E, 0>(t, i) * N] = v[i];
}

}  // namespace pa
