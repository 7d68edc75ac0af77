// Internal (non-ABI) interfaces between the translation units of libpyatm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pa {

// elements per thread of the register FFT, per precision (0 = complex64, 1 = complex128)
#ifndef PA_E32
#define PA_E32 16
#endif
constexpr int kE32 = PA_E32;
constexpr int kE64 = 8;
inline int elems_per_thread(int prec) { return prec == 0 ? kE32 : kE64; }

struct RowLaunch {
    void* field;          // cplx<T>* [rows_total][n]
    const void* tw;       // cplx<T>* twiddles
    const void* turns;    // T* or null
    double scale;
    int rows_total;
    bool in_perm, out_perm, src;
    const float* x;
    const float* y;
    double amp, aw, ac;
    const void* sep;      // separable start table (cplx<T>[n]) or null
    double* rowsums;      // fused reductions (in_perm && !out_perm only): per-row sums, or null
    const float* pupils;
    int npupil;
    int store;
    double sep_re, sep_im;
    bool use_tma;         // persistent TMA-fed variant (ignored where it does not exist)
    int num_sms;
};
struct ColLaunch {
    void* field;          // cplx<T>* [batch][n][n]
    const void* tw;
    const void* hp;       // cplx<T>* permuted transfer-function factor
    double alpha_re, alpha_im;
    int batch;
    const void* tmap;     // host pointer to a CUtensorMap of the field ([batch*n][2n] reals), or null -> direct-access kernel
    int num_sms;
    // split column pass (fft_split.cuh; set together, tmap then has the box of the inner transform's tiles)
    const void* hpy = nullptr;      // ky factors in the composite order of the split transform
    const void* tw_sub = nullptr;   // stage twiddles of the inner (n / 32)-point plan
    const void* otw = nullptr;      // outer-stage twiddles exp(-2 pi i t j / n), [n / 32][32]
    // inverse-only pass (direct kernel): column spectra in storage order -> alpha_re * IFFT_y, natural order out
    bool inv_only = false;
};

// implemented once per grid size in fft_n<N>.cu; return cudaError_t as int, or -1 for an unsupported size
int launch_rows(int prec, int n, const RowLaunch& a, cudaStream_t st);
int launch_cols(int prec, int n, const ColLaunch& a, cudaStream_t st);
bool fft_size_supported(int prec, int n);
bool fft_tma_supported(int prec, int n);
int fft_tma_cols_per_tile(int prec, int n);
// split column pass: outer radix (0 = this size/precision has no split variant) and columns per tile of the inner kernel
constexpr int kSplitRadix = 32;
int fft_split_radix(int prec, int n);
int fft_split_cols_per_tile(int prec, int n);
// number of CTAs / threads / smem of the two passes (reported through pa_fft_geometry for the roofline notes)
void fft_geometry(int prec, int n, int* rows_threads, int* rows_fpb, int* rows_smem, int* cols_threads, int* cols_tc, int* cols_smem);

#define PA_FFT_SIZES(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)
#define PA_DECL(N)                                                                   \
    int launch_rows_##N(int prec, const RowLaunch& a, cudaStream_t st);              \
    int launch_cols_##N(int prec, const ColLaunch& a, cudaStream_t st);              \
    void fft_geometry_##N(int prec, int* g);                                         \
    bool fft_tma_ok_##N(int prec);                                                   \
    int fft_tma_tc_##N(int prec);                                                    \
    int fft_split_tc_##N(int prec);
PA_FFT_SIZES(PA_DECL)
#undef PA_DECL

}  // namespace pa
