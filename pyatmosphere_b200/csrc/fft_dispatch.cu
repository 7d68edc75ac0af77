// Size dispatch for the FFT passes.
#include "internal.h"

namespace pa {

int launch_rows(int prec, int n, const RowLaunch& a, cudaStream_t st) {
    switch (n) {
#define PA_CASE(N) case N: return launch_rows_##N(prec, a, st);
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return -1;
    }
}
int launch_cols(int prec, int n, const ColLaunch& a, cudaStream_t st) {
    switch (n) {
#define PA_CASE(N) case N: return launch_cols_##N(prec, a, st);
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return -1;
    }
}
bool fft_size_supported(int prec, int n) {
    switch (n) {
#define PA_CASE(N) case N: return true;
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return false;
    }
}
bool fft_tma_supported(int prec, int n) {
    switch (n) {
#define PA_CASE(N) case N: return fft_tma_ok_##N(prec);
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return false;
    }
}
int fft_tma_cols_per_tile(int prec, int n) {
    switch (n) {
#define PA_CASE(N) case N: return fft_tma_tc_##N(prec);
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return 0;
    }
}
int fft_split_cols_per_tile(int prec, int n) {
    switch (n) {
#define PA_CASE(N) case N: return fft_split_tc_##N(prec);
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: return 0;
    }
}
int fft_split_radix(int prec, int n) { return fft_split_cols_per_tile(prec, n) > 0 ? kSplitRadix : 0; }
void fft_geometry(int prec, int n, int* rt, int* rf, int* rs, int* ct, int* cc, int* cs) {
    int g[6] = {0, 0, 0, 0, 0, 0};
    switch (n) {
#define PA_CASE(N) case N: fft_geometry_##N(prec, g); break;
        PA_FFT_SIZES(PA_CASE)
#undef PA_CASE
        default: break;
    }
    *rt = g[0]; *rf = g[1]; *rs = g[2]; *ct = g[3]; *cc = g[4]; *cs = g[5];
}

}  // namespace pa
