#define PA_N 8192
#include "fft_inst.inc"
