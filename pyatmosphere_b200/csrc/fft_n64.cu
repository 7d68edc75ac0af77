#define PA_N 64
#include "fft_inst.inc"
