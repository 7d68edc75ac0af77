#define PA_N 128
#include "fft_inst.inc"
