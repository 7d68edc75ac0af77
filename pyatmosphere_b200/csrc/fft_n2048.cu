#define PA_N 2048
#include "fft_inst.inc"
