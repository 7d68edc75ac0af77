#define PA_N 1024
#include "fft_inst.inc"
