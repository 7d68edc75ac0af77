#define PA_N 256
#include "fft_inst.inc"
