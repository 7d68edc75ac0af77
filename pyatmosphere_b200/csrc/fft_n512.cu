#define PA_N 512
#include "fft_inst.inc"
