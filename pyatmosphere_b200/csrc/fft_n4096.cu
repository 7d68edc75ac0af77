#define PA_N 4096
#include "fft_inst.inc"
